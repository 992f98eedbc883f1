// UpcRootFile.h -- a ROOT-less WRITER of the two kinds of .root files the reference produces on this path (and of TH1D
// objects, the form of the elementary cross sections sigma(m) the reference reads from cross_sections/*/cross_section_m.root):
//   * the two-photon-luminosity cache twoPhotonLumi[Pol].root with the TH2D hD2LDMDY[_s,_p]
//     (src/UpcCrossSection.cpp:493-507, :564-585), so that a table filled on the GPU is picked up by the reference
//     (and by this build) exactly like one the reference computed itself;
//   * events.root with the TTree "particles" and its nine branches (src/UpcGenerator.cpp:842-857, fill :595-608).
//
// The byte layout follows ROOT's file format (TFile header, TKey records, the directory / key-list / free-segment /
// streamer-info records, TBuffer streaming with byte counts and class versions) as ROOT 6.22-6.3x writes it; the
// TH1/TH2/TAxis member sequence was checked byte by byte against the histograms in the reference's own
// cross_sections/*.root files (written by ROOT 6.22/09).  Objects are written uncompressed unless SetCompression asks for LZ4.  ROOT itself is not
// available in this build environment: what reads these files back here is upcgen_b200/host/UpcRootHist.cpp and an
// independent Python parser (tests/test_root_file.py).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

class UpcRootFileWriter
{
 public:
  UpcRootFileWriter();  // takes the process-wide default compression setting (see UpcRootFileDefaultCompression)
  // a TH2D with uniform axes; cells: (nx + 2) * (ny + 2) doubles, x fastest, under-/overflow cells included
  void AddTH2D(const std::string& name, const std::string& title, int nx, double xlo, double xhi, int ny, double ylo,
               double yhi, const std::vector<double>& cells, double entries);

  // a TH1D with a uniform axis; cells: nx + 2 doubles, under-/overflow included
  void AddTH1D(const std::string& name, const std::string& title, int nx, double xlo, double xhi,
               const std::vector<double>& cells, double entries);

  // the same with variable bins: edges hold n + 1 increasing values per axis (TH2D(name, title, nx, xbins, ny, ybins));
  // stats: fTsumw, fTsumw2, fTsumwx, fTsumwx2 of the TH1D (null: zero, a histogram filled by SetBinContent)
  void AddTH2D(const std::string& name, const std::string& title, const std::vector<double>& xedges,
               const std::vector<double>& yedges, const std::vector<double>& cells, double entries);
  void AddTH1D(const std::string& name, const std::string& title, const std::vector<double>& xedges,
               const std::vector<double>& cells, double entries, const double* stats = nullptr);

  // a TTree of flat branches, one leaf each: type 'I' (Int_t) or 'D' (Double_t); all columns the same length.
  // Integer columns are given as doubles holding integral values.
  struct Column {
    std::string name;
    char type;  // 'I' or 'D'
    std::vector<double> values;
  };
  void AddTree(const std::string& name, const std::string& title, const std::vector<Column>& columns);

  // ROOT's compression setting (100 * algorithm + level).  0 (the default here): records are stored as they are.
  // 1xx: zlib "ZL" records (101 = ROOT's default setting; the framing of the reference's own cross_sections/*.root).
  // 4xx: LZ4 -- what the reference asks for in events.root (409, src/UpcGenerator.cpp:843): objects above 256 bytes
  // and basket buffers are stored as "L4" records (UpcLz4.h) when that makes them smaller, as ROOT does.  The blocks
  // come from a greedy matcher, not LZ4-HC level 9: any LZ4 decoder reads them, they are somewhat larger.
  void SetCompression(int setting) { compression_ = setting; }

  // writes the file; returns false and sets err on failure
  bool Write(const std::string& path, std::string& err);

 private:
  struct Record {
    std::string cls, name, title;
    std::vector<unsigned char> data;  // streamed object
    bool listed;                      // appears in the directory's key list (baskets do not)
    uint32_t seek = 0;
    uint32_t objlen = 0;              // set by Write(): length of the object before compression (fObjlen)
  };
  int compression_;
  std::vector<Record> records_;
  // baskets of the trees are written before the tree's own record; the tree streamer needs their positions, which
  // are fixed once the layout is known: trees are therefore streamed inside Write()
  struct Tree {
    std::string name, title;
    std::vector<Column> columns;
  };
  std::vector<Tree> trees_;
};

// The three histograms the reference adds to events.root at debug level > 0 (src/UpcGenerator.cpp:900-917): the
// nuclear cross section table "hNucCSYM" (TH2D over the y and m bin edges) and its projections "hNucCSYM_py" (over m)
// and "hNucCSYM_px" (over y), in the reference's order of writing.  nucCSYM: [ny][nm].
void UpcAddSigmaHists(UpcRootFileWriter& w, const std::vector<double>& yEdges, const std::vector<double>& mEdges,
                      const std::vector<std::vector<double>>& nucCSYM);

// The compression setting new writers start from: 0 unless set here or, at first use, by the environment variable
// UPCGEN_ROOT_COMPRESSION (e.g. 409 = the reference's LZ4 setting for events.root).  Returns the value in force;
// a negative argument only queries.
int UpcRootFileDefaultCompression(int setting = -1);
