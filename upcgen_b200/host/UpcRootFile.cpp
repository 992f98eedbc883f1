// UpcRootFile.cpp -- see UpcRootFile.h.  Nothing here links against or is copied from ROOT; the layout is ROOT's
// documented file format (TFile / TKey / TDirectory records, TBufferFile streaming).
#include "UpcRootFile.h"
#include "UpcLz4.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>

namespace {

constexpr uint32_t kByteCountMask = 0x40000000u;
constexpr uint32_t kNewClassTag = 0xFFFFFFFFu;
constexpr uint32_t kClassMask = 0x80000000u;
constexpr uint32_t kMapOffset = 2;

struct Out {
  std::vector<unsigned char> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { b.push_back((unsigned char)(v >> 8)); b.push_back((unsigned char)v); }
  void u32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
  void u64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
  void i16(int16_t v) { u16((uint16_t)v); }
  void i32(int32_t v) { u32((uint32_t)v); }
  void i64(int64_t v) { u64((uint64_t)v); }
  void f32(float v) { uint32_t x; std::memcpy(&x, &v, 4); u32(x); }
  void f64(double v) { uint64_t x; std::memcpy(&x, &v, 8); u64(x); }
  void str(const std::string& s)  // TString
  {
    if (s.size() < 255) u8((uint8_t)s.size());
    else { u8(255); u32((uint32_t)s.size()); }
    b.insert(b.end(), s.begin(), s.end());
  }
  void cstr(const std::string& s) { b.insert(b.end(), s.begin(), s.end()); b.push_back(0); }
  void raw(const std::vector<unsigned char>& v) { b.insert(b.end(), v.begin(), v.end()); }
  size_t size() const { return b.size(); }
  // byte count + class version in front of a streamed object; end() patches the count
  size_t begin(uint16_t version) { const size_t p = b.size(); u32(0); u16(version); return p; }
  void end(size_t p)
  {
    const uint32_t v = (uint32_t)(b.size() - p - 4) | kByteCountMask;
    b[p] = (unsigned char)(v >> 24); b[p + 1] = (unsigned char)(v >> 16); b[p + 2] = (unsigned char)(v >> 8); b[p + 3] = (unsigned char)v;
  }
};

uint32_t datime_now()
{
  std::time_t t = std::time(nullptr);
  std::tm lt{};
  localtime_r(&t, &lt);
  return (uint32_t)((lt.tm_year + 1900 - 1995) << 26 | (lt.tm_mon + 1) << 22 | lt.tm_mday << 17 | lt.tm_hour << 12 |
                    lt.tm_min << 6 | lt.tm_sec);
}

void tobject(Out& o, uint32_t bits)
{
  o.u16(1);  // TObject version
  o.u32(0);  // fUniqueID
  o.u32(bits);
}
void tnamed(Out& o, const std::string& name, const std::string& title, uint32_t bits = 0x03000000u)
{
  const size_t p = o.begin(1);
  tobject(o, bits);
  o.str(name);
  o.str(title);
  o.end(p);
}
void tattline(Out& o) { const size_t p = o.begin(2); o.i16(602); o.i16(1); o.i16(1); o.end(p); }
void tattfill(Out& o) { const size_t p = o.begin(2); o.i16(0); o.i16(1001); o.end(p); }
void tattmarker(Out& o) { const size_t p = o.begin(2); o.i16(1); o.i16(1); o.f32(1.0f); o.end(p); }

void taxis(Out& o, const std::string& name, int nbins, double lo, double hi, float title_offset,
           const std::vector<double>* edges = nullptr)
{
  const size_t p = o.begin(10);
  tnamed(o, name, "");
  {
    const size_t q = o.begin(4);  // TAttAxis
    o.i32(510);                   // fNdivisions
    o.i16(1); o.i16(1); o.i16(42);  // fAxisColor, fLabelColor, fLabelFont
    o.f32(0.005f); o.f32(0.035f); o.f32(0.03f); o.f32(title_offset); o.f32(0.035f);  // label offset/size, tick, title offset/size
    o.i16(1); o.i16(42);          // fTitleColor, fTitleFont
    o.end(q);
  }
  o.i32(nbins);
  o.f64(lo);
  o.f64(hi);
  if (edges && !edges->empty()) {  // fXbins (TArrayD): the nbins + 1 edges of an axis with variable bins
    o.i32((int32_t)edges->size());
    for (double e : *edges) o.f64(e);
  } else {
    o.i32(0);           // no variable edges
  }
  o.i32(0); o.i32(0);   // fFirst, fLast
  o.u16(0);             // fBits2
  o.u8(0);              // fTimeDisplay
  o.str("");            // fTimeFormat
  o.u32(0);             // fLabels (null)
  o.u32(0);             // fModLabs (null)
  o.end(p);
}

void empty_tlist(Out& o, uint32_t bits)
{
  const size_t p = o.begin(5);
  tobject(o, bits);
  o.str("");
  o.i32(0);
  o.end(p);
}

// TH1 (version 8) up to and including fStatOverflows
void th1(Out& o, const std::string& name, const std::string& title, int ncells, int nx, double xlo, double xhi, int ny, double ylo,
         double yhi, double entries, const std::vector<double>* xedges = nullptr, const std::vector<double>* yedges = nullptr,
         const double* stats = nullptr)
{
  const size_t p = o.begin(8);
  tnamed(o, name, title, 0x03000008u);
  tattline(o);
  tattfill(o);
  tattmarker(o);
  o.i32(ncells);
  taxis(o, "xaxis", nx, xlo, xhi, 1.0f, xedges);
  taxis(o, "yaxis", ny, ylo, yhi, 0.0f, yedges);
  taxis(o, "zaxis", 1, 0., 1., 1.0f);
  o.i16(0); o.i16(1000);  // fBarOffset, fBarWidth
  o.f64(entries);
  // fTsumw, fTsumw2, fTsumwx, fTsumwx2 (a histogram filled by SetBinContent: all zero)
  for (int i = 0; i < 4; i++) o.f64(stats ? stats[i] : 0.);
  o.f64(-1111.); o.f64(-1111.);            // fMaximum, fMinimum
  o.f64(0);                                // fNormFactor
  o.i32(0);                                // fContour
  o.i32(0);                                // fSumw2
  o.str("");                               // fOption
  empty_tlist(o, 0x03010000u);             // fFunctions
  o.i32(0);                                // fBufferSize
  o.u8(0);                                 // fBuffer (no array)
  o.i32(0);                                // fBinStatErrOpt
  o.i32(2);                                // fStatOverflows
  o.end(p);
}

// ---- TTree pieces --------------------------------------------------------------------------------------------
const unsigned char kTIOFeatures[] = {0x40, 0x00, 0x00, 0x07, 0x00, 0x00, 0x1a, 0xa1, 0x2f, 0x10, 0x00};  // ROOT::TIOFeatures: count, version 0, checksum, fIOBits 0

void tobjarray_header(Out& o, int n)
{
  tobject(o, 0x03000000u);
  o.str("");
  o.i32(n);
  o.i32(0);  // fLowerBound
}

struct BasketInfo {
  uint32_t seek, nbytes;
  int64_t first_entry;
};

}  // namespace

int UpcRootFileDefaultCompression(int setting)
{
  static int current = [] {
    const char* e = std::getenv("UPCGEN_ROOT_COMPRESSION");
    const int v = e ? std::atoi(e) : 0;
    return ((v / 100 == 4 || v / 100 == 1) && v % 100 > 0 && v % 100 <= 9) ? v : 0;
  }();
  if (setting >= 0) current = setting;
  return current;
}

UpcRootFileWriter::UpcRootFileWriter() : compression_(UpcRootFileDefaultCompression()) {}

void UpcRootFileWriter::AddTH2D(const std::string& name, const std::string& title, int nx, double xlo, double xhi, int ny,
                                double ylo, double yhi, const std::vector<double>& cells, double entries)
{
  const size_t ncells = (size_t)(nx + 2) * (ny + 2);
  if (cells.size() != ncells) throw std::invalid_argument("AddTH2D: cells must hold (nx + 2) * (ny + 2) values");
  Out o;
  const size_t p = o.begin(4);  // TH2D
  {
    const size_t q = o.begin(5);  // TH2
    th1(o, name, title, (int)ncells, nx, xlo, xhi, ny, ylo, yhi, entries);
    o.f64(1.0);                       // fScalefactor
    o.f64(0); o.f64(0); o.f64(0);     // fTsumwy, fTsumwy2, fTsumwxy
    o.end(q);
  }
  o.i32((int32_t)ncells);  // TArrayD
  o.b.reserve(o.b.size() + 8 * ncells + 16);
  for (double v : cells) o.f64(v);
  o.end(p);
  Record r;
  r.cls = "TH2D"; r.name = name; r.title = title; r.data = std::move(o.b); r.listed = true;
  records_.push_back(std::move(r));
}

// TH1D (version 3) = TH1 + TArrayD; the y and z axes of a one-dimensional histogram are ROOT's defaults (1 bin on [0, 1])
void UpcRootFileWriter::AddTH1D(const std::string& name, const std::string& title, int nx, double xlo, double xhi,
                                const std::vector<double>& cells, double entries)
{
  const size_t ncells = (size_t)nx + 2;
  if (cells.size() != ncells) throw std::invalid_argument("AddTH1D: cells must hold nx + 2 values");
  Out o;
  const size_t p = o.begin(3);  // TH1D
  th1(o, name, title, (int)ncells, nx, xlo, xhi, 1, 0., 1., entries);
  o.i32((int32_t)ncells);  // TArrayD
  for (double v : cells) o.f64(v);
  o.end(p);
  Record r;
  r.cls = "TH1D"; r.name = name; r.title = title; r.data = std::move(o.b); r.listed = true;
  records_.push_back(std::move(r));
}

namespace {
void check_edges(const std::vector<double>& e, const char* what)
{
  if (e.size() < 2) throw std::invalid_argument(std::string(what) + ": an axis needs at least two edges");
  for (size_t i = 1; i < e.size(); ++i)
    if (!(e[i] > e[i - 1])) throw std::invalid_argument(std::string(what) + ": edges must increase");
}
}  // namespace

// TH2D(name, title, nx, xbins, ny, ybins): TAxis::Set(n, edges) keeps the edges in fXbins and takes fXmin / fXmax from them
void UpcRootFileWriter::AddTH2D(const std::string& name, const std::string& title, const std::vector<double>& xedges,
                                const std::vector<double>& yedges, const std::vector<double>& cells, double entries)
{
  check_edges(xedges, "AddTH2D"); check_edges(yedges, "AddTH2D");
  const int nx = (int)xedges.size() - 1, ny = (int)yedges.size() - 1;
  const size_t ncells = (size_t)(nx + 2) * (ny + 2);
  if (cells.size() != ncells) throw std::invalid_argument("AddTH2D: cells must hold (nx + 2) * (ny + 2) values");
  Out o;
  const size_t p = o.begin(4);  // TH2D
  {
    const size_t q = o.begin(5);  // TH2
    th1(o, name, title, (int)ncells, nx, xedges.front(), xedges.back(), ny, yedges.front(), yedges.back(), entries, &xedges, &yedges);
    o.f64(1.0);                       // fScalefactor
    o.f64(0); o.f64(0); o.f64(0);     // fTsumwy, fTsumwy2, fTsumwxy
    o.end(q);
  }
  o.i32((int32_t)ncells);  // TArrayD
  o.b.reserve(o.b.size() + 8 * ncells + 16);
  for (double v : cells) o.f64(v);
  o.end(p);
  Record r;
  r.cls = "TH2D"; r.name = name; r.title = title; r.data = std::move(o.b); r.listed = true;
  records_.push_back(std::move(r));
}

void UpcRootFileWriter::AddTH1D(const std::string& name, const std::string& title, const std::vector<double>& xedges,
                                const std::vector<double>& cells, double entries, const double* stats)
{
  check_edges(xedges, "AddTH1D");
  const int nx = (int)xedges.size() - 1;
  const size_t ncells = (size_t)nx + 2;
  if (cells.size() != ncells) throw std::invalid_argument("AddTH1D: cells must hold nx + 2 values");
  Out o;
  const size_t p = o.begin(3);  // TH1D
  th1(o, name, title, (int)ncells, nx, xedges.front(), xedges.back(), 1, 0., 1., entries, &xedges, nullptr, stats);
  o.i32((int32_t)ncells);  // TArrayD
  for (double v : cells) o.f64(v);
  o.end(p);
  Record r;
  r.cls = "TH1D"; r.name = name; r.title = title; r.data = std::move(o.b); r.listed = true;
  records_.push_back(std::move(r));
}

// src/UpcGenerator.cpp:900-917 (debug > 0, ROOT output): the nuclear cross section table as TH2D "hNucCSYM" over the
// (y, m) bin edges, filled by SetBinContent(iy + 1, im + 1, cs[iy][im]) (entries = ny * nm, no statistics), and its two
// projections, written in the reference's order: ProjectionY() -> "hNucCSYM_py" (over m), ProjectionX() ->
// "hNucCSYM_px" (over y), then the table.  A projection is what TH2::DoProjection leaves behind: each cell the
// sequential sum over the integrated axis (under- and overflow cells included: zero here); the statistics reset and
// recomputed from the cells as TH1::ResetStats / TH1::GetStats do (bin centres, error^2 = (sqrt|w|)^2); the entries
// floor(total + 0.5).
void UpcAddSigmaHists(UpcRootFileWriter& w, const std::vector<double>& yEdges, const std::vector<double>& mEdges,
                      const std::vector<std::vector<double>>& nucCSYM)
{
  const int ny = (int)yEdges.size() - 1, nm = (int)mEdges.size() - 1;
  if (ny < 1 || nm < 1 || (int)nucCSYM.size() != ny) throw std::invalid_argument("UpcAddSigmaHists: table and edges disagree");
  for (const auto& row : nucCSYM)
    if ((int)row.size() != nm) throw std::invalid_argument("UpcAddSigmaHists: table and edges disagree");
  std::vector<double> cells((size_t)(ny + 2) * (nm + 2), 0.);  // x = y_pair (fastest), y = m_pair
  for (int iy = 0; iy < ny; ++iy)
    for (int im = 0; im < nm; ++im) cells[(size_t)(im + 1) * (ny + 2) + (iy + 1)] = nucCSYM[iy][im];
  auto projection = [&](bool onX, const std::vector<double>& edges, const char* name) {
    const int nout = (int)edges.size() - 1, nin = onX ? nm : ny;
    std::vector<double> c((size_t)nout + 2, 0.);
    double totcont = 0;
    for (int ob = 0; ob <= nout + 1; ++ob) {
      double cont = 0;
      for (int ib = 0; ib <= nin + 1; ++ib) {
        const int bx = onX ? ob : ib, by = onX ? ib : ob;
        cont += cells[(size_t)by * (ny + 2) + bx];
      }
      c[ob] = cont;
      totcont += cont;
    }
    double st[4] = {0, 0, 0, 0};
    for (int b = 1; b <= nout; ++b) {
      const double x = 0.5 * (edges[b - 1] + edges[b]);
      const double wgt = c[b];
      const double err = std::fabs(std::sqrt(std::fabs(wgt)));
      st[0] += wgt;
      st[1] += err * err;
      st[2] += wgt * x;
      st[3] += wgt * x * x;
    }
    w.AddTH1D(name, "", edges, c, std::floor(totcont + 0.5), st);
  };
  projection(false, mEdges, "hNucCSYM_py");
  projection(true, yEdges, "hNucCSYM_px");
  w.AddTH2D("hNucCSYM", "", yEdges, mEdges, cells, (double)ny * nm);
}

void UpcRootFileWriter::AddTree(const std::string& name, const std::string& title, const std::vector<Column>& columns)
{
  if (columns.empty()) throw std::invalid_argument("AddTree: no columns");
  for (const Column& c : columns) {
    if (c.type != 'I' && c.type != 'D') throw std::invalid_argument("AddTree: column type must be I or D");
    if (c.values.size() != columns[0].values.size()) throw std::invalid_argument("AddTree: columns differ in length");
  }
  trees_.push_back(Tree{name, title, columns});
}

bool UpcRootFileWriter::Write(const std::string& path, std::string& err)
{
  try {
    const uint32_t dt = datime_now();
    std::string fname = path;
    const size_t slash = fname.find_last_of('/');
    if (slash != std::string::npos) fname = fname.substr(slash + 1);
    const uint32_t kBegin = 100;
    unsigned char uuid[18] = {0, 1};
    {
      // any 16 bytes do; derived from the time and the path so that two files differ
      uint64_t h = 1469598103934665603ull ^ dt;
      for (char ch : path) h = (h ^ (unsigned char)ch) * 1099511628211ull;
      for (int i = 0; i < 8; i++) { uuid[2 + i] = (unsigned char)(h >> (8 * i)); uuid[10 + i] = (unsigned char)((h * 0x9E3779B97F4A7C15ull) >> (8 * i)); }
    }

    auto key_len = [](const std::string& cls, const std::string& nm, const std::string& ti) {
      return (uint16_t)(26 + 1 + cls.size() + 1 + nm.size() + 1 + ti.size());
    };
    auto key_header = [&](Out& o, uint32_t nbytes, uint32_t objlen, uint16_t keylen, uint32_t seek, const std::string& cls,
                          const std::string& nm, const std::string& ti) {
      o.i32((int32_t)nbytes);
      o.i16(4);  // key version (32-bit seeks)
      o.i32((int32_t)objlen);
      o.u32(dt);
      o.u16(keylen);
      o.i16(1);  // cycle
      o.i32((int32_t)seek);
      o.i32((int32_t)kBegin);  // fSeekPdir
      o.str(cls); o.str(nm); o.str(ti);
    };

    // ---- layout: [header 100][directory key][objects and baskets ...][key list][streamer info][free segments] ----
    const uint16_t dir_keylen = key_len("TFile", fname, "");
    const uint32_t dir_namelen = (uint32_t)(1 + fname.size() + 1);  // TNamed part: name + empty title
    const uint32_t dir_objlen = dir_namelen + 2 + 4 + 4 + 4 + 4 + 4 + 4 + 4 + 18 + 12;
    uint32_t pos = kBegin + dir_keylen + dir_objlen;

    std::vector<Record> recs;  // in file order
    for (Record& r : records_) recs.push_back(r);

    // trees: baskets first (their keys are not listed), then the TTree record, which needs the baskets' positions
    struct Placed { size_t idx; };
    std::vector<Record> out_recs;
    const int zalg = compression_ / 100, zlevel = compression_ % 100;
    const bool lz4 = zalg == 4 && zlevel > 0, zl = zalg == 1 && zlevel > 0;
    if (compression_ != 0 && !lz4 && !zl)
      throw std::invalid_argument("compression setting: 0, 1xx (zlib) and 4xx (LZ4) are written");
    // ROOT's zlib records: 'Z' 'L' 8 (Z_DEFLATED), 3 bytes compressed size, 3 bytes uncompressed size (little endian),
    // a zlib stream; at most 0xffffff input bytes per record (what the reference's own cross_sections/*.root hold)
    auto zlib_zip = [&](const unsigned char* src, size_t n, std::vector<unsigned char>& dst) {
      const size_t start = dst.size();
      for (size_t p = 0; p < n; p += 0xffffff) {
        const size_t cn = std::min<size_t>(n - p, 0xffffff);
        uLongf zn = compressBound((uLong)cn);
        const size_t hdr = dst.size();
        dst.resize(hdr + 9 + zn);
        if (compress2(dst.data() + hdr + 9, &zn, src + p, (uLong)cn, zlevel) != Z_OK || 9 + zn >= cn) { dst.resize(start); return false; }
        dst.resize(hdr + 9 + zn);
        unsigned char* h = dst.data() + hdr;
        h[0] = 'Z'; h[1] = 'L'; h[2] = 8;
        h[3] = (unsigned char)zn; h[4] = (unsigned char)(zn >> 8); h[5] = (unsigned char)(zn >> 16);
        h[6] = (unsigned char)cn; h[7] = (unsigned char)(cn >> 8); h[8] = (unsigned char)(cn >> 16);
      }
      if (dst.size() - start >= n) { dst.resize(start); return false; }
      return true;
    };
    // TKey's rule: an object above 256 bytes is stored compressed if that makes it smaller.  `head` bytes at the front
    // (the basket header, which belongs to the key) stay as they are.
    auto finish = [&](Record& r, size_t head) {
      r.objlen = (uint32_t)(r.data.size() - head);
      if ((!lz4 && !zl) || r.objlen <= 256) return;
      std::vector<unsigned char> z(r.data.begin(), r.data.begin() + head);
      if (lz4 ? upc_lz4::root_zip(r.data.data() + head, r.objlen, z) : zlib_zip(r.data.data() + head, r.objlen, z)) r.data.swap(z);
    };
    auto place = [&](Record& r) {
      finish(r, 0);
      r.seek = pos;
      const uint16_t kl = key_len(r.cls, r.name, r.title);
      pos += kl + (uint32_t)r.data.size();
      out_recs.push_back(r);
    };
    for (Record& r : recs) place(r);

    const int64_t kBasketEntries = 1 << 20;
    for (const Tree& t : trees_) {
      const int64_t n = (int64_t)t.columns[0].values.size();
      const int nb = (int)t.columns.size();
      std::vector<std::vector<BasketInfo>> baskets(nb);
      int64_t tot_bytes = 0, zip_bytes = 0;
      std::vector<int64_t> br_tot(nb, 0);
      for (int ib = 0; ib < nb; ++ib) {
        const Column& c = t.columns[ib];
        const int esz = c.type == 'I' ? 4 : 8;
        for (int64_t e0 = 0; e0 < n || (n == 0 && e0 == 0 && false); e0 += kBasketEntries) {
          const int64_t ne = std::min<int64_t>(kBasketEntries, n - e0);
          // TBasket: the key header is followed by the basket's own header; both count as fKeylen
          const uint16_t kl = (uint16_t)(key_len("TBasket", c.name, t.name) + 2 + 4 + 4 + 4 + 4 + 1);
          const uint32_t datalen = (uint32_t)(ne * esz);
          Out o;
          // (the key header proper is written by the common code below; what follows it belongs to the "object")
          o.i16(3);                   // TBasket version
          o.i32(32000);               // fBufferSize
          o.i32(esz);                 // fNevBufSize: fixed length of an entry
          o.i32((int32_t)ne);         // fNevBuf
          o.i32((int32_t)(kl + datalen));  // fLast
          o.u8(0);                    // flag: no entry-offset array, buffer not kept
          o.b.reserve(o.b.size() + datalen);
          for (int64_t e = e0; e < e0 + ne; ++e) {
            if (esz == 4) o.i32((int32_t)c.values[(size_t)e]);
            else o.f64(c.values[(size_t)e]);
          }
          Record r;
          r.cls = "TBasket"; r.name = c.name; r.title = t.name; r.data = std::move(o.b); r.listed = false;
          r.seek = pos;
          finish(r, 19);  // the 19 bytes of the basket header count as key, the buffer after them may be compressed
          BasketInfo bi{pos, (uint32_t)(kl - 19 + r.data.size()), e0};
          baskets[ib].push_back(bi);
          tot_bytes += kl + datalen;
          br_tot[ib] += kl + datalen;
          zip_bytes += bi.nbytes;
          pos += bi.nbytes;  // = key header + basket header + stored buffer
          out_recs.push_back(std::move(r));
        }
      }

      // the TTree object.  Tags inside a key's buffer count from the start of the key: offsets include fKeylen.
      const uint16_t tkl = key_len("TTree", t.name, t.title);
      Out o;
      const size_t p = o.begin(20);  // TTree
      tnamed(o, t.name, t.title, 0x03000008u);
      tattline(o);
      tattfill(o);
      tattmarker(o);
      o.i64(n);            // fEntries
      o.i64(tot_bytes);    // fTotBytes
      o.i64(zip_bytes);    // fZipBytes (= fTotBytes when nothing is compressed)
      o.i64(0);            // fSavedBytes
      o.i64(0);            // fFlushedBytes
      o.f64(1.0);          // fWeight
      o.i32(0);            // fTimerInterval
      o.i32(25);           // fScanField
      o.i32(0);            // fUpdate
      o.i32(1000);         // fDefaultEntryOffsetLen
      o.i32(0);            // fNClusterRange
      o.i64(1000000000000ll);  // fMaxEntries
      o.i64(1000000000000ll);  // fMaxEntryLoop
      o.i64(0);            // fMaxVirtualSize
      o.i64(0);            // fAutoSave (the reference calls SetAutoSave(0), src/UpcGenerator.cpp:857)
      o.i64(-30000000);    // fAutoFlush
      o.i64(1000000);      // fEstimate
      o.u8(0);             // fClusterRangeEnd (empty array)
      o.u8(0);             // fClusterSize (empty array)
      o.b.insert(o.b.end(), kTIOFeatures, kTIOFeatures + sizeof(kTIOFeatures));
      std::vector<uint32_t> leaf_tags(nb, 0);
      {
        // fBranches: TObjArray of TBranch
        const size_t q = o.begin(3);
        tobjarray_header(o, nb);
        uint32_t branch_class_tag = 0;
        uint32_t leaf_class_tag[2] = {0, 0};  // TLeafI, TLeafD
        for (int ib = 0; ib < nb; ++ib) {
          const Column& c = t.columns[ib];
          const std::vector<BasketInfo>& bk = baskets[ib];
          const int nbk = (int)bk.size();
          const int max_baskets = std::max(10, nbk + 1);
          const int esz = c.type == 'I' ? 4 : 8;
          int64_t br_bytes = 0;
          for (const BasketInfo& b : bk) br_bytes += b.nbytes;
          const size_t cnt = o.size();
          o.u32(0);  // byte count of the object-any record
          if (!branch_class_tag) {
            branch_class_tag = (uint32_t)(tkl + o.size()) + kMapOffset;
            o.u32(kNewClassTag);
            o.cstr("TBranch");
          } else {
            o.u32(branch_class_tag | kClassMask);
          }
          {
            const size_t bp = o.begin(13);  // TBranch
            tnamed(o, c.name, c.name + "/" + c.type);
            tattfill(o);
            o.i32(compression_); // fCompress
            o.i32(32000);       // fBasketSize
            o.i32(0);           // fEntryOffsetLen
            o.i32(nbk);         // fWriteBasket
            o.i64(n);           // fEntryNumber
            o.b.insert(o.b.end(), kTIOFeatures, kTIOFeatures + sizeof(kTIOFeatures));
            o.i32(0);           // fOffset
            o.i32(max_baskets); // fMaxBaskets
            o.i32(0);           // fSplitLevel
            o.i64(n);           // fEntries
            o.i64(0);           // fFirstEntry
            o.i64(br_tot[ib]);  // fTotBytes
            o.i64(br_bytes);    // fZipBytes
            {
              const size_t e = o.begin(3);  // fBranches: empty TObjArray
              tobjarray_header(o, 0);
              o.end(e);
            }
            {
              const size_t e = o.begin(3);  // fLeaves: one leaf
              tobjarray_header(o, 1);
              const size_t lcnt = o.size();
              leaf_tags[ib] = (uint32_t)(tkl + lcnt) + kMapOffset;  // the object's own tag: fLeaves of the tree refers to it
              o.u32(0);
              const int li = c.type == 'I' ? 0 : 1;
              if (!leaf_class_tag[li]) {
                leaf_class_tag[li] = (uint32_t)(tkl + o.size()) + kMapOffset;
                o.u32(kNewClassTag);
                o.cstr(c.type == 'I' ? "TLeafI" : "TLeafD");
              } else {
                o.u32(leaf_class_tag[li] | kClassMask);
              }
              {
                const size_t lp = o.begin(1);  // TLeafI / TLeafD
                {
                  const size_t tl = o.begin(2);  // TLeaf
                  tnamed(o, c.name, c.name);
                  o.i32(1);      // fLen
                  o.i32(esz);    // fLenType
                  o.i32(0);      // fOffset
                  o.u8(0);       // fIsRange
                  o.u8(0);       // fIsUnsigned
                  o.u32(0);      // fLeafCount (null)
                  o.end(tl);
                }
                if (c.type == 'I') { o.i32(0); o.i32(0); }   // fMinimum, fMaximum
                else { o.f64(0); o.f64(0); }
                o.end(lp);
              }
              {
                const uint32_t v = (uint32_t)(o.size() - lcnt - 4) | kByteCountMask;
                o.b[lcnt] = (unsigned char)(v >> 24); o.b[lcnt + 1] = (unsigned char)(v >> 16); o.b[lcnt + 2] = (unsigned char)(v >> 8); o.b[lcnt + 3] = (unsigned char)v;
              }
              o.end(e);
            }
            {
              const size_t e = o.begin(3);  // fBaskets: none kept in memory
              tobjarray_header(o, 0);
              o.end(e);
            }
            o.u8(1);  // fBasketBytes [fMaxBaskets]
            for (int i = 0; i < max_baskets; ++i) o.i32(i < nbk ? (int32_t)bk[i].nbytes : 0);
            o.u8(1);  // fBasketEntry [fMaxBaskets]: first entry of every basket, then the number of entries
            for (int i = 0; i < max_baskets; ++i) o.i64(i < nbk ? bk[i].first_entry : (i == nbk ? n : 0));
            o.u8(1);  // fBasketSeek [fMaxBaskets]
            for (int i = 0; i < max_baskets; ++i) o.i64(i < nbk ? (int64_t)bk[i].seek : 0);
            o.str("");  // fFileName
            o.end(bp);
          }
          {
            const uint32_t v = (uint32_t)(o.size() - cnt - 4) | kByteCountMask;
            o.b[cnt] = (unsigned char)(v >> 24); o.b[cnt + 1] = (unsigned char)(v >> 16); o.b[cnt + 2] = (unsigned char)(v >> 8); o.b[cnt + 3] = (unsigned char)v;
          }
        }
        o.end(q);
      }
      {
        // fLeaves: TObjArray of references to the leaves streamed inside the branches
        const size_t q = o.begin(3);
        tobjarray_header(o, nb);
        for (int ib = 0; ib < nb; ++ib) o.u32(leaf_tags[ib]);
        o.end(q);
      }
      o.u32(0);  // fAliases (null)
      o.i32(0);  // fIndexValues (TArrayD, empty)
      o.i32(0);  // fIndex (TArrayI, empty)
      o.u32(0);  // fTreeIndex
      o.u32(0);  // fFriends
      o.u32(0);  // fUserInfo
      o.u32(0);  // fBranchRef
      o.end(p);
      Record r;
      r.cls = "TTree"; r.name = t.name; r.title = t.title; r.data = std::move(o.b); r.listed = true;
      place(r);
    }

    // ---- key list, streamer info, free segments ----
    const uint32_t seek_keys = pos;
    Out keys;
    {
      int nlisted = 0;
      for (const Record& r : out_recs) nlisted += r.listed;
      keys.i32(nlisted);
      for (const Record& r : out_recs) {
        if (!r.listed) continue;
        const uint16_t kl = key_len(r.cls, r.name, r.title);
        key_header(keys, kl + (uint32_t)r.data.size(), r.objlen, kl, r.seek, r.cls, r.name, r.title);
      }
    }
    const uint32_t nbytes_keys = dir_keylen + (uint32_t)keys.size();
    pos += nbytes_keys;

    const uint32_t seek_info = pos;
    Out info;
    empty_tlist(info, 0x02000000u);  // no TStreamerInfo records: the classes are written in their current versions
    const uint16_t info_keylen = key_len("TList", "StreamerInfo", "Doubly linked list");
    const uint32_t nbytes_info = info_keylen + (uint32_t)info.size();
    pos += nbytes_info;

    const uint32_t seek_free = pos;
    const uint32_t nbytes_free = dir_keylen + 10;
    const uint32_t end = seek_free + nbytes_free;

    // ---- emit ----
    Out f;
    f.b.reserve(end + 16);
    f.b.insert(f.b.end(), {'r', 'o', 'o', 't'});
    f.i32(62206);  // fVersion: 6.22/06 layout, 32-bit seeks
    f.i32((int32_t)kBegin);
    f.i32((int32_t)end);
    f.i32((int32_t)seek_free);
    f.i32((int32_t)nbytes_free);
    f.i32(1);  // nfree
    f.i32((int32_t)(dir_keylen + dir_namelen));  // fNbytesName
    f.u8(4);   // fUnits
    f.i32(compression_);  // fCompress
    f.i32((int32_t)seek_info);
    f.i32((int32_t)nbytes_info);
    f.b.insert(f.b.end(), uuid, uuid + 18);
    while (f.size() < kBegin) f.u8(0);

    // the directory
    key_header(f, dir_keylen + dir_objlen, dir_objlen, dir_keylen, kBegin, "TFile", fname, "");
    f.b[f.size() - 1 - fname.size() - 1 - 5 - 1 - 4] = f.b[f.size() - 1 - fname.size() - 1 - 5 - 1 - 4];  // (fSeekPdir of the top directory is rewritten below)
    f.str(fname);
    f.str("");
    f.i16(5);  // TDirectory version
    f.u32(dt); f.u32(dt);
    f.i32((int32_t)nbytes_keys);
    f.i32((int32_t)(dir_keylen + dir_namelen));
    f.i32((int32_t)kBegin);  // fSeekDir
    f.i32(0);                // fSeekParent
    f.i32((int32_t)seek_keys);
    f.b.insert(f.b.end(), uuid, uuid + 18);
    for (int i = 0; i < 12; i++) f.u8(0);
    // the top directory has no parent: fSeekPdir = 0 in its own key
    {
      const size_t off = kBegin + 4 + 2 + 4 + 4 + 2 + 2 + 4;
      f.b[off] = f.b[off + 1] = f.b[off + 2] = f.b[off + 3] = 0;
    }

    for (const Record& r : out_recs) {
      if (f.size() != r.seek) throw std::runtime_error("internal: record offset mismatch");
      uint16_t kl = key_len(r.cls, r.name, r.title);
      const uint32_t nbytes = kl + (uint32_t)r.data.size();
      // the basket header (19 bytes) counts as part of the key: fKeylen includes it, fObjlen is the payload
      if (r.cls == "TBasket") kl = (uint16_t)(kl + 19);
      key_header(f, nbytes, r.objlen, kl, r.seek, r.cls, r.name, r.title);
      f.raw(r.data);
    }

    if (f.size() != seek_keys) throw std::runtime_error("internal: key list offset mismatch");
    key_header(f, nbytes_keys, (uint32_t)keys.size(), dir_keylen, seek_keys, "TFile", fname, "");
    f.raw(keys.b);
    key_header(f, nbytes_info, (uint32_t)info.size(), info_keylen, seek_info, "TList", "StreamerInfo", "Doubly linked list");
    f.raw(info.b);
    key_header(f, nbytes_free, 10, dir_keylen, seek_free, "TFile", fname, "");
    f.i16(1);  // TFree version
    f.i32((int32_t)end);
    f.i32(2000000000);
    if (f.size() != end) throw std::runtime_error("internal: file length mismatch");

    FILE* fp = std::fopen(path.c_str(), "wb");
    if (!fp) { err = "cannot create " + path; return false; }
    const size_t w = std::fwrite(f.b.data(), 1, f.b.size(), fp);
    std::fclose(fp);
    if (w != f.b.size()) { err = "short write to " + path; return false; }
    return true;
  } catch (const std::exception& e) {
    err = e.what();
    return false;
  }
}
