// UpcRootHist.h -- a ROOT-less reader of the TH1D / TH2D objects the reference keeps in .root files: the
// elementary cross sections of light-by-light scattering and pi0 pi0 production
// (cross_sections/{lbyl,pi0pi0}/cross_section_{m,zm}.root, src/UpcTwoPhotonLbyL.cpp:36-50) and the
// two-photon-luminosity caches (twoPhotonLumi[Pol].root, hD2LDMDY[_s,_p], src/UpcCrossSection.cpp:481-507).
//
// What is read: the TFile header, the chain of TKey records, the zlib-compressed object buffer ("ZL" blocks) and,
// inside it, the streamed TH1 / TH2: the three TAxis (fNbins, fXmin, fXmax, fXbins) and the TArrayD of cell
// contents; everything else is skipped by the byte counts ROOT writes in front of every object.  The class offers
// the handful of TH1 / TAxis calls the reference's plug-ins make (FindBin, GetBinContent, SetBinContent,
// GetNbinsX/Y), with ROOT's semantics: bin 0 is the underflow cell, bin n + 1 the overflow cell, a value on a bin's
// lower edge belongs to that bin.
#pragma once
#include <string>
#include <vector>

class UpcRootAxis
{
 public:
  int fNbins{0};
  double fXmin{0}, fXmax{0};
  std::vector<double> fXbins;  // variable bin edges (fNbins + 1 values) or empty
  int GetNbins() const { return fNbins; }
  // TAxis::FindBin
  int FindBin(double x) const;
};

class UpcRootHist
{
 public:
  int dim{0};  // 1: TH1D, 2: TH2D
  std::string name;
  UpcRootAxis fXaxis, fYaxis, fZaxis;
  std::vector<double> fArray;  // (nx + 2) [* (ny + 2)] cells, x fastest

  // reads object `name` (highest cycle) from the ROOT file `path`; on failure returns false and sets `err`
  bool Read(const std::string& path, const std::string& name, std::string& err);

  const UpcRootAxis* GetXaxis() const { return &fXaxis; }
  const UpcRootAxis* GetYaxis() const { return &fYaxis; }
  int GetNbinsX() const { return fXaxis.fNbins; }
  int GetNbinsY() const { return fYaxis.fNbins; }
  // TH1::GetBinContent(bin) / TH2::GetBinContent(binx, biny), with ROOT's clamping of out-of-range bins
  double GetBinContent(int bin) const;
  double GetBinContent(int binx, int biny) const;
  void SetBinContent(int bin, double v);
  void SetBinContent(int binx, int biny, double v);
};
