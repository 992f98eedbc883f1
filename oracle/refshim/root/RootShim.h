// ROOT shim for the reference-shim build (oracle/refshim): the handful of ROOT entry points that
// the reference's src/UpcCrossSection.cpp touches, with ROOT's documented semantics.  TEST
// INFRASTRUCTURE: lets the reference's OWN translation units run here without ROOT.
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <random>
#include <string>
#include <vector>

extern "C" double upco_tmath_besselK1(double x);

namespace TMath
{
inline double CosH(double x) { return std::cosh(x); }
inline double SinH(double x) { return std::sinh(x); }
inline double ACosH(double x) { return std::acosh(x); }  // ROOT: log(x + sqrt(x*x-1)) -- see note in refshim/README
inline double Sin(double x) { return std::sin(x); }
inline double Cos(double x) { return std::cos(x); }
inline double Sqrt(double x) { return std::sqrt(x); }
inline double Exp(double x) { return std::exp(x); }
inline double Log(double x) { return std::log(x); }
inline double Abs(double x) { return std::fabs(x); }
inline double Power(double x, double y) { return std::pow(x, y); }
template <class T> inline T Max(T a, T b) { return a >= b ? a : b; }
template <class T> inline T Min(T a, T b) { return a <= b ? a : b; }
inline double Pi() { return 3.14159265358979323846; }
inline double BesselK1(double x) { return upco_tmath_besselK1(x); }
}  // namespace TMath

class TString
{
 public:
  TString() = default;
  TString(const char* s) : s_(s) {}
  TString(const std::string& s) : s_(s) {}
  const char* Data() const { return s_.c_str(); }
  TString& operator+=(const char* o) { s_ += o; return *this; }
  TString& operator+=(const TString& o) { s_ += o.s_; return *this; }
  operator const char*() const { return s_.c_str(); }
  std::string s_;
};
inline TString operator+(const TString& a, const char* b) { return TString(a.s_ + b); }
inline TString operator+(const TString& a, const TString& b) { return TString(a.s_ + b.s_); }
inline std::ostream& operator<<(std::ostream& os, const TString& s) { return os << s.s_; }
inline const char* Form(const char* fmt, ...)
{
  static thread_local char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return buf;
}

class TRandom
{
 public:
  virtual ~TRandom() {}
  virtual void SetSeed(unsigned long s = 0) { rng.seed(s); }
  virtual double Rndm() { return (rng() >> 11) * (1.0 / 9007199254740992.0); }
  double Uniform(double a, double b) { return a + (b - a) * Rndm(); }
  double Uniform(double b = 1) { return b * Rndm(); }
  std::mt19937_64 rng{4357};
};
class TRandomMT64 : public TRandom {};
extern TRandom* gRandom;

class TObject
{
 public:
  virtual ~TObject() {}
};

class TH1D : public TObject
{
 public:
  TH1D() = default;
  TH1D(const char* name, const char*, int nb, double lo, double hi) : name_(name), n(nb), xlo(lo), xhi(hi), c(nb + 2, 0.) {}
  void SetDirectory(void*) {}
  void SetBinContent(int bin, double v) { c[bin] = v; integral.clear(); }
  double GetBinContent(int bin) const { return c[bin]; }
  int GetNbinsX() const { return n; }
  // TH1::ComputeIntegral + TH1::GetRandom
  double GetRandom()
  {
    if (integral.empty()) {
      integral.assign(n + 1, 0.);
      for (int b = 0; b < n; b++) integral[b + 1] = integral[b] + c[b + 1];
      if (integral[n] == 0) return 0;
      for (int b = 1; b <= n; b++) integral[b] /= integral[n];
    }
    if (integral[n] == 0) return 0;
    double r1 = gRandom->Rndm();
    int lo = 0, hi = n;
    while (hi - lo > 1) { int mid = (lo + hi) / 2; if (integral[mid] <= r1) lo = mid; else hi = mid; }
    double bw = (xhi - xlo) / n;
    double x = xlo + lo * bw;
    if (r1 > integral[lo]) x += bw * (r1 - integral[lo]) / (integral[lo + 1] - integral[lo]);
    return x;
  }
  std::string name_;
  int n{0};
  double xlo{0}, xhi{1};
  std::vector<double> c, integral;
};

class TFile;
class TH2D : public TObject
{
 public:
  TH2D(const char* name, const char*, int nx_, double, double, int ny_, double, double) : name_(name), nx(nx_), ny(ny_), c((size_t)(nx_ + 2) * (ny_ + 2), 0.) {}
  void SetBinContent(int ix, int iy, double v) { c[(size_t)iy * (nx + 2) + ix] = v; }
  double GetBinContent(int ix, int iy) const { return c[(size_t)iy * (nx + 2) + ix]; }
  TObject* Clone(const char* newname) const { auto* h = new TH2D(*this); h->name_ = newname; return h; }
  int Write();
  std::string name_;
  int nx, ny;
  std::vector<double> c;
};

// files live in memory (plus an empty marker on disk so that gSystem->AccessPathName sees them)
class TFile
{
 public:
  TFile(const char* fname, const char* mode = "", const char* = "", int = 0);
  TObject* Get(const char* name);
  void Close() {}
  void Write() {}
  std::string fname_;
  static std::map<std::string, std::map<std::string, TH2D*>>& store()
  {
    static std::map<std::string, std::map<std::string, TH2D*>> s;
    return s;
  }
  static TFile*& current() { static TFile* c = nullptr; return c; }
};
inline TFile::TFile(const char* fname, const char* mode, const char*, int) : fname_(fname)
{
  std::string m(mode);
  if (m == "recreate" || m == "RECREATE") {
    store()[fname_].clear();
    FILE* f = std::fopen(fname, "w");
    if (f) std::fclose(f);
  }
  current() = this;
}
inline TObject* TFile::Get(const char* name)
{
  auto& m = store()[fname_];
  auto it = m.find(name);
  return it == m.end() ? nullptr : new TH2D(*it->second);
}
inline int TH2D::Write()
{
  if (TFile::current()) TFile::store()[TFile::current()->fname_][name_] = new TH2D(*this);
  return 0;
}

class TSystem
{
 public:
  // ROOT: returns kFALSE if the path IS accessible
  bool AccessPathName(const char* path) { return access(path, F_OK) != 0; }
};
extern TSystem* gSystem;

namespace ROOT
{
inline void EnableThreadSafety() {}
}

class TLorentzVector
{
 public:
  void SetPxPyPzE(double x, double y, double z, double e) { fX = x; fY = y; fZ = z; fE = e; }
  double Px() const { return fX; }
  double Py() const { return fY; }
  double Pz() const { return fZ; }
  double E() const { return fE; }
  double fX{0}, fY{0}, fZ{0}, fE{0};
};
class TF1 {};
class TGraph {};
class TGraph2D {};
class TSpline3 {};
class TTree {};
class TClonesArray {};
class TParticle {};
