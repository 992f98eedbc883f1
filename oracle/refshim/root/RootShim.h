// ROOT shim for the reference-shim build (oracle/refshim): the handful of ROOT entry points that
// the reference's src/UpcCrossSection.cpp touches, with ROOT's documented semantics.  TEST
// INFRASTRUCTURE: lets the reference's OWN translation units run here without ROOT.
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../shim_tape.h"

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
  virtual double Rndm()
  {
    const double u = (rng() >> 11) * (1.0 / 9007199254740992.0);
    shim_tape().put(u, 0);
    return u;
  }
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

class TH1 : public TObject
{
 public:
  static void AddDirectory(bool) {}
};

// TAxis::FindBin on a uniform axis: 0 below, n + 1 at or above the upper edge, else 1 + int(n (x - lo) / (hi - lo))
class TAxis
{
 public:
  TAxis() = default;
  TAxis(int n_, double lo_, double hi_) : n(n_), lo(lo_), hi(hi_) {}
  int FindBin(double x) const
  {
    if (x < lo) return 0;
    if (!(x < hi)) return n + 1;
    return 1 + int(n * (x - lo) / (hi - lo));
  }
  int GetNbins() const { return n; }
  int n{1};
  double lo{0}, hi{1};
};

class TH1D : public TH1
{
 public:
  TH1D() = default;
  TH1D(const char* name, const char*, int nb, double lo, double hi) : name_(name), n(nb), xlo(lo), xhi(hi), c(nb + 2, 0.), xaxis_(nb, lo, hi) {}
  const TAxis* GetXaxis() const { return &xaxis_; }
  void SetDirectory(void*) {}
  int Write() { return 0; }
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
  TAxis xaxis_;
};

class TFile;
class TH2D : public TObject
{
 public:
  TH2D(const char* name, const char*, int nx_, double xlo, double xhi, int ny_, double ylo, double yhi)
      : name_(name), nx(nx_), ny(ny_), c((size_t)(nx_ + 2) * (ny_ + 2), 0.), xaxis_(nx_, xlo, xhi), yaxis_(ny_, ylo, yhi) {}
  const TAxis* GetXaxis() const { return &xaxis_; }
  const TAxis* GetYaxis() const { return &yaxis_; }
  int GetNbinsX() const { return nx; }
  int GetNbinsY() const { return ny; }
  void SetDirectory(void*) {}
  TH2D(const char* name, const char*, int nx_, const double*, int ny_, const double*) : name_(name), nx(nx_), ny(ny_), c((size_t)(nx_ + 2) * (ny_ + 2), 0.) {}
  // debug-only projections of UpcGenerator::generateEvents (src/UpcGenerator.cpp:902-908)
  TH1D* ProjectionX() const
  {
    auto* h = new TH1D((name_ + "_px").c_str(), "", nx, 0., 1.);
    for (int ix = 0; ix <= nx + 1; ix++) { double a = 0; for (int iy = 0; iy <= ny + 1; iy++) a += GetBinContent(ix, iy); h->SetBinContent(ix, a); }
    return h;
  }
  TH1D* ProjectionY() const
  {
    auto* h = new TH1D((name_ + "_py").c_str(), "", ny, 0., 1.);
    for (int iy = 0; iy <= ny + 1; iy++) { double a = 0; for (int ix = 0; ix <= nx + 1; ix++) a += GetBinContent(ix, iy); h->SetBinContent(iy, a); }
    return h;
  }
  void SetBinContent(int ix, int iy, double v) { c[(size_t)iy * (nx + 2) + ix] = v; }
  double GetBinContent(int ix, int iy) const { return c[(size_t)iy * (nx + 2) + ix]; }
  TObject* Clone(const char* newname) const { auto* h = new TH2D(*this); h->name_ = newname; return h; }
  int Write();
  std::string name_;
  int nx, ny;
  std::vector<double> c;
  TAxis xaxis_, yaxis_;
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
  // one-dimensional histograms (the sigma(m) tables of the light-by-light and pi0 pi0 plug-ins: injected by the tests)
  static std::map<std::string, std::map<std::string, TH1D*>>& store1()
  {
    static std::map<std::string, std::map<std::string, TH1D*>> s;
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
  if (it != m.end()) return new TH2D(*it->second);
  auto& m1 = store1()[fname_];
  auto it1 = m1.find(name);
  return it1 == m1.end() ? nullptr : new TH1D(*it1->second);
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

// TVector3 / TLorentzVector with ROOT's arithmetic (math/physics/src/TVector3.cxx, TLorentzVector.cxx), the members the
// reference's event stage calls (src/UpcGenerator.cpp:388-423, :526-587, :770-775; src/UpcCrossSection.cpp:1053-1074)
class TVector3
{
 public:
  TVector3() = default;
  TVector3(double x, double y, double z) : fX(x), fY(y), fZ(z) {}
  double X() const { return fX; }
  double Y() const { return fY; }
  double Z() const { return fZ; }
  double Mag2() const { return fX * fX + fY * fY + fZ * fZ; }
  double Mag() const { return std::sqrt(Mag2()); }
  double Perp() const { return std::sqrt(fX * fX + fY * fY); }
  double CosTheta() const { const double ptot = Mag(); return ptot == 0.0 ? 1.0 : fZ / ptot; }
  double PseudoRapidity() const
  {
    const double cosTheta = CosTheta();
    if (cosTheta * cosTheta < 1) return -0.5 * std::log((1.0 - cosTheta) / (1.0 + cosTheta));
    if (fZ == 0) return 0;
    if (fZ > 0) return 10e10;
    return -10e10;
  }
  void SetMagThetaPhi(double mag, double theta, double phi)
  {
    const double amag = std::fabs(mag);
    fX = amag * std::sin(theta) * std::cos(phi);
    fY = amag * std::sin(theta) * std::sin(phi);
    fZ = amag * std::cos(theta);
  }
  TVector3 Unit() const
  {
    const double tot2 = Mag2();
    const double tot = (tot2 > 0) ? 1.0 / std::sqrt(tot2) : 1.0;
    return TVector3(fX * tot, fY * tot, fZ * tot);
  }
  TVector3 operator-() const { return TVector3(-fX, -fY, -fZ); }
  void RotateUz(const TVector3& nu)
  {
    const double u1 = nu.fX, u2 = nu.fY, u3 = nu.fZ;
    double up = u1 * u1 + u2 * u2;
    if (up) {
      up = std::sqrt(up);
      const double px = fX, py = fY, pz = fZ;
      fX = (u1 * u3 * px - u2 * py + u1 * up * pz) / up;
      fY = (u2 * u3 * px + u1 * py + u2 * up * pz) / up;
      fZ = (u3 * u3 * px - px + u3 * up * pz) / up;
    } else if (u3 < 0.) {
      fX = -fX;
      fZ = -fZ;
    }
  }
  double fX{0}, fY{0}, fZ{0};
};

class TLorentzVector
{
 public:
  void SetPxPyPzE(double x, double y, double z, double e) { fP = TVector3(x, y, z); fE = e; }
  void SetXYZT(double x, double y, double z, double t) { SetPxPyPzE(x, y, z, t); }
  void SetXYZM(double x, double y, double z, double m)
  {
    if (m >= 0) SetXYZT(x, y, z, std::sqrt(x * x + y * y + z * z + m * m));
    else SetXYZT(x, y, z, std::sqrt(std::max((x * x + y * y + z * z - m * m), 0.)));
  }
  void SetVectM(const TVector3& v, double m) { SetXYZM(v.X(), v.Y(), v.Z(), m); }
  double Px() const { return fP.fX; }
  double Py() const { return fP.fY; }
  double Pz() const { return fP.fZ; }
  double E() const { return fE; }
  double X() const { return fP.fX; }
  double Y() const { return fP.fY; }
  double Z() const { return fP.fZ; }
  double T() const { return fE; }
  TVector3 Vect() const { return fP; }
  double Mag2() const { return fE * fE - fP.Mag2(); }
  double Mag() const { const double mm = Mag2(); return mm < 0.0 ? -std::sqrt(-mm) : std::sqrt(mm); }
  double M() const { return Mag(); }
  double Pt() const { return fP.Perp(); }
  double Eta() const { return fP.PseudoRapidity(); }
  TVector3 BoostVector() const { return TVector3(X() / T(), Y() / T(), Z() / T()); }
  void Boost(const TVector3& b) { Boost(b.X(), b.Y(), b.Z()); }
  void Boost(double bx, double by, double bz)
  {
    const double b2 = bx * bx + by * by + bz * bz;
    const double gamma = 1.0 / std::sqrt(1.0 - b2);
    const double bp = bx * X() + by * Y() + bz * Z();
    const double gamma2 = b2 > 0 ? (gamma - 1.0) / b2 : 0.0;
    const double nx = X() + gamma2 * bp * bx + gamma * bx * T();
    const double ny = Y() + gamma2 * bp * by + gamma * by * T();
    const double nz = Z() + gamma2 * bp * bz + gamma * bz * T();
    const double nt = gamma * (T() + bp);
    fP = TVector3(nx, ny, nz);
    fE = nt;
  }
  void RotateUz(const TVector3& nu) { fP.RotateUz(nu); }
  TVector3 fP;
  double fE{0};
};
// ---- what src/UpcPhotoNuclearVM.cpp uses of TF1 / TGraph / TSpline3 (restated from ROOT's documented behaviour) ----
extern "C" int upco_qags(double (*f)(double, void*), void* par, double a, double b, double epsabs, double epsrel,
                         size_t limit, double* result, double* abserr);
// TF1 over a C function f(x[], par[]): Eval, and Integral(a, b, epsrel = 1e-12) = TF1::IntegralOneDim, which hands the
// function to ROOT::Math::IntegratorOneDim with epsabs = epsrel: the adaptive-singular integrator (GSL QAGS) of a ROOT
// with MathMore.  (A ROOT without MathMore falls back to its Gauss-Legendre integrator; on these smooth integrands the
// two agree to the 1e-12 they are asked for.)
class TF1
{
 public:
  typedef double (*Fn)(double*, double*);
  TF1(const char*, Fn f, double, double, int npar) : fFn(f), fPar(npar > 0 ? npar : 1, 0.) {}
  void FixParameter(int i, double v) { fPar[i] = v; }
  void SetParameter(int i, double v) { fPar[i] = v; }
  double Eval(double x)
  {
    double xx[1] = {x};
    return fFn(xx, fPar.data());
  }
  double Integral(double a, double b, double epsrel = 1e-12)
  {
    double res = 0, err = 0;
    upco_qags(&TF1::Thunk, this, a, b, epsrel, epsrel, 1000, &res, &err);
    return res;
  }

 private:
  static double Thunk(double x, void* self) { return ((TF1*)self)->Eval(x); }
  Fn fFn;
  std::vector<double> fPar;
};

class TSpline3;
// TGraph: points, and Eval(x, spline, option): through the spline when one is given, else the straight line through
// the two points around x (beyond the ends: through the two end points), TGraph::Eval
class TGraph : public TObject
{
 public:
  enum { kIsSortedX = 1 << 19 };
  void SetPoint(int i, double x, double y)
  {
    if ((int)fX.size() <= i) { fX.resize(i + 1, 0.); fY.resize(i + 1, 0.); }
    fX[i] = x; fY[i] = y;
  }
  void SetBit(unsigned) {}
  int GetN() const { return (int)fX.size(); }
  const double* GetX() const { return fX.data(); }
  const double* GetY() const { return fY.data(); }
  inline double Eval(double x, TSpline3* spline = nullptr, const char* = "") const;
  std::vector<double> fX, fY;
};

// TSpline3(title, graph) with its default end conditions: ROOT's BuildCoeff is de Boor's CUBSPL with ibcbeg = ibcend =
// 0, the not-a-knot spline (the first two and the last two cubic pieces coincide); Eval picks the piece that holds x
// (the first or last piece beyond the ends) and returns y + d (b + d (c + d D)).
class TSpline3 : public TObject
{
 public:
  TSpline3(const char*, const TGraph* g, const char* = nullptr, double = 0, double = 0)
  {
    const int n = g->GetN();
    fX.assign(g->GetX(), g->GetX() + n);
    fY.assign(g->GetY(), g->GetY() + n);
    fB.assign(n, 0.); fC.assign(n, 0.); fD.assign(n, 0.);
    Build();
  }
  void SetBit(unsigned) {}
  double Eval(double x) const
  {
    const int n = (int)fX.size();
    int k = int(std::upper_bound(fX.begin(), fX.end(), x) - fX.begin()) - 1;
    if (k < 0) k = 0;
    if (k > n - 2) k = n - 2;
    const double d = x - fX[k];
    return fY[k] + d * (fB[k] + d * (fC[k] + d * fD[k]));
  }

 private:
  // second derivatives M_i from the interior continuity equations plus the two not-a-knot rows, solved by Gaussian
  // elimination with partial pivoting on the (n x n) system (n = 37 here); then the piecewise coefficients
  void Build()
  {
    const int n = (int)fX.size();
    std::vector<double> h(n - 1);
    for (int i = 0; i + 1 < n; ++i) h[i] = fX[i + 1] - fX[i];
    std::vector<std::vector<long double>> a(n, std::vector<long double>(n + 1, 0.L));
    for (int i = 1; i + 1 < n; ++i) {
      a[i][i - 1] = h[i - 1];
      a[i][i] = 2.L * ((long double)h[i - 1] + h[i]);
      a[i][i + 1] = h[i];
      a[i][n] = 6.L * (((long double)fY[i + 1] - fY[i]) / h[i] - ((long double)fY[i] - fY[i - 1]) / h[i - 1]);
    }
    // not-a-knot: the third derivative (M_{i+1} - M_i) / h_i is continuous across x_1 and across x_{n-2}
    a[0][0] = h[1]; a[0][1] = -((long double)h[0] + h[1]); a[0][2] = h[0];
    a[n - 1][n - 3] = h[n - 2]; a[n - 1][n - 2] = -((long double)h[n - 3] + h[n - 2]); a[n - 1][n - 1] = h[n - 3];
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int r = c + 1; r < n; ++r)
        if (fabsl(a[r][c]) > fabsl(a[piv][c])) piv = r;
      std::swap(a[c], a[piv]);
      for (int r = c + 1; r < n; ++r) {
        const long double f = a[r][c] / a[c][c];
        if (f == 0.L) continue;
        for (int k = c; k <= n; ++k) a[r][k] -= f * a[c][k];
      }
    }
    std::vector<long double> M(n);
    for (int r = n - 1; r >= 0; --r) {
      long double sacc = a[r][n];
      for (int k = r + 1; k < n; ++k) sacc -= a[r][k] * M[k];
      M[r] = sacc / a[r][r];
    }
    for (int i = 0; i + 1 < n; ++i) {
      fC[i] = double(M[i] / 2.L);
      fD[i] = double((M[i + 1] - M[i]) / (6.L * h[i]));
      fB[i] = double(((long double)fY[i + 1] - fY[i]) / h[i] - h[i] * (2.L * M[i] + M[i + 1]) / 6.L);
    }
  }
  std::vector<double> fX, fY, fB, fC, fD;
};

inline double TGraph::Eval(double x, TSpline3* spline, const char*) const
{
  if (spline) return spline->Eval(x);
  const int n = GetN();
  if (n == 0) return 0;
  if (n == 1) return fY[0];
  int low = int(std::upper_bound(fX.begin(), fX.end(), x) - fX.begin()) - 1, up;
  if (low < 0) { low = 0; up = 1; }
  else if (low >= n - 1) { up = n - 1; low = n - 2; }
  else up = low + 1;
  if (fX[low] == fX[up]) return fY[low];
  return fY[up] + (x - fX[up]) * (fY[low] - fY[up]) / (fX[low] - fX[up]);
}
class TGraph2D {};
// the output tree of generateEvents (src/UpcGenerator.cpp:842-857): branch addresses are recorded and every Fill()
// appends the current values, so that tests can read back what the reference would have written to events.root
class TTree
{
 public:
  TTree(const char* = "", const char* = "") {}
  struct Br { std::string name; void* addr; char type; };
  void Branch(const char* name, void* addr, const char* leaflist)
  {
    std::string l(leaflist);
    br.push_back(Br{name, addr, l.empty() ? 'D' : l.back()});
    cols.emplace_back();
  }
  int Fill()
  {
    for (size_t i = 0; i < br.size(); i++) {
      // "/I" leaves are 32-bit ints (eventNumber/I is bound to a long: ROOT reads its low 4 bytes -- Q10)
      cols[i].push_back(br[i].type == 'I' ? (double)*(int*)br[i].addr : *(double*)br[i].addr);
    }
    return 1;
  }
  void SetAutoSave(long long) {}
  int Write() { return 0; }
  std::vector<Br> br;
  std::vector<std::vector<double>> cols;
};
class TClonesArray
{
 public:
  TClonesArray(const char* = "") {}
  int GetEntriesFast() const { return 0; }
  TObject* At(int) const { return nullptr; }
};
class TParticle : public TObject
{
 public:
  TParticle() = default;
  TParticle(int pdg, int status, int mother1, int mother2, int daughter1, int daughter2, double px, double py, double pz,
            double etot, double vx, double vy, double vz, double time)
      : fPdgCode(pdg), fStatusCode(status), fPx(px), fPy(py), fPz(pz), fE(etot)
  {
    fMother[0] = mother1; fMother[1] = mother2; fDaughter[0] = daughter1; fDaughter[1] = daughter2;
    (void)vx; (void)vy; (void)vz; (void)time;
  }
  int GetPdgCode() const { return fPdgCode; }
  int GetStatusCode() const { return fStatusCode; }
  int GetFirstMother() const { return fMother[0]; }
  int GetFirstDaughter() const { return fDaughter[0]; }
  int GetLastDaughter() const { return fDaughter[1]; }
  double Px() const { return fPx; }
  double Py() const { return fPy; }
  double Pz() const { return fPz; }
  double Energy() const { return fE; }
  void Momentum(TLorentzVector& v) const { v.SetPxPyPzE(fPx, fPy, fPz, fE); }
  void Print() const {}
  int fPdgCode{0}, fStatusCode{0}, fMother[2]{0, 0}, fDaughter[2]{0, 0};
  double fPx{0}, fPy{0}, fPz{0}, fE{0};
};
