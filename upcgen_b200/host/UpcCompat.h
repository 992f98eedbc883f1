// UpcCompat.h -- the few ROOT / plog types the reference's public signatures mention
// (TLorentzVector in UpcGenerator::generateEvent, TString lumiFileDirectory, TParticle in
// getParticles, PLOG_* logging).  ROOT and plog are not in this image, so minimal stand-ins with
// the same member names are provided; building with -DUPC_WITH_ROOT uses the real headers and
// the facade compiles unchanged against them.
#pragma once

#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

#ifdef UPC_WITH_ROOT
#include "TLorentzVector.h"
#include "TParticle.h"
#include "TString.h"
#else

class TString : public std::string
{
 public:
  using std::string::string;
  TString() = default;
  TString(const std::string& s) : std::string(s) {}
  const char* Data() const { return c_str(); }
};

class TVector3
{
 public:
  TVector3(double x = 0, double y = 0, double z = 0) : fX(x), fY(y), fZ(z) {}
  double X() const { return fX; }
  double Y() const { return fY; }
  double Z() const { return fZ; }
  double Mag() const { return std::sqrt(fX * fX + fY * fY + fZ * fZ); }
  double fX, fY, fZ;
};

// arithmetic as ROOT's TLorentzVector (the subset the generator uses)
class TLorentzVector
{
 public:
  TLorentzVector(double x = 0, double y = 0, double z = 0, double t = 0) : fX(x), fY(y), fZ(z), fE(t) {}
  void SetPxPyPzE(double x, double y, double z, double e) { fX = x; fY = y; fZ = z; fE = e; }
  double Px() const { return fX; }
  double Py() const { return fY; }
  double Pz() const { return fZ; }
  double E() const { return fE; }
  double Pt() const { return std::sqrt(fX * fX + fY * fY); }
  double P() const { return std::sqrt(fX * fX + fY * fY + fZ * fZ); }
  double Mag2() const { return fE * fE - (fX * fX + fY * fY + fZ * fZ); }
  double Mag() const { double mm = Mag2(); return mm < 0.0 ? -std::sqrt(-mm) : std::sqrt(mm); }
  double M() const { return Mag(); }
  double Eta() const
  {
    double ptot = P();
    double cosTheta = ptot == 0.0 ? 1.0 : fZ / ptot;
    if (cosTheta * cosTheta < 1) return -0.5 * std::log((1.0 - cosTheta) / (1.0 + cosTheta));
    if (fZ == 0) return 0;
    return fZ > 0 ? 10e10 : -10e10;
  }
  double Rapidity() const { return 0.5 * std::log((fE + fZ) / (fE - fZ)); }
  TVector3 Vect() const { return TVector3(fX, fY, fZ); }
  TVector3 BoostVector() const { return TVector3(fX / fE, fY / fE, fZ / fE); }
  double fX, fY, fZ, fE;
};

// the fields UpcGenerator::generateEvent fills (src/UpcGenerator.cpp:817-829)
class TParticle
{
 public:
  TParticle() = default;
  TParticle(int pdg, int status, int mother1, int mother2, int daughter1, int daughter2, double px, double py, double pz,
            double etot, double vx, double vy, double vz, double time)
    : fPdgCode(pdg), fStatusCode(status), fPx(px), fPy(py), fPz(pz), fE(etot), fVx(vx), fVy(vy), fVz(vz), fVt(time)
  {
    fMother[0] = mother1; fMother[1] = mother2; fDaughter[0] = daughter1; fDaughter[1] = daughter2;
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
  int fPdgCode{0}, fStatusCode{0}, fMother[2]{0, 0}, fDaughter[2]{-1, -1};
  double fPx{0}, fPy{0}, fPz{0}, fE{0}, fVx{0}, fVy{0}, fVz{0}, fVt{0};
};
#endif  // UPC_WITH_ROOT

// plog-style logging macros (stream interface, severity prefix)
#ifndef PLOG_INFO
struct UpcLogLine {
  std::ostringstream os;
  const char* tag;
  explicit UpcLogLine(const char* t) : tag(t) {}
  ~UpcLogLine() { std::cerr << tag << os.str() << std::endl; }
  template <class T> UpcLogLine& operator<<(const T& v) { os << v; return *this; }
  UpcLogLine& operator<<(std::ios_base& (*f)(std::ios_base&)) { os << f; return *this; }
};
#define PLOG_INFO UpcLogLine("[INFO ] ")
#define PLOG_WARNING UpcLogLine("[WARN ] ")
#define PLOG_FATAL UpcLogLine("[FATAL] ")
#define PLOG_DEBUG UpcLogLine("[DEBUG] ")
#endif
