// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// Closed-form gamma gamma -> l+ l- cross sections, incl. the anomalous-moment terms.
// Formulas as evaluated by the reference (src/UpcTwoPhotonDilep.cpp:46-134); the operation order
// inside each expression is kept so that sigma(m) is bit-identical to a reference build, which
// the bit-exact sampler tables depend on.
#include "UpcTwoPhotonDilep.h"

#include <cmath>

#include "UpcPhysConstants.h"

using phys_consts::alpha;
using phys_consts::hc;

UpcTwoPhotonDilep::UpcTwoPhotonDilep(int pdg)
{
  partPDG = pdg;
  isCharged = true;
  switch (pdg) {
    case 11: mPart = phys_consts::mEl; break;
    case 13: mPart = phys_consts::mMu; break;
    case 15: mPart = phys_consts::mTau; break;
    default: break;
  }
}

// total unpolarised cross section in nb (:46-61)
double UpcTwoPhotonDilep::calcCrossSectionM(double m)
{
  const double s = m * m;
  const double x = 4 * mPart * mPart / s; // 1/gamma^2 of the lepton in the pair rest frame
  const double b = std::sqrt(1 - x);      // lepton velocity
  const double y = std::atanh(b);         // lepton rapidity
  const double a = aLep; // powers are applied factor by factor, as the reference does
  double cs = 0;
  cs += (2 + 2 * x - x * x) * y - b * (1 + x);
  cs += 4 * y * a;
  cs += (4 * b / x + y) * a * a;
  cs += (4 * b / x - 2 * y) * a * a * a;
  cs += ((7. / 12.) * b / x + (1. / 6.) * b / x / x - 0.5 * y) * a * a * a * a;
  cs *= 4 * hc * hc * 1e7 * alpha * alpha * M_PI / s;
  return cs;
}

// d sigma / dz in GeV^-2 (:63-83)
double UpcTwoPhotonDilep::calcCrossSectionZM(double z, double m)
{
  const double s = m * m;
  const double k = std::sqrt(s) / 2.;                  // photon energy in the pair rest frame
  const double p = std::sqrt(k * k - mPart * mPart);   // lepton momentum
  const double norm = 2 * M_PI * alpha * alpha / s * p / k;
  const double kt = -2 * k * (k - z * p) / mPart / mPart;
  const double ku = -2 * k * (k + z * p) / mPart / mPart;
  const double ks = kt + ku;
  const double kp = kt * ku;
  const double kq = 1. / kt + 1. / ku;
  const double kr = ku / kt + kt / ku;
  const double a = aLep;
  double cs = 0;
  cs += -8. * (4. * kq * kq + 4. * kq - kr);
  cs += 16. * (2. + kr) * a;
  cs += 4. * (2. - 4. * ks + kr) * a * a;
  cs += -8. * (2. + 2. * ks + kr) * a * a * a;
  cs += -4. * (4. + 2. * ks + 2. * kr - kp) * a * a * a * a;
  cs *= norm;
  return cs;
}

namespace
{
// common log term 2 ln(1/r + sqrt(1/r^2 - 1)) and velocity sqrt(1 - r^2) of the polarised totals
inline double polTotal(double m, double mPart, double c4, double c2)
{
  const double r = 2 * mPart / m;
  if (r > 1) return 0;
  return 4 * M_PI * alpha * alpha * hc * hc / m / m *
         ((1 + r * r - c4 * r * r * r * r) * 2 * std::log(1 / r + std::sqrt(1 / r / r - 1)) -
          (1 + c2 * r * r) * std::sqrt(1 - r * r));
}
} // namespace

// fm^2 (:85-95)
double UpcTwoPhotonDilep::calcCrossSectionMPolS(double m) { return polTotal(m, mPart, 3. / 4., 3. / 2.); }

// fm^2 (:111-121)
double UpcTwoPhotonDilep::calcCrossSectionMPolPS(double m) { return polTotal(m, mPart, 1. / 4., 1. / 2.); }

// GeV^-2 (:97-109)
double UpcTwoPhotonDilep::calcCrossSectionZMPolS(double z, double m)
{
  const double mLep2 = mPart * mPart, m2 = m * m, z2 = z * z;
  const double den = m2 * (1 - z2) + 4 * mLep2 * z2;
  double cs = 2 * M_PI * alpha * alpha;
  cs *= m2 - 4 * mLep2;
  cs *= std::sqrt(m2 - 4 * mLep2);
  cs *= (4 * mLep2 * (3 - 2 * z2 + z2 * z2)) + m2 * (1 - z2 * z2);
  cs /= m2 * m * den * den;
  return cs;
}

// GeV^-2 (:123-134)
double UpcTwoPhotonDilep::calcCrossSectionZMPolPS(double z, double m)
{
  const double mLep2 = mPart * mPart, m2 = m * m, z2 = z * z;
  const double den = m2 * (1 - z2) + 4 * mLep2 * z2;
  double cs = 2 * M_PI * alpha * alpha;
  cs *= std::sqrt(m2 - 4 * mLep2);
  cs *= m2 * m2 * (1 - z2 * z2) + 8 * m2 * mLep2 * (1 - z2 + z2 * z2) - 16 * mLep2 * mLep2 * (1 - z2) * (1 - z2);
  cs /= m2 * m * den * den;
  return cs;
}
