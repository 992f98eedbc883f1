#include "UpcTwoPhotonALP.h"

#include <cmath>

#include "UpcPhysConstants.h"

// sigma = 4 pi^2 Gamma / m_a^2 * (hc)^2 * 1e7 * alpha^2  [nb], independent of m
double UpcTwoPhotonALP::calcCrossSectionM(double /*m*/)
{
  double cs = 4. * M_PI * M_PI * width / (mPart * mPart);
  cs *= phys_consts::hc * phys_consts::hc * 1e7 * phys_consts::alpha * phys_consts::alpha;
  return cs;
}
