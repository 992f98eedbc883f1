// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
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
