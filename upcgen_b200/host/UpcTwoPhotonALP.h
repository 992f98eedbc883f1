// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// gamma gamma -> ALP in the narrow-resonance approximation (reference: include/UpcTwoPhotonALP.h,
// src/UpcTwoPhotonALP.cpp:28-33).  cos(theta) is not modelled (ignoreCSZ).
#pragma once
#include "UpcElemProcess.h"

class UpcTwoPhotonALP : public UpcElemProcess
{
 public:
  UpcTwoPhotonALP(double mass, double width, int spin = 0) : width{width}, spin{spin}
  {
    partPDG = 51; // spin-0 axion-like particle, PDG MC numbering
    mPart = mass;
    isCharged = false;
  }
  ~UpcTwoPhotonALP() override = default;

  double width{0.001};
  int spin{0};

  double calcCrossSectionM(double m) override;
  double calcCrossSectionZM(double, double) override { return 0; }
  double calcCrossSectionMPolS(double) override { return 0; }
  double calcCrossSectionZMPolS(double, double) override { return 0; }
  double calcCrossSectionMPolPS(double) override { return 0; }
  double calcCrossSectionZMPolPS(double, double) override { return 0; }
};
