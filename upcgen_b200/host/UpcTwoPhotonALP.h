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
