// gamma gamma -> l+ l-  (reference: include/UpcTwoPhotonDilep.h, src/UpcTwoPhotonDilep.cpp)
#pragma once
#include "UpcElemProcess.h"

class UpcTwoPhotonDilep : public UpcElemProcess
{
 public:
  explicit UpcTwoPhotonDilep(int partPDG);
  ~UpcTwoPhotonDilep() override = default;

  double aLep{0}; // anomalous magnetic moment

  double calcCrossSectionM(double m) override;
  double calcCrossSectionZM(double z, double m) override;
  double calcCrossSectionMPolS(double m) override;
  double calcCrossSectionZMPolS(double z, double m) override;
  double calcCrossSectionMPolPS(double m) override;
  double calcCrossSectionZMPolPS(double z, double m) override;
};
