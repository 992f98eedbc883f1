// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
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
