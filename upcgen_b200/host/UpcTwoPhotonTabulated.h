// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// Elementary processes whose cross sections the reference ships as histograms in ROOT files:
//   UpcTwoPhotonLbyL    gamma gamma -> gamma gamma   (include/UpcTwoPhotonLbyL.h, src/UpcTwoPhotonLbyL.cpp)
//   UpcTwoPhotonDipion  gamma gamma -> pi0 pi0       (include/UpcTwoPhotonDipion.h, src/UpcTwoPhotonDipion.cpp)
// Same class names, constructor arguments, public histogram members and look-up semantics
// (sigma(m) = content of the bin TAxis::FindBin(m) selects, [nb]; dsigma/dz likewise from the (z, m) histogram; the
// DO_M_CUT workaround zeroes the mass bins outside [lowMCut, hiMCut]); the histograms are read by UpcRootHist
// instead of TFile / TH1D / TH2D.  Polarised cross sections do not exist for these processes (they return 0 as in the
// reference).
//
// Where the files are looked for: $UPCGEN_CROSS_SEC_DIR, else the compile-time CROSS_SEC_DIR (the reference's
// CMake definition), else ./cross_sections .
#pragma once
#include <string>

#include "UpcElemProcess.h"
#include "UpcRootHist.h"

std::string upcCrossSecDir();

class UpcTwoPhotonTabulated : public UpcElemProcess
{
 public:
  UpcTwoPhotonTabulated(const std::string& subdir, bool doMassCut, double lowMCut, double hiMCut);
  ~UpcTwoPhotonTabulated() override
  {
    delete hCrossSectionM;
    delete hCrossSectionZM;
  }

  // for these processes, cross sections are stored in files
  UpcRootHist* hCrossSectionM{nullptr};
  UpcRootHist* hCrossSectionZM{nullptr};
  bool ok{false};          // both histograms were read
  std::string error;       // why not

  double calcCrossSectionM(double m) override;
  double calcCrossSectionZM(double z, double m) override;
  double calcCrossSectionMPolS(double) override { return 0; }
  double calcCrossSectionZMPolS(double, double) override { return 0; }
  double calcCrossSectionMPolPS(double) override { return 0; }
  double calcCrossSectionZMPolPS(double, double) override { return 0; }
};

class UpcTwoPhotonLbyL : public UpcTwoPhotonTabulated
{
 public:
  UpcTwoPhotonLbyL(bool doMassCut, double lowMCut, double hiMCut) : UpcTwoPhotonTabulated("lbyl", doMassCut, lowMCut, hiMCut)
  {
    mPart = 0.0;
    partPDG = 22;
    isCharged = false;
  }
};

class UpcTwoPhotonDipion : public UpcTwoPhotonTabulated
{
 public:
  UpcTwoPhotonDipion(bool doMassCut, double lowMCut, double hiMCut) : UpcTwoPhotonTabulated("pi0pi0", doMassCut, lowMCut, hiMCut)
  {
    mPart = 0.1349770;  // pi0 mass from PDG
    partPDG = 111;
    isCharged = false;
  }
};
