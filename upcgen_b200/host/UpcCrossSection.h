// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// UpcCrossSection -- the reference's nuclear-physics numerics class (include/UpcCrossSection.h)
// with the same public data members and methods, implemented over the C-ABI of the CUDA library
// (include/upcgpu.h).  Tables, fluxes, the luminosity grid and the sigma fold run on the GPU; the
// elementary-process plug-in (elemProcess) stays host code exactly as in the reference.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/upcgpu.h"
#include "UpcCompat.h"
#include "UpcElemProcess.h"
#include "UpcPhysConstants.h"
#include "UpcTwoPhotonALP.h"
#include "UpcTwoPhotonDilep.h"

class UpcCrossSection
{
 public:
  UpcCrossSection();
  ~UpcCrossSection();

  // Woods-Saxon parameters
  inline static double rho0{0.}; // fm^-3
  inline static double R{6.68};  // fm
  inline static double a{0.447}; // fm

  // parameters of the nucleus
  inline static int Z{82};
  inline static int A{208};
  inline static double mNucl{(Z * phys_consts::mProt + (A - Z) * phys_consts::mNeut) / A};

  // beam parameters
  inline static double sqrts{5020.};
  inline static double g1{sqrts / (2. * phys_consts::mProt)};
  inline static double g2{sqrts / (2. * phys_consts::mProt)};

  // photon luminosity calculation parameters
  TString lumiFileDirectory{"."};
  void setLumiFileDirectory(TString directory) { lumiFileDirectory = directory; };
  const int nb1{120};
  const int nb2{120};

  // cross sections binning
  double zmin{-1};
  double zmax{1};
  int nz{100};
  double mmin{3.56};
  double mmax{50.};
  int nm{1001};
  double ymin{-6.};
  double ymax{6.};
  int ny{121};

  bool doMassCut{false};
  double lowMCut{0};
  double hiMCut{9999};

  double factor{Z * Z * phys_consts::alpha / M_PI / M_PI / phys_consts::hc / phys_consts::hc};
  static const int nb{200};
  double gtot; // fixed in the constructor (from the default sqrts), as the reference does

  bool isPoint{true};
  bool useNonzeroGamPt{true};
  bool usePolarizedCS{false};
  int breakupMode{1};
  int shadowingOption{0};
  int dghtPDG{};

  inline static int debug{0};
  int numThreads{1}; // kept for interface compatibility; the GPU path does not use host threads

  void setElemProcess(int procID);
  UpcElemProcess* elemProcess{nullptr};

  double alpMass{1.};
  double alpWidth{0.010};

  void init();

  template <typename ArrayType>
  static double simpson(int n, ArrayType* v, double h);
  double calcWSRho();
  double fluxPoint(double b, double k);
  static double calcFormFac(double Q2);
  double fluxForm(double b, double k);
  double calcTwoPhotonLumi(double M, double Y);
  void calcTwoPhotonLumiPol(double& ns, double& np, double M, double Y);
  double calcPhotonFlux(double M, double Y);
  // NB: as in the reference, the declaration names (mmin.. first) and the definition's
  // (zmin.. first) differ; the call site follows the definition
  void fillCrossSectionZM(std::vector<std::vector<double>>& crossSectionZM, double zmin, double zmax, int nz,
                          double mmin, double mmax, int nm, int flag);
  void calcNucCrossSectionYM(std::vector<std::vector<double>>& crossSectionYM,
                             std::vector<std::vector<double>>& polCSRatio, double& totCS);
  void calcNucCrossSectionY(std::vector<std::vector<double>>& crossSectionY,
                            std::vector<std::vector<double>>& csYRatio, double& totCS);
  double calcBreakupProb(double b, int mode);
  double getPhotonPt(double ePhot);
  void getPairMomentum(double mPair, double yPair, TLorentzVector& pPair);
  void getMomentumVM(double m, double y, int target, TLorentzVector& pPair);

  void prepareGAA();
  void prepareBreakupProb();
  void prepareFormFac();
  void prepareTwoPhotonLumi();

  // ---- additions of the GPU build (not in the reference) ----
  int device{0};                 // CUDA device of this instance (first device when numGpus > 1)
  int numGpus{1};                // devices used for the table stage and the event stage (-ngpus; devices device .. device+numGpus-1)
  upcgpu_ctx* gpu() { return ctx; }
  const std::vector<double>& lumiTable() const { return lumi; }     // [nm][ny], x dm dy
  const std::vector<double>& lumiTableS() const { return lumiS; }
  const std::vector<double>& lumiTablePs() const { return lumiPs; }
  upcgpu_params makeParams() const;   // the parameter block handed to the C-ABI
  // event-stage knobs forwarded by UpcGenerator into the parameter block
  bool evIsPair{false}, evIsSingle{false}, evIgnoreCSZ{false};
  int evDecayUniformPDG{0};
  bool evDoPtCut{false}, evDoEtaCut{false};
  double evMinPt{0}, evMinEta{0}, evMaxEta{0};

 private:
  upcgpu_ctx* ctx{nullptr};
  bool tablesReady{false};
  std::vector<double> lumi, lumiS, lumiPs;
  std::map<int, std::vector<double>> photPtCdf; // per-MeV cache, as photPtDistrMap
  void ensureContext();
  void ensureTables();
  [[noreturn]] void fail(const char* what, int rc);
};
