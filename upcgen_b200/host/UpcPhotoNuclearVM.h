// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// UpcPhotoNuclearVM -- elementary process of the vector-meson path: coherent photoproduction of J/psi, psi(2S) and
// Upsilon(1S) off a nucleus (reference: include/UpcPhotoNuclearVM.h, src/UpcPhotoNuclearVM.cpp).  Same class name,
// constructor and virtuals.  Host code, like every elementary-process plug-in; the b-integrated photon flux it is
// folded with (UpcCrossSection::calcPhotonFlux) runs on the GPU.
//
// sigma(y) = cAcP^2 * dsigma/dt(W_gp)|_{t=0} * Rg(x)^2 * Phi_A(t_min) * 1e-6  (src/UpcPhotoNuclearVM.cpp:340-381)
//   dsigma/dt: the power-law fit of :41-52;  Phi_A = integral of the squared nuclear form factor over [t_min, t_min + 1]
//   (the reference uses TF1::Integral; here an adaptive Gauss-Kronrod rule to 1e-12);  Rg: gluon shadowing --
//   SHADOWING 0 (impulse approximation, Rg = 1) and 4 (leading-twist approximation, Guzey-Zhalov tables
//   cross_sections/vm/lta/LT2013_pb208_cteq6l1_m12_Q2_{3,4}.dat, read from $UPCGEN_CROSS_SEC_DIR; TSpline3 = a
//   not-a-knot cubic spline through the 37 points), 2 / 3 (FGS10 weak / strong shadowing for psi(2S) and Upsilon:
//   cross_sections/vm/lta/QCDEvolution_pb208proton_2009_model{2,1}.dat through the bicubic interpolation of
//   gsl_spline2d; for J/psi the reference's condition mu^2 >= 4 leaves the impulse approximation in place, and so
//   does this).  SHADOWING 1 needs EPS09 grids the reference does not ship: not provided here.
#pragma once
#include <string>
#include <vector>

#include "UpcElemProcess.h"

class UpcPhotoNuclearVM : public UpcElemProcess
{
 public:
  explicit UpcPhotoNuclearVM(int partPDG, int shadowingOpt, int dghtPDG);
  ~UpcPhotoNuclearVM() override = default;

  double calcCrossSectionY(double y) override;
  double calcCrossSectionZM(double, double) override { return 0.; }
  double calcCrossSectionMPolS(double) override { return 0.; }
  double calcCrossSectionZMPolS(double, double) override { return 0.; }
  double calcCrossSectionMPolPS(double) override { return 0.; }
  double calcCrossSectionZMPolPS(double, double) override { return 0.; }

  bool ok{true};
  std::string error;

  // pieces, public for the tests
  double dsdt(double Wgp) const;
  static double integrateFormFactorSq(double tmin, double tmax);
  double getRgLtaVG(double x);
  double getRgLta(int type, double x);

 private:
  int fShadowing{0};
  double fMu2{1.};
  double fC0{0}, fPw{0.4}, fMmin{0};
  // the LTA table and its not-a-knot spline (TSpline3 of the reference)
  std::vector<double> fX, fY, fB, fC, fD;
  bool fLtaInit{false};
  // the FGS10 glue ratio on its (x, Q^2) grid with the derivatives of the bicubic interpolation; row = Q^2 index
  std::vector<double> fGx, fGq, fGz, fGzx, fGzy, fGzxy;
  int fFgsType{-1};
};
