// Interface compatibility notice: the class, member and method names declared in this file reproduce the public
// interface of Upcgen (https://github.com/nburmaso/upcgen), Copyright (C) 2021-2025 Nazar Burmasov, Evgeny Kryshen,
// distributed under the GNU General Public License, version 3 or later (see LICENSE-UPCGEN-NOTICE.md at the
// repository root).  They are kept identical so that code written against the reference compiles against this
// drop-in; the implementation behind them is this project's own.
// Elementary-process plug-in interface: the reference's only "plugin API"
// (include/UpcElemProcess.h:26-62).  Same member names, same virtuals, so that a process class
// written against the reference compiles against this header unchanged.  The plug-ins stay host
// code: they are evaluated nm (or nm*nz) times per run and their results cross the C-ABI as
// plain arrays (upcgpu_fold_sigma / upcgpu_sampler_build).
#pragma once

class UpcElemProcess
{
 public:
  UpcElemProcess() = default;
  virtual ~UpcElemProcess() {}

  double mPart{};   // mass of the produced particle
  double mDght{};   // mass of a decay daughter
  int dghtPDG{};
  int partPDG{};
  bool isCharged{};

  // unpolarised cross sections
  virtual double calcCrossSectionY(double /*y*/) { return 0.; }
  virtual double calcCrossSectionM(double /*m*/) { return 0.; }
  virtual double calcCrossSectionZM(double z, double m) = 0;

  // polarised cross sections: scalar and pseudoscalar parts
  virtual double calcCrossSectionMPolS(double m) = 0;
  virtual double calcCrossSectionZMPolS(double z, double m) = 0;
  virtual double calcCrossSectionMPolPS(double m) = 0;
  virtual double calcCrossSectionZMPolPS(double z, double m) = 0;
};
