// Exercises the vector-meson plug-in (upcgen_b200/host/UpcPhotoNuclearVM.cpp) without a GPU: prints a JSON line with
// sigma(y), the integrated squared form factor and the shadowing factor for tests/test_vm_path.py.
//   vm_check <pdg> <shadowing> <dght> <Z> <A> <R> <a> <sqrts> <rho0>
#include <cstdio>
#include <cstdlib>

#include "UpcCrossSection.h"
#include "UpcPhotoNuclearVM.h"

int main(int argc, char** argv)
{
  if (argc < 10) return 2;
  const int pdg = std::atoi(argv[1]), shad = std::atoi(argv[2]), dght = std::atoi(argv[3]);
  UpcCrossSection::Z = std::atoi(argv[4]);
  UpcCrossSection::A = std::atoi(argv[5]);
  UpcCrossSection::R = std::atof(argv[6]);
  UpcCrossSection::a = std::atof(argv[7]);
  UpcCrossSection::sqrts = std::atof(argv[8]);
  UpcCrossSection::rho0 = std::atof(argv[9]);
  UpcCrossSection::mNucl = (UpcCrossSection::Z * phys_consts::mProt + (UpcCrossSection::A - UpcCrossSection::Z) * phys_consts::mNeut) / UpcCrossSection::A;
  UpcPhotoNuclearVM vm(pdg, shad, dght);
  if (!vm.ok) { std::printf("{\"error\": \"%s\"}\n", vm.error.c_str()); return 0; }
  std::printf("{\"mPart\": %.17g, \"mDght\": %.17g, \"sigma\": [", vm.mPart, vm.mDght);
  for (int i = 0; i <= 24; i++) std::printf("%s%.17g", i ? ", " : "", vm.calcCrossSectionY(-6. + 0.5 * i));
  std::printf("], \"phi\": [");
  const double tm[5] = {1e-8, 1e-5, 1e-3, 1e-2, 0.3};
  for (int i = 0; i < 5; i++) std::printf("%s%.17g", i ? ", " : "", UpcPhotoNuclearVM::integrateFormFactorSq(tm[i], tm[i] + 1.));
  std::printf("], \"dsdt\": [%.17g, %.17g, %.17g]", vm.dsdt(3.0), vm.dsdt(10.), vm.dsdt(100.));
  if (shad == 4) {
    std::printf(", \"rg\": [");
    const double xs[8] = {1e-6, 5e-6, 2e-5, 1e-4, 1.3e-3, 2e-2, 0.09, 0.5};
    for (int i = 0; i < 8; i++) std::printf("%s%.17g", i ? ", " : "", vm.getRgLtaVG(xs[i]));
    std::printf("]");
  }
  if (shad == 2 || shad == 3) {
    // the FGS10 patch at the table's own nodes (first, an inner one, last in x; mu^2 of the plug-in) and between them
    std::printf(", \"rg\": [");
    const double xs[8] = {1e-6, 9.99999975e-6, 3e-5, 1.2e-4, 1.3e-3, 2e-2, 0.4, 0.99};
    for (int i = 0; i < 8; i++) std::printf("%s%.17g", i ? ", " : "", vm.getRgLta(shad == 2 ? 0 : 1, xs[i]));
    std::printf("]");
    if (!vm.ok) std::printf(", \"error\": \"%s\"", vm.error.c_str());
  }
  std::printf("}\n");
  return 0;
}
