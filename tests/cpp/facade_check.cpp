// Exercises the C++ facade (UpcCrossSection / UpcSampler / UpcGenerator with the reference's
// method names) and prints a JSON line that tests/test_cpp_facade.py compares with the oracle.
#include <cstdio>
#include <vector>

#include "UpcCrossSection.h"
#include "UpcGenerator.h"
#include "UpcSampler.h"

int main()
{
  UpcCrossSection cs;
  cs.isPoint = false;
  cs.breakupMode = 2;
  cs.nm = 24; cs.ny = 10;
  cs.mmin = 3.56; cs.mmax = 50.;
  cs.setElemProcess(13);
  cs.evIsPair = true;
  cs.init();
  upc_host::registerSamplerContext(cs.gpu());
  std::printf("{\"rho0\": %.17g, \"fluxPoint\": %.17g, \"fluxForm\": %.17g, \"breakup\": %.17g, \"lumi\": %.17g, ",
              UpcCrossSection::rho0, cs.fluxPoint(10., 1.), cs.fluxForm(5., 1.), cs.calcBreakupProb(15., 2),
              cs.calcTwoPhotonLumi(10., 0.5));
  std::vector<std::vector<double>> csYM(cs.ny, std::vector<double>(cs.nm, 0.)), ratio;
  double totCS = 0;
  cs.calcNucCrossSectionYM(csYM, ratio, totCS);
  std::printf("\"totCS\": %.17g, \"cs00\": %.17g, \"cs_last\": %.17g, ", totCS, csYM[0][0], csYM[cs.ny - 1][cs.nm - 1]);
  std::vector<double> ye(cs.ny + 1), me(cs.nm + 1);
  for (int i = 0; i <= cs.ny; i++) ye[i] = cs.ymin + (cs.ymax - cs.ymin) / cs.ny * i;
  for (int i = 0; i <= cs.nm; i++) me[i] = cs.mmin + (cs.mmax - cs.mmin) / cs.nm * i;
  UpcSampler2D s2(csYM, ye, me, 12345);
  std::printf("\"sum_last\": %.17g, \"samples\": [", s2.sum.back());
  for (int i = 0; i < 6; i++) {
    double y, m;
    s2(y, m);
    std::printf("%s[%.17g, %.17g, %d, %d]", i ? ", " : "", y, m, s2.getBinX(y), s2.getBinY(m));
  }
  std::vector<double> d1 = {1., 2., 0., 4., 3.}, e1 = {0., 1., 2., 3., 4., 5.};
  UpcSampler1D s1(d1, e1, 7);
  std::printf("], \"s1\": [%.17g, %.17g, %.17g], \"s1sum\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g]}\n", s1(), s1(), s1(),
              s1.sum[0], s1.sum[1], s1.sum[2], s1.sum[3], s1.sum[4], s1.sum[5]);
  return 0;
}
