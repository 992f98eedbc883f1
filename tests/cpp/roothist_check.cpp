// CPU check of UpcRootAxis / UpcRootHist look-up semantics (TAxis::FindBin, TH1/TH2::GetBinContent):
// uniform and variable-width axes, edges, under- and overflow, out-of-range clamping.
#include <cstdio>
#include <cstdlib>

#include "UpcRootHist.h"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main()
{
  UpcRootAxis u;
  u.fNbins = 10; u.fXmin = -1.; u.fXmax = 4.;
  CHECK(u.FindBin(-1.) == 1);          // the lower edge belongs to the first bin
  CHECK(u.FindBin(-1.0000001) == 0);   // underflow
  CHECK(u.FindBin(4.) == 11);          // the upper edge is overflow
  CHECK(u.FindBin(3.9999) == 10);
  CHECK(u.FindBin(0.) == 3);           // 1 + int(10 * 1 / 5)
  CHECK(u.FindBin(1e30) == 11);
  UpcRootAxis v;
  v.fNbins = 4; v.fXmin = 0.; v.fXmax = 10.;
  v.fXbins = {0., 1., 2.5, 6., 10.};
  CHECK(v.FindBin(0.) == 1 && v.FindBin(0.999) == 1 && v.FindBin(1.) == 2 && v.FindBin(2.49) == 2);
  CHECK(v.FindBin(2.5) == 3 && v.FindBin(5.99) == 3 && v.FindBin(6.) == 4 && v.FindBin(9.99) == 4);
  CHECK(v.FindBin(10.) == 5 && v.FindBin(-0.1) == 0);

  UpcRootHist h;
  h.dim = 2;
  h.fXaxis = u;                        // 10 x bins -> 12 cells per row
  h.fYaxis = v;                        // 4 y bins  -> 6 rows
  h.fArray.assign(12 * 6, 0.);
  for (int by = 0; by < 6; ++by)
    for (int bx = 0; bx < 12; ++bx) h.fArray[bx + 12 * by] = 100. * by + bx;
  CHECK(h.GetBinContent(3, 2) == 203.);
  CHECK(h.GetBinContent(h.GetXaxis()->FindBin(0.), h.GetYaxis()->FindBin(1.)) == 203.);
  CHECK(h.GetBinContent(-5, 99) == 500.);      // clamped to the under- / overflow cells
  h.SetBinContent(3, 2, -7.);
  CHECK(h.GetBinContent(3, 2) == -7.);
  h.SetBinContent(40, 2, 1.);                  // out of range: ignored
  CHECK(h.GetNbinsX() == 10 && h.GetNbinsY() == 4);
  std::string err;
  UpcRootHist none;
  CHECK(!none.Read("/nonexistent/file.root", "h", err) && !err.empty());
  std::printf("ROOTHIST_OK\n");
  return 0;
}
