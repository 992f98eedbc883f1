#include "UpcTwoPhotonTabulated.h"

#include <cstdlib>

#ifndef CROSS_SEC_DIR
#define CROSS_SEC_DIR ""
#endif

std::string upcCrossSecDir()
{
  if (const char* e = std::getenv("UPCGEN_CROSS_SEC_DIR"))
    if (*e) return e;
  const std::string built = CROSS_SEC_DIR;
  return built.empty() ? std::string("cross_sections") : built;
}

UpcTwoPhotonTabulated::UpcTwoPhotonTabulated(const std::string& subdir, bool doMassCut, double lowMCut, double hiMCut)
{
  const std::string dir = upcCrossSecDir() + "/" + subdir;
  hCrossSectionM = new UpcRootHist();
  hCrossSectionZM = new UpcRootHist();
  ok = hCrossSectionM->Read(dir + "/cross_section_m.root", "hCrossSectionM", error) &&
       hCrossSectionZM->Read(dir + "/cross_section_zm.root", "hCrossSectionZM", error);
  if (ok && (hCrossSectionM->dim != 1 || hCrossSectionZM->dim != 2)) {
    ok = false;
    error = dir + ": hCrossSectionM must be a TH1D and hCrossSectionZM a TH2D";
  }
  if (!ok) return;
  // the reference's "temporary workaround": mass bins outside the cut are emptied in both histograms
  if (doMassCut) {
    const int lowBin = hCrossSectionM->GetXaxis()->FindBin(lowMCut);
    const int hiBin = hCrossSectionM->GetXaxis()->FindBin(hiMCut);
    for (int i = 1; i <= hCrossSectionM->GetNbinsX(); i++) {
      if (i >= lowBin && i <= hiBin) continue;
      hCrossSectionM->SetBinContent(i, 0.);
      for (int j = 1; j <= hCrossSectionZM->GetNbinsX(); j++) hCrossSectionZM->SetBinContent(j, i, 0.);
    }
  }
}

double UpcTwoPhotonTabulated::calcCrossSectionM(double m)
{
  if (!ok) return 0.;
  return hCrossSectionM->GetBinContent(hCrossSectionM->GetXaxis()->FindBin(m));  // [nb]
}

double UpcTwoPhotonTabulated::calcCrossSectionZM(double z, double m)
{
  if (!ok) return 0.;
  return hCrossSectionZM->GetBinContent(hCrossSectionZM->GetXaxis()->FindBin(z), hCrossSectionZM->GetYaxis()->FindBin(m));
}
