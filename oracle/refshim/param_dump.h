// param_dump.h -- what configGeneratorFromFile left in a UpcGenerator, as "KEY value" lines.  The member names are the
// reference's (include/UpcGenerator.h, include/UpcCrossSection.h); the host facade keeps them, so this one function is
// compiled against either set of headers: by oracle/refshim/gen_capi.cpp over the REFERENCE's parser and by
// tests/cpp/parse_check.cpp over the drop-in's.  Test infrastructure.  Include after UpcGenerator.h with its private
// members opened (#define private public around the include).
#pragma once
#include <cstdio>
#include <string>

template <class Gen>
std::string upc_param_dump(Gen& g)
{
  char b[256];
  std::string s;
  auto add = [&](const char* k, double v) { std::snprintf(b, sizeof b, "%s %.17g\n", k, v); s += b; };
  auto* cs = g.nucProcessCS;
  add("NUCLEUS_Z", cs->Z); add("NUCLEUS_A", cs->A); add("WS_R", cs->R); add("WS_A", cs->a);
  add("SQRTS", cs->sqrts); add("G1", cs->g1); add("G2", cs->g2);
  add("PROC_ID", g.procID); add("LEP_A", g.aLep); add("ALP_MASS", cs->alpMass); add("ALP_WIDTH", cs->alpWidth);
  add("NEVENTS", (double)g.nEvents);
  add("DO_PT_CUT", g.doPtCut); add("PT_MIN", g.minPt); add("DO_ETA_CUT", g.doEtaCut); add("ETA_MIN", g.minEta);
  add("ETA_MAX", g.maxEta);
  add("ZMIN", cs->zmin); add("ZMAX", cs->zmax); add("MMIN", cs->mmin); add("MMAX", cs->mmax); add("YMIN", cs->ymin);
  add("YMAX", cs->ymax); add("BINS_Z", cs->nz); add("BINS_M", cs->nm); add("BINS_Y", cs->ny);
  add("FLUX_POINT", cs->isPoint); add("BREAKUP_MODE", cs->breakupMode); add("NON_ZERO_GAM_PT", cs->useNonzeroGamPt);
  add("USE_POLARIZED_CS", g.usePolarizedCS); add("CS_USE_POLARIZED_CS", cs->usePolarizedCS);  // the cross-section object's own copy of the flag
  add("PYTHIA_VERSION", g.pythiaVersion); add("PYTHIA8_FSR", g.doFSR); add("PYTHIA8_DECAYS", g.doDecays);
  add("SEED", (double)g.seed); add("USE_ROOT_OUTPUT", g.useROOTOut); add("USE_HEPMC_OUTPUT", g.useHepMCOut);
  add("DO_M_CUT", cs->doMassCut); add("LOW_M_CUT", cs->lowMCut); add("HIGH_M_CUT", cs->hiMCut);
  add("SHADOWING", cs->shadowingOption); add("DECAY_PDG", cs->dghtPDG);
  return s;
}
