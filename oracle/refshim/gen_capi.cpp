// C entry points over the REFERENCE's own UpcGenerator / UpcSampler (src/UpcGenerator.cpp, include/UpcSampler.h,
// compiled unmodified from /root/reference against the GSL/ROOT shim): the event stage of the reference -- sampler
// construction, generateEvent, getPairMomentum / getPhotonPt, pair / single production, the uniform decay, the cuts,
// the HepMC writer -- runs here as the reference wrote it, driven by its own MT19937 streams.  TEST INFRASTRUCTURE:
// the statistical-compatibility tests draw their reference sample from this.
#include <pthread.h>

#include <cstdio>
#include <cstring>
#include <string>

#include <fstream>
#include <iomanip>
#include <sstream>

#include "UpcCrossSection.h"  // everything UpcGenerator.h includes, with its own access control ...
#include "UpcPythia6Helper.h"
#include "UpcPythia8Helper.h"
#include "UpcSampler.h"
#define private public         // ... so that only UpcGenerator's members open up: the tests read its tables and samplers
#include "UpcGenerator.h"
#undef private
#include "param_dump.h"

static UpcGenerator* g_gen = nullptr;

struct GenInit {
  const char* parfile;
  const char* lumi_dir;
};

static void* gen_init_thread(void* vp)
{
  const GenInit* gi = (const GenInit*)vp;
  // main.cpp:102-110
  g_gen = new UpcGenerator();
  g_gen->setDebugLevel(0);
  g_gen->setNumThreads(1);
  g_gen->setParFile(gi->parfile);
  g_gen->configGeneratorFromFile();
  g_gen->setLumiFileDirectory(gi->lumi_dir);
  g_gen->init();
  return nullptr;
}

extern "C" {

// parameters.in through the REFERENCE's parser only (new UpcGenerator, setParFile, configGeneratorFromFile -- no init):
// the resulting parameter block as "KEY value" lines.  Returns the text's length (truncated to cap - 1).
long upcrefgen_parse(const char* parfile, char* out, long cap)
{
  auto* g = new UpcGenerator();
  g->setDebugLevel(0);
  g->setParFile(parfile);
  g->configGeneratorFromFile();
  const std::string s = upc_param_dump(*g);
  const long n = (long)s.size() < cap - 1 ? (long)s.size() : cap - 1;
  std::memcpy(out, s.data(), n);
  out[n] = 0;
  return n;   // (the generator is leaked on purpose: ~UpcGenerator deletes gRandom)
}

// Places a luminosity table into the shim's in-memory "ROOT file" <dir>/twoPhotonLumi[Pol].root as the TH2D(s) the
// reference writes (src/UpcCrossSection.cpp:503-506, :564-571: bin (im + 1, iy + 1) = table[im][iy]), and the marker
// file that makes prepareTwoPhotonLumi take its "found pre-calculated luminosity" branch (:481-491).
int upcrefgen_put_lumi(const char* dir, int pol, int nm, int ny, double mmin, double mmax, double ymin, double ymax,
                       const double* lumi, const double* lumi_s, const double* lumi_p)
{
  std::string fname = std::string(dir) + "/twoPhotonLumi" + (pol ? "Pol.root" : ".root");
  TFile f(fname.c_str(), "recreate");
  const char* names[2] = {pol ? "hD2LDMDY_s" : "hD2LDMDY", "hD2LDMDY_p"};
  const double* tabs[2] = {pol ? lumi_s : lumi, lumi_p};
  for (int t = 0; t < (pol ? 2 : 1); t++) {
    if (!tabs[t]) return -1;
    TH2D h(names[t], "", nm, mmin, mmax, ny, ymin, ymax);
    for (int im = 0; im < nm; im++)
      for (int iy = 0; iy < ny; iy++) h.SetBinContent(im + 1, iy + 1, tabs[t][(size_t)im * ny + iy]);
    h.Write();
  }
  return 0;
}

// new UpcGenerator + configGeneratorFromFile + init, exactly as main.cpp does it (on a thread with a large stack:
// prepareBreakupProb keeps 16 MB of arrays on it)
int upcrefgen_create(const char* parfile, const char* lumi_dir)
{
  if (g_gen) { delete g_gen; g_gen = nullptr; gRandom = new TRandomMT64(); }  // ~UpcGenerator deletes gRandom
  GenInit gi{parfile, lumi_dir};
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, (size_t)256 << 20);
  pthread_t th;
  if (pthread_create(&th, &at, gen_init_thread, &gi)) return -1;
  pthread_join(th, nullptr);
  return g_gen ? 0 : -1;
}

double upcrefgen_totcs() { return g_gen->totNuclX(); }
void upcrefgen_grid(int* nm, int* ny, int* nz)
{
  *nm = g_gen->nucProcessCS->nm; *ny = g_gen->nucProcessCS->ny; *nz = g_gen->nucProcessCS->nz;
}
// the generator's folded table nucCSYM [ny][nm] and, when polarised, polCSRatio
void upcrefgen_cs(double* cs, double* ratio)
{
  const int ny = g_gen->nucProcessCS->ny, nm = g_gen->nucProcessCS->nm;
  for (int iy = 0; iy < ny; iy++)
    for (int im = 0; im < nm; im++) {
      cs[(size_t)iy * nm + im] = g_gen->nucCSYM[iy][im];
      if (ratio && !g_gen->polCSRatio.empty()) ratio[(size_t)iy * nm + im] = g_gen->polCSRatio[iy][im];
    }
}
// cumulative tables of the reference's samplers: hpdf->sum of samplerCsYM [ny*nm+1] and of samplersCsZ[im] [nz+1]
void upcrefgen_cdfs(double* sum2d, double* sumz)
{
  const int ny = g_gen->nucProcessCS->ny, nm = g_gen->nucProcessCS->nm, nz = g_gen->nucProcessCS->nz;
  if (sum2d) std::memcpy(sum2d, g_gen->samplerCsYM->hpdf->sum, ((size_t)ny * nm + 1) * sizeof(double));
  if (sumz)
    for (size_t im = 0; im < g_gen->samplersCsZ.size(); im++)
      std::memcpy(sumz + im * (nz + 1), g_gen->samplersCsZ[im]->hpdf->sum, (nz + 1) * sizeof(double));
}

// n calls of UpcGenerator::generateEvent (src/UpcGenerator.cpp:715-832).  Per candidate: npart (0 = rejected by the
// cuts), pdg/status/mother[6], p4[6][4] = (px, py, pz, E) (six slots: a pi0 pair and its four photons).  Returns the number
// of accepted events.
long upcrefgen_generate(long n, int* npart, int* pdg, int* status, int* mother, double* p4)
{
  std::vector<int> pdgs, statuses, mothers;
  std::vector<TLorentzVector> particles;
  long acc = 0;
  for (long i = 0; i < n; i++) {
    const long ok = g_gen->generateEvent(pdgs, statuses, mothers, particles);
    const int np = ok == 1 ? (int)particles.size() : 0;
    npart[i] = np;
    for (int j = 0; j < 6; j++) {
      const bool v = j < np;
      pdg[i * 6 + j] = v ? pdgs[j] : 0;
      status[i * 6 + j] = v ? statuses[j] : 0;
      mother[i * 6 + j] = v ? mothers[j] : 0;
      double* q = p4 + ((size_t)i * 6 + j) * 4;
      q[0] = v ? particles[j].Px() : 0; q[1] = v ? particles[j].Py() : 0;
      q[2] = v ? particles[j].Pz() : 0; q[3] = v ? particles[j].E() : 0;
    }
    acc += ok == 1;
  }
  return acc;
}

// The reference seeds its 2-D sampler and all nm 1-D samplers with the SAME seed (src/UpcGenerator.cpp:688-700, Q6):
// the k-th draw of EVERY mass bin's z sampler uses the same uniform, so the pooled cos(theta) sample of a run is not
// i.i.d. (with 1001 bins, the first draw of each bin puts 1001 events on one quantile).  For two-sample tests the
// streams are decorrelated by re-seeding sampler i with seed0 + 1 + i; the code that draws stays the reference's.
void upcrefgen_reseed_z(unsigned long seed0)
{
  for (size_t i = 0; i < g_gen->samplersCsZ.size(); i++) gsl_rng_set(g_gen->samplersCsZ[i]->rng, seed0 + 1 + i);
  for (size_t i = 0; i < g_gen->samplersCsSZ.size(); i++) gsl_rng_set(g_gen->samplersCsSZ[i]->rng, seed0 + 100001 + i);
  for (size_t i = 0; i < g_gen->samplersCsPsZ.size(); i++) gsl_rng_set(g_gen->samplersCsPsZ[i]->rng, seed0 + 200001 + i);
}

// tape of the uniforms drawn by the reference's code (shim_tape.h): switch on / off (clears), read back
void upcrefgen_tape(int on) { shim_tape().on = on != 0; shim_tape().v.clear(); shim_tape().tag.clear(); }
long upcrefgen_tape_read(double* v, int* tag, long cap)
{
  const long n = (long)shim_tape().v.size();
  for (long i = 0; i < n && i < cap; i++) { if (v) v[i] = shim_tape().v[i]; if (tag) tag[i] = shim_tape().tag[i]; }
  return n;
}

// single draws of the reference's samplers / photon-pT generator (their own MT19937 / gRandom streams)
void upcrefgen_sample_ym(long n, double* y, double* m, int* ybin, int* mbin)
{
  for (long i = 0; i < n; i++) {
    (*g_gen->samplerCsYM)(y[i], m[i]);
    ybin[i] = g_gen->samplerCsYM->getBinX(y[i]);
    mbin[i] = g_gen->samplerCsYM->getBinY(m[i]);
  }
}
void upcrefgen_sample_z(int mbin, long n, double* z)
{
  for (long i = 0; i < n; i++) z[i] = (*g_gen->samplersCsZ[mbin])();
}
void upcrefgen_photon_pt(double e_phot, long n, double* pt)
{
  for (long i = 0; i < n; i++) pt[i] = g_gen->nucProcessCS->getPhotonPt(e_phot);
}

// UpcGenerator::generateEvents with the HepMC writer (writes events.hepmc into the current directory): the byte
// format our writer must reproduce
void upcrefgen_generate_events_hepmc(long n_events)
{
  g_gen->nEvents = n_events;
  g_gen->useROOTOut = false;
  g_gen->useHepMCOut = true;
  g_gen->generateEvents();
}
// ... and with the ROOT tree: returns the rows the reference filled into `particles` (9 branches, :842-857);
// out [n_rows][9] in branch order; n_rows via the return value (call with out = NULL to size)
long upcrefgen_generate_events_tree(long n_events, double* out, long cap_rows)
{
  g_gen->nEvents = n_events;
  g_gen->useROOTOut = true;
  g_gen->useHepMCOut = false;
  g_gen->generateEvents();
  TTree* t = g_gen->mOutTree;
  const long rows = t->cols.empty() ? 0 : (long)t->cols[0].size();
  if (out)
    for (long r = 0; r < rows && r < cap_rows; r++)
      for (int b = 0; b < 9; b++) out[r * 9 + b] = t->cols[b][r];
  return rows;
}
}
