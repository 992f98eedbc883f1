// C entry points over the REFERENCE's own UpcCrossSection (compiled unmodified from
// /root/reference/src against the GSL/ROOT shim), so that tests can run the reference's code.
#include <pthread.h>

#include <cstring>
#include <stdexcept>

#include "UpcCrossSection.h"

TRandom* gRandom = new TRandomMT64();
TSystem* gSystem = new TSystem();

// (every elementary process is the reference's own translation unit: dileptons, ALP, vector mesons, and the two that
// read ROOT histogram files -- light-by-light, pi0 pi0 --, whose files the tests inject: upcref_put_hist1/2)
extern gsl_spline* gslSplineGAA;
extern gsl_spline* gslSplineFormFac;
extern gsl_spline* gslSplineBreakP;

struct RefParams {
  int Z, A; double R, a, sqrts;
  int is_point, breakup_mode, use_pol, nm, ny; double mmin, mmax, ymin, ymax;
  int proc_id; double a_lep, alp_mass, alp_width;
  int with_breakup_table;  // run prepareBreakupProb (1e6 knots, ~1 min, 16 MB of stack)
};

static UpcCrossSection* g_cs = nullptr;

static void* init_thread(void* vp)
{
  const RefParams* p = (const RefParams*)vp;
  // UpcGenerator::UpcGenerator + setParameterValue, src/UpcGenerator.cpp:30-40, 183-305
  g_cs = new UpcCrossSection();   // gtot fixed here from the defaults (Q1)
  UpcCrossSection::sqrts = p->sqrts;
  UpcCrossSection::g1 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
  UpcCrossSection::g2 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
  UpcCrossSection::Z = p->Z; UpcCrossSection::A = p->A; UpcCrossSection::R = p->R; UpcCrossSection::a = p->a;
  g_cs->isPoint = p->is_point; g_cs->breakupMode = p->breakup_mode; g_cs->usePolarizedCS = p->use_pol;
  g_cs->nm = p->nm; g_cs->ny = p->ny; g_cs->mmin = p->mmin; g_cs->mmax = p->mmax; g_cs->ymin = p->ymin; g_cs->ymax = p->ymax;
  g_cs->alpMass = p->alp_mass; g_cs->alpWidth = p->alp_width;
  g_cs->setElemProcess(p->proc_id);
  if (p->proc_id >= 11 && p->proc_id <= 15) ((UpcTwoPhotonDilep*)g_cs->elemProcess)->aLep = p->a_lep;
  // UpcCrossSection::init without prepareTwoPhotonLumi, src/UpcCrossSection.cpp:116-131
  g_cs->factor = UpcCrossSection::Z * UpcCrossSection::Z * phys_consts::alpha / M_PI / M_PI / phys_consts::hc / phys_consts::hc;
  UpcCrossSection::mNucl = (UpcCrossSection::Z * phys_consts::mProt + (UpcCrossSection::A - UpcCrossSection::Z) * phys_consts::mNeut) / UpcCrossSection::A;
  UpcCrossSection::rho0 = g_cs->calcWSRho();
  g_cs->prepareGAA();
  g_cs->prepareFormFac();
  if (p->breakup_mode > 1 && p->with_breakup_table) g_cs->prepareBreakupProb();
  return nullptr;
}

extern "C" {

int upcref_init(const RefParams* p)
{
  // prepareBreakupProb keeps 16 MB on the stack (src/UpcCrossSection.cpp:422-423): run with a big one
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, (size_t)256 << 20);
  pthread_t th;
  if (pthread_create(&th, &at, init_thread, (void*)p)) return -1;
  pthread_join(th, nullptr);
  return 0;
}
double upcref_rho0() { return UpcCrossSection::rho0; }
double upcref_gtot() { return g_cs->gtot; }
double upcref_factor() { return g_cs->factor; }
void upcref_gaa(double* y, double* c) { std::memcpy(y, gslSplineGAA->y.data(), 200 * 8); std::memcpy(c, gslSplineGAA->c.data(), 200 * 8); }
double upcref_formfac(double q2) { return UpcCrossSection::calcFormFac(q2); }
double upcref_formfac_knot(int i) { return gslSplineFormFac->y[i]; }
double upcref_flux_point(double b, double k) { return g_cs->fluxPoint(b, k); }
double upcref_flux_form(double b, double k) { return g_cs->fluxForm(b, k); }
double upcref_breakup_raw(double b, int mode) { return g_cs->calcBreakupProb(b, mode); }
double upcref_breakup_spline(double b) { return gsl_spline_eval(gslSplineBreakP, b, nullptr); }
double upcref_breakup_knot(int i, double* c) { if (c) *c = gslSplineBreakP->c[i]; return gslSplineBreakP->y[i]; }
double upcref_lumi(double M, double Y) { return g_cs->calcTwoPhotonLumi(M, Y); }
double upcref_photon_flux(double M, double Y) { return g_cs->calcPhotonFlux(M, Y); }
void upcref_lumi_pol(double M, double Y, double* s, double* p) { g_cs->calcTwoPhotonLumiPol(*s, *p, M, Y); }
double upcref_sigma_m(double m) { return g_cs->elemProcess->calcCrossSectionM(m); }
double upcref_sigma_zm(double z, double m) { return g_cs->elemProcess->calcCrossSectionZM(z, m); }
double upcref_sigma_m_pol(double m, int ps) { return ps ? g_cs->elemProcess->calcCrossSectionMPolPS(m) : g_cs->elemProcess->calcCrossSectionMPolS(m); }
double upcref_sigma_zm_pol(double z, double m, int ps) { return ps ? g_cs->elemProcess->calcCrossSectionZMPolPS(z, m) : g_cs->elemProcess->calcCrossSectionZMPolS(z, m); }
// Histograms of the "files" the light-by-light and pi0 pi0 plug-ins open (src/UpcTwoPhotonLbyL.cpp:40-47,
// src/UpcTwoPhotonDipion.cpp:44-51): placed into the shim's in-memory TFile store under the path the plug-in will ask
// for (CROSS_SEC_DIR/<process>/cross_section_[z]m.root).  cells: all (nx + 2) [x (ny + 2)] bins, x fastest.
const char* upcref_cross_sec_dir() { return CROSS_SEC_DIR; }
void upcref_put_hist1(const char* path, const char* name, int nx, double xlo, double xhi, const double* cells)
{
  auto* h = new TH1D(name, "", nx, xlo, xhi);
  for (int i = 0; i < nx + 2; i++) h->SetBinContent(i, cells[i]);
  TFile::store1()[path][name] = h;
}
void upcref_put_hist2(const char* path, const char* name, int nx, double xlo, double xhi, int ny, double ylo, double yhi,
                      const double* cells)
{
  auto* h = new TH2D(name, "", nx, xlo, xhi, ny, ylo, yhi);
  for (int iy = 0; iy < ny + 2; iy++)
    for (int ix = 0; ix < nx + 2; ix++) h->SetBinContent(ix, iy, cells[(size_t)iy * (nx + 2) + ix]);
  TFile::store()[path][name] = h;
}
// the reference's vector-meson plug-in (src/UpcPhotoNuclearVM.cpp, compiled unmodified): sigma(y) of
// calcCrossSectionY for n rapidities; needs upcref_init (calcFormFac and the statics sqrts, mNucl, R, a, rho0).
// getRgLtaVG keeps its table in function-local statics: one (PDG, SHADOWING 4) combination per process.
int upcref_vm_sigma_y(int pdg, int shadowing, int dght_pdg, const double* y, int n, double* out, double* m_part)
{
  if (!g_cs) return -1;
  try {
    UpcPhotoNuclearVM vm(pdg, shadowing, dght_pdg);
    for (int i = 0; i < n; ++i) out[i] = vm.calcCrossSectionY(y[i]);
    if (m_part) *m_part = vm.mPart;
  } catch (const std::exception&) {
    return -2;
  }
  return 0;
}
// the reference's grid driver + fold on its own small grid: prepareTwoPhotonLumi (OpenMP region, TH2D,
// "file") followed by calcNucCrossSectionYM; cs [ny][nm]
double upcref_grid_and_fold(int nthreads, const char* dir, double* cs, double* ratio)
{
  g_cs->numThreads = nthreads;
  g_cs->setLumiFileDirectory(dir);
  g_cs->prepareTwoPhotonLumi();
  std::vector<std::vector<double>> csYM(g_cs->ny, std::vector<double>(g_cs->nm, 0.)), r(g_cs->ny, std::vector<double>(g_cs->nm, 0.));
  double tot = 0;
  g_cs->calcNucCrossSectionYM(csYM, r, tot);
  for (int iy = 0; iy < g_cs->ny; iy++)
    for (int im = 0; im < g_cs->nm; im++) {
      cs[(size_t)iy * g_cs->nm + im] = csYM[iy][im];
      if (ratio) ratio[(size_t)iy * g_cs->nm + im] = r[iy][im];
    }
  return tot;
}
}
