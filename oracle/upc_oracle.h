/*
 * upc_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (nburmaso/upcgen) algorithm for the hot path:
 * lookup tables, photon fluxes, two-photon luminosity quadrature, sigma fold, inverse-CDF
 * samplers and event kinematics.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (upcgen_b200/) never includes, links or calls it.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path
 *  and cannot be built here as-is (needs ROOT + GSL).  The oracle is
 * pinned in three ways (tests/test_oracle_*.py):
 *   (1) its restatements of the third-party numerics are checked against independent
 *       implementations present in this image: QUADPACK dqagse via scipy.integrate.quad
 *       (result, neval, last), scipy CubicSpline(natural), scipy.special k0/k1/j1, mpmath;
 *   (2) the reference's OWN translation units (src/UpcCrossSection.cpp etc.) are compiled
 *       unmodified against a small shim of the GSL/ROOT entry points they use
 *       (oracle/refshim -> oracle/_ref/libupcref.so) and run against this oracle;
 *   (3) physics identities (form-factor flux -> point flux for b >> R, F(0)=A, ...).
 * Third-party arithmetic (GSL/ROOT) is restated, not linked: see DESIGN.md "Oracle".
 */
#ifndef UPC_ORACLE_H
#define UPC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors UpcCrossSection.h:73-164 + the UpcGenerator members used on the path */
typedef struct {
  int Z, A;
  double R, a;          /* Woods-Saxon */
  double sqrts;         /* GeV */
  double g1, g2;        /* beam Lorentz factors sqrts/(2 mProt) */
  double gtot;          /* Q1: fixed in the UpcCrossSection constructor from the DEFAULT sqrts */
  int is_point;         /* FLUX_POINT */
  int breakup_mode;     /* 1 none, 2 XNXN, 3 0N0N, 4 0NXN */
  int use_pol;          /* USE_POLARIZED_CS */
  int nonzero_gam_pt;   /* NON_ZERO_GAM_PT */
  int nm, ny, nz;
  double mmin, mmax, ymin, ymax, zmin, zmax;
  int nb1, nb2;         /* 120, 120 */
  /* elementary process (UpcGenerator::init) */
  int proc_id;          /* 11, 13, 15 dilepton; 51 ALP */
  double a_lep;         /* LEP_A */
  double alp_mass, alp_width;
  /* kinematic cuts */
  int do_pt_cut, do_eta_cut;
  double pt_min, eta_min, eta_max;
} upco_params;

typedef struct upco_ctx upco_ctx;

/* number of breakup-spline knots actually tabulated (reference: 1e6; only b<=20 is ever read) */
#define UPCO_NBC_DEFAULT 21001

upco_ctx* upco_create(const upco_params* p, int nbc_used);
void upco_destroy(upco_ctx*);
void upco_set_threads(upco_ctx*, int nthreads);

/* --- special functions (test hooks) --- */
double upco_bessel_K0(double x);
double upco_bessel_K1(double x);
double upco_bessel_J1(double x);
double upco_tmath_besselK1(double x);
double upco_tmath_besselI1(double x);

/* --- generic numerics (test hooks) --- */
void upco_cspline_init(const double* x, const double* y, int n, double* c);
double upco_cspline_eval(const double* x, const double* y, const double* c, int n, double xv);
/* QAGS with GK21 on f(x) = x^2 F(t)/t J1(b x/hc), the fluxForm integrand; returns result and
   fills neval/last/abserr/ier */
double upco_qags_fluxform(upco_ctx*, double b, double k, double* abserr, int* neval, int* last, int* ier);
/* analysis aid: the bisected intervals of that integral, in order, as (level << 24 | position) */
int upco_qags_fluxform_trace(upco_ctx*, double b, double k, unsigned* trace, int cap, int* neval);
/* generic QAGS (GK21) on a C callback; returns the GSL error code */
int upco_qags(double (*f)(double, void*), void* par, double a, double b, double epsabs, double epsrel, size_t limit,
              double* result, double* abserr);
/* QAGS on a small family of analytic test integrands (kind 0..5), for pinning against scipy */
double upco_qags_test(int kind, double alpha, double a, double b, double epsabs, double epsrel,
                      double* abserr, int* neval, int* last, int* ier);

/* --- tables (T1-T4) --- */
double upco_rho0(upco_ctx*);
double upco_sigma_nn(upco_ctx*);
void upco_get_gaa(upco_ctx*, double* b, double* gaa, double* c, double* ta);  /* 200 each */
double upco_formfac(upco_ctx*, double Q2);                               /* analytic calcFormFac */
double upco_formfac_spline(upco_ctx*, double Q2);                        /* 1e6-knot spline eval */
void upco_get_formfac_table(upco_ctx*, int i0, int n, double* y, double* c);
double upco_breakup_raw(upco_ctx*, double b, int mode);                  /* calcBreakupProb */
double upco_breakup_spline(upco_ctx*, double b);                         /* spline eval, b<=20 */
void upco_get_breakup_table(upco_ctx*, int i0, int n, double* y, double* c);
int upco_breakup_nknots_energy(upco_ctx*);

/* --- fluxes (F1-F3) --- */
double upco_flux_point(upco_ctx*, double b, double k);
double upco_flux_form(upco_ctx*, double b, double k);
void upco_flux_form_batch(upco_ctx*, const double* b, const double* k, size_t n, double* out,
                          int* neval);

/* --- luminosity (L1-L3) --- */
double upco_lumi(upco_ctx*, double M, double Y);
void upco_lumi_pol(upco_ctx*, double M, double Y, double* ns, double* np);
/* fills cells (im, iy) for im in [im0, im1) step im_step, iy in [0,ny) step iy_step; others
   untouched.  lumi[im*ny+iy] (x dm dy).  Polarised: lumi_s/lumi_p.  neval (may be NULL)
   receives the QAGS evaluation count per cell.  OpenMP static m-slabs as the reference. */
void upco_fill_lumi(upco_ctx*, int im0, int im1, int im_step, int iy_step, double* lumi,
                    double* lumi_s, double* lumi_p, long long* neval);

/* --- elementary cross sections (P1) --- */
double upco_sigma_m(upco_ctx*, double m);
double upco_sigma_zm(upco_ctx*, double z, double m);
double upco_sigma_m_pol(upco_ctx*, double m, int ps);
double upco_sigma_zm_pol(upco_ctx*, double z, double m, int ps);

/* --- fold (X1, X2) --- */
void upco_fold(upco_ctx*, const double* lumi, const double* lumi_s, const double* lumi_p,
               double* cs /*[ny][nm]*/, double* ratio, double* totcs_mb);
void upco_fill_cs_zm(upco_ctx*, int flag, double* cszm /*[nm][nz]*/);

/* --- samplers (S1-S3) --- */
void upco_pdf_init(const double* bin, size_t n, double* sum /*n+1*/);
/* returns k (or -1 when r lies outside [sum[0], sum[n])) */
long long upco_pdf_find(const double* sum, size_t n, double r);
void upco_sample2d(const double* sum, int nx, int ny_, const double* xe, const double* ye,
                   double r1, double r2, long long* k, double* x, double* y);
double upco_sample1d(const double* sum, int n, const double* edges, double r);
int upco_get_bin(int nbins, double x, double lo, double hi);

/* --- photon pT (E3) --- */
void upco_photon_pt_cdf(upco_ctx*, double ePhot, double* cdf /*5001*/);
double upco_photon_pt_sample(upco_ctx*, const double* cdf, double r);

/* --- Philox4x32-10 (the product's counter-based generator, restated for event parity) --- */
void upco_philox(uint64_t seed, uint64_t ctr, uint32_t block, double* u0, double* u1);

/* --- event generation with the product's uniform slot map (E1-E5) ---
   cs_sum: 2-D CDF (ny*nm+1); z_sum: nm CDFs of nz+1 (or pol: s then ps); ratio [ny][nm].
   Output arrays sized 4 particles/event (ALP: 3 used; pairs: 2).  returns 1 accepted/0. */
double upco_photon_flux(upco_ctx* c, double M, double Y);
int upco_generate_event_u(upco_ctx* c, const double* u, const double* cs_sum, const double* z_sum,
                          const double* z_sum_ps, const double* ratio, int* npart, int* pdg, int* status, int* mother,
                          double* p4, double* aux);
int upco_generate_event(upco_ctx*, uint64_t seed, uint64_t candidate, const double* cs_sum,
                        const double* z_sum, const double* z_sum_ps, const double* ratio,
                        int* npart, int* pdg, int* status, int* mother, double* p4 /*[4][4]*/,
                        double* aux /*y,m,z,pt1,pt2*/);

#ifdef __cplusplus
}
#endif
#endif
