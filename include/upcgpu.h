/*
 * upcgpu.h -- C-ABI of the B200-native (sm_100a, FP64) two-photon-luminosity / sigma-table /
 * event-sampling path.  This is the drop-in boundary: plain C, opaque handle, int status,
 * caller-owned host buffers, no exceptions, no torch/CUDA types in any signature.
 *
 * The reference (nburmaso/upcgen) has no FFI; its boundary is the public C++ surface of
 * UpcCrossSection / UpcSampler / UpcGenerator.  Each entry point below names
 * the reference interface it replaces (paths relative to the reference root).  The C++ facade
 * with the reference's own class/method names lives in upcgen_b200/host/ and calls only this
 * header; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns UPCGPU_OK (0) or a negative UPCGPU_E* code; the message of the
 *     last failure on a context is available from upcgpu_last_error().
 *   - one context = one CUDA device (upcgpu_create) or several devices behind one handle (upcgpu_create_multi);
 *     a context is thread-compatible, not thread-safe.
 *   - there is NO CPU fallback: without a usable CUDA device upcgpu_create fails.
 *   - tables are FP64.  lumi tables are [nm][ny] (im-major, as TH2D hD2LDMDY x=M, y=Y);
 *     sigma tables are [ny][nm] (transposed, as the reference's crossSectionYM).
 */
#ifndef UPCGPU_H
#define UPCGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPCGPU_OK 0
#define UPCGPU_EINVAL (-1)   /* bad argument / state */
#define UPCGPU_ECUDA (-2)    /* CUDA runtime error (see upcgpu_last_error) */
#define UPCGPU_ENODEV (-3)   /* no CUDA device */
#define UPCGPU_EQAGS (-4)    /* device QAGS hit an error state the reference would abort on */
#define UPCGPU_ERANGE (-5)   /* uniform outside the cumulative pdf (GSL: "cannot find r1") */

typedef struct upcgpu_ctx upcgpu_ctx;

/* Parameter block.  Mirrors the public data members of UpcCrossSection
 * (include/UpcCrossSection.h:73-85 statics Z,A,R,a,sqrts,g1,g2; :114-164 instance fields) plus
 * the UpcGenerator members used by the event stage (include/UpcGenerator.h:63-76).
 * gtot is passed explicitly because the reference fixes it in the UpcCrossSection constructor
 * from the default sqrts (src/UpcCrossSection.cpp:51-54). */
typedef struct {
  int Z, A;
  double R, a;
  double sqrts, g1, g2, gtot;
  int is_point;        /* FLUX_POINT        -> UpcCrossSection::isPoint */
  int breakup_mode;    /* BREAKUP_MODE      -> UpcCrossSection::breakupMode (1..4) */
  int use_pol;         /* USE_POLARIZED_CS  -> UpcCrossSection::usePolarizedCS */
  int nonzero_gam_pt;  /* NON_ZERO_GAM_PT   -> UpcCrossSection::useNonzeroGamPt */
  int nm, ny, nz;
  double mmin, mmax, ymin, ymax, zmin, zmax;
  int nb1, nb2;        /* UpcCrossSection::nb1, nb2 (120) */
  /* event stage */
  int part_pdg;        /* UpcElemProcess::partPDG */
  double m_part;       /* UpcElemProcess::mPart */
  int is_charged;      /* UpcElemProcess::isCharged */
  int is_pair;         /* UpcGenerator::isPairProduction */
  int is_single;       /* UpcGenerator::isSingleProduction */
  int ignore_csz;      /* UpcGenerator::ignoreCSZ */
  int decay_uniform_pdg; /* !=0: twoPartDecayUniform into two of this pdg (22) of the single particle (ALP, :806-808) or of
                          * BOTH particles of a pair (pi0 pi0, src/UpcGenerator.cpp:799-803) */
  int do_pt_cut, do_eta_cut;
  double pt_min, eta_min, eta_max;
} upcgpu_params;

enum { UPCGPU_TABLE_GAA = 0, UPCGPU_TABLE_TA = 1, UPCGPU_TABLE_FORMFAC = 2, UPCGPU_TABLE_BREAKUP = 3 };

/* Scalars computed while preparing the tables. */
typedef struct {
  double rho0;       /* UpcCrossSection::calcWSRho     src/UpcCrossSection.cpp:152-163 */
  double sigma_nn;   /* csNN in prepareGAA              src/UpcCrossSection.cpp:368-369 */
  double factor;     /* UpcCrossSection::factor         src/UpcCrossSection.cpp:120 */
  double breakup_p20;/* P(b=20): the clamp value of     src/UpcCrossSection.cpp:260 */
  int n_breakup_energy_knots;
  double gaa_zero_below; /* b [fm] below which the G_AA spline is <= 1e-20 in magnitude: such (b1,b2,phi)
                          * points are not evaluated by the cell quadrature (they add < 1e-19 relative) */
} upcgpu_table_info;

/* Work counters of the last lumi fill (for the roofline arithmetic; not needed by callers). */
typedef struct {
  long long qags_integrals;   /* number of form-factor flux integrals evaluated */
  long long qags_evals;       /* integrand evaluations inside them (21 per GK21 call) */
  long long qags_overflow;    /* integrals that needed the large-workspace pass */
  long long qags_errors;      /* integrals ending in a GSL error state (reference: abort) */
  long long flux_rows;        /* photon-energy rows tabulated */
  long long band_pairs;       /* (b1,b2) pairs with some b < 20 fm that were evaluated */
  double ms_tables, ms_flux, ms_cells, ms_total; /* device time of the stages, CUDA events */
  double ms_qags;             /* of ms_flux: the QAGS kernels (head + row-cooperative) */
  double ms_qags_head;        /* of ms_qags: k_flux_qags_head alone (one thread per integral on tabulated intervals) */
  long long qags_head_evals;  /* integrand evaluations made by k_flux_qags_head */
  long long qags_head_done;   /* integrals that converged inside the head */
  long long qags_table_evals; /* of qags_head_evals: those whose J1 factor came from the common-grid table */
  long long cells_evaluated;  /* cells the quadrature kernel evaluated (the others are mirror images, Y -> -Y) */
} upcgpu_fill_stats;

/* ---- lifetime ------------------------------------------------------------------------ */
/* replaces: new UpcCrossSection() + parameter assignment (src/UpcGenerator.cpp:30-40,183-305) */
int upcgpu_create(const upcgpu_params* params, int device, upcgpu_ctx** out);
/* The same on n_gpus devices of this node, driven from this one process (devices: n_gpus ordinals, or NULL for
 * 0 .. n_gpus-1).  Replaces the OpenMP team of the reference's grid driver (-nthreads; static m-slabs,
 * src/UpcCrossSection.cpp:536-539): behind the returned handle stands one member context per device with its own host
 * thread; an NCCL communicator over the devices (ncclCommInitAll; NCCL is loaded with dlopen) and peer access
 * between them are set up here.  With such a handle
 *   upcgpu_prepare_tables   builds the lookup tables on every device,
 *   upcgpu_fill_lumi        deals the m rows to the devices (blocks of 32 rows, in a snake), exchanges the shards
 *                           (stores of every finished cell into all devices' tables from inside the cell kernel where
 *                           the devices have peer access -- the default --, or an NCCL all-gather over NVLink:
 *                           upcgpu_group_set_exchange(ctx, 0)) and leaves the FULL table on every device; the host
 *                           buffers are filled from device 0,
 *   upcgpu_fold_sigma, upcgpu_sampler_build   run on every device (each samples events from its own copy),
 *   upcgpu_generate[_device]   split the candidate range into contiguous ranges, one per device; Philox counters make
 *                           the result independent of n_gpus,
 * and every other entry point acts on device 0.  n_gpus = 1 is upcgpu_create. */
int upcgpu_create_multi(const upcgpu_params* params, int n_gpus, const int* devices, upcgpu_ctx** out);
/* number of devices behind a handle; the member context of `rank` (borrowed: do not destroy), e.g. to read a
 * member's copy of a table; the exchange in use (0 NCCL all-gather, 1 peer stores) and a one-line description */
int upcgpu_group_size(const upcgpu_ctx* ctx);
int upcgpu_group_member(upcgpu_ctx* ctx, int rank, upcgpu_ctx** member);
int upcgpu_group_set_exchange(upcgpu_ctx* ctx, int mode);
int upcgpu_group_describe(const upcgpu_ctx* ctx, char* buf, size_t cap);
void upcgpu_destroy(upcgpu_ctx* ctx);
/* message of the last failure (ctx may be NULL for create failures) */
const char* upcgpu_last_error(const upcgpu_ctx* ctx);
/* library/ABI version, and the device the context runs on */
int upcgpu_abi_version(void);
int upcgpu_device_name(const upcgpu_ctx* ctx, char* buf, size_t cap);
/* the CUDA stream all work of this context is issued on, as an integer (cudaStream_t), so that a
 * caller can record its own timing events on it; and the number of kernels of this library
 * launched on it so far */
int upcgpu_stream_handle(const upcgpu_ctx* ctx, uint64_t* stream);
long long upcgpu_launch_count(const upcgpu_ctx* ctx);

/* ---- tables T1-T4 -------------------------------------------------------------------- */
/* replaces UpcCrossSection::init's calcWSRho/prepareGAA/prepareFormFac/prepareBreakupProb
 * (src/UpcCrossSection.cpp:116-131, :152-163, :364-461).  All on the device. */
int upcgpu_prepare_tables(upcgpu_ctx* ctx);
int upcgpu_get_table_info(const upcgpu_ctx* ctx, upcgpu_table_info* info);
/* marks the tables stale so that the next upcgpu_prepare_tables recomputes them (bench: the
 * table stage is part of every timed step) */
int upcgpu_invalidate_tables(upcgpu_ctx* ctx);
/* copies knots i0..i0+n-1 of a spline table: x (knot), y (value), c (GSL c-coefficient);
 * any of x,y,c may be NULL.  Replaces reading gslSplineGAA/FormFac/BreakP (:31-38). */
int upcgpu_get_table(upcgpu_ctx* ctx, int which, size_t i0, size_t n, double* x, double* y, double* c);
/* spline evaluation on the device, test hook for gsl_spline_eval call sites (:188,:260,:262) */
int upcgpu_eval_table(upcgpu_ctx* ctx, int which, const double* x, size_t n, double* out);
/* un-splined breakup probability, test hook for UpcCrossSection::calcBreakupProb(b, mode)
 * (src/UpcCrossSection.cpp:752-1019); needs breakup_mode > 1 at create time */
int upcgpu_breakup_raw(upcgpu_ctx* ctx, const double* b, int mode, size_t n, double* out);

/* ---- fluxes F1-F3 -------------------------------------------------------------------- */
/* replaces UpcCrossSection::fluxPoint / fluxForm (src/UpcCrossSection.cpp:166-218) on n
 * (b,k) points; neval (may be NULL) receives the QAGS evaluation count (0 for point flux). */
int upcgpu_flux_point(upcgpu_ctx* ctx, const double* b, const double* k, size_t n, double* out);
int upcgpu_flux_form(upcgpu_ctx* ctx, const double* b, const double* k, size_t n, double* out, int* neval);

/* ---- luminosity L1-L3 ---------------------------------------------------------------- */
/* replaces UpcCrossSection::prepareTwoPhotonLumi (src/UpcCrossSection.cpp:463-592): fills the
 * whole grid, values already multiplied by dm*dy (:546-550).  Unpolarised: lumi[nm*ny];
 * polarised (use_pol): lumi_s, lumi_p.  Unused pointers may be NULL.  Host buffers. */
int upcgpu_fill_lumi(upcgpu_ctx* ctx, double* lumi, double* lumi_s, double* lumi_p);
/* Same, but QUEUES only the m-rows of one shard and returns without waiting for the device: the result is collected
 * by the next entry point that reads results on the host (upcgpu_get_fill_stats, upcgpu_lumi_download,
 * upcgpu_fold_sigma ...), which is also where a QAGS error state of the fill is reported.
 * Computes only the m-rows of one shard -- blocks of B consecutive rows dealt round-robin: row im belongs
 * to shard (im / B) % nshards, B = 32 or, on small grids, the largest power of two with nm >= 2 B nshards -- keeps the result
 * on the device (see upcgpu_lumi_device) and copies nothing.  One rank per GPU calls this with
 * its rank; the exchange between ranks is the caller's (NCCL all-gather of the packed shard). */
int upcgpu_fill_lumi_shard(upcgpu_ctx* ctx, int shard, int nshards);
/* single cells, test hook for calcTwoPhotonLumi / calcTwoPhotonLumiPol (:221-335): n pairs
 * (M,Y) -> out (unpol) or out_s/out_p, NOT multiplied by dm*dy */
int upcgpu_lumi_cells(upcgpu_ctx* ctx, const double* M, const double* Y, size_t n, double* out,
                      double* out_s, double* out_p);
int upcgpu_get_fill_stats(upcgpu_ctx* ctx, upcgpu_fill_stats* st);
/* replaces UpcCrossSection::calcPhotonFlux (src/UpcCrossSection.cpp:700-722), the b-integrated photon flux of the
 * vector-meson path: flux_pos[i] = calcPhotonFlux(M[i], +Y[i]), flux_neg[i] = calcPhotonFlux(M[i], -Y[i]) -- the two
 * values calcNucCrossSectionY (:724-748) needs per rapidity.  Either output may be NULL. */
int upcgpu_photon_flux(upcgpu_ctx* ctx, const double* M, const double* Y, size_t n, double* flux_pos, double* flux_neg);

/* Device-resident buffers for the multi-GPU exchange (device pointers as integers so that no
 * CUDA type appears here).  packed shard: [rows_per_shard][ny] per table, the shard's rows in ascending
 * im (rows_per_shard = the row count of shard 0, the largest; padding rows zero).  which: 0 unpol, 1 scalar, 2 pseudo. */
int upcgpu_lumi_shard_buffer(upcgpu_ctx* ctx, int which, uint64_t* dev_ptr, size_t* n_doubles);
/* gathered buffer [nshards][rows_per_shard][ny] to be filled by the caller's all-gather, then
 * un-permuted into the full [nm][ny] device table by upcgpu_lumi_unpack */
int upcgpu_lumi_gather_buffer(upcgpu_ctx* ctx, int which, int nshards, uint64_t* dev_ptr, size_t* n_doubles);
int upcgpu_lumi_unpack(upcgpu_ctx* ctx, int nshards);
/* The exchange folded into the cell kernel, between PROCESSES (one rank per GPU on one node): every rank exports the
 * CUDA IPC handles of its full tables (3 x 64 bytes: unpolarised, scalar, pseudoscalar; unused ones zero), the ranks
 * swap them with whatever they communicate with, every rank imports all of them (handles of rank d at
 * handles + d * 3 * 64), and upcgpu_fill_lumi_shard_peers fills this rank's m rows with a cell kernel that stores
 * every finished cell straight into the table of EVERY rank over NVLink -- no gather buffer, no un-permute.  Queued
 * like upcgpu_fill_lumi_shard; before the fold the caller orders the ranks behind it with any collective on
 * upcgpu_stream_handle (a one-element all-reduce).  Inside one process upcgpu_create_multi + exchange 1 does the same. */
int upcgpu_lumi_ipc_export(upcgpu_ctx* ctx, void* handles);
int upcgpu_lumi_ipc_import(upcgpu_ctx* ctx, int nshards, int rank, const void* handles);
int upcgpu_fill_lumi_shard_peers(upcgpu_ctx* ctx);
/* copy the full device table to / from the host (which as above) */
int upcgpu_lumi_download(upcgpu_ctx* ctx, int which, double* host);
int upcgpu_lumi_upload(upcgpu_ctx* ctx, int which, const double* host);

/* ---- sigma fold X1 ------------------------------------------------------------------- */
/* replaces UpcCrossSection::calcNucCrossSectionYM (src/UpcCrossSection.cpp:594-698) on the
 * device-resident lumi table.  sig_m[nm] = elemProcess->calcCrossSectionM(m_im) (unpol, nb) or
 * sig_s/sig_p[nm] = calcCrossSectionMPolS/PS (pol, fm^2), evaluated by the caller's plug-in.
 * cs[ny*nm] (nb), ratio[ny*nm] (pol only) and totcs_mb may be NULL (results stay on device). */
int upcgpu_fold_sigma(upcgpu_ctx* ctx, const double* sig_m, const double* sig_s, const double* sig_p,
                      double* cs, double* ratio, double* totcs_mb);

/* ---- elementary-process plug-ins P1 (host code) -------------------------------------- */
/* C access to the built-in plug-ins UpcTwoPhotonDilep (proc_id 11/13/15), UpcTwoPhotonALP (51) and the
 * file-based UpcTwoPhotonLbyL (22) / UpcTwoPhotonDipion (111), which read the reference's histograms from
 * $UPCGEN_CROSS_SEC_DIR/{lbyl,pi0pi0}/cross_section_{m,zm}.root (UPCGPU_EINVAL if they cannot be read)
 * (reference src/UpcTwoPhotonDilep.cpp:46-134, src/UpcTwoPhotonALP.cpp:28-33, src/UpcTwoPhotonLbyL.cpp:32-80), for callers that
 * cannot instantiate the C++ classes.  which: 0 calcCrossSectionM, 1 ...MPolS, 2 ...MPolPS. */
int upcgpu_elem_sigma_m(int proc_id, double a_lep, double alp_mass, double alp_width, int which, const double* m,
                        size_t n, double* out);
/* UpcCrossSection::fillCrossSectionZM (src/UpcCrossSection.cpp:337-362): out[nm][nz] =
 * dsigma/dz(z_iz, m_im) * (hc)^2 * 1e7 / dm; flag 0 unpolarised, 1 scalar, 2 pseudoscalar. */
int upcgpu_elem_fill_cs_zm(int proc_id, double a_lep, double alp_mass, double alp_width, int flag, double zmin,
                           double zmax, int nz, double mmin, double mmax, int nm, double* out);

/* Reads a TH1D / TH2D from a ROOT file without ROOT (upcgen_b200/host/UpcRootHist.cpp): the reference's elementary
 * cross-section histograms and its luminosity caches twoPhotonLumi[Pol].root (hD2LDMDY[_s,_p]: bin (im + 1, iy + 1)
 * holds table[im][iy], src/UpcCrossSection.cpp:503-506, :564-571).  cells: (nx + 2) x (ny + 2) doubles (x fastest,
 * under-/overflow cells included; ny = 0 and one row for a TH1D), or NULL to query the sizes only.  No GPU is
 * involved.  Returns UPCGPU_EINVAL if the file or object cannot be read or `cap` is too small. */
int upcgpu_root_hist_read(const char* path, const char* name, int* dim, int* nx, double* xlo, double* xhi, int* ny,
                          double* ylo, double* yhi, double* cells, size_t cap, size_t* n_cells);

/* a TH1D with a uniform axis (the form of cross_sections/<process>/cross_section_m.root, which UpcTwoPhotonLbyL /
 * UpcTwoPhotonDipion read): cells[nx + 2] with under- and overflow */
int upcgpu_root_write_th1d(const char* path, const char* name, int nx, double xlo, double xhi, const double* cells,
                           double entries);
/* The writer side (upcgen_b200/host/UpcRootFile.cpp), also without ROOT and without a GPU.
 * upcgpu_root_write_th2d: a file with n_hist TH2D objects of common uniform axes -- the luminosity cache
 * twoPhotonLumi[Pol].root as the reference writes it (src/UpcCrossSection.cpp:493-507, :578-585): names[i],
 * cells[i] = (nx + 2) x (ny + 2) doubles, x fastest, under-/overflow cells included.
 * upcgpu_root_write_tree: a file with one TTree of flat branches -- events.root with the tree "particles"
 * (src/UpcGenerator.cpp:842-857): n_cols columns of n_rows values, types[i] = 'I' (Int_t) or 'D' (Double_t); integer
 * columns are passed as doubles holding integral values. */
int upcgpu_root_write_th2d(const char* path, int n_hist, const char* const* names, int nx, double xlo, double xhi, int ny,
                           double ylo, double yhi, const double* const* cells);
int upcgpu_root_write_tree(const char* path, const char* tree, const char* title, int n_cols, const char* const* names,
                           const char* types, const double* const* columns, size_t n_rows);
/* upcgpu_root_set_compression: ROOT's compression setting (100 * algorithm + level) for the files written by the
 * entry points above and below.  0 (default): uncompressed.  1xx: zlib "ZL" records (101 is ROOT's default, what the
 * reference's luminosity cache is written with).  4xx: LZ4 "L4" records, what the reference asks for in events.root
 * (4 * 100 + 9, src/UpcGenerator.cpp:843).  Levels 1-9; other algorithms: UPCGPU_EINVAL.  Returns the previous
 * setting through *previous when that is not NULL. */
int upcgpu_root_set_compression(int setting, int* previous);
/* upcgpu_root_write_sigma_hists: what the reference adds to events.root at debug level > 0
 * (src/UpcGenerator.cpp:900-917) -- the nuclear cross section table as the TH2D "hNucCSYM" over the ny + 1 rapidity
 * and nm + 1 mass bin edges (variable-bin axes, bin (iy + 1, im + 1) = cs[iy][im]) and its projections
 * "hNucCSYM_py" (TH1D over m) and "hNucCSYM_px" (TH1D over y), with the entries and statistics TH2::ProjectionX/Y
 * leave in them.  cs: [ny][nm], the table upcgpu_fold_sigma returns. */
int upcgpu_root_write_sigma_hists(const char* path, int ny, const double* y_edges, int nm, const double* m_edges,
                                  const double* cs);

/* ---- samplers S1-S3 ------------------------------------------------------------------ */
/* replaces the UpcSampler2D / UpcSampler1D constructors (include/UpcSampler.h:40-59, :81-109,
 * i.e. gsl_histogram[2d]_pdf_init) built in UpcGenerator::computeNuclXsection
 * (src/UpcGenerator.cpp:684-700).  cs: [ny][nm] or NULL to use the folded table on the device;
 * cszm [nm][nz] (unpol) or cszm_s/cszm_ps (pol); NULL when ignore_csz. */
int upcgpu_sampler_build(upcgpu_ctx* ctx, const double* cs, const double* cszm, const double* cszm_s,
                         const double* cszm_ps);
/* diagnostics of the CDF build (S1): blocks of bins done by the speculate-and-verify recurrences, blocks that needed a
 * sequential run, verification rounds -- out6[0..2] for the running mean, out6[3..5] for the cumulative sum; summed over
 * this context's builds */
int upcgpu_sampler_spec_stats(upcgpu_ctx* ctx, unsigned long long* out6);
/* copy the cumulative tables back (test hook: hpdf->sum of the reference's samplers) */
int upcgpu_sampler_get_cdf(upcgpu_ctx* ctx, double* sum2d /*ny*nm+1*/, double* sumz /*nm*(nz+1)*/,
                           double* sumz_ps);
/* replaces UpcSampler2D::operator() + getBinX/getBinY (include/UpcSampler.h:118-133) with
 * INJECTED uniforms u[2*i], u[2*i+1]: bit-exact flat bin k, bins and (y, m). */
int upcgpu_sample_ym(upcgpu_ctx* ctx, const double* u, size_t n, long long* k, int* ybin, int* mbin,
                     double* y, double* m);
/* replaces UpcSampler1D::operator() (:68-71) for sampler index mbin[i] with injected u[i];
 * ps selects the pseudoscalar set */
int upcgpu_sample_z(upcgpu_ctx* ctx, const int* mbin, const double* u, size_t n, int ps, double* z);

/* Generic forms of S1/S2 for the UpcSampler1D / UpcSampler2D classes (include/UpcSampler.h),
 * which take an arbitrary histogram: gsl_histogram[2d]_pdf_init on bins[n] -> sum[n+1], and
 * gsl_histogram2d_pdf_sample / gsl_histogram_pdf_sample with injected uniforms (host buffers). */
int upcgpu_hist_pdf_init(upcgpu_ctx* ctx, const double* bins, size_t n, double* sum);
int upcgpu_hist_sample2d(upcgpu_ctx* ctx, const double* sum, int nx, int ny, const double* xedges,
                         const double* yedges, const double* u, size_t n, long long* k, double* x, double* y);
int upcgpu_hist_sample1d(upcgpu_ctx* ctx, const double* sum, int n, const double* edges, const double* u,
                         size_t nsamp, double* x);

/* ---- events E1-E5 -------------------------------------------------------------------- */
#define UPCGPU_MAX_PART 6
/* replaces the body of UpcGenerator::generateEvent (src/UpcGenerator.cpp:715-832) incl.
 * getPairMomentum/getPhotonPt (src/UpcCrossSection.cpp:1021-1074), pairProduction,
 * singleProduction, twoPartDecayUniform and checkKinCuts, for candidates
 * [first_candidate, first_candidate + n_candidates) of the Philox4x32-10 stream keyed by seed.
 * Output (host, SoA, sized n_candidates): npart[i] (0 = rejected by the cuts),
 * pdg/status/mother[i*UPCGPU_MAX_PART+j], p4[(i*UPCGPU_MAX_PART+j)*4 + {px,py,pz,E}].  aux (may be NULL) [i*5+..] =
 * yPair, mPair, cos(theta), pT(gamma1), pT(gamma2). */
int upcgpu_generate(upcgpu_ctx* ctx, uint64_t seed, uint64_t first_candidate, size_t n_candidates,
                    int* npart, int* pdg, int* status, int* mother, double* p4, double* aux,
                    uint64_t* n_accepted);
/* The same with the particle arrays packed to part_stride slots per candidate instead of UPCGPU_MAX_PART:
 * pdg/status/mother[i*part_stride+j], p4[(i*part_stride+j)*4+..]; part_stride must be at least
 * upcgpu_particles_per_event() -- 2 for pair production, 1 for single production, + 2 per uniform two-body decay
 * (ALP -> gamma gamma: 3; pi0 pi0 -> 4 gamma: 6) -- so an accepted event always fits.  Fewer bytes cross PCIe: 92 B per candidate for lepton
 * pairs instead of 180.  With PINNED host buffers the copies of one chunk of candidates overlap the kernels of the
 * next (pageable buffers work, without the overlap). */
int upcgpu_particles_per_event(const upcgpu_ctx* ctx);
int upcgpu_generate_packed(upcgpu_ctx* ctx, uint64_t seed, uint64_t first_candidate, size_t n_candidates, int part_stride,
                           int* npart, int* pdg, int* status, int* mother, double* p4, double* aux,
                           uint64_t* n_accepted);
/* device-only variant for throughput measurement: generates and keeps results on the device,
 * returns the accepted count */
int upcgpu_generate_device(upcgpu_ctx* ctx, uint64_t seed, uint64_t first_candidate, size_t n_candidates,
                           uint64_t* n_accepted);
/* photon-pT pdf of getPhotonPt for one photon energy: cdf[5001] (TH1 integral), test hook */
int upcgpu_photon_pt_cdf(upcgpu_ctx* ctx, double e_phot, double* cdf);
/* measurement aid: dependent-free DFMA loop on every SM; returns the FP64 FMA rate in TFLOP/s
 * (2 flop per DFMA) and the kernel time.  Used as the roofline denominator (MEASURED_PEAKS.json
 * carries no FP64 figure). */
int upcgpu_fp64_peak(upcgpu_ctx* ctx, int iters, double* tflops, double* ms);
/* Philox4x32-10 uniforms, test hook: out[2*i], out[2*i+1] for counter ctr0+i, block */
int upcgpu_philox(uint64_t seed, uint64_t ctr0, uint32_t block, size_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif
