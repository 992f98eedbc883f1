// upc_tables.cu -- device construction of the lookup tables T1-T4:
//   rho0 (calcWSRho), T_A and G_AA (prepareGAA), the 1e6-knot form-factor spline
//   (prepareFormFac) and the breakup-probability spline (prepareBreakupProb/calcBreakupProb),
// reference src/UpcCrossSection.cpp:152-163, :364-461, :752-1019.
// All of it runs once per context in ~1 ms; nothing here is on the CPU.
#include <cstring>
#include <cstddef>
#include <cstdio>

#include "upc_ctx.h"
#include <vector>

#include "upc_internal.h"

namespace upc {

// GL10 positive abscissas / weights, include/UpcCrossSection.h:90-111.  cos(pi x_k) is computed
// on the host with the platform libm (as the reference does) and passed in.
struct Gl5 {
  double w[5], c[5], s[5];
};

// ---------------------------------------------------------------------------------------------
// src/UpcCrossSection.cpp:139-150, literal (n = 200: index 199 is an end point, 198 gets weight 2)
__device__ inline double simpson_seq(int n, const double* v, double h)
{
  double sum = v[0] + v[n - 1];
  for (int i = 1; i < n - 1; i += 2) sum += 4. * v[i];
  for (int i = 2; i < n - 1; i += 2) sum += 2. * v[i];
  return sum * h / 3.;
}

// T1: calcWSRho, src/UpcCrossSection.cpp:152-163
__global__ void k_rho0(double R, double a, int A, double* out)
{
  __shared__ double v[kNB];
  const double bmax = 20.;
  const double db = bmax / (kNB - 1.);
  for (int ib = threadIdx.x; ib < kNB; ib += blockDim.x) {
    double r = ib * db;
    v[ib] = r * r / (1. + exp((r - R) / a));
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = A / simpson_seq(kNB, v, db) / 4. / kPi;
}

// T2a: T_A(b) = 2 int rho dz, src/UpcCrossSection.cpp:374-383.  One block per b.
__global__ void k_ta(double R, double a, const double* rho0p, double* xb, double* ta)
{
  __shared__ double v[kNB];
  const double bmax = 20.;
  const double db = bmax / (kNB - 1);
  const double rho0 = rho0p[0];
  const int ib = blockIdx.x;
  const double b = ib * db;
  for (int iz = threadIdx.x; iz < kNB; iz += blockDim.x) {
    double z = iz * db;
    double r = sqrt(b * b + z * z);
    v[iz] = rho0 / (1 + exp((r - R) / a));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ta[ib] = 2. * simpson_seq(kNB, v, db);
    xb[ib] = b;
  }
}

// GSL cspline_init for a small table (gsl_linalg_solve_symm_tridiag, LDL^T), one CTA.  The three recurrences are
// sequential and walked by one thread in GSL's order -- but on shared memory, with the matrix entries and right-hand
// sides formed in parallel beforehand (a single thread reading global memory at every step took 71 / 173 us for the
// two 200-knot tables: the critical path of the table stage).
__global__ void __launch_bounds__(256) k_spline_small(const double* xa, const double* ya, int size, double* c)
{
  __shared__ double sx[kNB], sy[kNB], sdiag[kNB], soff[kNB], srhs[kNB], gamma[kNB], alpha[kNB], z[kNB], xs[kNB];
  const int tid = threadIdx.x;
  const int max_index = size - 1;
  const int N = max_index - 1;
  for (int i = tid; i < size; i += blockDim.x) { sx[i] = xa[i]; sy[i] = ya[i]; }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    sdiag[i] = 2.0 * ((sx[i + 2] - sx[i + 1]) + (sx[i + 1] - sx[i]));
    soff[i] = sx[i + 2] - sx[i + 1];
    const double h_i = sx[i + 1] - sx[i], h_ip1 = sx[i + 2] - sx[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    srhs[i] = 3.0 * ((sy[i + 2] - sy[i + 1]) * g_ip1 - (sy[i + 1] - sy[i]) * g_i);
  }
  __syncthreads();
  if (tid == 0) {
    alpha[0] = sdiag[0];
    gamma[0] = soff[0] / alpha[0];
    for (int i = 1; i < N - 1; i++) {
      alpha[i] = sdiag[i] - soff[i - 1] * gamma[i - 1];
      gamma[i] = soff[i] / alpha[i];
    }
    if (N > 1) alpha[N - 1] = sdiag[N - 1] - soff[N - 2] * gamma[N - 2];
    z[0] = srhs[0];
    for (int i = 1; i < N; i++) z[i] = srhs[i] - gamma[i - 1] * z[i - 1];
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) z[i] = z[i] / alpha[i];  // the quotients of the back substitution
  __syncthreads();
  if (tid == 0) {
    xs[N - 1] = z[N - 1];
    for (int i = N - 2; i >= 0; i--) xs[i] = z[i] - gamma[i] * xs[i + 1];
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) c[i + 1] = xs[i];
  if (tid == 0) { c[0] = 0.; c[max_index] = 0.; }
}

// T2b: G_AA(b) = exp(-sigma_NN T_AA(b)), src/UpcCrossSection.cpp:389-406.  One block per b,
// one thread per s.
__global__ void k_gaa(const double* xb, const double* ta, const double* ta_c, double csNN, Gl5 gl, double* gaa)
{
  __shared__ double vs[kNB];
  const double bmax = 20.;
  const double db = bmax / (kNB - 1);
  const double inv_db = (kNB - 1) / bmax;
  const int ib = blockIdx.x;
  const double b = ib * db;
  for (int is = threadIdx.x; is < kNB; is += blockDim.x) {
    double s = is * db;
    double sum_phi = 0;
    for (int k = 0; k < 5; k++) {
      double r = sqrt(b * b + s * s + 2. * b * s * gl.c[k]);
      sum_phi += 2. * kPi * gl.w[k] * spline_eval_raw(xb, ta, ta_c, kNB, 0., inv_db, r < bmax ? r : bmax);
    }
    vs[is] = 2. * s * spline_eval_raw(xb, ta, ta_c, kNB, 0., inv_db, s < bmax ? s : bmax) * sum_phi;
  }
  __syncthreads();
  if (threadIdx.x == 0) gaa[ib] = exp(-csNN * simpson_seq(kNB, vs, db));
}

// evaluation-form segments for a table given by arrays
__global__ void k_segs_from_arrays(const double* xa, const double* ya, const double* ca, int n, SplineSeg* seg,
                                   double tail_value)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n - 1) seg[i] = make_seg(xa[i], xa[i + 1], ya[i], ya[i + 1], ca[i], ca[i + 1]);
  if (i == n - 1) seg[i] = SplineSeg{tail_value, 0., 0., 0.};
}

// (knot(x0, dx, i): a knot of a uniform table exactly as the reference forms it, upc_math.cuh)
// ... except the breakup table, which is written `bmin + db * i` (same thing)

__global__ void k_segs_uniform(double x0, double dx, const double* ya, const double* ca, int n, SplineSeg* seg,
                               int n_seg_out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n - 1 && i < n_seg_out)
    seg[i] = make_seg(knot(x0, dx, i), knot(x0, dx, i + 1), ya[i], ya[i + 1], ca[i], ca[i + 1]);
}

// T3: calcFormFac on the 1e6 knots, src/UpcCrossSection.cpp:436-456
__device__ inline double calc_formfac(double Q2, double R, double a, double rho0)
{
  double Q = sqrt(Q2) / kHc;
  double coshVal = cosh(kPi * Q * a);
  double sinhVal = sinh(kPi * Q * a);
  double ff = 4 * kPi * kPi * rho0 * a * a * a / (Q * a * Q * a * sinhVal * sinhVal) *
              (kPi * Q * a * coshVal * sin(Q * R) - Q * R * cos(Q * R) * sinhVal);
  ff += 8 * kPi * rho0 * a * a * a * exp(-R / a) / (1 + Q * Q * a * a) / (1 + Q * Q * a * a);
  return ff;
}

__global__ void k_ff_y(double R, double a, const double* rho0p, double* y)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kNQ2) y[i] = calc_formfac(knot(kQ2min, kDQ2, i), R, a, rho0p[0]);
}

// natural cubic spline c-coefficients of a long uniform table.  The (1,4,1)-type system is
// strictly diagonally dominant, so the influence of a far knot decays as (2-sqrt 3)^distance
// (1e-37 after 64 knots): each thread solves the system on its chunk plus a 64-knot halo with
// the same LDL^T recurrences GSL uses and keeps the chunk.
// (rcp_fast: upc_math.cuh)
constexpr int kSpChunk = 48;   // short chunks: the kernel is a latency chain per thread (192: 0.33 ms for the 10^6-knot table)
constexpr int kSpHalo = 64;
constexpr int kSpThreads = 32;
constexpr int kSpWin = kSpThreads * kSpChunk + 2 * kSpHalo + 2;
__global__ void __launch_bounds__(kSpThreads) k_spline_windowed(double x0, double dx, const double* ya, int size, double* c)
{
  // Everything a step of the recurrence needs except the recurrence itself -- the knot spacings h_i and the right-hand
  // sides (two divisions each) -- is formed beforehand, in parallel, into shared memory; a step of a thread's chain is
  // then one FMA and one division (alpha -> gamma) with the second division (z / alpha) beside it.  With the spacings,
  // right-hand sides and their three loads inside the chain a step took ~1100 cycles: 100 us for the 20 200-knot
  // breakup table on 7 CTAs, the critical path of the table stage.
  __shared__ double sy[kSpWin], sh[kSpWin], srhs[kSpWin];
  const int N = size - 2;  // unknowns u = 0..N-1  <->  c[u+1]
  const int sb = blockIdx.x * kSpThreads * kSpChunk;
  const int wsb = max(0, sb - kSpHalo), web = min(N, sb + kSpThreads * kSpChunk + kSpHalo);
  const int nwin = web - wsb;  // unknowns of the block's window; knots wsb .. web + 1
  for (int i = threadIdx.x; i < nwin + 2; i += kSpThreads) sy[i] = ya[wsb + i];
  for (int i = threadIdx.x; i < nwin + 1; i += kSpThreads) sh[i] = __dsub_rn(knot(x0, dx, wsb + i + 1), knot(x0, dx, wsb + i));
  __syncthreads();
  for (int i = threadIdx.x; i < nwin; i += kSpThreads) {
    const double h_i = sh[i], h_ip1 = sh[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    srhs[i] = 3.0 * ((sy[i + 2] - sy[i + 1]) * g_ip1 - (sy[i + 1] - sy[i]) * g_i);
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = t * kSpChunk;
  if (t == 0) {
    c[0] = 0.;
    c[size - 1] = 0.;
  }
  if (s >= N) return;
  const int e = min(N, s + kSpChunk);
  const int ws = max(0, s - kSpHalo), we = min(N, e + kSpHalo);
  const int W = we - ws;
  double gamma[kSpChunk + 2 * kSpHalo], z[kSpChunk + 2 * kSpHalo];
  const double* h = sh - wsb;      // h[i] = x_{i+1} - x_i
  const double* rhs = srhs - wsb;  // right-hand side of unknown i
  // forward: alpha_i = diag_i - off_{i-1} gamma_{i-1}; gamma_i = off_i/alpha_i; z_i = b_i - gamma_{i-1} z_{i-1}
  double gamma_prev = 0, z_prev = 0;
#pragma unroll 4
  for (int j = 0; j < W; j++) {
    const int i = ws + j;
    const double diag = 2.0 * (h[i + 1] + h[i]);
    const double off = h[i + 1];
    const double alpha = (j == 0) ? diag : diag - h[i] * gamma_prev;  // off_{i-1} = h(i)
    const double ra = rcp_fast(alpha);
    const double g = off * ra;
    const double zz = (j == 0) ? rhs[i] : rhs[i] - gamma_prev * z_prev;
    gamma[j] = g;
    z[j] = zz * ra;  // store c_i = z_i/alpha_i
    gamma_prev = g; z_prev = zz;
  }
  // back substitution x_i = c_i - gamma_i x_{i+1}
  double xn = z[W - 1];
  if (we - 1 >= s && we - 1 < e) c[we - 1 + 1] = xn;
  for (int j = W - 2; j >= 0; j--) {
    xn = z[j] - gamma[j] * xn;
    const int i = ws + j;
    if (i >= s && i < e) c[i + 1] = xn;
  }
}

// T4: photo-nuclear table of calcBreakupProb, src/UpcCrossSection.cpp:853-969 (one thread)
// per-context (the reference keeps these in function-local statics: one generator per process)
struct BkTable {
  double ee[700], se[700];
  double zcon, o0, gammatarg, omaxx;
  int n;
};

__device__ const double bk_e1[23] = {0., 103., 106., 112., 119., 127., 132., 145., 171., 199., 230., 235.,
                                     254., 280., 300., 320., 330., 333., 373., 390., 420., 426., 440.};
__device__ const double bk_s1[23] = {0., 12.0, 11.5, 12.0, 12.0, 12.0, 15.0, 17.0, 28.0, 33.0, 52.0, 60.0,
                                     70.0, 76.0, 85.0, 86.0, 89.0, 89.0, 75.0, 76.0, 69.0, 59.0, 61.0};
__device__ const double bk_e2[12] = {0., 2000., 3270., 4100., 4810., 6210., 6600., 7790., 8400., 9510., 13600., 16400.};
__device__ const double bk_s2[12] = {0., .1266, .1080, .0805, .1017, .0942, .0844, .0841, .0755, .0827, .0626, .0740};
__device__ const double bk_e3[29] = {0., 26., 28., 30., 32., 34., 36., 38., 40., 44., 46., 48., 50., 52., 55.,
                                     57., 62., 64., 66., 69., 72., 74., 76., 79., 82., 86., 92., 98., 103.};
__device__ const double bk_s3[29] = {0., 30., 21.5, 22.5, 18.5, 17.5, 15., 14.5, 19., 17.5, 16., 14., 20., 16.5, 17.5,
                                     17., 15.5, 18., 15.5, 15.5, 15., 13.5, 18., 14.5, 15.5, 12.5, 13., 13., 12.};
// Armstrong et al. gamma-p / gamma-n totals, entries 0..70 (the loop at :932 reads 9..70)
__device__ const double bk_sigt[71] = {
  0., .4245, .4870, .5269, .4778, .4066, .3341, .2444, .2245, .2005, .1783, .1769, .1869, .1940, .2117, .2226,
  .2327, .2395, .2646, .2790, .2756, .2607, .2447, .2211, .2063, .2137, .2088, .2017, .2050, .2015, .2121, .2175,
  .2152, .1917, .1911, .1747, .1650, .1587, .1622, .1496, .1486, .1438, .1556, .1468, .1536, .1544, .1536, .1468,
  .1535, .1442, .1515, .1559, .1541, .1461, .1388, .1565, .1502, .1503, .1454, .1389, .1445, .1425, .1415, .1424,
  .1432, .1486, .1539, .1354, .1480, .1443, .1435};
__device__ const double bk_sigtn[71] = {
  0., .3125, .3930, .4401, .4582, .3774, .3329, .2996, .2715, .2165, .2297, .1861, .1551, .2020, .2073, .2064,
  .2193, .2275, .2384, .2150, .2494, .2133, .2023, .1969, .1797, .1693, .1642, .1463, .1280, .1555, .1489, .1435,
  .1398, .1573, .1479, .1493, .1417, .1403, .1258, .1354, .1394, .1420, .1364, .1325, .1455, .1326, .1397, .1286,
  .1260, .1314, .1378, .1353, .1264, .1471, .1650, .1311, .1261, .1348, .1277, .1518, .1297, .1452, .1453, .1598,
  .1323, .1234, .1212, .1333, .1434, .1380, .1330};

__global__ void k_bk_init(double beam_gamma, BkTable* T)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* ee = T->ee;
  double* se = T->se;
  int zp = 82, ap = 208;  // Q2: the reference hard-codes lead (:758-759, :882-883)
  const double hbarcmev = 197.3269718;
  const double pi = 3.14159;  // :767
  double gammatarg = 2. * beam_gamma * beam_gamma - 1.;
  double si1 = 640., g1 = 4.05, o1 = 13.42, o0 = 7.4;
  double delo = .05;
  double scon = .1 * g1 * g1 * si1;
  double zcon = zp / (gammatarg * (pi) * (hbarcmev)) * zp / (gammatarg * (pi) * (hbarcmev)) / 137.04;
  int ne = int((25. - o0) / delo) + 1;
  for (int i = 1; i <= ne; i++) {
    ee[i] = o0 + (i - 1) * delo;
    se[i] = scon * ee[i] * ee[i] / (((o1 * o1 - ee[i] * ee[i]) * (o1 * o1 - ee[i] * ee[i])) + ee[i] * ee[i] * g1 * g1);
  }
  int ij = ne;
  for (int j = 1; j <= 27; j++) { ij++; ee[ij] = bk_e3[j]; se[ij] = .1 * ap * bk_s3[j] / 208.; }
  for (int j = 1; j <= 22; j++) { ij++; ee[ij] = bk_e1[j]; se[ij] = .1 * ap * bk_s1[j] / 208.; }
  for (int j = 9; j <= 70; j++) {
    ij++;
    ee[ij] = ee[ij - 1] + 25.;
    se[ij] = .1 * (zp * bk_sigt[j] + (ap - zp) * bk_sigtn[j]);
  }
  for (int j = 1; j <= 11; j++) { ij++; ee[ij] = bk_e2[j]; se[ij] = .1 * ap * bk_s2[j]; }
  double x = .0677, y = .129, eps = .0808, eta = .4525, em = .94;
  double exx = pow(10., .05);
  double s = .002 * em * ee[ij];
  int ictr = 100;
  if (gammatarg > (2. * 150. * 150.)) ictr = 150;
  for (int j = 1; j <= ictr; j++) {
    ij++;
    s = s * exx;
    ee[ij] = 1000. * .5 * (s - em * em) / em;
    double pom = x * pow(s, eps);
    double vec = y * pow(s, (-eta));
    se[ij] = .1 * .65 * ap * (pom + vec);
  }
  ee[ij + 1] = 99999999999.;
  T->n = ij;
  T->zcon = zcon;
  T->o0 = o0;
  T->gammatarg = gammatarg;
  T->omaxx = beam_gamma > 500. ? 1.E10 : 1.E7;
}

// calcBreakupProb for one b, src/UpcCrossSection.cpp:972-1018 (the unused one-neutron sum omitted)
__device__ inline double calc_breakup(const BkTable* __restrict__ T, double b, int mode)
{
  const double hbarcmev = 197.3269718;
  const double* ee = T->ee;
  const double* se = T->se;
  const double gammatarg = T->gammatarg, zcon = T->zcon;
  double prob = 0.;
  double pxn = 0.;
  double omax = fmin(T->omaxx, 4. * gammatarg * (hbarcmev) / b);
  if (omax < T->o0) return prob;
  double gk1m = tmath_bessel_k1(ee[1] * b / ((hbarcmev)*gammatarg));
  int k = 2;
  while (ee[k] < omax) {
    double gk1 = tmath_bessel_k1(ee[k] * b / ((hbarcmev)*gammatarg));
    // explicit non-fused arithmetic in the reference's order
    double t1 = __dmul_rn(__dmul_rn(__dmul_rn(se[k - 1], ee[k - 1]), gk1m), gk1m);
    double t2 = __dmul_rn(__dmul_rn(__dmul_rn(se[k], ee[k]), gk1), gk1);
    double term = __dmul_rn(__dmul_rn(__dmul_rn(zcon, __dsub_rn(ee[k], ee[k - 1])), .5), __dadd_rn(t1, t2));
    pxn = __dadd_rn(pxn, term);
    k = k + 1;
    gk1m = gk1;
  }
  if (mode == 1) prob = 1.;
  if (mode == 2) prob = (1 - exp(-1 * pxn)) * (1 - exp(-1 * pxn));
  if (mode == 3) prob = exp(-2 * pxn);
  if (mode == 4) prob = 2. * exp(-pxn) * (1. - exp(-pxn));
  return prob;
}

// the table of P(b): one WARP per knot.  The lanes form K1 at the energy knots and the trapezoid terms in parallel
// (the same expressions as calc_breakup), then lane 0 adds the terms in the reference's order -- the sum keeps its
// rounding, the ~600 Bessel evaluations of a knot are spread over 32 lanes (one thread per knot left 4 warps per SM
// busy for 0.34 ms; this is the critical path of the table stage).
constexpr int kBkWarps = 4;
__global__ void __launch_bounds__(32 * kBkWarps) k_bk_prob(const BkTable* __restrict__ T, int mode, int n, double* __restrict__ y)
{
  __shared__ double sg[kBkWarps][700];   // K1 at knot k, then the term of step k
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * kBkWarps + w;
  if (i >= n) return;
  const double b = knot(kBkBmin, kBkDb, i);
  const double hbarcmev = 197.3269718;
  const double* ee = T->ee;
  const double* se = T->se;
  const double gammatarg = T->gammatarg, zcon = T->zcon;
  const double omax = fmin(T->omaxx, 4. * gammatarg * (hbarcmev) / b);
  if (omax < T->o0) {
    if (lane == 0) y[i] = 0.;
    return;
  }
  // K = first k >= 2 with ee[k] >= omax: the loop of calc_breakup stops there (ee[n + 1] is a sentinel above any omax)
  const int nk = T->n;
  int kmin = nk + 1;
  for (int k = 2 + lane; k <= nk; k += 32)
    if (!(ee[k] < omax)) kmin = min(kmin, k);
  const int K = __reduce_min_sync(0xffffffffu, kmin);
  double* g = sg[w];
  for (int k = 1 + lane; k < K; k += 32) g[k] = tmath_bessel_k1(ee[k] * b / ((hbarcmev)*gammatarg));
  __syncwarp();
  double term[20];  // terms of k = 2 + lane + 32 j (K <= 700: at most 22 per lane; bounded below)
  int nt = 0;
  for (int k = 2 + lane; k < K && nt < 20; k += 32, ++nt) {
    const double gk1m = g[k - 1], gk1 = g[k];
    const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(se[k - 1], ee[k - 1]), gk1m), gk1m);
    const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(se[k], ee[k]), gk1), gk1);
    term[nt] = __dmul_rn(__dmul_rn(__dmul_rn(zcon, __dsub_rn(ee[k], ee[k - 1])), .5), __dadd_rn(t1, t2));
  }
  __syncwarp();
  nt = 0;
  for (int k = 2 + lane; k < K && nt < 20; k += 32, ++nt) g[k] = term[nt];
  __syncwarp();
  if (lane == 0) {
    double pxn = 0.;
#pragma unroll 8
    for (int k = 2; k < K; ++k) pxn = __dadd_rn(pxn, g[k]);
    double prob = 0.;
    if (mode == 1) prob = 1.;
    if (mode == 2) prob = (1 - exp(-1 * pxn)) * (1 - exp(-1 * pxn));
    if (mode == 3) prob = exp(-2 * pxn);
    if (mode == 4) prob = 2. * exp(-pxn) * (1. - exp(-pxn));
    y[i] = prob;
  }
}

__global__ void k_bk_raw(const BkTable* T, const double* b, int mode, size_t n, double* out)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = calc_breakup(T, b[i], mode);
}

// scalars finishing the tables, one launch, so that the host needs one read-back:
//   out[0] = ff_last = F(Q2max - dQ2), out[1] = P20 = P(20), out[2] = number of photo-nuclear energy knots,
//   out[3] = number of leading G_AA segments bounded by 1e-20 (the inner cut of the cell quadrature);
// and the clamp segment {P(20), 0, 0, 0} of the breakup table at index i20.
__global__ void __launch_bounds__(256) k_table_scalars(const SplineSeg* ff_seg, const SplineSeg* gaa_seg, double* out)
{
  __shared__ int s_first;
  if (threadIdx.x == 0) s_first = kNB - 1;
  __syncthreads();
  {
    // inner cut: the leading G_AA segments whose magnitude is bounded by 1e-20 on the whole segment
    // (|y| + |b| h + |c| h^2 + |d| h^3); see prepare_tables.  One thread per segment; the first that fails.
    const double hh = 20. / (kNB - 1);
    const int i = threadIdx.x;
    if (i < kNB - 1) {
      const SplineSeg sg = gaa_seg[i];
      const double bound = fabs(sg.y) + hh * (fabs(sg.b) + hh * (fabs(sg.c) + hh * fabs(sg.d)));
      if (!(bound <= 1e-20)) atomicMin(&s_first, i);
    }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  out[3] = (double)s_first;
  {
    // src/UpcCrossSection.cpp:188: gsl_spline_eval(.., Q2max - dQ2): the last knot exactly
    double x = kQ2max - kDQ2;
    int i = kNQ2 - 2;
    out[0] = seg_eval(ff_seg[i], x - knot(kQ2min, kDQ2, i));
  }
}

// the breakup table's part, at the end of ITS chain (which runs beside the flux stage, see prepare_tables):
// out[1] = P(20), out[2] = energy knots, and the clamp segment {P(20), 0, 0, 0} at index i20
__global__ void k_bk_scalars(SplineSeg* bk_seg, int use_breakup, int i20, const BkTable* bk_table, double* out)
{
  if (use_breakup) {
    double x = 20.;
    int i = (int)((x - kBkBmin) / kBkDb);
    while (knot(kBkBmin, kBkDb, i) > x) --i;
    while (knot(kBkBmin, kBkDb, i + 1) <= x) ++i;
    const double p20 = seg_eval(bk_seg[i], x - knot(kBkBmin, kBkDb, i));
    out[1] = p20;
    out[2] = (double)bk_table->n;
    bk_seg[i20] = SplineSeg{p20, 0., 0., 0.};  // b >= 20 (index clamp): P(20), :260
  } else {
    out[1] = 1.;
    out[2] = 0.;
  }
}

// generic spline evaluation hook (GSL bsearch semantics), which: see upcgpu.h
__global__ void k_eval_uniform(const SplineSeg* seg, int nseg, double x0, double dx, const double* x, size_t n,
                               double* out)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  double xv = x[t];
  int i = (int)((xv - x0) / dx);
  i = max(0, min(i, nseg - 1));
  while (i > 0 && knot(x0, dx, i) > xv) --i;
  while (i < nseg - 1 && knot(x0, dx, i + 1) <= xv) ++i;
  out[t] = seg_eval(seg[i], xv - knot(x0, dx, i));
}

__global__ void k_eval_arrays(const double* xa, const SplineSeg* seg, int n, double inv_dx, const double* x, size_t cnt,
                              double* out)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  double xv = x[t];
  int i = (int)((xv - xa[0]) * inv_dx);
  i = max(0, min(i, n - 2));
  while (i > 0 && xa[i] > xv) --i;
  while (i < n - 2 && xa[i + 1] <= xv) ++i;
  out[t] = seg_eval(seg[i], xv - xa[i]);
}

// ---------------------------------------------------------------------------------------------
static Gl5 make_gl5(bool pol_sign)
{
  // include/UpcCrossSection.h:90-111, positive half; cos/sin by the host libm as the reference
  static const double w[5] = {0.2955242247147529, 0.2692667193099963, 0.2190863625159820, 0.1494513491505806,
                              0.0666713443086881};
  static const double x[5] = {0.1488743389816312, 0.4333953941292472, 0.6794095682990244, 0.8650633666889845,
                              0.9739065285171717};
  Gl5 g;
  for (int k = 0; k < 5; k++) {
    g.w[k] = w[k];
    g.c[k] = cos(M_PI * x[k]);
    g.s[k] = sin(M_PI * x[k]);
  }
  (void)pol_sign;
  return g;
}

int prepare_tables(upcgpu_ctx* c)
{
  const upcgpu_params& p = c->p;
  cudaStream_t st = c->stream;
  if (c->tables_pending) {  // an earlier, uncollected table stage: collect it before its events are reused
    const int frc = finish_tables(c);
    if (frc) return frc;
  }
  if (!c->tab_ev[0]) {
    UPC_CUDA(c, cudaEventCreate(&c->tab_ev[0]));
    UPC_CUDA(c, cudaEventCreate(&c->tab_ev[1]));
  }
  cudaEventRecord(c->tab_ev[0], st);

  if (!c->d_scal) {
  UPC_CUDA(c, cudaMalloc(&c->d_scal, 64 * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->gaa_x, kNB * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->gaa_y, kNB * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->gaa_c, kNB * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->ta_y, kNB * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->ta_c, kNB * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->gaa_seg, (kNB + 1) * sizeof(SplineSeg)));
  UPC_CUDA(c, cudaMalloc(&c->ff_y, kNQ2 * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->ff_c, kNQ2 * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&c->ff_seg, (size_t)kNQ2 * sizeof(SplineSeg)));
  }

  // sigma_NN, PDG 2016 fit, src/UpcCrossSection.cpp:368-369 (scalar, host libm like the reference)
  double ssm = pow(p.sqrts, 2) / pow(2 * kMProt + 2.1206, 2);
  double csNN = 0.1 * (34.41 + 0.2720 * pow(log(ssm), 2) + 13.07 * pow(ssm, -0.4473) - 7.394 * pow(ssm, -0.5486));
  c->info.sigma_nn = csNN;
  c->info.factor = p.Z * p.Z * kAlpha / M_PI / M_PI / kHc / kHc;  // :120

  // T2 (G_AA), T3 (form factor) and T4 (breakup) are independent chains of small, latency-bound kernels:
  // G_AA stays on the main stream, the other two run beside it on side streams (T3 needs rho0 only)
  if (!c->aux[0]) {
    UPC_CUDA(c, create_stream(&c->aux[0], /*high_priority=*/true));   // form factor: the flux stage waits for it
    UPC_CUDA(c, create_stream(&c->aux[1], /*high_priority=*/false));  // breakup: runs beside the flux stage
    for (int i = 0; i < 4; ++i) UPC_CUDA(c, cudaEventCreateWithFlags(&c->aux_ev[i], cudaEventDisableTiming));
  }
  cudaStream_t st_ff = c->aux[0], st_bk = c->aux[1];
  cudaEventRecord(c->aux_ev[0], st);
  cudaStreamWaitEvent(st_bk, c->aux_ev[0], 0);

  Gl5 gl = make_gl5(false);
  UPC_K(c), k_rho0<<<1, 256, 0, st>>>(p.R, p.a, p.A, c->d_scal);
  cudaEventRecord(c->aux_ev[1], st);
  cudaStreamWaitEvent(st_ff, c->aux_ev[1], 0);
  UPC_K(c), k_ta<<<kNB, 256, 0, st>>>(p.R, p.a, c->d_scal, c->gaa_x, c->ta_y);
  UPC_K(c), k_spline_small<<<1, 256, 0, st>>>(c->gaa_x, c->ta_y, kNB, c->ta_c);
  UPC_K(c), k_gaa<<<kNB, 256, 0, st>>>(c->gaa_x, c->ta_y, c->ta_c, csNN, gl, c->gaa_y);
  UPC_K(c), k_spline_small<<<1, 256, 0, st>>>(c->gaa_x, c->gaa_y, kNB, c->gaa_c);
  UPC_K(c), k_segs_from_arrays<<<1, 256, 0, st>>>(c->gaa_x, c->gaa_y, c->gaa_c, kNB, c->gaa_seg, 1.0);

  UPC_K(c), k_ff_y<<<(kNQ2 + 255) / 256, 256, 0, st_ff>>>(p.R, p.a, c->d_scal, c->ff_y);
  {
    int nthr = (kNQ2 - 2 + kSpChunk - 1) / kSpChunk;
    UPC_K(c), k_spline_windowed<<<(nthr + kSpThreads - 1) / kSpThreads, kSpThreads, 0, st_ff>>>(kQ2min, kDQ2, c->ff_y, kNQ2, c->ff_c);
  }
  UPC_K(c), k_segs_uniform<<<(kNQ2 + 255) / 256, 256, 0, st_ff>>>(kQ2min, kDQ2, c->ff_y, c->ff_c, kNQ2, c->ff_seg, kNQ2 - 1);
  cudaEventRecord(c->aux_ev[2], st_ff);

  int use_bk = p.breakup_mode > 1;
  if (use_bk) {
    // knots up to b = 20.2 fm: 20 001 are reachable, 200 more isolate the artificial right end
    c->bk_nknots = 20200;
    if (!c->bk_y) {
    UPC_CUDA(c, cudaMalloc(&c->bk_y, c->bk_nknots * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&c->bk_c, c->bk_nknots * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&c->bk_seg, (size_t)(c->bk_nknots + 1) * sizeof(SplineSeg)));
    UPC_CUDA(c, cudaMalloc(&c->bk_table, sizeof(BkTable)));
    }
    if (c->bk_table_g1 != p.g1) {  // depends on the beam gamma only; built once, like the reference's statics
      UPC_K(c), k_bk_init<<<1, 1, 0, st_bk>>>(p.g1, (BkTable*)c->bk_table);
      c->bk_table_g1 = p.g1;
    }
    UPC_K(c), k_bk_prob<<<(c->bk_nknots + kBkWarps - 1) / kBkWarps, 32 * kBkWarps, 0, st_bk>>>((const BkTable*)c->bk_table, p.breakup_mode, c->bk_nknots,
                                                          c->bk_y);
    int nthr = (c->bk_nknots - 2 + kSpChunk - 1) / kSpChunk;
    UPC_K(c), k_spline_windowed<<<(nthr + kSpThreads - 1) / kSpThreads, kSpThreads, 0, st_bk>>>(kBkBmin, kBkDb, c->bk_y, c->bk_nknots, c->bk_c);
    UPC_K(c), k_segs_uniform<<<(c->bk_nknots + 255) / 256, 256, 0, st_bk>>>(kBkBmin, kBkDb, c->bk_y, c->bk_c, c->bk_nknots,
                                                              c->bk_seg, c->bk_nknots - 1);
  }
  // segments 0..i20-1 of the breakup table cover [bmin, > 20); index i20 is the clamp segment
  const int i20 = (int)((20. - kBkBmin) / kBkDb) + 1;
  if (!c->h_scal) UPC_CUDA(c, cudaMallocHost(&c->h_scal, 8 * sizeof(double)));
  // the breakup chain ends with its scalars and their copy; its event is what the cell stage waits for
  UPC_K(c), k_bk_scalars<<<1, 1, 0, st_bk>>>(c->bk_seg, use_bk, i20, (const BkTable*)c->bk_table, c->d_scal + 1);
  UPC_CUDA(c, cudaMemcpyAsync(c->h_scal + 2, c->d_scal + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, st_bk));
  cudaEventRecord(c->aux_ev[3], st_bk);
  cudaStreamWaitEvent(st, c->aux_ev[2], 0);
  UPC_K(c), k_table_scalars<<<1, 256, 0, st>>>(c->ff_seg, c->gaa_seg, c->d_scal + 1);
  UPC_CUDA(c, cudaMemcpyAsync(c->h_scal, c->d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  UPC_CUDA(c, cudaMemcpyAsync(c->h_scal + 4, c->d_scal + 4, sizeof(double), cudaMemcpyDeviceToHost, st));
  if (c->scal_cached) {
    // Not the first table stage of this context: the flux stage needs G_AA's inner cut and the form factor only, so the
    // main stream does NOT wait for the breakup chain (0.21 ms, two thirds of the table stage) -- it runs beside the
    // flux rows, and the cell stage waits for its event (run_slab; finish_tables when no fill follows).
    c->bk_deferred = true;
  } else {
    cudaStreamWaitEvent(st, c->aux_ev[3], 0);
  }
  cudaEventRecord(c->tab_ev[1], st);
  c->tables_pending = true;
  if (!c->scal_cached) {
    // first table stage of this context: the one host wait; the scalars go into the cache
    const int frc = finish_tables(c);
    if (frc) return frc;
  }
  // later stages: nothing waits here -- the kernels that follow are queued behind the tables with the cached scalars as
  // launch parameters, and finish_tables (at the step's host wait) checks them against what the device produced

  const double* h = c->scal_cache;  // rho0, ff_last, P20, energy knots, leading zero segments of G_AA
  c->info.rho0 = h[0];
  c->info.breakup_p20 = h[2];
  c->info.n_breakup_energy_knots = (int)h[3];
  c->tab.bk_n = use_bk ? i20 : 0;
  {
    // inner cut of the cell quadrature: the leading G_AA segments whose magnitude is bounded by
    // 1e-20 on the whole segment (|y| + |b| h + |c| h^2 + |d| h^3); P(b) <= 1.  What is skipped adds less than
    // 1e-19 of a cell's sum (the far pairs alone, with G_AA = 1, are a fifth of the total weight): four orders below
    // one ulp.  (1e-30 put the cut at 10.6 fm for Pb-Pb 5.02 TeV, 1e-20 puts it at 12.4 fm: 20 % fewer look-ups.)
    const double hh = 20. / (kNB - 1);
    const double b_in = (int)h[4] * hh;
    c->tab.b_in2 = b_in * b_in;
    c->tab.gaa_i0 = std::max(0, (int)h[4] - 1);  // one segment of margin below the cut (b is compared with b_in * (1 - 1e-12))
    c->info.gaa_zero_below = b_in;
  }
  c->tab.gaa_seg = c->gaa_seg;
  c->tab.gaa_db = 20. / (kNB - 1);
  c->tab.gaa_inv_db = (kNB - 1) / 20.;
  c->tab.bk_seg = c->bk_seg;
  c->tab.p20 = h[2];
  c->tab.use_breakup = use_bk;
  c->tab.ff_seg = c->ff_seg;
  c->tab.ff_last = h[1];
  c->tables_ready = true;
  return UPCGPU_OK;
}

int finish_tables(upcgpu_ctx* c)
{
  if (!c->tables_pending) return UPCGPU_OK;
  c->tables_pending = false;
  cudaSetDevice(c->device);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->bk_deferred) {  // no cell stage took the breakup chain's event: wait for the chain here
    UPC_CUDA(c, cudaStreamSynchronize(c->aux[1]));
    c->bk_deferred = false;
  }
  UPC_CUDA(c, cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->tab_ev[0], c->tab_ev[1]);
  c->stats.ms_tables = ms;
  if (!c->scal_cached) {
    std::memcpy(c->scal_cache, c->h_scal, sizeof(c->scal_cache));
    c->scal_cached = true;
  } else if (std::memcmp(c->scal_cache, c->h_scal, sizeof(c->scal_cache)) != 0) {
    c->tables_ready = false;
    c->err = "prepare_tables: the table scalars differ from the ones of this context's first table stage";
    return UPCGPU_ECUDA;
  }
  return UPCGPU_OK;
}

int eval_table(upcgpu_ctx* c, int which, const double* x, size_t n, double* out)
{
  double *dx = nullptr, *dout = nullptr;
  UPC_CUDA(c, cudaMalloc(&dx, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dout, n * sizeof(double)));
  UPC_CUDA(c, cudaMemcpy(dx, x, n * sizeof(double), cudaMemcpyHostToDevice));
  unsigned g = (unsigned)((n + 127) / 128);
  if (which == UPCGPU_TABLE_GAA) {
    UPC_K(c), k_eval_arrays<<<g, 128, 0, c->stream>>>(c->gaa_x, c->gaa_seg, kNB, (kNB - 1) / 20., dx, n, dout);
  } else if (which == UPCGPU_TABLE_FORMFAC) {
    UPC_K(c), k_eval_uniform<<<g, 128, 0, c->stream>>>(c->ff_seg, kNQ2 - 1, kQ2min, kDQ2, dx, n, dout);
  } else if (which == UPCGPU_TABLE_BREAKUP && c->bk_seg) {
    UPC_K(c), k_eval_uniform<<<g, 128, 0, c->stream>>>(c->bk_seg, c->tab.bk_n, kBkBmin, kBkDb, dx, n, dout);
  } else {
    cudaFree(dx); cudaFree(dout);
    c->err = "eval_table: table not available";
    return UPCGPU_EINVAL;
  }
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(dx);
  cudaFree(dout);
  return UPCGPU_OK;
}

int breakup_raw(upcgpu_ctx* c, const double* b, int mode, size_t n, double* out)
{
  double *db = nullptr, *dout = nullptr;
  UPC_CUDA(c, cudaMalloc(&db, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dout, n * sizeof(double)));
  UPC_CUDA(c, cudaMemcpy(db, b, n * sizeof(double), cudaMemcpyHostToDevice));
  UPC_K(c), k_bk_raw<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>((const BkTable*)c->bk_table, db, mode, n, dout);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaMemcpy(out, dout, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(db);
  cudaFree(dout);
  return UPCGPU_OK;
}

}  // namespace upc
