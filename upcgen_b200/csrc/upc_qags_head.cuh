// upc_qags_head.cuh -- stage A.3a: the form-factor flux integrals, one thread per integral, for as long as
// gsl_integration_qags stays on the intervals it almost always visits.
//
// gsl_integration_qags on [0, 10] with this integrand (src/UpcCrossSection.cpp:181-212) is very regular.
// The integrand lives at small k_perp, so QAGS bisects [0, 10], [0, 5], [0, 2.5], [0, 1.25] and (98.9 %)
// [0, 0.625]; after that the oracle's trace over the cfg2 grid (tools/qags_interval_stats.py: 87 253
// integrals, 163 642 later bisections) shows 14 distinct intervals in all, nine of which take 99.65 % of
// the bisections: (level, position) = (5,0) 30 %, (4,1) 26 %, (6,0) 19 %, (7,0) 10 %, (5,1) 5 %, (8,0) 4 %,
// (5,2) 3 %, (1,1) 1 %, (9,0) 1 %.  So g -- the row-only factor of the integrand -- is tabulated once per row
// on the GK21 nodes of the 31 intervals that 15 such bisections produce (k_head_tables), and every integral
// runs the reference's QAGS in its own thread (same Qags<> state machine, state in shared memory strided by
// thread) for as long as the interval it bisects next is one of the 15 and its state fits.
//
// Two passes of one kernel template: pass 1 takes all integrals with an 11-interval state (9 bisections: 92 %
// finish; three CTAs per SM), pass 2 resumes the others from their HeadState with the 16-interval state of the
// fallback (two CTAs per SM).  What pass 2 leaves (an interval outside the table: 0.1 %) goes to
// k_flux_qags_rows with its complete QAGS state.  Nothing is speculated: every evaluation made here is one the
// reference makes, in the reference's order.
//
// The lanes of a warp are the same b index of 32 neighbouring m rows at one y row (HeadItemValid): nearly the
// same photon energy and b, so the same rounds, intervals and branch of J1; the g table is laid out
// [y row][node][m row] so that their loads coalesce.  Rows with k >= g1 hc / R share one b grid: their J1
// values on the tabulated nodes come from a table (k_head_j1_table).
#pragma once
#include "upc_hot.cuh"
#include "upc_qags.cuh"

namespace upc {

constexpr int kHdThreads = 128;
constexpr int kHdPar = 15;                // bisections the head knows (parents of tabulated intervals)
constexpr int kHdIv = 1 + 2 * kHdPar;     // tabulated intervals: [0,10], then (left, right) of each known bisection
constexpr int kHdG = kHdIv * 21;          // g values per row
constexpr int kHdCap = 11;                // interval-list capacity of the first pass: up to 9 bisections
constexpr int kHdEps = 13;                // epsilon-table capacity of the first pass (1 + 9 entries + 2 scratch, +1)
constexpr int kHsCap = 16;                // ... of the second pass (and of HeadState): up to 14 bisections
constexpr int kHsEps = 16;

// heap index ((1 << level) + position) of the known bisections, in table order: interval 1 + 2j is the left
// half of parent j, 2 + 2j the right half
__constant__ unsigned kHdParent[kHdPar] = {1, 2, 4, 8, 16, 17, 32, 64, 128, 33, 256, 34, 3, 512, 1024};
__host__ __device__ __forceinline__ int head_child_slot(unsigned heap)
{
  switch (heap) {
    case 1: return 1;
    case 2: return 3;
    case 4: return 5;
    case 8: return 7;
    case 16: return 9;
    case 17: return 11;
    case 32: return 13;
    case 64: return 15;
    case 128: return 17;
    case 33: return 19;
    case 256: return 21;
    case 34: return 23;
    case 3: return 25;
    case 512: return 27;
    case 1024: return 29;
    default: return -1;
  }
}

// QAGS state of an integral leaving the head (everything Qags<Store> holds; see upc_qags.cuh)
struct HeadState {
  double sc[11];
  double eps[kHsEps];
  double rl[kHsCap], el[kHsCap];
  unsigned hp[kHsCap];
  unsigned char od[kHsCap];
  int size, nrmax, i, maximum_level, ktmin, roundoff_type1, roundoff_type2, roundoff_type3, error_type, error_type2,
      iteration, tab_n, tab_nres, flags, neval, pad;
};
static_assert(sizeof(HeadState) % 8 == 0, "HeadState layout");
enum { kHdPositive = 1, kHdExtrapolate = 2, kHdDisallow = 4 };
constexpr unsigned kHdNoSlot = 0xffffffffu;  // state_slot[item]: no hand-over state (pool full)

template <int CAP, int EPS>
struct HdSharedT {
  double fv[21][kHdThreads];            // GK21 function values [node][thread]
  double rl[CAP][kHdThreads], el[CAP][kHdThreads];
  double ep[EPS][kHdThreads];
  double sc[11][kHdThreads];
  unsigned hp[CAP][kHdThreads];
  unsigned char od[CAP][kHdThreads];
};
using HdShared = HdSharedT<kHdCap, kHdEps>;   // first pass: three CTAs per SM
using HdShared2 = HdSharedT<kHsCap, kHsEps>;  // second pass: two
static_assert(3 * (sizeof(HdShared) + 1024) <= 227 * 1024, "three CTAs per SM");
static_assert(2 * (sizeof(HdShared2) + 1024) <= 227 * 1024, "two CTAs per SM");

// strided shared-memory store (same interval encoding as QagsSharedStore: heap indices)
template <int CAP, int EPS>
struct QagsHeadStoreT {
  HdSharedT<CAP, EPS>* sh;
  int slot;
  static constexpr int cap = CAP;
  static constexpr int eps_cap = EPS;
  __device__ __forceinline__ double& R(int k) { return sh->rl[k][slot]; }
  __device__ __forceinline__ double& E(int k) { return sh->el[k][slot]; }
  __device__ __forceinline__ int ord(int k) const { return sh->od[k][slot]; }
  __device__ __forceinline__ void set_ord(int k, int v) { sh->od[k][slot] = (unsigned char)v; }
  __device__ __forceinline__ int lvl(int k) const { return 31 - __clz(sh->hp[k][slot]); }
  static __device__ __forceinline__ double pow2(int e) { return __hiloint2double((1023 + e) << 20, 0); }
  __device__ __forceinline__ void get_iv(int k, double& a, double& b) const
  {
    const unsigned heap = sh->hp[k][slot];
    const int level = 31 - __clz(heap);
    const unsigned pos = heap - (1u << level);
    const double width = 10. * pow2(-level);
    a = pos * width;
    b = (pos + 1.) * width;
  }
  __device__ __forceinline__ bool set_iv(int k, double a, double /*b*/, int level)
  {
    sh->hp[k][slot] = (1u << level) + __double2uint_rn(a * (0.1 * pow2(level)));
    return true;
  }
  __device__ __forceinline__ double& eps(int k) { return sh->ep[k][slot]; }
  __device__ __forceinline__ double& sc(int k) { return sh->sc[k][slot]; }
};
using QagsHeadStore = QagsHeadStoreT<kHdCap, kHdEps>;

// bounds of tabulated interval iv (exact: dyadic sub-intervals of [0, 10])
__device__ __forceinline__ void head_interval(int iv, double& a, double& b)
{
  const unsigned heap = iv == 0 ? 1u : 2u * kHdParent[(iv - 1) >> 1] + ((iv - 1) & 1);
  const int level = 31 - __clz(heap);
  const unsigned pos = heap - (1u << level);
  const double width = 10. * QagsHeadStore::pow2(-level);
  a = pos * width;
  b = (pos + 1.) * width;
}

__device__ __forceinline__ double head_g(double x, double c0, const SplineSeg* __restrict__ ff, double ff_last)
{
  const double x2 = x * x;
  const double t = x2 + c0;
  double F = ff_last;
  if (t < kQ2max) {
    int idx = (int)((t - kQ2min) * (1. / kDQ2));
    idx = max(0, min(idx, kNQ2 - 2));
    const double delx = t - fma((double)idx, kDQ2, kQ2min);
    F = seg_eval(ld_seg(ff + idx), delx);
  }
  return x2 * F / t;
}

// g on the kHdG tabulated nodes of every row.  Layout hg[ir][e][iml] (row = iml * rows_per_m + ir): the lanes of a
// head warp are the same grid point i of 32 consecutive m rows at one y row ir, so a load of node e is one
// contiguous 256 B segment.
__global__ void k_head_tables(int n_m, int rows_per_m, const RowInfo* __restrict__ rows, double g1, DevTables tab,
                              double* __restrict__ hg)
{
  // one thread per (m row, y row, interval): 21 values each (a thread per row with all 651 values in sequence took
  // 0.33 ms however few rows there were -- it is a chain of dependent gathers -- and did not shrink with the shard)
  const int iml = blockIdx.x * blockDim.x + threadIdx.x;
  const int ir = blockIdx.y;
  const int iv = blockIdx.z;
  if (iml >= n_m) return;
  const double k = rows[(size_t)iml * rows_per_m + ir].k;
  const double c0 = k * k / g1 / g1;  // w*w/g/g, :187
  double* out = hg + ((size_t)ir * kHdG + (size_t)iv * 21) * n_m + iml;
  double a, b;
  head_interval(iv, a, b);
  const double center = 0.5 * (a + b), half = 0.5 * (b - a);
#pragma unroll 3
  for (int n = 0; n < 21; ++n) out[(size_t)n * n_m] = head_g(fma(half, kGkNode[n], center), c0, tab.ff_seg, tab.ff_last);
}

// The order in which the head takes the integrals: flat index q = (ir * nb + i) * n_m + iml, i.e. for one y row ir
// and one grid point i the integrals of consecutive m rows are neighbours.  M changes by dm from one m row to the
// next (0.1-1 % of k), so the 32 integrals of a warp have almost the same photon energy and the same b index:
// the same number of QAGS rounds, the same intervals, the same branch of J1 (lanes that are consecutive b of
// ONE row differ by a factor 3 in b: 66 % of the lanes were busy, here 9x %).  Selected: i < nq(row).
struct HeadItemValid {
  const int* nq;
  int n_m, nb, rows_per_m;
  __device__ __forceinline__ bool operator()(unsigned q) const
  {
    const unsigned per_ir = (unsigned)nb * (unsigned)n_m;
    const unsigned ir = q / per_ir, rem = q - ir * per_ir;
    const unsigned i = rem / (unsigned)n_m, iml = rem - i * (unsigned)n_m;
    return (int)i < nq[(size_t)iml * rows_per_m + ir];
  }
};

// J1 on the tabulated nodes of the COMMON b grid.  bmax = max(5 g1 hc / k, 5R) (:228-229) is 5R for every
// photon energy k >= g1 hc / R, so all those rows (34-38 % of the rows, 47-51 % of the integrals of the
// cfg2 / cfg4 grids: they are the rows with the most b <= 2R points) share one b grid, and with it the
// arguments beta_i * x of J1 on the tabulated nodes.  j1h[node][i] holds these values, formed by the same
// instruction sequence as head_gk21's own evaluation (same j1_3 site, same products), so a row of the
// common grid multiplies g by a table entry instead of evaluating J1: same bits, ~1/100 of the work.
constexpr int kJ1hStride = 128;           // i stride of j1h (nb <= 128)

__global__ void k_head_j1_table(int nb, double R, double* __restrict__ j1h)
{
  const int i = threadIdx.x;
  const int iv = blockIdx.x;
  if (i >= nb) return;
  RowInfo ri;  // the common grid, formed as k_rows_setup forms it when bmax = 5R
  ri.k = 0.; ri.bmin = 0.05 * R; ri.ld = (log(5. * R) - log(ri.bmin)) / nb; ri.nq = nb; ri.pad = 1;
  double b, w;
  grid_point(ri, i, b, w);
  const double beta = b * (1. / kHc);
  double a, bb;
  head_interval(iv, a, bb);
  const double center = 0.5 * (a + bb), half = 0.5 * (bb - a);
#pragma unroll 1
  for (int n = 0; n < 21; n += 3) {
    const D3 j = j1_3(D3{{beta * fma(half, kGkNode[n], center), beta * fma(half, kGkNode[n + 1], center),
                          beta * fma(half, kGkNode[n + 2], center)}});
    j1h[(size_t)(iv * 21 + n) * kJ1hStride + i] = j.v[0];
    j1h[(size_t)(iv * 21 + n + 1) * kJ1hStride + i] = j.v[1];
    j1h[(size_t)(iv * 21 + n + 2) * kJ1hStride + i] = j.v[2];
  }
}

// GK21 rule on tabulated interval iv = [center - half, center + half] for this thread's integral:
// f = g * J1(beta x), three nodes at a time; g = the row's table + 21 iv;
// jt != nullptr: the row is on the common b grid, J1 comes from the table (jt = j1h + i)
template <class SH>
__device__ __forceinline__ GkOut head_gk21(SH& sh, int iv, double center, double half, double beta,
                                           const double* __restrict__ g, size_t gs, const double* __restrict__ jt, int tid)
{
  const double* gi = g + (size_t)(iv * 21) * gs;
  double* fvp = &sh.fv[0][tid];
  if (jt) {
    const double* ji = jt + (size_t)(iv * 21) * kJ1hStride;
#pragma unroll
    for (int n = 0; n < 21; ++n) fvp[n * kHdThreads] = gi[n * gs] * ji[n * kJ1hStride];
  } else {
#pragma unroll 1
    for (int n = 0; n < 21; n += 3, gi += 3 * gs, fvp += 3 * kHdThreads) {
      const double g0 = gi[0], g1 = gi[gs], g2 = gi[2 * gs];
      const D3 j = j1_3(D3{{beta * fma(half, kGkNode[n], center), beta * fma(half, kGkNode[n + 1], center),
                            beta * fma(half, kGkNode[n + 2], center)}});
      fvp[0] = g0 * j.v[0];
      fvp[kHdThreads] = g1 * j.v[1];
      fvp[2 * kHdThreads] = g2 * j.v[2];
    }
  }
  return gk21_sums_strided<kHdThreads>(&sh.fv[0][tid], half);  // one site, fully unrolled
}

struct HeadCounters {
  unsigned long long evals;       // total evaluations (S.neval) of the integrals finished by the head passes
  unsigned long long errors;
  unsigned long long left;        // integrals the LAST pass hands to k_flux_qags_rows
  unsigned long long evals_made;  // evaluations made by the head passes themselves
  unsigned long long evals_tab;   // ... of which J1 came from the common-grid table
  unsigned long long left1;       // integrals the first pass hands to the second
};

// One thread per integral.  RESUME = false (first pass): all integrals, in the order of HeadItemValid (order[]: the
// selected flat indices), interval list up to CAP = 11 (three CTAs per SM).  RESUME = true (second pass): the
// integrals the first pass left (order[] compacted by its `left_flag`, *n_order of them), resumed from their
// HeadState with the capacity of the reference-sized fallback (CAP = 16, two CTAs per SM).
template <int CAP, int EPS, bool RESUME>
__global__ void __launch_bounds__(kHdThreads, RESUME ? 2 : 3)
k_flux_qags_head(const long long* __restrict__ n_items_dev, const int* __restrict__ n_order, int n_m, int rows_per_m, int nb,
                 const RowInfo* __restrict__ rows, const long long* __restrict__ item_off,
                 const unsigned* __restrict__ order, const double* __restrict__ hg, const double* __restrict__ j1h,
                 FluxConsts fc, double* __restrict__ W, int* __restrict__ neval_out, HeadCounters* __restrict__ ctr,
                 HeadState* __restrict__ state, unsigned* __restrict__ state_slot, unsigned* __restrict__ slot_ctr,
                 unsigned state_cap, long long* __restrict__ overflow_items, unsigned long long* __restrict__ overflow_ctr,
                 unsigned char* __restrict__ done_flag, unsigned char* __restrict__ left_flag)
{
  using SH = HdSharedT<CAP, EPS>;
  extern __shared__ __align__(16) unsigned char hd_smem[];
  SH& sh = *reinterpret_cast<SH*>(hd_smem);
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;

  // The first pass is launched with one thread per integral (the host sizes the grid from the integral count of
  // the previous identical fill, the device count is authoritative); the second pass has a grid for a quarter of the
  // integrals and strides over what the first left.
  const long long n_mine = RESUME ? (long long)*n_order : *n_items_dev;
  const long long stride = (long long)gridDim.x * kHdThreads;
  double my_evals = 0, my_made = 0, my_tab = 0;
  unsigned my_err = 0, my_left = 0;
#pragma unroll 1
  for (long long t = (long long)blockIdx.x * kHdThreads + tid; t < n_mine; t += stride) {
    const unsigned q = order[t];
    const unsigned per_ir = (unsigned)nb * (unsigned)n_m;
    const unsigned ir = q / per_ir, rem = q - ir * per_ir;
    const int i = (int)(rem / (unsigned)n_m), iml = (int)(rem - (unsigned)i * (unsigned)n_m);
    const size_t row = (size_t)iml * rows_per_m + ir;
    const long long item = item_off[row] + i;          // row-major index: done_flag, state_slot (k_flux_qags_rows)
    const RowInfo ri = rows[row];
    const double* g = hg + (size_t)ir * kHdG * n_m + iml;
    const size_t gs = (size_t)n_m;
    double b, w;
    grid_point(ri, i, b, w);
    const double beta = b * (1. / kHc);
    const double* jt = ri.pad ? j1h + i : nullptr;     // row on the common b grid
    Qags<QagsHeadStoreT<CAP, EPS>> S;
    S.sh = &sh;
    S.slot = tid;
    // ONE GK21 site for every rule (the kernel's code must stay inside the 32 KB instruction cache: with
    // three inlined sites `no_instruction` was the largest stall).  phase 0: the rule on [0, 10];
    // phase 1 / 2: the left / right half of the interval being bisected.
    bool done = false;
    GkOut ga{};
    int iv = 0, phase = 0;
    double center = 5., half = 5., c2 = 0., h2 = 0.;
    // the pass goes on while QAGS bisects an interval whose halves are tabulated and the state fits
    auto next_step = [&]() -> bool {
      const int cs = head_child_slot(sh.hp[S.i][tid]);
      if (cs < 0 || S.size + 1 >= CAP || S.tab_n + 2 >= EPS) return false;
      double a1, b1, a2, b2;
      int level;
      S.pre_step(a1, b1, a2, b2, level);
      iv = cs; phase = 1;
      center = 0.5 * (a1 + b1); half = 0.5 * (b1 - a1);
      c2 = 0.5 * (a2 + b2); h2 = 0.5 * (b2 - a2);
      return true;
    };
    bool go = true;
    int neval_in = 0;
    // hand-over states live in a pool of state_cap slots (8 % of the integrals of cfg2 need one): slot kHdNoSlot =
    // the first pass found the pool full, the integral starts again from the first rule
    unsigned sl = RESUME ? state_slot[item] : kHdNoSlot;
    if (RESUME && sl != kHdNoSlot) {
      const HeadState& hs = state[sl];
#pragma unroll
      for (int k = 0; k < 11; ++k) sh.sc[k][tid] = hs.sc[k];
#pragma unroll
      for (int k = 0; k < kHdEps; ++k) sh.ep[k][tid] = hs.eps[k];
#pragma unroll
      for (int k = 0; k < kHdCap; ++k) {
        sh.rl[k][tid] = hs.rl[k]; sh.el[k][tid] = hs.el[k];
        sh.hp[k][tid] = hs.hp[k]; sh.od[k][tid] = hs.od[k];
      }
      S.size = hs.size; S.nrmax = hs.nrmax; S.i = hs.i; S.maximum_level = hs.maximum_level; S.ktmin = hs.ktmin;
      S.roundoff_type1 = hs.roundoff_type1; S.roundoff_type2 = hs.roundoff_type2; S.roundoff_type3 = hs.roundoff_type3;
      S.error_type = hs.error_type; S.error_type2 = hs.error_type2; S.iteration = hs.iteration;
      S.tab_n = hs.tab_n; S.tab_nres = hs.tab_nres;
      S.positive_integrand = (hs.flags & kHdPositive) != 0;
      S.extrapolate = (hs.flags & kHdExtrapolate) != 0;
      S.disallow_extrapolation = (hs.flags & kHdDisallow) != 0;
      S.overflow = false;
      S.neval = hs.neval;
      S.result = 0; S.abserr = 0; S.ier = 0;
      neval_in = hs.neval;
      go = next_step();
    } else {
      S.begin(0., 10.);                                // :209
    }
#pragma unroll 1
    while (go) {
      const GkOut gk = head_gk21(sh, iv, center, half, beta, g, gs, jt, tid);
      if (phase == 1) {
        ga = gk;
        ++iv; center = c2; half = h2; phase = 2;
        continue;
      }
      done = phase == 0 ? S.post_first(gk) : S.post_step(ga, gk);
      if (done) break;
      go = next_step();
    }
    if (!RESUME) left_flag[t] = done ? 0 : 1;
    const double made = S.neval - neval_in;
    my_made += made;
    if (jt) my_tab += made;
    bool to_overflow = false;
    if (done) {
      const double Q = S.result / fc.A;                  // :214
      const double flux = fc.factor * Q * Q / ri.k;      // :215
      W[(size_t)row * nb + i] = flux * (b * w);
      if (neval_out) neval_out[(size_t)row * nb + i] = S.neval;
      my_evals += S.neval;
      if (S.ier != 0) my_err++;
    } else {
      if (sl == kHdNoSlot) {
        sl = atomicAdd(slot_ctr, 1u);
        if (sl >= state_cap) sl = kHdNoSlot;
        state_slot[item] = sl;
      }
      if (sl != kHdNoSlot) {
        HeadState& hs = state[sl];
#pragma unroll
        for (int k = 0; k < 11; ++k) hs.sc[k] = sh.sc[k][tid];
#pragma unroll
        for (int k = 0; k < EPS; ++k) hs.eps[k] = sh.ep[k][tid];
#pragma unroll
        for (int k = 0; k < CAP; ++k) {
          hs.rl[k] = sh.rl[k][tid]; hs.el[k] = sh.el[k][tid];
          hs.hp[k] = sh.hp[k][tid]; hs.od[k] = sh.od[k][tid];
        }
        hs.size = S.size; hs.nrmax = S.nrmax; hs.i = S.i; hs.maximum_level = S.maximum_level; hs.ktmin = S.ktmin;
        hs.roundoff_type1 = S.roundoff_type1; hs.roundoff_type2 = S.roundoff_type2; hs.roundoff_type3 = S.roundoff_type3;
        hs.error_type = S.error_type; hs.error_type2 = S.error_type2; hs.iteration = S.iteration;
        hs.tab_n = S.tab_n; hs.tab_nres = S.tab_nres;
        hs.flags = (S.positive_integrand ? kHdPositive : 0) | (S.extrapolate ? kHdExtrapolate : 0) |
                   (S.disallow_extrapolation ? kHdDisallow : 0);
        hs.neval = S.neval;
      } else if (RESUME) {
        // no slot even now: the integral is redone from scratch by the large-workspace pass (k_flux_qags_overflow);
        // its evaluations here are not counted twice
        to_overflow = true;
        overflow_items[atomicAdd(overflow_ctr, 1ull)] = item;
        my_made -= made;
        if (jt) my_tab -= made;
      }
      if (!to_overflow) my_left++;
    }
    // done_flag = 0 sends the integral to k_flux_qags_rows, which needs its state
    done_flag[item] = (done || to_overflow) ? 1 : 0;
  }
  const double ev = warp_sum(my_evals), evm = warp_sum(my_made), evt = warp_sum(my_tab);
  const unsigned er = __reduce_add_sync(0xffffffffu, my_err);
  const unsigned lf = __reduce_add_sync(0xffffffffu, my_left);
  if (lane == 0) {
    if (ev > 0) atomicAdd(&ctr->evals, (unsigned long long)ev);
    if (er) atomicAdd(&ctr->errors, (unsigned long long)er);
    if (lf) atomicAdd(RESUME ? &ctr->left : &ctr->left1, (unsigned long long)lf);
    if (evm > 0) atomicAdd(&ctr->evals_made, (unsigned long long)evm);
    if (evt > 0) atomicAdd(&ctr->evals_tab, (unsigned long long)evt);
  }
}

// the integrals of each row the head hands over, in order of b
__global__ void k_head_compact(int n_rows, const RowInfo* __restrict__ rows, const long long* __restrict__ item_off,
                               const unsigned char* __restrict__ done_flag, int* __restrict__ left_idx,
                               int* __restrict__ nq_left)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const long long i0 = item_off[row];
  const int nq = rows[row].nq;
  int n = 0;
  for (int i = 0; i < nq; ++i)
    if (!done_flag[i0 + i]) left_idx[i0 + n++] = i;
  nq_left[row] = n;
}

}  // namespace upc
