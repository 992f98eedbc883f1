// upc_qags_rows.cuh -- stage A.3b, the form-factor flux rows (F2/F3: fluxFormIntegrand / fluxForm,
// src/UpcCrossSection.cpp:181-218) as a row-cooperative, warp-specialised persistent kernel.
//
// Role today: the general fallback.  It takes the integrals the head passes (upc_qags_head.cuh) hand over --
// QAGS wants an interval outside the head's table, or more than 14 bisections: 0.1 % of the cfg2 grid -- and,
// unlike the head, caches g for ANY dyadic interval.  It was written when it carried the whole stage (DESIGN.md
// section 7, v3-v11), which is why it is built for throughput.
//
// The reference's integrand is   x^2 F(x^2 + (k/gamma)^2) / (x^2 + (k/gamma)^2) * J1(b x / hc):
// the first factor ("g") depends on the ROW (photon energy k) only, the Bessel factor on the
// integral (b).  Every integral starts on [0, 10] and QAGS bisects, so all intervals are dyadic
// and the ~30-100 integrals of a row visit the same ~20-30 of them (measured: 1100 GK21 rules per
// row on 29 distinct intervals).  One CTA per SM, 640 threads in five warpgroups:
//   * OWNERS (warpgroups 0 and 1 = four groups of 64 threads, one integral per thread): the QAGS
//     bookkeeping, which follows gsl_integration_qags decision for decision (upc_qags.cuh).
//     Interval lists, epsilon tables and the FP64 driver scalars live in SHARED memory, strided by
//     slot (bank = slot whatever entry a lane touches): the bookkeeping is a chain of dependent
//     look-ups that was latency-bound on L2 while that state sat in local memory.  A slot whose
//     integral is finished takes the next one of the global row queue.  Owners publish the 1-2
//     pending intervals of each integral as tasks.
//   * EVALUATORS (warpgroups 2 to 4): the integrand evaluations of a round (21 per task),
//     flattened over 384 threads whatever integral they belong to, two nodes per thread and
//     trip, sorted by the branch of J1 they take, followed by the GK21 sums in GSL's summation
//     order.  They serve the owner groups in turn, so that the bookkeeping of three groups runs
//     under the evaluations of the fourth and the FP64 pipe always has dense, convergent J1 work.
//     Register budgets follow the roles (setmaxnreg: 72 / 112).
//   * g on the 21 GK nodes of an interval is evaluated ONCE per (row, interval) -- the 10^6-knot
//     form-factor spline gather and the division leave the hot loop -- into an L2-resident
//     per-group table keyed by the interval's heap index (level, position).
// The value g * J1 is formed exactly as the reference's  x*x*F/t * J1  (left to right), so the
// QAGS decisions and results are the ones of the per-thread path (k_flux_list) bit for bit.
#pragma once
#include "upc_hot.cuh"
#include "upc_qags.cuh"
#include "upc_qags_head.cuh"

namespace upc {

constexpr int kRcThreads = 640;
constexpr int kRcGroups = 4;     // owner groups of two warps (warpgroups 0 and 1)
constexpr int kRcSlots = 64;     // integrals in flight per group (multiple of 32: conflict-free strides)
constexpr int kRcEval = 384;     // evaluator threads (three warpgroups)
constexpr int kRcCap = 16;       // interval-list capacity per integral (largest seen: 15)
constexpr int kRcEps = 16;       // epsilon-table capacity (largest index touched so far: 12)
constexpr int kRcCtx = 8;        // rows in flight per group
constexpr int kRcChunk = 4;      // rows per pop of the global queue
constexpr int kRcGrant = 4;      // row segments handed to idle slots per round
constexpr int kRcGE = 64;        // cached intervals per row in flight
constexpr int kRcG = kRcCtx * kRcGE;  // cached (row, interval) entries per group
constexpr int kRcMaxLevel = 26;  // deeper intervals are evaluated uncached
constexpr unsigned kRcEmpty = 0xffffffffu;
constexpr int kRcRegsOwner = 72, kRcRegsEval = 112;  // 256 * 72 + 384 * 112 = 60 K registers (launched with 96)

// state of one owner group
struct RcGroup {
  double gk[4][2][kRcSlots];           // GK21 sums of each task: result, abserr, resabs, resasc
  double rl[kRcCap][kRcSlots];         // QAGS interval list: integral estimates ...
  double el[kRcCap][kRcSlots];         // ... error estimates ...
  double ep[kRcEps][kRcSlots];         // Wynn epsilon table
  double sc[11][kRcSlots];             // FP64 scalars of the QAGS driver
  double t_center[2][kRcSlots];        // pending intervals of each slot (A = left/only, B = right)
  double t_half[2][kRcSlots];
  double beta[kRcSlots];               // b / hc
  double k_cur[kRcSlots], bw_cur[kRcSlots];  // photon energy, b * width of the slot's grid point
  double ctx_c0[kRcCtx];               // (k/gamma)^2 of each row in flight
  unsigned hp[kRcCap][kRcSlots];       // ... heap index (1 << level) + position of each interval ...
  unsigned gkeys[kRcG];
  short t_g[2][kRcSlots];              // cache entry of each pending interval (-1: uncached)
  unsigned char od[kRcCap][kRcSlots];  // ... and the error-sorted permutation
  unsigned char slot_ctx[kRcSlots];
  int t_off[2][kRcSlots];              // position of the task's evaluations in the round's list (small << 16 | large)
  unsigned char t_small[2][kRcSlots];  // number of nodes on J1's small-argument branch (255: no task)
  unsigned short newlist[kRcG];        // entries inserted this round (to be filled)
  int n_new[2];                        // by round parity
  int cnt[3][kRcSlots / 32];           // per-warp counts of the round: active slots, evaluations (packed), idle slots
  int n_cls[2];                        // evaluations handed to the evaluators: large-argument, small-argument
  int fin;                             // the group has no more work
  // rows in flight: a row keeps its context (c0, cached intervals) until its last integral is done
  int ctx_left[kRcCtx];                // integrals of the row not finished yet (0: context free)
  // the chunk of rows being handed out, and this round's grants to idle slots
  int chunk_row0, chunk_n, cur_row, cur_i, cur_ctx, exhausted;
  int chunk_nq[kRcChunk];
  int n_grant, clear_mask;
  int grant_rank0[kRcGrant], grant_n[kRcGrant], grant_row[kRcGrant], grant_i0[kRcGrant], grant_ctx[kRcGrant];
};

struct RcShared {
  RcGroup g[kRcGroups];
  double fv[2][kRcSlots][21];          // integrand values of the group-round in service [half][slot][node] (odd stride:
                                       // conflict-free both ways)
  unsigned short elist[kRcSlots * 42]; // the evaluations of the group-round in service (slot | half << 6 | node << 7):
                                       // J1 large-argument ones from the front, small-argument ones from the back
  double hot[H_END];                   // the J1 / sincos coefficient block (upc_hot) for LDS access
  double node[24];                     // signed GK21 abscissas (kGkNode)
  double snode[24];                    // the same in ascending order ...
  int sid[24];                         // ... and their index in kGkNode
};

static_assert(sizeof(RcShared) <= 227 * 1024, "shared memory of one SM");
static_assert(kRcCap == kHsCap && kRcEps <= kHsEps, "the row-cooperative kernel takes the head's state as it is");

// named barriers (0 is __syncthreads)
enum { kBarOwn0 = 1, kBarFull0 = 1 + kRcGroups, kBarDone0 = 1 + 2 * kRcGroups, kBarEval = 1 + 3 * kRcGroups };
static_assert(kBarEval < 16, "named barriers");
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// QAGS state of one slot in shared memory (see the store interface in upc_qags.cuh).  Intervals
// are dyadic sub-intervals of [0, 10] and are stored as heap indices.
struct QagsSharedStore {
  RcGroup* sh;
  int slot;
  static constexpr int cap = kRcCap;
  static constexpr int eps_cap = kRcEps;
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
    const double width = 10. * pow2(-level);  // exact
    a = pos * width;                          // exact: pos * 10 < 2^53
    b = (pos + 1.) * width;
  }
  static __device__ __forceinline__ unsigned heap_of(double a, int level)
  {
    // a = pos * 10 * 2^-level exactly, so a * (0.1 * 2^level) rounds to pos
    return (1u << level) + __double2uint_rn(a * (0.1 * pow2(level)));
  }
  __device__ __forceinline__ bool set_iv(int k, double a, double /*b*/, int level)
  {
    if (level > 30) return false;
    sh->hp[k][slot] = heap_of(a, level);
    return true;
  }
  __device__ __forceinline__ double& eps(int k) { return sh->ep[k][slot]; }
  __device__ __forceinline__ double& sc(int k) { return sh->sc[k][slot]; }
};

// g(x) = x^2 F(t) / t, t = x^2 + c0: the row-only factor of fluxFormIntegrand (:186-190) with
// the form-factor spline look-up and its clamp (:188)
__device__ __forceinline__ double rc_g(double x, double c0, const SplineSeg* __restrict__ ff, double ff_last)
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

// The same sums split over two threads of a task (contiguous fv[0..20]), fully unrolled: role 0
// forms the Kronrod, Gauss and |f| sums, role 1 re-forms the Kronrod sum (same operations, same
// order: same bits) for the mean and then the |f - mean| sum; gk21_finish combines them (role 0,
// after a shuffle).  The Gauss sum runs over the 5 Gauss pairs only, as in qk.c (the zero weights
// of gk21_tri add +0.0: the same value).
__device__ __forceinline__ void gk21_sums_pair(const double* fv, double half_length, int role, GkOut& o, double& asc)
{
  double f[21];
#pragma unroll
  for (int n = 0; n < 21; ++n) f[n] = fv[n];
  const double f_center = f[20];
  double result_kronrod = f_center * kGkWkC;
  if (role == 0) {
    double result_gauss = 0;
    double result_abs = fabs(result_kronrod);
#pragma unroll
    for (int p = 0; p < 10; ++p) {
      const double fsum = f[2 * p] + f[2 * p + 1];
      if (p < 5) result_gauss += kGkWg[p] * fsum;
      result_kronrod += kGkWk[p] * fsum;
      result_abs += kGkWk[p] * (fabs(f[2 * p]) + fabs(f[2 * p + 1]));
    }
    o.abserr = (result_kronrod - result_gauss) * half_length;  // raw error, rescaled in gk21_finish
    o.result = result_kronrod * half_length;
    o.resabs = result_abs * fabs(half_length);
  } else {
#pragma unroll
    for (int p = 0; p < 10; ++p) result_kronrod += kGkWk[p] * (f[2 * p] + f[2 * p + 1]);
    const double mean = result_kronrod * 0.5;
    double result_asc = kGkWkC * fabs(f_center - mean);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int p = (j & 1) ? (j >> 1) : (5 + (j >> 1));
      result_asc += kGkWk[p] * (fabs(f[2 * p] - mean) + fabs(f[2 * p + 1] - mean));
    }
    asc = result_asc * fabs(half_length);
  }
}

// rescale_error of integration/qk.c
__device__ __forceinline__ void gk21_finish(GkOut& o, double result_asc, double /*half_length*/)
{
  double err = fabs(o.abserr);
  if (result_asc != 0 && err != 0) {
    double s = 200 * err / result_asc;
    double scale = s * sqrt(s);
    err = scale < 1 ? result_asc * scale : result_asc;
  }
  if (o.resabs > DBL_MIN / (50 * DBL_EPSILON)) {
    double min_err = 50 * DBL_EPSILON * o.resabs;
    if (min_err > err) err = min_err;
  }
  o.abserr = err;
  o.resasc = result_asc;
}

// look up / insert the interval (heap index) in the table of row context ctx; returns the entry
// or -1 when the table is full or the interval too deep (the task is then evaluated uncached)
__device__ __forceinline__ int rc_entry(RcGroup& sh, int parity, int ctx, int level, unsigned heap)
{
  if (level > kRcMaxLevel) return -1;
  unsigned* keys = sh.gkeys + ctx * kRcGE;
  unsigned h = (heap * 2654435761u) >> 26;  // 6 bits
#pragma unroll 1
  for (int probe = 0; probe < kRcGE; ++probe) {
    const unsigned old = atomicCAS(&keys[h], kRcEmpty, heap);
    if (old == kRcEmpty) {
      const int slot = atomicAdd(&sh.n_new[parity], 1);
      sh.newlist[slot] = (unsigned short)(ctx * kRcGE + h);
      return ctx * kRcGE + (int)h;
    }
    if (old == heap) return ctx * kRcGE + (int)h;
    h = (h + 1) & (kRcGE - 1);
  }
  return -1;
}

// EVALUATORS: the integrand evaluations of one round of one owner group, flattened over the 256
// evaluator threads whatever integral they belong to, three per thread and trip.  The owners have
// sorted them by the branch of J1 they take (argument <= 8: polynomial; > 8: modulus/phase form),
// so every warp runs one branch with all lanes.
constexpr int kRcWide = 2;  // evaluations per evaluator thread and trip.  Measured on cfg2 (8 evaluator warps): 1: 40.8 ms,
                            // 2: 34.2, 3: 36.7, 4: 43.0, 6: 60; with 12 warps x 2: 32.1, 16 warps x 1: 33.2 -- short
                            // trips quantise better over the ~2700 evaluations of a group-round

template <bool LARGE>
__device__ __forceinline__ void rc_eval_trip(RcGroup& sh, double (*fv)[kRcSlots][21], const unsigned short* elist,
                                             const double* node, const HotShared& hs, int e, int n_ev,
                                             const double* __restrict__ gbuf, const SplineSeg* __restrict__ ff,
                                             double ff_last)
{
  constexpr int kLast = kRcSlots * 42 - 1;
  constexpr int N = kRcWide;
  int slot[N], half[N], nd[N], ent[N];
  DV<N> g, x, z;
  UPC_FOR_N {
    const int idx = min(e + i_, n_ev - 1);  // tail: repeats (same value, same place)
    const unsigned c = elist[LARGE ? idx : kLast - idx];
    slot[i_] = c & 63; half[i_] = (c >> 6) & 1; nd[i_] = c >> 7;
    ent[i_] = sh.t_g[half[i_]][slot[i_]];
    // issued first: the L2 latency is covered by the J1 arithmetic below
    if (ent[i_] >= 0) g.v[i_] = __ldcg(gbuf + ent[i_] * 21 + nd[i_]);
  }
  UPC_FOR_N {
    x.v[i_] = fma(sh.t_half[half[i_]][slot[i_]], node[nd[i_]], sh.t_center[half[i_]][slot[i_]]);
    if (ent[i_] < 0) g.v[i_] = rc_g(x.v[i_], sh.ctx_c0[sh.slot_ctx[slot[i_]]], ff, ff_last);
    z.v[i_] = sh.beta[slot[i_]] * x.v[i_];
  }
  const DV<N> j = LARGE ? j1_largeN<N>(z, hs) : j1_smallN<N>(z, hs);
  UPC_FOR_N fv[half[i_]][slot[i_]][nd[i_]] = g.v[i_] * j.v[i_];
}

// trips of kRcWide evaluations: the large-argument ones, padded to whole warps, then the small ones
__device__ __forceinline__ void rc_eval_round(RcGroup& sh, double (*fv)[kRcSlots][21], const unsigned short* elist,
                                              const double* node, const HotShared& hs, int etid, int n_large, int n_small,
                                              const double* __restrict__ gbuf, const SplineSeg* __restrict__ ff,
                                              double ff_last)
{
  const int t_large = (n_large + kRcWide - 1) / kRcWide, t_small = (n_small + kRcWide - 1) / kRcWide;
  const int t_pad = (t_large + 31) & ~31;
  const int n_trips = t_pad + t_small;
#pragma unroll 1
  for (int t = etid; t < n_trips; t += kRcEval) {
    if (t < t_pad) {  // warp-uniform
      if (t < t_large) rc_eval_trip<true>(sh, fv, elist, node, hs, kRcWide * t, n_large, gbuf, ff, ff_last);
    } else {
      rc_eval_trip<false>(sh, fv, elist, node, hs, kRcWide * (t - t_pad), n_small, gbuf, ff, ff_last);
    }
  }
}

// OWNERS, single thread: hand the next integrals of the row queue to this round's idle slots.
// Rows are taken in order; a row entering service gets a free context (its table is cleared by
// the group afterwards); if none is free the remaining idle slots wait for a later round.
__device__ __forceinline__ void rc_grant(RcGroup& sh, int n_idle, int n_rows, const int* __restrict__ nq_left,
                                         QagsCounters* __restrict__ ctr)
{
  int n_grant = 0, clear_mask = 0, given = 0;
  while (given < n_idle && n_grant < kRcGrant) {
    if (sh.cur_row >= sh.chunk_n) {  // next chunk of consecutive rows
      if (sh.exhausted) break;
      const unsigned long long q = atomicAdd(&ctr->next, 1ull);
      const long long r0 = (long long)q * kRcChunk;
      if (r0 >= n_rows) { sh.exhausted = 1; break; }
      sh.chunk_row0 = (int)r0;
      sh.chunk_n = (int)min((long long)kRcChunk, n_rows - r0);
      for (int j = 0; j < sh.chunk_n; ++j) sh.chunk_nq[j] = nq_left[r0 + j];  // integrals the head handed over
      sh.cur_row = 0; sh.cur_i = 0; sh.cur_ctx = -1;
    }
    const int nq = sh.chunk_nq[sh.cur_row];
    if (sh.cur_i >= nq) { ++sh.cur_row; sh.cur_i = 0; sh.cur_ctx = -1; continue; }
    if (sh.cur_ctx < 0) {  // the row enters service
      int c = -1;
      for (int j = 0; j < kRcCtx; ++j)
        if (sh.ctx_left[j] == 0) { c = j; break; }
      if (c < 0) break;
      sh.cur_ctx = c;
      sh.ctx_left[c] = nq;
      clear_mask |= 1 << c;
    }
    const int take = min(nq - sh.cur_i, n_idle - given);
    sh.grant_rank0[n_grant] = given;
    sh.grant_n[n_grant] = take;
    sh.grant_row[n_grant] = sh.chunk_row0 + sh.cur_row;
    sh.grant_i0[n_grant] = sh.cur_i;
    sh.grant_ctx[n_grant] = sh.cur_ctx;
    ++n_grant;
    given += take;
    sh.cur_i += take;
  }
  sh.n_grant = n_grant;
  sh.clear_mask = clear_mask;
}

// OWNERS: one group of 64 slots.  A slot whose integral is finished takes the next integral of the
// row queue in the following round, so the rounds stay full until the queue runs dry.
__device__ __forceinline__ void rc_owner(RcGroup& sh, const double* node, const double* snode, const int* sid, int grp, int ltid, int n_rows, int nb,
                                         const RowInfo* __restrict__ rows, const long long* __restrict__ item_off,
                                         const FluxConsts& fc, const SplineSeg* __restrict__ ff, double ff_last,
                                         double* __restrict__ W, int* __restrict__ neval_out,
                                         QagsCounters* __restrict__ ctr, long long* __restrict__ overflow_items,
                                         double* __restrict__ gbuf, const HeadState* __restrict__ head_state,
                                         const unsigned* __restrict__ state_slot, const int* __restrict__ left_idx,
                                         const int* __restrict__ nq_left)
{
  const unsigned lane = ltid & 31, warp = ltid >> 5;
  const unsigned lt = (1u << lane) - 1;
  const int bar_own = kBarOwn0 + grp, bar_full = kBarFull0 + grp, bar_done = kBarDone0 + grp;
  Qags<QagsSharedStore> S;
  S.sh = &sh;
  S.slot = ltid;
  double my_evals = 0;
  unsigned my_err = 0;
  int parity = 0;
  bool active = false;
  int my_ctx = 0, my_row = 0, my_i = 0;
  if (ltid == 0) {
    sh.n_new[0] = 0; sh.n_new[1] = 0; sh.fin = 0;
    sh.chunk_n = 0; sh.cur_row = 0; sh.cur_i = 0; sh.cur_ctx = -1; sh.exhausted = 0;
  }
  if (ltid < kRcCtx) sh.ctx_left[ltid] = 0;

  while (true) {
    // ---- R0: refill idle slots from the row queue ----
    const unsigned bal_idle = __ballot_sync(0xffffffffu, !active);
    if (lane == 0) sh.cnt[2][warp] = __popc(bal_idle);
    bar_sync(bar_own, kRcSlots);
    int n_idle = 0, base_idle = 0;
#pragma unroll
    for (int w = 0; w < kRcSlots / 32; ++w) {
      if (w < (int)warp) base_idle += sh.cnt[2][w];
      n_idle += sh.cnt[2][w];
    }
    if (n_idle > 0) {  // group-uniform
      if (ltid == 0) rc_grant(sh, n_idle, n_rows, nq_left, ctr);
      bar_sync(bar_own, kRcSlots);
      const int clear_mask = sh.clear_mask;
      if (clear_mask) {
        for (int e = ltid; e < kRcG; e += kRcSlots)
          if ((clear_mask >> (e / kRcGE)) & 1) sh.gkeys[e] = kRcEmpty;
      }
      if (!active) {
        const int rank = base_idle + __popc(bal_idle & lt);
        const int n_grant = sh.n_grant;
        for (int j = 0; j < n_grant; ++j) {
          const int r0 = sh.grant_rank0[j];
          if (rank >= r0 && rank < r0 + sh.grant_n[j]) {
            my_row = sh.grant_row[j];
            const int pos = sh.grant_i0[j] + (rank - r0);  // position among the row's handed-over integrals
            const long long item0 = item_off[my_row];
            my_i = left_idx[item0 + pos];
            my_ctx = sh.grant_ctx[j];
            const RowInfo ri = rows[my_row];
            double b, w;
            grid_point(ri, my_i, b, w);
            sh.k_cur[ltid] = ri.k;
            sh.bw_cur[ltid] = b * w;
            sh.beta[ltid] = b * (1. / kHc);
            sh.slot_ctx[ltid] = (unsigned char)my_ctx;
            if (pos == 0) sh.ctx_c0[my_ctx] = ri.k * ri.k / fc.g1 / fc.g1;  // w*w/g/g, :187
            // the QAGS state the head left (upc_qags_head.cuh)
            const HeadState& hs = head_state[state_slot[item0 + my_i]];  // pool slot of this integral
#pragma unroll
            for (int k = 0; k < 11; ++k) sh.sc[k][ltid] = hs.sc[k];
#pragma unroll
            for (int k = 0; k < kRcEps; ++k) sh.ep[k][ltid] = hs.eps[k];
#pragma unroll
            for (int k = 0; k < kRcCap; ++k) {
              sh.rl[k][ltid] = hs.rl[k]; sh.el[k][ltid] = hs.el[k];
              sh.hp[k][ltid] = hs.hp[k]; sh.od[k][ltid] = hs.od[k];
            }
            S.size = hs.size; S.nrmax = hs.nrmax; S.i = hs.i; S.maximum_level = hs.maximum_level; S.ktmin = hs.ktmin;
            S.roundoff_type1 = hs.roundoff_type1; S.roundoff_type2 = hs.roundoff_type2;
            S.roundoff_type3 = hs.roundoff_type3; S.error_type = hs.error_type; S.error_type2 = hs.error_type2;
            S.iteration = hs.iteration; S.tab_n = hs.tab_n; S.tab_nres = hs.tab_nres;
            S.positive_integrand = (hs.flags & kHdPositive) != 0;
            S.extrapolate = (hs.flags & kHdExtrapolate) != 0;
            S.disallow_extrapolation = (hs.flags & kHdDisallow) != 0;
            S.overflow = false;
            S.neval = hs.neval;
            S.result = 0; S.abserr = 0; S.ier = 0;
            active = true;
          }
        }
      }
      bar_sync(bar_own, kRcSlots);
    }

    // ---- R1 (thread = integral): pending intervals, cache entries, evaluation counts ----
    // z = beta * x is monotonic over the (ascending) nodes, so the evaluations of a task that take
    // J1's small-argument branch (z <= 8, j1_smallN) are its first n_small nodes: a bisection with
    // the very arithmetic of the evaluators
    bool has_b = false;
    int small_a = 0, small_b = 0;
    if (active) {
      int level;
      double a1, b1, a2, b2;
      S.pre_step(a1, b1, a2, b2, level);
      has_b = true;
      const unsigned heap = QagsSharedStore::heap_of(a1, level);
      const double beta = sh.beta[ltid];
      {
        const double c = 0.5 * (a1 + b1), h = 0.5 * (b1 - a1);
        sh.t_center[0][ltid] = c;
        sh.t_half[0][ltid] = h;
        sh.t_g[0][ltid] = (short)rc_entry(sh, parity, my_ctx, level, heap);
        int lo = 0, hi = 21;  // first node with z > 8
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (beta * fma(h, snode[mid], c) > 8.) hi = mid; else lo = mid + 1;
        }
        small_a = lo;
      }
      if (has_b) {
        const double c = 0.5 * (a2 + b2), h = 0.5 * (b2 - a2);
        sh.t_center[1][ltid] = c;
        sh.t_half[1][ltid] = h;
        sh.t_g[1][ltid] = (short)rc_entry(sh, parity, my_ctx, level, heap + 1);
        int lo = 0, hi = 21;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (beta * fma(h, snode[mid], c) > 8.) hi = mid; else lo = mid + 1;
        }
        small_b = lo;
      }
    }
    // packed (small << 16 | large) counts, inclusive warp scan
    const int n_mine = active ? (has_b ? 42 : 21) : 0;
    const int n_small = small_a + small_b;
    const int packed = (n_small << 16) | (n_mine - n_small);
    int incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += v;
    }
    const unsigned bal_a = __ballot_sync(0xffffffffu, active);
    if (lane == 31) { sh.cnt[0][warp] = __popc(bal_a); sh.cnt[1][warp] = incl; }
    bar_sync(bar_own, kRcSlots);

    // ---- R2: this round's evaluation list, fill of the new table entries, hand-over ----
    int n_a = 0, tot = 0, base = 0;
#pragma unroll
    for (int w = 0; w < kRcSlots / 32; ++w) {
      if (w < (int)warp) base += sh.cnt[1][w];
      n_a += sh.cnt[0][w];
      tot += sh.cnt[1][w];
    }
    if (n_a == 0) break;  // group-uniform: nothing in flight and nothing left to hand out
    {
      const int excl = base + incl - packed;  // (small << 16 | large) evaluations of the slots before this one
      sh.t_small[0][ltid] = active ? (unsigned char)small_a : 255;
      sh.t_off[0][ltid] = excl;
      sh.t_small[1][ltid] = has_b ? (unsigned char)small_b : 255;
      sh.t_off[1][ltid] = excl + ((small_a << 16) | (21 - small_a));
    }
    {
      const int n_fill = sh.n_new[parity] * 21;
      for (int e = ltid; e < n_fill; e += kRcSlots) {
        const int ent = sh.newlist[e / 21], n = e - (e / 21) * 21;
        const unsigned heap = sh.gkeys[ent];
        const int level = 31 - __clz(heap);
        const unsigned pos = heap - (1u << level);
        const double width = 10. * QagsSharedStore::pow2(-level);
        const double a = pos * width, b = (pos + 1.) * width;
        const double x = fma(0.5 * (b - a), node[n], 0.5 * (a + b));
        gbuf[ent * 21 + n] = rc_g(x, sh.ctx_c0[ent / kRcGE], ff, ff_last);
      }
    }
    if (ltid == 0) {
      sh.n_cls[0] = tot & 0xffff; sh.n_cls[1] = tot >> 16;
      sh.n_new[parity ^ 1] = 0;  // last read one round ago, next used one round ahead
    }
    parity ^= 1;
    __threadfence_block();
    bar_arrive(bar_full, kRcSlots + kRcEval);   // tasks published
    bar_sync(bar_done, kRcSlots + kRcEval);     // ... and evaluated

    // ---- R5 (thread = integral): QAGS bookkeeping on the GK21 sums the evaluators formed ----
    if (active) {
      bool done;
      const GkOut ga{sh.gk[0][0][ltid], sh.gk[1][0][ltid], sh.gk[2][0][ltid], sh.gk[3][0][ltid]};
      const GkOut gb{sh.gk[0][1][ltid], sh.gk[1][1][ltid], sh.gk[2][1][ltid], sh.gk[3][1][ltid]};
      done = S.post_step(ga, gb);
      if (done) {
        const double Q = S.result / fc.A;                         // :214
        const double flux = fc.factor * Q * Q / sh.k_cur[ltid];   // :215
        const size_t out_idx = (size_t)my_row * nb + my_i;
        W[out_idx] = flux * sh.bw_cur[ltid];
        if (neval_out) neval_out[out_idx] = S.neval;
        if (S.overflow) {
          const unsigned long long o = atomicAdd(&ctr->overflow, 1ull);
          overflow_items[o] = item_off[my_row] + my_i;
        } else {
          my_evals += S.neval;
          if (S.ier != 0) my_err++;
        }
        atomicSub(&sh.ctx_left[my_ctx], 1);  // the row's context is free after its last integral
        active = false;
      }
    }
  }
  // tell the evaluators that this group is finished
  if (ltid == 0) sh.fin = 1;
  __threadfence_block();
  bar_arrive(bar_full, kRcSlots + kRcEval);

  const double ev = warp_sum(my_evals);
  const unsigned er = __reduce_add_sync(0xffffffffu, my_err);
  if (lane == 0) {
    atomicAdd(&ctr->evals, (unsigned long long)ev);
    if (er) atomicAdd(&ctr->errors, (unsigned long long)er);
  }
}

__global__ void __launch_bounds__(kRcThreads, 1)
k_flux_qags_rows(int n_rows, int nb, const RowInfo* __restrict__ rows, const long long* __restrict__ item_off,
                 FluxConsts fc, DevTables tab, double* __restrict__ W, int* __restrict__ neval_out,
                 QagsCounters* __restrict__ ctr, long long* __restrict__ overflow_items, double* __restrict__ gbuf_all,
                 const HeadState* __restrict__ head_state, const unsigned* __restrict__ state_slot,
                 const int* __restrict__ left_idx, const int* __restrict__ nq_left)
{
  extern __shared__ __align__(16) unsigned char rc_smem[];
  RcShared& sh = *reinterpret_cast<RcShared*>(rc_smem);
  const int tid = threadIdx.x;
  const int wg = tid >> 7;
  for (int e = tid; e < H_END; e += kRcThreads) sh.hot[e] = upc_hot[e];
  if (tid < 21) {
    sh.node[tid] = kGkNode[tid];
    const double v = kGkNode[tid];  // rank of node tid in ascending order (the abscissas are distinct)
    int rank = 0;
    for (int j = 0; j < 21; ++j) rank += kGkNode[j] < v;
    sh.snode[rank] = v;
    sh.sid[rank] = tid;
  }
  __syncthreads();
  double* const gbuf = gbuf_all + (size_t)blockIdx.x * (kRcGroups * kRcG * 21);

  if (wg < 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRcRegsOwner));
    const int grp = tid / kRcSlots;
    rc_owner(sh.g[grp], sh.node, sh.snode, sh.sid, grp, tid - grp * kRcSlots, n_rows, nb, rows, item_off, fc, tab.ff_seg, tab.ff_last, W,
             neval_out, ctr, overflow_items, gbuf + (size_t)grp * (kRcG * 21), head_state, state_slot, left_idx, nq_left);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRcRegsEval));
    const int etid = tid - 256;
    unsigned alive = (1u << kRcGroups) - 1;
    int g = 0;
    while (alive) {
      if ((alive >> g) & 1) {
        bar_sync(kBarFull0 + g, kRcSlots + kRcEval);
        if (sh.g[g].fin) {
          alive &= ~(1u << g);
        } else {
          RcGroup& G = sh.g[g];
          const double* gb = gbuf + (size_t)g * (kRcG * 21);
          // two evaluator threads per task (64 slots x 2 halves x 2 = 256)
          const int role = etid & 1, task = etid >> 1;
          const int half = task >= kRcSlots, slot = task - half * kRcSlots;
          const int n_small = task < 2 * kRcSlots ? G.t_small[half][slot] : 255;
          // expand the tasks into the class-sorted evaluation list: ascending nodes 0-10 / 11-20
          if (n_small != 255) {
            const int off = G.t_off[half][slot];
            const unsigned code = slot | (half << 6);
            const int j0 = role ? 11 : 0, j1 = role ? 21 : 11;
            // node j of the task is its j-th small evaluation (j < n_small) or its (j - n_small)-th large one
            const int base_s = kRcSlots * 42 - 1 - (off >> 16), base_l = (off & 0xffff) - n_small;
            for (int j = j0; j < j1; ++j) {
              const unsigned short c = (unsigned short)(code | (sh.sid[j] << 7));
              if (j < n_small) sh.elist[base_s - j] = c; else sh.elist[base_l + j] = c;
            }
          }
          bar_sync(kBarEval, kRcEval);
          rc_eval_round(G, sh.fv, sh.elist, sh.node, HotShared{sh.hot}, etid, G.n_cls[0], G.n_cls[1], gb, tab.ff_seg, tab.ff_last);
          bar_sync(kBarEval, kRcEval);
          // GK21 sums in GSL's order, split over the two threads of the task
          {
            GkOut o{};
            double asc = 0;
            const bool live = n_small != 255;
            if (live) gk21_sums_pair(&sh.fv[half][slot][0], G.t_half[half][slot], role, o, asc);
            asc = __shfl_xor_sync(0xffffffffu, asc, 1);
            if (live && role == 0) {
              gk21_finish(o, asc, G.t_half[half][slot]);
              G.gk[0][half][slot] = o.result; G.gk[1][half][slot] = o.abserr;
              G.gk[2][half][slot] = o.resabs; G.gk[3][half][slot] = o.resasc;
            }
          }
          __threadfence_block();
          bar_arrive(kBarDone0 + g, kRcSlots + kRcEval);
        }
      }
      g = g + 1 == kRcGroups ? 0 : g + 1;
    }
  }
}

}  // namespace upc
