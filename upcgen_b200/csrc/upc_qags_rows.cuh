// upc_qags_rows.cuh -- stage A.3, the form-factor flux rows (F2/F3: fluxFormIntegrand / fluxForm,
// src/UpcCrossSection.cpp:181-218) as a row-cooperative, warp-specialised persistent kernel.
//
// The reference's integrand is   x^2 F(x^2 + (k/gamma)^2) / (x^2 + (k/gamma)^2) * J1(b x / hc):
// the first factor ("g") depends on the ROW (photon energy k) only, the Bessel factor on the
// integral (b).  Every integral starts on [0, 10] and QAGS bisects, so all intervals are dyadic
// and the ~30-100 integrals of a row visit the same ~20-30 of them (measured: 1100 GK21 rules per
// row on 29 distinct intervals).  One CTA per SM, 512 threads in four warpgroups:
//   * OWNERS (warpgroups 0 and 1, 96 integrals each, one integral per thread): the QAGS
//     bookkeeping, which follows gsl_integration_qags decision for decision (upc_qags.cuh).
//     Interval lists, epsilon tables and the FP64 driver scalars live in SHARED memory, strided by
//     slot (bank = slot whatever entry a lane touches): the bookkeeping is a chain of dependent
//     look-ups that was latency-bound on L2 while that state sat in local memory.  Owners publish
//     the 1-2 pending intervals of each integral as tasks and form the GK21 sums in GSL's order.
//   * EVALUATORS (warpgroups 2 and 3): the integrand evaluations of a round (21 per task),
//     flattened over 256 threads whatever integral they belong to, three nodes per thread and
//     trip.  They alternate between the two owner groups, so that the bookkeeping of one group
//     runs under the evaluations of the other and the FP64 pipe always has dense, convergent J1
//     work.  Register budgets follow the roles (setmaxnreg: 88 / 168).
//   * g on the 21 GK nodes of an interval is evaluated ONCE per (row, interval) -- the 10^6-knot
//     form-factor spline gather and the division leave the hot loop -- into an L2-resident
//     per-group table keyed by the interval's heap index (level, position).
// The value g * J1 is formed exactly as the reference's  x*x*F/t * J1  (left to right), so the
// QAGS decisions and results are the ones of the per-thread path (k_flux_list) bit for bit.
#pragma once
#include "upc_hot.cuh"
#include "upc_qags.cuh"

namespace upc {

constexpr int kRcThreads = 512;
constexpr int kRcGroups = 2;     // owner groups (one warpgroup each; 3 of its 4 warps own slots)
constexpr int kRcSlots = 96;     // integrals in flight per group (multiple of 32: conflict-free strides)
constexpr int kRcEval = 256;     // evaluator threads
constexpr int kRcCap = 16;       // interval-list capacity per integral (largest seen: 15)
constexpr int kRcEps = 16;       // epsilon-table capacity (largest index touched so far: 12)
constexpr int kRcCtx = 8;        // rows in flight per group
constexpr int kRcChunk = 4;      // rows per pop of the global queue
constexpr int kRcGrant = 4;      // row segments handed to idle slots per round
constexpr int kRcGE = 64;        // cached intervals per row in flight
constexpr int kRcG = kRcCtx * kRcGE;  // cached (row, interval) entries per group
constexpr int kRcMaxLevel = 26;  // deeper intervals are evaluated uncached
constexpr unsigned kRcEmpty = 0xffffffffu;
constexpr int kRcRegsOwner = 88, kRcRegsEval = 168;  // 256 * 88 + 256 * 168 = 64 K registers

// state of one owner group
struct RcGroup {
  double fv[2][21][kRcSlots];          // integrand values [half][node][slot]
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
  unsigned char list[2][kRcSlots];     // compacted tasks of the round
  unsigned short newlist[kRcG];        // entries inserted this round (to be filled)
  int n_new[2];                        // by round parity
  int cnt[3][kRcSlots / 32];           // per-warp counts of the round: A tasks, B tasks, idle slots
  int n_task[2];                       // tasks handed to the evaluators (A, B)
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
  double node[24];                     // signed GK21 abscissas (kGkNode)
};

// named barriers (0 is __syncthreads)
enum { kBarOwn0 = 1, kBarFull0 = 3, kBarDone0 = 5 };
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

// GK21 sums from the 21 stored integrand values (node order of kGkNode), in GSL's order
// (integration/qk.c); same arithmetic as the tail of gk21_tri.
__device__ __forceinline__ GkOut gk21_sums(const double* fv, int fv_stride, double half_length)
{
  const double abs_half_length = fabs(half_length);
  const double f_center = fv[20 * fv_stride];
  double result_gauss = 0;
  double result_kronrod = f_center * kGkWkC;
  double result_abs = fabs(result_kronrod);
#pragma unroll 1
  for (int p = 0; p < 10; ++p) {
    const double fval1 = fv[(2 * p) * fv_stride], fval2 = fv[(2 * p + 1) * fv_stride];
    const double fsum = fval1 + fval2;
    result_gauss += kGkWg[p] * fsum;
    result_kronrod += kGkWk[p] * fsum;
    result_abs += kGkWk[p] * (fabs(fval1) + fabs(fval2));
  }
  const double mean = result_kronrod * 0.5;
  double result_asc = kGkWkC * fabs(f_center - mean);
#pragma unroll 1
  for (int j = 0; j < 10; ++j) {
    const int p = (j & 1) ? (j >> 1) : (5 + (j >> 1));
    result_asc += kGkWk[p] * (fabs(fv[(2 * p) * fv_stride] - mean) + fabs(fv[(2 * p + 1) * fv_stride] - mean));
  }
  double err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double s = 200 * err / result_asc;
    double scale = s * sqrt(s);
    err = scale < 1 ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  GkOut o;
  o.result = result_kronrod;
  o.abserr = err;
  o.resabs = result_abs;
  o.resasc = result_asc;
  return o;
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
// evaluator threads: trip e -> (task, 3 consecutive nodes); A tasks first, then B tasks
__device__ __forceinline__ void rc_eval(RcGroup& sh, const double* node, int etid, const double* __restrict__ gbuf,
                                        const SplineSeg* __restrict__ ff, double ff_last)
{
  const int n_a = sh.n_task[0], n_t = n_a + sh.n_task[1];
  const int n_trips = 7 * n_t;
#pragma unroll 1
  for (int e = etid; e < n_trips; e += kRcEval) {
    const int q = e / n_t, rank = e - q * n_t;
    const int half = rank >= n_a;
    const int slot = sh.list[half][rank - (half ? n_a : 0)];
    const int ent = sh.t_g[half][slot];
    const int n = 3 * q;
    double g0, g1, g2;
    if (ent >= 0) {  // issued first: the L2 latency is covered by the J1 arithmetic below
      const double* gv = gbuf + ent * 21 + n;
      g0 = __ldcg(gv); g1 = __ldcg(gv + 1); g2 = __ldcg(gv + 2);
    }
    const double center = sh.t_center[half][slot], hl = sh.t_half[half][slot];
    const double beta = sh.beta[slot];
    const double x0 = fma(hl, node[n], center), x1 = fma(hl, node[n + 1], center), x2 = fma(hl, node[n + 2], center);
    if (ent < 0) {
      const double c0 = sh.ctx_c0[sh.slot_ctx[slot]];
      g0 = rc_g(x0, c0, ff, ff_last); g1 = rc_g(x1, c0, ff, ff_last); g2 = rc_g(x2, c0, ff, ff_last);
    }
    const D3 j = j1_3(D3{beta * x0, beta * x1, beta * x2});
    double* fo = &sh.fv[half][n][slot];
    fo[0] = g0 * j.a;
    fo[kRcSlots] = g1 * j.b;
    fo[2 * kRcSlots] = g2 * j.c;
  }
}

// OWNERS, single thread: hand the next integrals of the row queue to this round's idle slots.
// Rows are taken in order; a row entering service gets a free context (its table is cleared by
// the group afterwards); if none is free the remaining idle slots wait for a later round.
__device__ __forceinline__ void rc_grant(RcGroup& sh, int n_idle, int n_rows, const RowInfo* __restrict__ rows,
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
      for (int j = 0; j < sh.chunk_n; ++j) sh.chunk_nq[j] = rows[r0 + j].nq;
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

// OWNERS: one group of 96 slots.  A slot whose integral is finished takes the next integral of the
// row queue in the following round, so the rounds stay full until the queue runs dry.
__device__ __forceinline__ void rc_owner(RcGroup& sh, const double* node, int grp, int ltid, int n_rows, int nb,
                                         const RowInfo* __restrict__ rows, const long long* __restrict__ item_off,
                                         const FluxConsts& fc, const SplineSeg* __restrict__ ff, double ff_last,
                                         double* __restrict__ W, int* __restrict__ neval_out,
                                         QagsCounters* __restrict__ ctr, long long* __restrict__ overflow_items,
                                         double* __restrict__ gbuf)
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
  bool active = false, first = false;
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
      if (ltid == 0) rc_grant(sh, n_idle, n_rows, rows, ctr);
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
            my_i = sh.grant_i0[j] + (rank - r0);
            my_ctx = sh.grant_ctx[j];
            const RowInfo ri = rows[my_row];
            double b, w;
            grid_point(ri, my_i, b, w);
            sh.k_cur[ltid] = ri.k;
            sh.bw_cur[ltid] = b * w;
            sh.beta[ltid] = b * (1. / kHc);
            sh.slot_ctx[ltid] = (unsigned char)my_ctx;
            if (my_i == 0) sh.ctx_c0[my_ctx] = ri.k * ri.k / fc.g1 / fc.g1;  // w*w/g/g, :187
            S.begin(0., 10.);                                              // :209
            active = true;
            first = true;
          }
        }
      }
      bar_sync(bar_own, kRcSlots);
    }

    // ---- R1 (thread = integral): pending intervals, cache entries, task counts ----
    bool has_b = false;
    if (active) {
      int level = 0;
      double a1 = 0., b1 = 10., a2 = 10., b2 = 10.;
      if (!first) {
        S.pre_step(a1, b1, a2, b2, level);
        has_b = true;
      }
      const unsigned heap = QagsSharedStore::heap_of(a1, level);
      sh.t_center[0][ltid] = 0.5 * (a1 + b1);
      sh.t_half[0][ltid] = 0.5 * (b1 - a1);
      sh.t_g[0][ltid] = (short)rc_entry(sh, parity, my_ctx, level, heap);
      if (has_b) {
        sh.t_center[1][ltid] = 0.5 * (a2 + b2);
        sh.t_half[1][ltid] = 0.5 * (b2 - a2);
        sh.t_g[1][ltid] = (short)rc_entry(sh, parity, my_ctx, level, heap + 1);
      }
    }
    const unsigned bal_a = __ballot_sync(0xffffffffu, active);
    const unsigned bal_b = __ballot_sync(0xffffffffu, has_b);
    if (lane == 0) { sh.cnt[0][warp] = __popc(bal_a); sh.cnt[1][warp] = __popc(bal_b); }
    bar_sync(bar_own, kRcSlots);

    // ---- R2: compacted task lists (slot order), fill of the new table entries, hand-over ----
    int n_a = 0, n_b = 0, base_a = 0, base_b = 0;
#pragma unroll
    for (int w = 0; w < kRcSlots / 32; ++w) {
      if (w < (int)warp) { base_a += sh.cnt[0][w]; base_b += sh.cnt[1][w]; }
      n_a += sh.cnt[0][w];
      n_b += sh.cnt[1][w];
    }
    if (n_a == 0) break;  // group-uniform: nothing in flight and nothing left to hand out
    if (active) sh.list[0][base_a + __popc(bal_a & lt)] = (unsigned char)ltid;
    if (has_b) sh.list[1][base_b + __popc(bal_b & lt)] = (unsigned char)ltid;
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
      sh.n_task[0] = n_a; sh.n_task[1] = n_b;
      sh.n_new[parity ^ 1] = 0;  // last read one round ago, next used one round ahead
    }
    parity ^= 1;
    __threadfence_block();
    bar_arrive(bar_full, kRcSlots + kRcEval);   // tasks published
    bar_sync(bar_done, kRcSlots + kRcEval);     // ... and evaluated

    // ---- R4/R5 (thread = integral): GK21 sums in GSL's order, QAGS bookkeeping ----
    if (active) {
      bool done;
      const GkOut ga = gk21_sums(&sh.fv[0][0][ltid], kRcSlots, sh.t_half[0][ltid]);
      if (first) {
        done = S.post_first(ga);
        first = false;
      } else {
        const GkOut gb = gk21_sums(&sh.fv[1][0][ltid], kRcSlots, sh.t_half[1][ltid]);
        done = S.post_step(ga, gb);
      }
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
                 QagsCounters* __restrict__ ctr, long long* __restrict__ overflow_items, double* __restrict__ gbuf_all)
{
  extern __shared__ __align__(16) unsigned char rc_smem[];
  RcShared& sh = *reinterpret_cast<RcShared*>(rc_smem);
  const int tid = threadIdx.x;
  const int wg = tid >> 7;
  if (tid < 21) sh.node[tid] = kGkNode[tid];
  __syncthreads();
  double* const gbuf = gbuf_all + (size_t)blockIdx.x * (kRcGroups * kRcG * 21);

  if (wg < kRcGroups) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRcRegsOwner));
    const int ltid = tid & 127;
    if (ltid >= kRcSlots) return;  // the 4th warp of an owner warpgroup owns no slots
    rc_owner(sh.g[wg], sh.node, wg, ltid, n_rows, nb, rows, item_off, fc, tab.ff_seg, tab.ff_last, W, neval_out, ctr,
             overflow_items, gbuf + (size_t)wg * (kRcG * 21));
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRcRegsEval));
    const int etid = tid - kRcGroups * 128;
    bool alive[kRcGroups] = {true, true};
    int g = 0;
    while (alive[0] || alive[1]) {
      if (alive[g]) {
        bar_sync(kBarFull0 + g, kRcSlots + kRcEval);
        if (sh.g[g].fin) {
          alive[g] = false;
        } else {
          rc_eval(sh.g[g], sh.node, etid, gbuf + (size_t)g * (kRcG * 21), tab.ff_seg, tab.ff_last);
          __threadfence_block();
          bar_arrive(kBarDone0 + g, kRcSlots + kRcEval);
        }
      }
      g ^= 1;
    }
  }
}

}  // namespace upc
