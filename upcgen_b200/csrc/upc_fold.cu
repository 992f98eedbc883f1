// upc_fold.cu -- sigma fold (X1) and the inverse-CDF samplers (S1-S3).
// Reference: src/UpcCrossSection.cpp:594-698 (calcNucCrossSectionYM) and include/UpcSampler.h
// (GSL gsl_histogram[2d]_pdf_init / _pdf_sample, getBinX/getBinY).
//
// Everything that decides an integer bin is evaluated with explicit non-fused FP64 operations
// (__dadd_rn/__dmul_rn/__ddiv_rn) in the reference's operation order, so that the cumulative
// tables and the selected bins are bit-identical to a generic x86-64 build of the reference.
#include <cstdio>

#include <cub/cub.cuh>

#include "upc_ctx.h"
#include "upc_internal.h"
#include "upc_sampler.cuh"

namespace upc {

// cs[iy][im] = sigma(m_im) * lumi[im][iy]  (transposed), :643-659
__global__ void k_fold(int nm, int ny, int pol, const double* __restrict__ lumi, const double* __restrict__ lumi_s,
                       const double* __restrict__ lumi_p, const double* __restrict__ sig_m,
                       const double* __restrict__ sig_s, const double* __restrict__ sig_p, double* __restrict__ cs,
                       double* __restrict__ ratio)
{
  // 32x32 tile transpose through shared memory: coalesced reads along iy, writes along im
  __shared__ double tile[32][33];
  __shared__ double tile2[32][33];
  const int im0 = blockIdx.x * 32, iy0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int im = im0 + r, iy = iy0 + threadIdx.x;
    if (im < nm && iy < ny) {
      const size_t src = (size_t)im * ny + iy;
      if (!pol) {
        tile[r][threadIdx.x] = __dmul_rn(sig_m[im], lumi[src]);  // :648
      } else {
        const double nuccs_s = __dmul_rn(lumi_s[src], sig_s[im]);  // :654
        const double nuccs_p = __dmul_rn(lumi_p[src], sig_p[im]);  // :655
        tile[r][threadIdx.x] = __dmul_rn(__dadd_rn(nuccs_s, nuccs_p), 1e7);  // :656-657
        tile2[r][threadIdx.x] = __ddiv_rn(nuccs_s, nuccs_p);                 // :658
      }
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int iy = iy0 + r, im = im0 + threadIdx.x;
    if (im < nm && iy < ny) {
      const size_t dst = (size_t)iy * nm + im;
      cs[dst] = tile[threadIdx.x][r];
      if (pol) ratio[dst] = tile2[threadIdx.x][r];
    }
  }
}

// totCS: fixed-shape pairwise tree, independent of grid/launch geometry and of the number of
// GPUs (the reference's own sum order depends on OpenMP scheduling).
__global__ void k_sum_blocks(const double* __restrict__ x, size_t n, double* __restrict__ partial)
{
  __shared__ double sm[256];
  const size_t base = (size_t)blockIdx.x * 4096;
  double acc = 0;
  // each thread sums 16 strided elements in a fixed order
  for (int r = 0; r < 16; r++) {
    size_t i = base + (size_t)r * 256 + threadIdx.x;
    if (i < n) acc += x[i];
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

int fold_sigma(upcgpu_ctx* c, const double* sig_m, const double* sig_s, const double* sig_p, double* cs, double* ratio,
               double* totcs_mb)
{
  const upcgpu_params& p = c->p;
  const size_t n = (size_t)p.nm * p.ny;
  if (!c->lumi_ready) { c->err = "fold_sigma: lumi table not filled (or not gathered)"; return UPCGPU_EINVAL; }
  if (p.use_pol ? (!sig_s || !sig_p) : !sig_m) { c->err = "fold_sigma: missing sigma array"; return UPCGPU_EINVAL; }
  cudaStream_t st = c->stream;
  if (!c->cs) UPC_CUDA(c, cudaMalloc(&c->cs, n * sizeof(double)));
  if (p.use_pol && !c->ratio) UPC_CUDA(c, cudaMalloc(&c->ratio, n * sizeof(double)));
  const size_t nb0 = (n + 4095) / 4096, nb1 = (nb0 + 4095) / 4096;
  if (!c->fold_ws) UPC_CUDA(c, cudaMalloc(&c->fold_ws, (3 * (size_t)p.nm + nb0 + nb1) * sizeof(double)));
  double* dsig = c->fold_ws;
  if (sig_m) UPC_CUDA(c, cudaMemcpyAsync(dsig, sig_m, p.nm * sizeof(double), cudaMemcpyHostToDevice, st));
  if (sig_s) UPC_CUDA(c, cudaMemcpyAsync(dsig + p.nm, sig_s, p.nm * sizeof(double), cudaMemcpyHostToDevice, st));
  if (sig_p) UPC_CUDA(c, cudaMemcpyAsync(dsig + 2 * p.nm, sig_p, p.nm * sizeof(double), cudaMemcpyHostToDevice, st));
  dim3 grid((p.nm + 31) / 32, (p.ny + 31) / 32), block(32, 8);
  UPC_K(c), k_fold<<<grid, block, 0, st>>>(p.nm, p.ny, p.use_pol, c->lumi[0], c->lumi[1], c->lumi[2], dsig, dsig + p.nm,
                                 dsig + 2 * p.nm, c->cs, c->ratio);
  // totCS
  double total = 0;
  {
    size_t cur = n;
    double *a = c->cs, *b0 = c->fold_ws + 3 * (size_t)p.nm, *b1 = b0 + nb0;
    double* outb = b0;
    while (true) {
      size_t nblk = (cur + 4095) / 4096;
      UPC_K(c), k_sum_blocks<<<(unsigned)nblk, 256, 0, st>>>(a, cur, outb);
      if (nblk == 1) break;
      a = outb;
      outb = (outb == b0) ? b1 : b0;
      cur = nblk;
    }
    UPC_CUDA(c, cudaMemcpyAsync(&total, outb, sizeof(double), cudaMemcpyDeviceToHost, st));
    UPC_CUDA(c, cudaStreamSynchronize(st));  // the one host wait of a sharded step: a queued fill is collected here
  }
  UPC_CUDA(c, cudaGetLastError());
  if (c->fill_pending) {
    const int frc = finish_fill(c);
    if (frc) return frc;
  }
  if (c->tables_pending) {  // a table stage queued without a fill behind it
    const int trc = finish_tables(c);
    if (trc) return trc;
  }
  if (totcs_mb) *totcs_mb = total * 1e-6;  // :696
  if (cs) UPC_CUDA(c, cudaMemcpy(cs, c->cs, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (ratio && p.use_pol) UPC_CUDA(c, cudaMemcpy(ratio, c->ratio, n * sizeof(double), cudaMemcpyDeviceToHost));
  c->fold_ready = true;
  return UPCGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// S1: gsl_histogram2d_pdf_init.  The running mean and the cumulative sum are sequential
// recurrences whose rounding the bin selection depends on: they are walked in the reference's order by
// ONE lane, while the other lanes of its warp keep it fed -- bins staged in shared memory by coalesced
// loads and the reciprocals 1/(i+1) formed in parallel.  (The per-bin divisions bin/mean/n have no
// dependency and run in parallel.)
//
// x / d for an INTEGER d = i + 1 < 2^40 with y = RN(1/d) given: q0 = RN(x y) is within 2 ulp of x/d; r = x - d q0 is
// exact in the FMA (a multiple of ulp(q0)/2 below 2^42 ulp); q1 = RN(q0 + r y) = RN(x/d + e) with |e| <= 2 ulp 2^-53.
// x/d cannot lie that close to a rounding boundary m = (k + 1/2) ulp: x - d m is a non-zero multiple of ulp(q)/2
// (non-zero because d (2k+1) spans 54 bits or more and x has 53), so |x/d - m| >= ulp/(2 d) >= 2^-41 ulp >> |e|.
// Hence q1 is the correctly rounded quotient -- the value __ddiv_rn returns -- on a chain of three operations.  (Normal
// range only: among subnormals exact ties exist.  The caller takes the library division while the running mean is
// below kMeanSafe: with bins >= 0 and mean >= kMeanSafe, x - mean is 0 or at least one ulp of the mean in size,
// far above the subnormals, and step i lowers the mean by the factor 1 - 1/(i + 1) at most.)
__device__ __forceinline__ double div_by_known(double x, double d, double y)
{
  const double q = __dmul_rn(x, y);
  return __fma_rn(__fma_rn(-q, d, x), y, q);
}

constexpr int kSeqChunk = 1024;
constexpr double kMeanSafe = 0x1p-900;

// tools/micro/seq_chain.cu measures the pieces on B200: a dependent DFMA is 8.2 cycles, this five-operation recurrence
// (sub, mul, fma, fma, add) runs at 53 cycles per bin, with __ddiv_rn in it at 137.
__global__ void __launch_bounds__(32) k_running_mean(const double* __restrict__ bin, size_t n, double* __restrict__ mean_out)
{
  __shared__ double sb[2][kSeqChunk], sr[2][kSeqChunk];
  __shared__ int sneg[2];  // a negative bin in the chunk (GSL refuses such a histogram): library division throughout
  const int lane = threadIdx.x;
  double mean = 0;
  int buf = 0;
  if (lane < 2) sneg[lane] = 0;
  __syncwarp();
  {
    bool neg = false;
    for (int j = lane; j < kSeqChunk && (size_t)j < n; j += 32) {
      const double v = bin[j];
      neg |= v < 0.;
      sb[0][j] = v;
      sr[0][j] = __drcp_rn((double)(j + 1));
    }
    if (neg) sneg[0] = 1;
  }
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kSeqChunk, buf ^= 1) {
    const size_t nxt = i0 + kSeqChunk;
    if (lane == 0) {
      const int m = (int)min((size_t)kSeqChunk, n - i0);
      const double* b = sb[buf];
      const double* r = sr[buf];
      // mean += (bin[i] - mean) / (i + 1).  The fast division needs mean >= kMeanSafe throughout: decided once per
      // chunk (from the second chunk on a chunk lowers the mean by a factor 2 at most; the first one starts at 0)
      if (sneg[buf] != 0 || !(mean >= 2. * kMeanSafe) || i0 == 0) {
#pragma unroll 1
        for (int j = 0; j < m; ++j) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[j], mean), (double)(i0 + j + 1)));
      } else {
        int j = 0;
        for (; j + 8 <= m; j += 8) {  // eight bins per trip
          double bv[8], rv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { bv[k] = b[j + k]; rv[k] = r[j + k]; }
          const double base = (double)(i0 + j);  // one conversion per trip; base + k is exact
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mean = __dadd_rn(mean, div_by_known(__dsub_rn(bv[k], mean), base + (double)(k + 1), rv[k]));
        }
        for (; j < m; ++j) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[j], mean), (double)(i0 + j + 1)));
      }
      sneg[buf] = 0;
    } else {
      // lanes 1..31 stage the next chunk meanwhile
      bool neg = false;
      for (size_t j = nxt + (lane - 1); j < nxt + kSeqChunk && j < n; j += 31) {
        const double v = bin[j];
        neg |= v < 0.;
        sb[buf ^ 1][j - nxt] = v;
        sr[buf ^ 1][j - nxt] = __drcp_rn((double)(j + 1));
      }
      if (neg) sneg[buf ^ 1] = 1;
    }
    __syncwarp();
  }
  if (lane == 0) mean_out[0] = mean;
}

__global__ void k_pdf_terms(const double* __restrict__ bin, size_t n, const double* __restrict__ mean,
                            double* __restrict__ term)
{
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) term[i] = __ddiv_rn(__ddiv_rn(bin[i], mean[0]), (double)n);  // (bin/mean)/n
}

__global__ void __launch_bounds__(32) k_seq_cumsum(const double* __restrict__ term, size_t n, double* __restrict__ sum)
{
  __shared__ double st[2][kSeqChunk], so[kSeqChunk];
  const int lane = threadIdx.x;
  double s = 0;
  int buf = 0;
  if (lane == 0) sum[0] = 0;
  for (int j = lane; j < kSeqChunk && (size_t)j < n; j += 32) st[0][j] = term[j];
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kSeqChunk, buf ^= 1) {
    const size_t nxt = i0 + kSeqChunk;
    const int m = (int)min((size_t)kSeqChunk, n - i0);
    if (lane == 0) {
      const double* t = st[buf];
      // sum[k + 1] = sum[k] + term[k]: sixteen terms into registers, sixteen chained additions, sixteen stores
      int j = 0;
      for (; j + 16 <= m; j += 16) {
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = t[j + k];
#pragma unroll
        for (int k = 0; k < 16; ++k) { s = __dadd_rn(s, v[k]); v[k] = s; }
#pragma unroll
        for (int k = 0; k < 16; ++k) so[j + k] = v[k];
      }
      for (; j < m; ++j) {
        s = __dadd_rn(s, t[j]);
        so[j] = s;
      }
    } else {
      for (size_t j = nxt + (lane - 1); j < nxt + kSeqChunk && j < n; j += 31) st[buf ^ 1][j - nxt] = term[j];
    }
    __syncwarp();
    for (int j = lane; j < m; j += 32) sum[i0 + j + 1] = so[j];
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// The two recurrences of gsl_histogram2d_pdf_init beyond their first kSeqHead elements: speculate and verify.
//
// v[k] = step(v[k-1], x[k]) must be reproduced bit for bit, and it is a dependent chain.  But deep into a long table the
// INCREMENT of a step hardly depends on v: for the running mean it is rint((x - v) / (n ulp)) ulps, which changes by one
// when v moves by n ulps; for the cumulative sum it is rint(t / ulp) ulps whatever v is (ties and a change of binade
// aside).  So a block of B elements is done like this:
//   candidates c[j] (all v_start to begin with);
//   round: every element takes the TRUE step from its predecessor's candidate, nxt[j] = step(c[j-1], x[j]), in parallel;
//          if nxt[j] == c[j] for every j the candidates ARE the sequence (induction over j: c[-1] = v_start is exact, and
//          each c[j] is the exact step from an exact predecessor) -- done;
//          otherwise the steps' increments, as integers in ulps of v_start, are prefix-summed and give new candidates.
// Nothing is assumed about how good the candidates are: exactness rests on the verification alone, which uses the very
// operations of the sequential code.  When a block does not verify within kSpecRounds rounds (a change of binade, increments
// that are not whole ulps, values near zero) its verified PREFIX is kept -- everything before the first mismatch is exact by
// the same induction --, one thread walks kSpecSeqRun elements sequentially, and speculation resumes behind them with a
// short block.  On smooth tables two or three rounds verify; the block length grows with the element index (the running mean's increments are insensitive to v only when
// B << n).  cfg4's 12 M bins: 0.43 s of chain -> see DESIGN.md.
constexpr int kSpecThreads = 1024;
constexpr int kSpecItemsMax = 8;
constexpr int kSpecRounds = 8;
constexpr int kSpecSeqRun = 128;  // elements walked sequentially past a spot where the candidates did not verify
constexpr size_t kSeqHead = 4096;
// a thread walks `items` consecutive elements of the block, so neighbouring lanes are `items` (up to 8) doubles apart in shared
// memory: one padding double per eight keeps them on different banks
#define SPX(j) ((j) + ((j) >> 3))
constexpr int kSpecPadded = kSpecThreads * kSpecItemsMax + kSpecThreads * kSpecItemsMax / 8 + 8;  // elements left to the sequential kernels (k_running_mean, k_seq_cumsum)

struct StepMean {  // mean += (bin - mean) / (i + 1)
  static __device__ __forceinline__ double step(double v, double x, double n) { return __dadd_rn(v, __ddiv_rn(__dsub_rn(x, v), n)); }
  // first candidates: the recurrence in real arithmetic, v_j = (i0 v_start + x_0 + .. + x_j) / (i0 + j + 1).  (A step moves
  // the mean by ~(x - v) / (n ulp) ulps -- 1e10 on a typical table -- so "all v_start" would be 1e14 ulps off at the end of
  // a block and the rounds, which contract an error by ~B / n each, would need nine of them; from this guess, two.)
  static constexpr bool kLinearGuess = true;
  // the step with the division by the known count done on its reciprocal (div_by_known): equal to step() when v >= kMeanSafe
  // and x >= 0 -- the caller's "bad" flag covers both
  static __device__ __forceinline__ double step_fast(double v, double x, double n, double rn) { return __dadd_rn(v, div_by_known(__dsub_rn(x, v), n, rn)); }
  static constexpr bool kHasFast = true;
  // a sequential run by one thread; fast_ok: v >= 2 kMeanSafe and no negative bin in the run (see div_by_known)
  static __device__ __forceinline__ double run(double v, const double* sx, double* cand, int j0, int j1, size_t i0, bool fast_ok)
  {
    if (fast_ok) {
      for (int j = j0; j < j1; ++j) {
        const double nn = (double)(i0 + j + 1);
        v = __dadd_rn(v, div_by_known(__dsub_rn(sx[SPX(j)], v), nn, __drcp_rn(nn)));
        cand[SPX(j)] = v;
      }
    } else {
      for (int j = j0; j < j1; ++j) { v = step(v, sx[SPX(j)], (double)(i0 + j + 1)); cand[SPX(j)] = v; }
    }
    return v;
  }
};
struct StepSum {  // sum += term
  static __device__ __forceinline__ double step(double v, double x, double) { return __dadd_rn(v, x); }
  static constexpr bool kLinearGuess = false;  // the increment of a step does not depend on v: "all v_start" is one round off
  static __device__ __forceinline__ double step_fast(double v, double x, double, double) { return __dadd_rn(v, x); }
  static constexpr bool kHasFast = false;
  static __device__ __forceinline__ double run(double v, const double* sx, double* cand, int j0, int j1, size_t, bool)
  {
    for (int j = j0; j < j1; ++j) { v = __dadd_rn(v, sx[SPX(j)]); cand[SPX(j)] = v; }
    return v;
  }
};

struct SpecStats {
  unsigned long long blocks, fallback_blocks, rounds;
};

template <class STEP, bool WRITE_ALL>
__global__ void __launch_bounds__(kSpecThreads) k_seq_spec(const double* __restrict__ x, size_t n, size_t start, double* __restrict__ out,
                                                          SpecStats* __restrict__ stats)
{
  // out: WRITE_ALL ? out[k + 1] = v after element k (out[start] holds v_start) : out[0] = v, read and written
  extern __shared__ __align__(16) unsigned char spec_smem[];
  double* sx = reinterpret_cast<double*>(spec_smem);            // [B]
  double* cand = sx + kSpecPadded;                               // [B]
  typedef cub::BlockScan<long long, kSpecThreads> Scan;
  typedef cub::BlockScan<double, kSpecThreads> ScanD;
  __shared__ union {
    typename Scan::TempStorage i;
    typename ScanD::TempStorage d;
  } scan_tmp;
  __shared__ double s_cur;
  __shared__ int s_first_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_cur = WRITE_ALL ? out[start] : out[0];
  __syncthreads();
  unsigned long long n_blocks = 0, n_fallback = 0, n_rounds = 0;
  size_t i0 = start;
  int items_now = kSpecItemsMax;  // shrinks to 1 after a block that did not verify, doubles after one that did
  int fail_streak = 0;
  while (i0 < n) {
    const double cur = s_cur;
    size_t it = (i0 + 1) / (4 * (size_t)kSpecThreads);
    const int items = (int)min((size_t)items_now, it < 1 ? (size_t)1 : it);
    const int B = (int)min((size_t)kSpecThreads * items, n - i0);
    bool neg = false;
    for (int j = tid; j < B; j += kSpecThreads) {
      const double v = x[i0 + j];
      neg |= v < 0.;
      sx[SPX(j)] = v;
      cand[SPX(j)] = cur;
    }
    const bool any_neg = __syncthreads_or(neg);
    // ulp grid of the block: that of cur's binade
    const int e = (int)((__double2hiint(cur) >> 20) & 0x7ff);
    int accepted = 0;  // candidates [0, accepted) are known to be the exact sequence
    if (e > 60 && e < 2040 && cur > 0.) {
      const int j0 = tid * items;
      if (STEP::kLinearGuess) {
        double loc[kSpecItemsMax], tsum = 0;
#pragma unroll
        for (int k = 0; k < kSpecItemsMax; ++k) {
          const int j = j0 + k;
          if (k < items && j < B) tsum += sx[SPX(j)];
          loc[k] = tsum;
        }
        double before;
        ScanD(scan_tmp.d).ExclusiveSum(tsum, before);
        const double base = (double)i0 * cur;
#pragma unroll
        for (int k = 0; k < kSpecItemsMax; ++k) {
          const int j = j0 + k;
          if (k < items && j < B) cand[SPX(j)] = (base + (before + loc[k])) / (double)(i0 + j + 1);
        }
        __syncthreads();
      }
      const double inv_u = __hiloint2double((2098 - e) << 20, 0);  // 2^(52 - (e - 1023)) = 1 / ulp(cur)
      const double u = __hiloint2double((e - 52) << 20, 0);
      // the counts and their reciprocals once per block (the rounds reuse them)
      const bool fast_step = STEP::kHasFast && !any_neg;
      double nn[kSpecItemsMax], rn[kSpecItemsMax];
#pragma unroll
      for (int k = 0; k < kSpecItemsMax; ++k) {
        nn[k] = (double)(i0 + j0 + k + 1);
        rn[k] = fast_step ? __drcp_rn(nn[k]) : 0.;
      }
      for (int r = 0; r < kSpecRounds; ++r) {
        ++n_rounds;
        if (tid == 0) s_first_bad = B;
        __syncthreads();
        long long d[kSpecItemsMax];
        bool bad = false;
        int my_first = B;
        long long tot = 0;
#pragma unroll
        for (int k = 0; k < kSpecItemsMax; ++k) {
          d[k] = 0;
          const int j = j0 + k;
          if (k < items && j < B) {
            const double pred = j == 0 ? cur : cand[SPX(j - 1)];
            const double nxt = fast_step ? STEP::step_fast(pred, sx[SPX(j)], nn[k], rn[k]) : STEP::step(pred, sx[SPX(j)], nn[k]);
            if (fast_step) bad |= !(pred >= kMeanSafe);  // (a candidate outside the range where step_fast is the exact step)
            if (nxt != cand[SPX(j)] && j < my_first) my_first = j;
            const double q = (nxt - pred) * inv_u;  // the step's increment in ulps of cur: whole, and small, or this path is not for it
            const long long qi = __double2ll_rn(q);
            bad |= !(fabs(q) < 1e15) || (double)qi != q;
            d[k] = qi;
            tot += qi;
          }
        }
        if (my_first < B) atomicMin(&s_first_bad, my_first);
        const bool any_bad = __syncthreads_or(bad);  // (also the barrier behind the atomicMin)
        accepted = s_first_bad;  // every candidate before the first mismatch is the exact step from an exact predecessor
        if (accepted == B || any_bad || r == kSpecRounds - 1) break;
        long long before;
        Scan(scan_tmp.i).ExclusiveSum(tot, before);
        long long acc = before;
#pragma unroll
        for (int k = 0; k < kSpecItemsMax; ++k) {
          const int j = j0 + k;
          if (k < items && j < B) {
            acc += d[k];
            cand[SPX(j)] = __dadd_rn(cur, __dmul_rn((double)acc, u));
          }
        }
        __syncthreads();
      }
    }
    int done = B;
    if (accepted < B) {  // uniform over the block: keep the verified prefix, walk a short run sequentially, speculate again
      // ... with runs that double while the failures follow each other (a table whose rounding decisions all hang on the
      // last ulp -- an exactly linear ramp is one -- is then walked at the sequential kernels' pace)
      ++n_fallback;
      const int seq = min(B - accepted, kSpecSeqRun << min(fail_streak, 6));
      if (tid == 0) {
        const double v0 = accepted ? cand[SPX(accepted - 1)] : cur;
        STEP::run(v0, sx, cand, accepted, accepted + seq, i0, !any_neg && v0 >= 2. * kMeanSafe && (size_t)seq <= i0);
      }
      __syncthreads();
      done = accepted + seq;
      ++fail_streak;
      items_now = fail_streak > 2 ? kSpecItemsMax : 1;  // (long runs need long blocks to be loaded)
    } else {
      fail_streak = 0;
      items_now = min(2 * items_now, kSpecItemsMax);
    }
    ++n_blocks;
    if (WRITE_ALL)
      for (int j = tid; j < done; j += kSpecThreads) out[i0 + j + 1] = cand[SPX(j)];
    if (tid == 0) s_cur = cand[SPX(done - 1)];
    __syncthreads();
    i0 += done;
  }
  if (tid == 0) {
    if (!WRITE_ALL) out[0] = s_cur;
    if (stats) { stats->blocks += n_blocks; stats->fallback_blocks += n_fallback; stats->rounds += n_rounds; }
  }
}

// blocks done, blocks that needed a sequential run, verification rounds -- for the running mean, then for the cumulative
// sum; summed over the builds of this context (diagnostics)
int sampler_spec_stats(upcgpu_ctx* c, unsigned long long out[6])
{
  for (int i = 0; i < 6; ++i) out[i] = 0;
  if (!c->spec_stats) return UPCGPU_OK;
  SpecStats h[2];
  UPC_CUDA(c, cudaMemcpy(h, c->spec_stats, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 2; ++k) { out[3 * k] = h[k].blocks; out[3 * k + 1] = h[k].fallback_blocks; out[3 * k + 2] = h[k].rounds; }
  return UPCGPU_OK;
}

constexpr size_t kSpecSmem = 2 * (size_t)kSpecPadded * sizeof(double);

// gsl_histogram2d_pdf_init on the device: mean (1 double), term [n] scratch, sum [n + 1]
static int pdf_init_device(upcgpu_ctx* c, const double* bin, size_t n, double* mean, double* term, double* sum, cudaStream_t st)
{
  if (!c->spec_attr_set) {
    cudaFuncSetAttribute(k_seq_spec<StepMean, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpecSmem);
    cudaFuncSetAttribute(k_seq_spec<StepSum, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpecSmem);
    c->spec_attr_set = true;
  }
  if (!c->spec_stats) {
    UPC_CUDA(c, cudaMalloc(&c->spec_stats, 2 * sizeof(SpecStats)));  // [0] running mean, [1] cumulative sum
    UPC_CUDA(c, cudaMemsetAsync(c->spec_stats, 0, 2 * sizeof(SpecStats), st));
  }
  const size_t head = std::min(n, kSeqHead);
  UPC_K(c), k_running_mean<<<1, 32, 0, st>>>(bin, head, mean);
  if (n > head) UPC_K(c), k_seq_spec<StepMean, false><<<1, kSpecThreads, kSpecSmem, st>>>(bin, n, head, mean, c->spec_stats);
  UPC_K(c), k_pdf_terms<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bin, n, mean, term);
  UPC_K(c), k_seq_cumsum<<<1, 32, 0, st>>>(term, head, sum);
  if (n > head) UPC_K(c), k_seq_spec<StepSum, true><<<1, kSpecThreads, kSpecSmem, st>>>(term, n, head, sum, c->spec_stats + 1);
  return UPCGPU_OK;
}

// gsl_histogram_pdf_init for the nm z-samplers: one thread per sampler
__global__ void k_pdf_init_rows(const double* __restrict__ bin, int nrows, int n, double* __restrict__ sum)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const double* b = bin + (size_t)r * n;
  double* s = sum + (size_t)r * (n + 1);
  double mean = 0;
  for (int i = 0; i < n; i++) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[i], mean), (double)(i + 1)));
  double acc = 0;
  s[0] = 0;
  for (int i = 0; i < n; i++) {
    acc = __dadd_rn(acc, __ddiv_rn(__ddiv_rn(b[i], mean), (double)n));
    s[i + 1] = acc;
  }
}

__global__ void k_edges(double lo, double d, int n, double* e)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) e[i] = __dadd_rn(lo, __dmul_rn(d, (double)i));  // mmin + dm * i, UpcGenerator.cpp:675-682
}

int sampler_build(upcgpu_ctx* c, const double* cs, const double* cszm, const double* cszm_s, const double* cszm_ps)
{
  const upcgpu_params& p = c->p;
  const size_t n = (size_t)p.nm * p.ny;
  cudaStream_t st = c->stream;
  if (cs) {
    if (!c->cs) UPC_CUDA(c, cudaMalloc(&c->cs, n * sizeof(double)));
    UPC_CUDA(c, cudaMemcpyAsync(c->cs, cs, n * sizeof(double), cudaMemcpyHostToDevice, st));
  } else if (!c->fold_ready) {
    c->err = "sampler_build: no cross-section table (call fold_sigma or pass cs)";
    return UPCGPU_EINVAL;
  }
  if (!c->sum2d) UPC_CUDA(c, cudaMalloc(&c->sum2d, (n + 1) * sizeof(double)));
  if (!c->edges_y) {
    UPC_CUDA(c, cudaMalloc(&c->edges_y, (p.ny + 1) * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&c->edges_m, (p.nm + 1) * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&c->edges_z, (p.nz + 1) * sizeof(double)));
  }
  const double dm = (p.mmax - p.mmin) / p.nm, dy = (p.ymax - p.ymin) / p.ny, dz = (p.zmax - p.zmin) / p.nz;
  UPC_K(c), k_edges<<<(p.ny + 128) / 128, 128, 0, st>>>(p.ymin, dy, p.ny, c->edges_y);
  UPC_K(c), k_edges<<<(p.nm + 128) / 128, 128, 0, st>>>(p.mmin, dm, p.nm, c->edges_m);
  UPC_K(c), k_edges<<<(p.nz + 128) / 128, 128, 0, st>>>(p.zmin, dz, p.nz, c->edges_z);
  if (!c->samp_term) UPC_CUDA(c, cudaMalloc(&c->samp_term, n * sizeof(double)));
  if (!c->samp_mean) UPC_CUDA(c, cudaMalloc(&c->samp_mean, sizeof(double)));
  double *term = c->samp_term, *mean = c->samp_mean;
  {
    const int prc = pdf_init_device(c, c->cs, n, mean, term, c->sum2d, st);
    if (prc) return prc;
  }
  // z samplers
  const size_t nzm = (size_t)p.nm * p.nz, nsz = (size_t)p.nm * (p.nz + 1);
  const double* first = p.use_pol ? cszm_s : cszm;
  if (!p.ignore_csz) {
    if (!first || (p.use_pol && !cszm_ps)) {
      c->err = "sampler_build: missing z cross-section table";
      return UPCGPU_EINVAL;
    }
    if (!c->samp_dz) UPC_CUDA(c, cudaMalloc(&c->samp_dz, nzm * sizeof(double)));
    double* dz_in = c->samp_dz;
    if (!c->sumz) UPC_CUDA(c, cudaMalloc(&c->sumz, nsz * sizeof(double)));
    UPC_CUDA(c, cudaMemcpyAsync(dz_in, first, nzm * sizeof(double), cudaMemcpyHostToDevice, st));
    UPC_K(c), k_pdf_init_rows<<<(p.nm + 63) / 64, 64, 0, st>>>(dz_in, p.nm, p.nz, c->sumz);
    if (p.use_pol) {
      if (!c->sumz_ps) UPC_CUDA(c, cudaMalloc(&c->sumz_ps, nsz * sizeof(double)));
      UPC_CUDA(c, cudaStreamSynchronize(st));
      UPC_CUDA(c, cudaMemcpyAsync(dz_in, cszm_ps, nzm * sizeof(double), cudaMemcpyHostToDevice, st));
      UPC_K(c), k_pdf_init_rows<<<(p.nm + 63) / 64, 64, 0, st>>>(dz_in, p.nm, p.nz, c->sumz_ps);
    }
  }
  UPC_CUDA(c, cudaStreamSynchronize(st));
  UPC_CUDA(c, cudaGetLastError());
  c->sampler_ready = true;
  return UPCGPU_OK;
}

__global__ void k_sample_ym(const double* __restrict__ u, size_t n, const double* __restrict__ sum, int ny, int nm,
                            const double* __restrict__ ye, const double* __restrict__ me, long long* __restrict__ k,
                            int* __restrict__ ybin, int* __restrict__ mbin, double* __restrict__ y,
                            double* __restrict__ m)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long kk; double yy, mm;
  sample_ym_dev(sum, ny, nm, ye, me, u[2 * t], u[2 * t + 1], kk, yy, mm);
  k[t] = kk; y[t] = yy; m[t] = mm;
  if (kk >= 0) {
    ybin[t] = get_bin(ny, yy, ye[0], ye[ny]);
    mbin[t] = get_bin(nm, mm, me[0], me[nm]);
  } else {
    ybin[t] = -1; mbin[t] = -1;
  }
}

__global__ void k_sample_z(const int* __restrict__ mbin, const double* __restrict__ u, size_t n,
                           const double* __restrict__ sumz, int nm, int nz, const double* __restrict__ ze,
                           double* __restrict__ z)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  int mb = mbin[t];
  z[t] = (mb >= 0 && mb < nm) ? sample_1d_dev(sumz + (size_t)mb * (nz + 1), nz, ze, u[t]) : nan("");
}

int sample_ym(upcgpu_ctx* c, const double* u, size_t n, long long* k, int* ybin, int* mbin, double* y, double* m)
{
  const upcgpu_params& p = c->p;
  if (!c->sampler_ready) { c->err = "sample_ym: samplers not built"; return UPCGPU_EINVAL; }
  double *du, *dy, *dm; long long* dk; int *dyb, *dmb;
  UPC_CUDA(c, cudaMalloc(&du, 2 * n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dy, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dm, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dk, n * sizeof(long long)));
  UPC_CUDA(c, cudaMalloc(&dyb, n * sizeof(int)));
  UPC_CUDA(c, cudaMalloc(&dmb, n * sizeof(int)));
  UPC_CUDA(c, cudaMemcpy(du, u, 2 * n * sizeof(double), cudaMemcpyHostToDevice));
  UPC_K(c), k_sample_ym<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(du, n, c->sum2d, p.ny, p.nm, c->edges_y, c->edges_m, dk,
                                                                 dyb, dmb, dy, dm);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  if (k) UPC_CUDA(c, cudaMemcpy(k, dk, n * sizeof(long long), cudaMemcpyDeviceToHost));
  if (ybin) UPC_CUDA(c, cudaMemcpy(ybin, dyb, n * sizeof(int), cudaMemcpyDeviceToHost));
  if (mbin) UPC_CUDA(c, cudaMemcpy(mbin, dmb, n * sizeof(int), cudaMemcpyDeviceToHost));
  if (y) UPC_CUDA(c, cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (m) UPC_CUDA(c, cudaMemcpy(m, dm, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(du); cudaFree(dy); cudaFree(dm); cudaFree(dk); cudaFree(dyb); cudaFree(dmb);
  return UPCGPU_OK;
}

int sample_z(upcgpu_ctx* c, const int* mbin, const double* u, size_t n, int ps, double* z)
{
  const upcgpu_params& p = c->p;
  const double* tabz = ps ? c->sumz_ps : c->sumz;
  if (!c->sampler_ready || !tabz) { c->err = "sample_z: z samplers not built"; return UPCGPU_EINVAL; }
  double *du, *dz; int* dmb;
  UPC_CUDA(c, cudaMalloc(&du, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dz, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dmb, n * sizeof(int)));
  UPC_CUDA(c, cudaMemcpy(du, u, n * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(dmb, mbin, n * sizeof(int), cudaMemcpyHostToDevice));
  UPC_K(c), k_sample_z<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(dmb, du, n, tabz, p.nm, p.nz, c->edges_z, dz);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  UPC_CUDA(c, cudaMemcpy(z, dz, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(du); cudaFree(dz); cudaFree(dmb);
  return UPCGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// generic histogram samplers (UpcSampler1D/2D facade)
int hist_pdf_init(upcgpu_ctx* c, const double* bins, size_t n, double* sum)
{
  double *db, *dt, *dm, *ds;
  UPC_CUDA(c, cudaMalloc(&db, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dt, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dm, sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&ds, (n + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMemcpyAsync(db, bins, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {
    const int prc = pdf_init_device(c, db, n, dm, dt, ds, c->stream);
    if (prc) { cudaFree(db); cudaFree(dm); cudaFree(dt); cudaFree(ds); return prc; }
  }
  UPC_CUDA(c, cudaMemcpyAsync(sum, ds, (n + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  cudaFree(db); cudaFree(dt); cudaFree(dm); cudaFree(ds);
  return UPCGPU_OK;
}

__global__ void k_hist_sample2d(const double* __restrict__ u, size_t n, const double* __restrict__ sum, int nx, int ny,
                                const double* __restrict__ xe, const double* __restrict__ ye, long long* __restrict__ k,
                                double* __restrict__ x, double* __restrict__ y)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  long long kk; double xx, yy;
  sample_ym_dev(sum, nx, ny, xe, ye, u[2 * t], u[2 * t + 1], kk, xx, yy);
  if (k) k[t] = kk;
  x[t] = xx; y[t] = yy;
}

__global__ void k_hist_sample1d(const double* __restrict__ u, size_t n, const double* __restrict__ sum, int nb,
                                const double* __restrict__ e, double* __restrict__ x)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n) x[t] = sample_1d_dev(sum, nb, e, u[t]);
}

int hist_sample2d(upcgpu_ctx* c, const double* sum, int nx, int ny, const double* xe, const double* ye, const double* u,
                  size_t n, long long* k, double* x, double* y)
{
  const size_t nb = (size_t)nx * ny;
  double *ds, *dxe, *dye, *du, *dx, *dy; long long* dk;
  UPC_CUDA(c, cudaMalloc(&ds, (nb + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dxe, (nx + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dye, (ny + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&du, 2 * n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dx, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dy, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dk, n * sizeof(long long)));
  UPC_CUDA(c, cudaMemcpy(ds, sum, (nb + 1) * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(dxe, xe, (nx + 1) * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(dye, ye, (ny + 1) * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(du, u, 2 * n * sizeof(double), cudaMemcpyHostToDevice));
  UPC_K(c), k_hist_sample2d<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(du, n, ds, nx, ny, dxe, dye, dk, dx, dy);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  if (k) UPC_CUDA(c, cudaMemcpy(k, dk, n * sizeof(long long), cudaMemcpyDeviceToHost));
  UPC_CUDA(c, cudaMemcpy(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost));
  UPC_CUDA(c, cudaMemcpy(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(ds); cudaFree(dxe); cudaFree(dye); cudaFree(du); cudaFree(dx); cudaFree(dy); cudaFree(dk);
  return UPCGPU_OK;
}

int hist_sample1d(upcgpu_ctx* c, const double* sum, int nb, const double* edges, const double* u, size_t n, double* x)
{
  double *ds, *de, *du, *dx;
  UPC_CUDA(c, cudaMalloc(&ds, (nb + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&de, (nb + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&du, n * sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dx, n * sizeof(double)));
  UPC_CUDA(c, cudaMemcpy(ds, sum, (nb + 1) * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(de, edges, (nb + 1) * sizeof(double), cudaMemcpyHostToDevice));
  UPC_CUDA(c, cudaMemcpy(du, u, n * sizeof(double), cudaMemcpyHostToDevice));
  UPC_K(c), k_hist_sample1d<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(du, n, ds, nb, de, dx);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  UPC_CUDA(c, cudaMemcpy(x, dx, n * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(ds); cudaFree(de); cudaFree(du); cudaFree(dx);
  return UPCGPU_OK;
}

}  // namespace upc
