// upc_fold.cu -- sigma fold (X1) and the inverse-CDF samplers (S1-S3).
// Reference: src/UpcCrossSection.cpp:594-698 (calcNucCrossSectionYM) and include/UpcSampler.h
// (GSL gsl_histogram[2d]_pdf_init / _pdf_sample, getBinX/getBinY).
//
// Everything that decides an integer bin is evaluated with explicit non-fused FP64 operations
// (__dadd_rn/__dmul_rn/__ddiv_rn) in the reference's operation order, so that the cumulative
// tables and the selected bins are bit-identical to a generic x86-64 build of the reference.
#include <cstdio>

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
  UPC_K(c), k_running_mean<<<1, 32, 0, st>>>(c->cs, n, mean);
  UPC_K(c), k_pdf_terms<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->cs, n, mean, term);
  UPC_K(c), k_seq_cumsum<<<1, 32, 0, st>>>(term, n, c->sum2d);
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
  UPC_K(c), k_running_mean<<<1, 32, 0, c->stream>>>(db, n, dm);
  UPC_K(c), k_pdf_terms<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(db, n, dm, dt);
  UPC_K(c), k_seq_cumsum<<<1, 32, 0, c->stream>>>(dt, n, ds);
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
