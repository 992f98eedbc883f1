// Sequential FP64 recurrences on ONE lane (the running mean / cumulative sum of gsl_histogram2d_pdf_init): what bounds
// them on sm_100a?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o seq_chain seq_chain.cu && ./seq_chain
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kChunk = 1024;

__device__ __forceinline__ double div3(double x, double d, double y)
{
  const double q = __dmul_rn(x, y);
  return __fma_rn(__fma_rn(-q, d, x), y, q);
}

// V: 0 = divisor converted from the 64-bit index per element (I2F.F64.U64), 1 = divisor staged as a double next to its
// reciprocal, 2 = divisor as base + k (one conversion per block), 3 = library division
template <int V>
__global__ void __launch_bounds__(32) k_mean(const double* __restrict__ bin, size_t n, double* __restrict__ out)
{
  __shared__ double sb[2][kChunk], sr[2][kChunk], sd[2][kChunk];
  const int lane = threadIdx.x;
  double mean = 0;
  int buf = 0;
  for (int j = lane; j < kChunk && (size_t)j < n; j += 32) { sb[0][j] = bin[j]; sr[0][j] = __drcp_rn((double)(j + 1)); sd[0][j] = (double)(j + 1); }
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kChunk, buf ^= 1) {
    const size_t nxt = i0 + kChunk;
    if (lane == 0) {
      const int m = (int)min((size_t)kChunk, n - i0);
      const double *b = sb[buf], *r = sr[buf], *dd = sd[buf];
      for (int j = 0; j + 8 <= m; j += 8) {
        double bv[8], rv[8], dv[8];
        const double base = (double)(i0 + j);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          bv[k] = b[j + k]; rv[k] = r[j + k];
          dv[k] = V == 0 ? (double)(i0 + j + k + 1) : V == 1 ? dd[j + k] : base + (double)(k + 1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double d = __dsub_rn(bv[k], mean);
          mean = __dadd_rn(mean, V == 3 ? __ddiv_rn(d, dv[k]) : div3(d, dv[k], rv[k]));
        }
      }
    } else {
      for (size_t j = nxt + (lane - 1); j < nxt + kChunk && j < n; j += 31) {
        sb[buf ^ 1][j - nxt] = bin[j]; sr[buf ^ 1][j - nxt] = __drcp_rn((double)(j + 1)); sd[buf ^ 1][j - nxt] = (double)(j + 1);
      }
    }
    __syncwarp();
  }
  if (lane == 0) out[0] = mean;
}

// pure dependent chains in registers: OPS dependent DFMAs per "element"
template <int OPS>
__global__ void __launch_bounds__(32) k_chain(size_t n, double a, double b, double* out)
{
  double x = threadIdx.x;
  if (threadIdx.x == 0)
    for (size_t i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < OPS; ++k) x = fma(x, a, b);
    }
  out[threadIdx.x] = x;
}
// the same with DADD
__global__ void __launch_bounds__(32) k_chain_add(size_t n, double b, double* out)
{
  double x = threadIdx.x;
  if (threadIdx.x == 0)
    for (size_t i = 0; i < n; i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) x = __dadd_rn(x, b);
    }
  out[threadIdx.x] = x;
}

// cumulative sum: U terms into registers, U chained additions, U stores
template <int U, bool ALL_LANES_WRITE>
__global__ void __launch_bounds__(32) k_cumsum(const double* __restrict__ term, size_t n, double* __restrict__ sum)
{
  __shared__ double st[2][kChunk], so[kChunk];
  const int lane = threadIdx.x;
  double s = 0;
  int buf = 0;
  if (lane == 0) sum[0] = 0;
  for (int j = lane; j < kChunk && (size_t)j < n; j += 32) st[0][j] = term[j];
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kChunk, buf ^= 1) {
    const size_t nxt = i0 + kChunk;
    const int m = (int)min((size_t)kChunk, n - i0);
    if (lane == 0) {
      const double* t = st[buf];
      for (int j = 0; j + U <= m; j += U) {
        double v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = t[j + k];
#pragma unroll
        for (int k = 0; k < U; ++k) { s = __dadd_rn(s, v[k]); v[k] = s; }
#pragma unroll
        for (int k = 0; k < U; ++k) { if (ALL_LANES_WRITE) so[j + k] = v[k]; else sum[i0 + j + k + 1] = v[k]; }
      }
    } else {
      for (size_t j = nxt + (lane - 1); j < nxt + kChunk && j < n; j += 31) st[buf ^ 1][j - nxt] = term[j];
    }
    __syncwarp();
    if (ALL_LANES_WRITE) {
      for (int j = lane; j < m; j += 32) sum[i0 + j + 1] = so[j];
      __syncwarp();
    }
  }
}


// software-pipelined: the next eight elements' operands are loaded into registers BEFORE this block's chain starts
template <bool CHECK>
__global__ void __launch_bounds__(32) k_mean_pipe(const double* __restrict__ bin, size_t n, double* __restrict__ out)
{
  __shared__ double sb[2][kChunk], sr[2][kChunk];
  const int lane = threadIdx.x;
  double mean = 0;
  int buf = 0;
  for (int j = lane; j < kChunk && (size_t)j < n; j += 32) { sb[0][j] = bin[j]; sr[0][j] = __drcp_rn((double)(j + 1)); }
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kChunk, buf ^= 1) {
    const size_t nxt = i0 + kChunk;
    if (lane == 0) {
      const int m = (int)min((size_t)kChunk, n - i0);
      const double *b = sb[buf], *r = sr[buf];
      const int m8 = m & ~7;
      double bn[8], rn[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { bn[k] = b[k]; rn[k] = r[k]; }
      for (int j = 0; j < m8; j += 8) {
        double bv[8], rv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { bv[k] = bn[k]; rv[k] = rn[k]; }
        const int jn = (j + 8 < m8) ? j + 8 : j;
#pragma unroll
        for (int k = 0; k < 8; ++k) { bn[k] = b[jn + k]; rn[k] = r[jn + k]; }
        const double base = (double)(i0 + j);
        const double mean0 = mean;
        bool tiny = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const double d = __dsub_rn(bv[k], mean);
          if (CHECK) tiny |= fabs(d) < 0x1p-960;
          mean = __dadd_rn(mean, div3(d, base + (double)(k + 1), rv[k]));
        }
        if (CHECK && tiny) {
          mean = mean0;
#pragma unroll 1
          for (int k = 0; k < 8; ++k) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[j + k], mean), (double)(i0 + j + k + 1)));
        }
      }
      for (int j = m8; j < m; ++j) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[j], mean), (double)(i0 + j + 1)));
    } else {
      for (size_t j = nxt + (lane - 1); j < nxt + kChunk && j < n; j += 31) {
        sb[buf ^ 1][j - nxt] = bin[j]; sr[buf ^ 1][j - nxt] = __drcp_rn((double)(j + 1));
      }
    }
    __syncwarp();
  }
  if (lane == 0) out[0] = mean;
}

// cumulative sum, software-pipelined: lane 0 keeps the chain; results go to shared memory and all lanes write them out
// one chunk later (so the write-out overlaps the next chunk's chain)
template <int U>
__global__ void __launch_bounds__(32) k_cumsum_pipe(const double* __restrict__ term, size_t n, double* __restrict__ sum)
{
  __shared__ double st[2][kChunk], so[2][kChunk];
  const int lane = threadIdx.x;
  double s = 0;
  int buf = 0;
  if (lane == 0) sum[0] = 0;
  for (int j = lane; j < kChunk && (size_t)j < n; j += 32) st[0][j] = term[j];
  __syncwarp();
  size_t prev0 = 0; int prevm = 0;
  for (size_t i0 = 0; i0 < n; i0 += kChunk, buf ^= 1) {
    const size_t nxt = i0 + kChunk;
    const int m = (int)min((size_t)kChunk, n - i0);
    if (lane == 0) {
      const double* t = st[buf];
      double* o = so[buf];
      const int mu = m - m % U;
      double vn[U];
#pragma unroll
      for (int k = 0; k < U; ++k) vn[k] = t[k];
      for (int j = 0; j < mu; j += U) {
        double v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = vn[k];
        const int jn = (j + U < mu) ? j + U : j;
#pragma unroll
        for (int k = 0; k < U; ++k) vn[k] = t[jn + k];
#pragma unroll
        for (int k = 0; k < U; ++k) { s = __dadd_rn(s, v[k]); v[k] = s; }
#pragma unroll
        for (int k = 0; k < U; ++k) o[j + k] = v[k];
      }
      for (int j = mu; j < m; ++j) { s = __dadd_rn(s, t[j]); o[j] = s; }
    } else {
      // lanes 1..31: stage the next chunk and write out the previous chunk's results
      for (size_t j = nxt + (lane - 1); j < nxt + kChunk && j < n; j += 31) st[buf ^ 1][j - nxt] = term[j];
      for (int j = lane - 1; j < prevm; j += 31) sum[prev0 + j + 1] = so[buf ^ 1][j];
    }
    __syncwarp();
    prev0 = i0; prevm = m;
  }
  for (int j = lane; j < prevm; j += 32) sum[prev0 + j + 1] = so[buf ^ 1][j];
}

// ---- the library's kernel, verbatim (upcgen_b200/csrc/upc_fold.cu) ----
__device__ __forceinline__ double div_lib(double x, double d, double y)
{
  const double q = __dmul_rn(x, y);
  return __fma_rn(__fma_rn(-q, d, x), y, q);
}

constexpr int kChunkLib = 1024;
constexpr double kMeanSafe = 0x1p-900;

// tools/micro/seq_chain.cu measures the pieces on B200: a dependent DFMA is 8.2 cycles, this five-operation recurrence
// (sub, mul, fma, fma, add) runs at 53 cycles per bin, with __ddiv_rn in it at 137.
__global__ void __launch_bounds__(32) k_mean_lib(const double* __restrict__ bin, size_t n, double* __restrict__ mean_out)
{
  __shared__ double sb[2][kChunkLib], sr[2][kChunkLib];
  __shared__ int sneg[2];  // a negative bin in the chunk (GSL refuses such a histogram): library division throughout
  const int lane = threadIdx.x;
  double mean = 0;
  int buf = 0;
  if (lane < 2) sneg[lane] = 0;
  __syncwarp();
  {
    bool neg = false;
    for (int j = lane; j < kChunkLib && (size_t)j < n; j += 32) {
      const double v = bin[j];
      neg |= v < 0.;
      sb[0][j] = v;
      sr[0][j] = __drcp_rn((double)(j + 1));
    }
    if (neg) sneg[0] = 1;
  }
  __syncwarp();
  for (size_t i0 = 0; i0 < n; i0 += kChunkLib, buf ^= 1) {
    const size_t nxt = i0 + kChunkLib;
    if (lane == 0) {
      const int m = (int)min((size_t)kChunkLib, n - i0);
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
            mean = __dadd_rn(mean, div_lib(__dsub_rn(bv[k], mean), base + (double)(k + 1), rv[k]));
        }
        for (; j < m; ++j) mean = __dadd_rn(mean, __ddiv_rn(__dsub_rn(b[j], mean), (double)(i0 + j + 1)));
      }
      sneg[buf] = 0;
    } else {
      // lanes 1..31 stage the next chunk meanwhile
      bool neg = false;
      for (size_t j = nxt + (lane - 1); j < nxt + kChunkLib && j < n; j += 31) {
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


template <class F>
float timeit(F f)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main()
{
  const size_t n = 1 << 20;
  double *bin, *out, *sum;
  cudaMalloc(&bin, n * 8); cudaMalloc(&out, 32 * 8); cudaMalloc(&sum, (n + 1) * 8);
  double* h = new double[n];
  for (size_t i = 0; i < n; i++) h[i] = 1.0 + (i % 977) * 1e-3;
  cudaMemcpy(bin, h, n * 8, cudaMemcpyHostToDevice);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ghz = clk * 1e-6;
  auto rep = [&](const char* name, float ms, int ops) {
    printf("%-46s %8.3f ms  %6.1f ns/elem  %6.1f cycles/elem  %5.1f cycles/op\n", name, ms, ms * 1e6 / n, ms * 1e6 / n * ghz,
           ms * 1e6 / n * ghz / ops);
  };
  rep("chain 1 DFMA", timeit([&] { k_chain<1><<<1, 32>>>(n, 0.999, 1e-3, out); }), 1);
  rep("chain 5 DFMA", timeit([&] { k_chain<5><<<1, 32>>>(n, 0.999, 1e-3, out); }), 5);
  rep("chain 1 DADD (x8 unrolled)", timeit([&] { k_chain_add<<<1, 32>>>(n, 1e-3, out); }), 1);
  rep("mean V0 (I2F.U64 per element)", timeit([&] { k_mean<0><<<1, 32>>>(bin, n, out); }), 5);
  rep("mean V1 (divisor staged as double)", timeit([&] { k_mean<1><<<1, 32>>>(bin, n, out); }), 5);
  rep("mean V2 (divisor = base + k)", timeit([&] { k_mean<2><<<1, 32>>>(bin, n, out); }), 5);
  rep("mean V3 (__ddiv_rn)", timeit([&] { k_mean<3><<<1, 32>>>(bin, n, out); }), 5);
  rep("mean pipelined, no check", timeit([&] { k_mean_pipe<false><<<1, 32>>>(bin, n, out); }), 5);
  rep("mean pipelined, tiny check", timeit([&] { k_mean_pipe<true><<<1, 32>>>(bin, n, out); }), 5);
  rep("cumsum pipelined U=8", timeit([&] { k_cumsum_pipe<8><<<1, 32>>>(bin, n, sum); }), 1);
  rep("cumsum pipelined U=16", timeit([&] { k_cumsum_pipe<16><<<1, 32>>>(bin, n, sum); }), 1);
  rep("cumsum U=16 via shared, all lanes write", timeit([&] { k_cumsum<16, true><<<1, 32>>>(bin, n, sum); }), 1);
  rep("cumsum U=16 lane 0 writes", timeit([&] { k_cumsum<16, false><<<1, 32>>>(bin, n, sum); }), 1);
  rep("cumsum U=32 via shared, all lanes write", timeit([&] { k_cumsum<32, true><<<1, 32>>>(bin, n, sum); }), 1);
  {
    const size_t n2 = 121121;
    for (int rep2 = 0; rep2 < 3; ++rep2) {
      float ms = timeit([&] { k_mean<2><<<1, 32>>>(bin, n2, out); });
      printf("mean V2 at n = 121121: %.3f ms = %.1f ns/elem\n", ms, ms * 1e6 / n2);
      ms = timeit([&] { k_mean_lib<<<1, 32>>>(bin, n2, out); });
      printf("library kernel at n = 121121: %.3f ms = %.1f ns/elem\n", ms, ms * 1e6 / n2);
      ms = timeit([&] { k_mean_pipe<false><<<1, 32>>>(bin, n2, out); });
      printf("mean pipelined at n = 121121: %.3f ms = %.1f ns/elem\n", ms, ms * 1e6 / n2);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
