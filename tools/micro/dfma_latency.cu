// DFMA pipe probe for sm_100a: throughput as a function of independent chains per thread and warps
// per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters, double a, double b)
{
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(int warps_per_sm, double* d)
{
  int threads = warps_per_sm * 32, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<148, threads>>>(d, 16, 0.999999, 1e-7);
  cudaEventRecord(e0);
  k<CH><<<148, threads>>>(d, iters, 0.999999, 1e-7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double dfma_per_warp = (double)iters * 16 * CH;
  double cycles = ms * 1e-3 * 1.965e9;
  // per SMSP: warps_per_sm/4 warps
  printf("chains %d warps/SM %2d : %.2f cycles per DFMA per warp, SMSP pipe util %.1f%%  (%.2f TFLOP/s)\n", CH, warps_per_sm,
         cycles / dfma_per_warp, 100. * dfma_per_warp * (warps_per_sm / 4.) * 2. / cycles,
         2. * dfma_per_warp * 32 * warps_per_sm * 148 / (ms * 1e-3) / 1e12);
}
int main()
{
  double* d;
  cudaMalloc(&d, 148 * 1024 * sizeof(double));
  for (int w : {4, 8, 12, 16, 32}) {
    run<1>(w, d); run<2>(w, d); run<3>(w, d); run<4>(w, d); run<6>(w, d); run<8>(w, d);
  }
  return 0;
}
