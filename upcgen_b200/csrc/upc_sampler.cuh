// upc_sampler.cuh -- device-side inverse-CDF sampling (S2/S3), shared by the sampling hooks and
// the event kernel.  Reference: include/UpcSampler.h:68-71, :118-133 over GSL's
// gsl_histogram[2d]_pdf_sample and histogram/find.c.  All arithmetic that feeds an integer bin
// uses explicit non-fused FP64 operations in the reference's order (bit-exact selection).
#pragma once
#include "upc_math.cuh"

namespace upc {

// S2: gsl_histogram2d_pdf_sample / find (histogram/find.c): the unique k with
// sum[k] <= r < sum[k+1]
__device__ __forceinline__ long long pdf_find(const double* __restrict__ range, size_t n, double x)
{
  if (x < range[0] || x >= range[n]) return -1;
  size_t lower = 0, upper = n;
  while (upper - lower > 1) {
    size_t mid = (upper + lower) / 2;
    if (x >= range[mid]) lower = mid; else upper = mid;
  }
  return (long long)lower;
}

// S3: UpcSampler2D::getBinX/getBinY, include/UpcSampler.h:124-133
__device__ __forceinline__ int get_bin(int nbins, double x, double lo, double hi)
{
  int inner = (int)__dmul_rn((double)nbins, __dsub_rn(x, lo));
  return (int)__ddiv_rn((double)inner, __dsub_rn(hi, lo));
}

__device__ __forceinline__ void sample_ym_dev(const double* __restrict__ sum, int ny, int nm,
                                              const double* __restrict__ ye, const double* __restrict__ me, double r1,
                                              double r2, long long& k, double& y, double& m)
{
  if (r2 == 1.0) r2 = 0.0;
  if (r1 == 1.0) r1 = 0.0;
  k = pdf_find(sum, (size_t)ny * nm, r1);
  if (k < 0) { y = nan(""); m = nan(""); return; }
  const size_t i = (size_t)k / nm;
  const size_t j = (size_t)k - i * nm;
  const double delta = __ddiv_rn(__dsub_rn(r1, sum[k]), __dsub_rn(sum[k + 1], sum[k]));
  y = __dadd_rn(ye[i], __dmul_rn(delta, __dsub_rn(ye[i + 1], ye[i])));
  m = __dadd_rn(me[j], __dmul_rn(r2, __dsub_rn(me[j + 1], me[j])));
}

__device__ __forceinline__ double sample_1d_dev(const double* __restrict__ sum, int n, const double* __restrict__ edges,
                                                double r)
{
  if (r == 1.0) r = 0.0;
  long long i = pdf_find(sum, (size_t)n, r);
  if (i < 0) return 0.;
  const double delta = __ddiv_rn(__dsub_rn(r, sum[i]), __dsub_rn(sum[i + 1], sum[i]));
  return __dadd_rn(edges[i], __dmul_rn(delta, __dsub_rn(edges[i + 1], edges[i])));
}


}  // namespace upc
