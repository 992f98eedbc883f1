// upc_math.cuh -- FP64 device math shared by the table, flux, luminosity and event kernels.
//
// Everything here is plain CUDA-core FP64 (DFMA/DADD/DMUL + MUFU seeds); no tensor cores are
// involved anywhere on this path (transcendental-heavy quadrature, not a contraction).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "upc_bessel_coeffs.h"

namespace upc {

// include/UpcPhysConstants.h:26-32 of the reference
constexpr double kAlpha = 1.0 / 137.035999074;
constexpr double kHc = 0.1973269718;
constexpr double kMProt = 0.9382720813;
constexpr double kMNeut = 0.939565346;
constexpr double kPi = 3.14159265358979323846;

// form-factor table grid, src/UpcCrossSection.cpp:46-49
constexpr double kQ2min = 1e-9;
constexpr double kQ2max = 2.;
constexpr int kNQ2 = 1000000;
constexpr double kDQ2 = (kQ2max - kQ2min) / kNQ2;
// breakup table grid, src/UpcCrossSection.cpp:418-421
constexpr double kBkBmin = 1e-6;
constexpr double kBkDb = (1000 - 1e-6) / 1000000;
// G_AA grid, src/UpcCrossSection.cpp:366-367
constexpr int kNB = 200;

template <int N>
__device__ __forceinline__ double horner(const double (&c)[N], double u)
{
  double r = c[N - 1];
#pragma unroll
  for (int i = N - 2; i >= 0; --i) r = fma(r, u, c[i]);
  return r;
}

// K0 and K1 together (fluxPoint needs both; they share log/exp/sqrt).  Replaces
// gsl_sf_bessel_K0/K1 at src/UpcCrossSection.cpp:171-172.  Q7: where GSL would raise an
// underflow error (x >~ 707) the value is simply 0 here.
__device__ __forceinline__ void bessel_k0k1(double x, double& k0, double& k1)
{
  if (x <= 2.) {
    double u = fma(0.5 * x, x, -1.);
    double lg = log(0.5 * x);
    k0 = fma(-lg, horner(UPC_K0_Q, u), horner(UPC_K0_P, u));
    k1 = fma(lg * x, horner(UPC_K1_Q, u), horner(UPC_K1_P, u) / x);
  } else {
    double e = exp(-x) * rsqrt(x);
    if (x <= 8.) {
      double u = (16. / x - 5.) * (1. / 3.);
      k0 = e * horner(UPC_K0_A, u);
      k1 = e * horner(UPC_K1_A, u);
    } else {
      double u = 16. / x - 1.;
      k0 = e * horner(UPC_K0_B, u);
      k1 = e * horner(UPC_K1_B, u);
    }
  }
}

// (J1, which replaces gsl_sf_bessel_J1 at src/UpcCrossSection.cpp:189, lives in upc_hot.cuh: j1_3 and its two branches.)

// ROOT TMath::BesselI1 / BesselK1 polynomials (A&S 9.8.3-9.8.8), used by calcBreakupProb
// (src/UpcCrossSection.cpp:982-999).  H2: these 1e-7-accurate forms ARE the reference.
__device__ __forceinline__ double tmath_bessel_i1(double x)
{
  const double p1 = 0.5, p2 = 0.87890594, p3 = 0.51498869, p4 = 0.15084934, p5 = 2.658733e-2,
               p6 = 3.01532e-3, p7 = 3.2411e-4;
  const double q1 = 0.39894228, q2 = -3.988024e-2, q3 = -3.62018e-3, q4 = 1.63801e-3,
               q5 = -1.031555e-2, q6 = 2.282967e-2, q7 = -2.895312e-2, q8 = 1.787654e-2,
               q9 = -4.20059e-3;
  const double k1 = 3.75;
  double ax = fabs(x);
  if (ax < k1) {
    double xx = x / k1;
    double y = xx * xx;
    return x * (p1 + y * (p2 + y * (p3 + y * (p4 + y * (p5 + y * (p6 + y * p7))))));
  }
  double y = k1 / ax;
  double r = (exp(ax) / sqrt(ax)) *
             (q1 + y * (q2 + y * (q3 + y * (q4 + y * (q5 + y * (q6 + y * (q7 + y * (q8 + y * q9))))))));
  return x < 0 ? -r : r;
}

__device__ __forceinline__ double tmath_bessel_k1(double x)
{
  const double p1 = 1., p2 = 0.15443144, p3 = -0.67278579, p4 = -0.18156897, p5 = -1.919402e-2,
               p6 = -1.10404e-3, p7 = -4.686e-5;
  const double q1 = 1.25331414, q2 = 0.23498619, q3 = -3.655620e-2, q4 = 1.504268e-2,
               q5 = -7.80353e-3, q6 = 3.25614e-3, q7 = -6.8245e-4;
  if (x <= 0) return 0;
  if (x <= 2) {
    double y = x * x / 4;
    return (log(x / 2.) * tmath_bessel_i1(x)) +
           (1. / x) * (p1 + y * (p2 + y * (p3 + y * (p4 + y * (p5 + y * (p6 + y * p7))))));
  }
  double y = 2 / x;
  return (exp(-x) / sqrt(x)) * (q1 + y * (q2 + y * (q3 + y * (q4 + y * (q5 + y * (q6 + y * q7))))));
}

// 1 / a for a normal, positive a: hardware seed and two Newton steps (error ~1 ulp) -- six dependent operations where
// the IEEE division is ~35.  The tridiagonal recurrences below are ONE dependent chain per solve (alpha -> 1 / alpha
// -> gamma -> next alpha), so the division's latency is the latency of the whole table stage; the spline
// coefficients move by ~1e-16 relative, eleven orders below anything the path resolves (tests: <= 1e-9).
__device__ __forceinline__ double rcp_fast(double a)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  r = fma(r, fma(-a, r, 1.0), r);
  r = fma(r, fma(-a, r, 1.0), r);
  return r;
}

// knot of a uniform table exactly as the reference forms it: x0 + i*dx (two roundings)
__device__ __forceinline__ double knot(double x0, double dx, int i) { return __dadd_rn(x0, __dmul_rn((double)i, dx)); }

// One spline interval in evaluation form: S(x) = y + d*(b + d*(c + d*dd)), d = x - x_i, with
// b, dd computed exactly as GSL's cspline coeff_calc does from (y_i, y_i+1, c_i, c_i+1, dx).
struct __align__(32) SplineSeg {
  double y, b, c, d;
};

__device__ __forceinline__ SplineSeg make_seg(double x_lo, double x_hi, double y_lo, double y_hi, double c_i,
                                              double c_ip1)
{
  // gsl interpolation/cspline.c coeff_calc + cspline_eval; explicit non-fused ops so that the
  // coefficients equal the ones GSL forms on the fly
  double dx = __dsub_rn(x_hi, x_lo);
  double dy = __dsub_rn(y_hi, y_lo);
  SplineSeg s;
  s.y = y_lo;
  s.b = __dsub_rn(__ddiv_rn(dy, dx), __ddiv_rn(__dmul_rn(dx, __dadd_rn(c_ip1, __dmul_rn(2.0, c_i))), 3.0));
  s.c = c_i;
  s.d = __ddiv_rn(__dsub_rn(c_ip1, c_i), __dmul_rn(3.0, dx));
  return s;
}

// one 256-bit read-only load per segment (sm_100: LDG.E.256): half the LSU requests of 2 x 128 bit
__device__ __forceinline__ SplineSeg ld_seg(const SplineSeg* p)
{
  SplineSeg s;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s.y), "=d"(s.b), "=d"(s.c), "=d"(s.d) : "l"(p));
  return s;
}

__device__ __forceinline__ double seg_eval(const SplineSeg& s, double delx)
{
  return fma(delx, fma(delx, fma(delx, s.d, s.c), s.b), s.y);
}

// natural cubic spline evaluated GSL-style from the raw (x, y, c) arrays with a bsearch-equal
// index (used by the small set-up kernels only)
__device__ inline double spline_eval_raw(const double* xa, const double* ya, const double* ca, int n, double x0,
                                         double inv_dx, double x)
{
  int idx = (int)((x - x0) * inv_dx);
  idx = max(0, min(idx, n - 2));
  while (idx > 0 && xa[idx] > x) --idx;
  while (idx < n - 2 && xa[idx + 1] <= x) ++idx;
  SplineSeg s = make_seg(xa[idx], xa[idx + 1], ya[idx], ya[idx + 1], ca[idx], ca[idx + 1]);
  double delx = x - xa[idx];
  return s.y + delx * (s.b + delx * (s.c + delx * s.d));
}

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter = (ctr lo, ctr hi, block, 0),
// key = (seed lo, seed hi).  Two 53-bit uniforms in [0,1) per block.
__host__ __device__ inline void philox4x32_10(uint64_t seed, uint64_t ctr, uint32_t block, double& u0, double& u1)
{
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = block, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  uint64_t a = ((uint64_t)c0 << 32) | c1;
  uint64_t b = ((uint64_t)c2 << 32) | c3;
  u0 = (double)(a >> 11) * 0x1.0p-53;
  u1 = (double)(b >> 11) * 0x1.0p-53;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace upc
