// upc_hot.cuh -- the innermost arithmetic of the form-factor flux (F2: fluxFormIntegrand with
// J1), written three evaluations wide.
//
// Why it looks like this (measured, profiles/r01_v0|v1_ncu_qags_cfg2.txt): on sm_100 a DFMA takes
// its third operand from a register or a UNIFORM register, not from a constant bank, and the
// integrand touches ~80 distinct FP64 coefficients.  Left to itself ptxas hoists them all out of
// the loop and then spills uniform registers around every DFMA (R2UR/MOV.SPILL = 30 % of the
// issued instructions), or, for literals, builds each one with two UMOVs (33 %).  Here every
// coefficient is fetched at its point of use with a volatile ld.const -- ptxas turns neighbouring
// ones into a single LDCU.128 -- and is consumed by THREE independent evaluations, so the FP64
// pipe sees ~6 DFMAs per uniform load and three interleaved dependency chains per lane.
//
// Include from exactly one translation unit (defines an extern "C" __constant__ table).
#pragma once
#include "upc_math.cuh"

namespace upc {

// ---- one __constant__ block holding everything the integrand needs ----
enum HotOff {
  H_J1P = 0,                         // 17: J1(x)/x on x <= 8, centred in x^2/32 - 1
  H_J1M = H_J1P + UPC_J1_P_N,        // 14: modulus
  H_J1T = H_J1M + UPC_J1_M_N,        // 16: phase
  H_SC = H_J1T + UPC_J1_T_N,         // 16: angle reduction by pi/2 and the fdlibm sin / cos kernels
  H_EPS = H_SC + 16,                 // 8 : 3/8, spare
  H_MISC = H_EPS + 8,                // 8 : 1/32, 2/pi, 1/sqrt2, Q2min, 1/dQ2, dQ2, 64, sqrt(2/pi)
  H_END = H_MISC + 8
};

}  // namespace upc

extern "C" {
__constant__ double upc_hot[upc::H_END] = {
  UPC_J1_P_VALUES, UPC_J1_M_VALUES, UPC_J1_T_VALUES,
  // sincos: 2/pi, pi/2 hi, mid, lo, S6..S1, C6..C1
  0.63661977236758134308, 1.57079632679489655800e+00, 6.12323399573676603587e-17, -1.49738490485916983693e-33,
  1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
  -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
  -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
  2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
  // 3/8 (rsqrt refinement), 7 spare
  0.375, 0., 0., 0., 0., 0., 0., 0.,
  // misc
  1. / 32., 0.63661977236758134308, 0.70710678118654752440, upc::kQ2min, 1. / upc::kDQ2, upc::kDQ2, 64.,
  0.79788456080286535588 /* sqrt(2/pi) */};
}

namespace upc {

template <int I>
__device__ __forceinline__ double hot()
{
  double v;
  asm volatile("ld.const.f64 %0, [upc_hot+%1];" : "=d"(v) : "n"(I * 8));
  return v;
}

// where the coefficients come from: the __constant__ block (volatile ld.const, see above) ...
struct HotConst {
  template <int I>
  __device__ __forceinline__ double at() const { return hot<I>(); }
};
// ... or a copy of it in shared memory: plain LDS into ordinary registers.  In a kernel that keeps
// other FP64 constants in uniform registers the ld.const stream overflows the uniform register
// file (R2UR / MOV.SPILL around the DFMAs); LDS broadcasts do not.
struct HotShared {
  const double* p;
  template <int I>
  __device__ __forceinline__ double at() const { return p[I]; }
};

// N evaluations side by side in one thread: every coefficient fetch feeds N independent DFMA chains
template <int N>
struct DV {
  double v[N];
};
using D3 = DV<3>;

#define UPC_FOR_N _Pragma("unroll") for (int i_ = 0; i_ < N; ++i_)

// r = r*u + K[I] on N values, descending I from I0+N-2 to I0
template <int N, int I0, int I, class H>
struct HornerN {
  static __device__ __forceinline__ void run(DV<N>& r, const DV<N>& u, const H& h)
  {
    const double k = h.template at<I0 + I>();
    UPC_FOR_N r.v[i_] = fma(r.v[i_], u.v[i_], k);
    HornerN<N, I0, I - 1, H>::run(r, u, h);
  }
};
template <int N, int I0, class H>
struct HornerN<N, I0, -1, H> {
  static __device__ __forceinline__ void run(DV<N>&, const DV<N>&, const H&) {}
};
template <int N, int I0, int LEN, class H>
__device__ __forceinline__ DV<N> hornerN(const DV<N>& u, const H& h)
{
  const double top = h.template at<I0 + LEN - 1>();
  DV<N> r;
  UPC_FOR_N r.v[i_] = top;
  HornerN<N, I0, LEN - 2, H>::run(r, u, h);
  return r;
}

// J1 on N arguments, all > 8: modulus/phase form
//     J1(x) = sqrt(2/(pi x)) M(u) sin(x - pi/4 + eps),  eps = T(u)/x,  u = 128/x^2 - 1   (|eps| <= 0.047),
// with the angle reduced in one go: n = rint((x + eps) 2/pi - 1/2), r = x - (n + 1/2) pi/2 + eps in
// [-pi/4, pi/4] (three-term Cody-Waite with FMAs, then + eps), sin(r + n pi/2) from the fdlibm minimax
// kernels of sin and cos on that interval.  1/x and 1/sqrt(x) come from one rsqrt.  ~75 FP64 instructions
// per value (the earlier form -- 1/x, sqrt, sincos(x), then sin/cos(eps) series and the addition theorem --
// took ~100); same ~1e-16 absolute accuracy relative to the amplitude as gsl_sf_bessel_J1.
template <int N, class H>
__device__ __forceinline__ DV<N> j1_largeN(const DV<N>& x, const H& h)
{
  DV<N> rs, rx, u;
  const double k128 = 128.;
  {
    // 1/sqrt(x): MUFU.RSQ64H seed (~2^-22) + one third-order step, the sequence rsqrt() itself uses, without
    // its special-case branch (x is positive and normal here; an argument that belongs to the other branch
    // of J1 may give inf/NaN in a lane whose value is discarded)
    const double k375 = h.template at<H_EPS + 0>();
    UPC_FOR_N {
      double y;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x.v[i_]));
      const double e = fma(-x.v[i_], y * y, 1.);
      rs.v[i_] = fma(fma(e, k375, 0.5), y * e, y);
    }
  }
  UPC_FOR_N {
    rx.v[i_] = rs.v[i_] * rs.v[i_];
    u.v[i_] = fma(k128 * rx.v[i_], rx.v[i_], -1.);
  }
  const DV<N> m = hornerN<N, H_J1M, UPC_J1_M_N>(u, h);
  const DV<N> t = hornerN<N, H_J1T, UPC_J1_T_N>(u, h);
  const double sqrt_two_over_pi = h.template at<H_MISC + 7>();
  const double kMagic = 6755399441055744.0;  // 2^52 + 2^51: low word of (w + magic) = rint(w)
  DV<N> ampl, eps, r, f;
  int n[N];
  {
    const double two_over_pi = h.template at<H_SC + 0>();
    UPC_FOR_N {
      ampl.v[i_] = m.v[i_] * (rs.v[i_] * sqrt_two_over_pi);
      eps.v[i_] = t.v[i_] * rx.v[i_];
      const double q = fma(x.v[i_] + eps.v[i_], two_over_pi, -0.5) + kMagic;
      n[i_] = __double2loint(q);
      f.v[i_] = (q - kMagic) + 0.5;          // n + 1/2, exact
    }
  }
  {
    const double p = h.template at<H_SC + 1>();
    UPC_FOR_N r.v[i_] = fma(-f.v[i_], p, x.v[i_]);
  }
  {
    const double p = h.template at<H_SC + 2>();
    UPC_FOR_N r.v[i_] = fma(-f.v[i_], p, r.v[i_]);
  }
  {
    const double p = h.template at<H_SC + 3>();
    UPC_FOR_N r.v[i_] = fma(-f.v[i_], p, r.v[i_]) + eps.v[i_];
  }
  DV<N> z;
  UPC_FOR_N z.v[i_] = r.v[i_] * r.v[i_];
  // the table stores S6..S1 and C6..C1 in evaluation order (ascending index)
  DV<N> sp, cp;
  {
    const double k6 = h.template at<H_SC + 4>(), k5 = h.template at<H_SC + 5>();
    UPC_FOR_N sp.v[i_] = fma(z.v[i_], k6, k5);
    const double k4 = h.template at<H_SC + 6>();
    UPC_FOR_N sp.v[i_] = fma(z.v[i_], sp.v[i_], k4);
    const double k3 = h.template at<H_SC + 7>();
    UPC_FOR_N sp.v[i_] = fma(z.v[i_], sp.v[i_], k3);
    const double k2 = h.template at<H_SC + 8>();
    UPC_FOR_N sp.v[i_] = fma(z.v[i_], sp.v[i_], k2);
    const double k1 = h.template at<H_SC + 9>();
    UPC_FOR_N sp.v[i_] = fma(z.v[i_], sp.v[i_], k1);
  }
  {
    const double k6 = h.template at<H_SC + 10>(), k5 = h.template at<H_SC + 11>();
    UPC_FOR_N cp.v[i_] = fma(z.v[i_], k6, k5);
    const double k4 = h.template at<H_SC + 12>();
    UPC_FOR_N cp.v[i_] = fma(z.v[i_], cp.v[i_], k4);
    const double k3 = h.template at<H_SC + 13>();
    UPC_FOR_N cp.v[i_] = fma(z.v[i_], cp.v[i_], k3);
    const double k2 = h.template at<H_SC + 14>();
    UPC_FOR_N cp.v[i_] = fma(z.v[i_], cp.v[i_], k2);
    const double k1 = h.template at<H_SC + 15>();
    UPC_FOR_N cp.v[i_] = fma(z.v[i_], cp.v[i_], k1);
  }
  DV<N> res;
  UPC_FOR_N {
    const double sr = fma(z.v[i_] * r.v[i_], sp.v[i_], r.v[i_]);
    const double cr = fma(z.v[i_] * z.v[i_], cp.v[i_], fma(z.v[i_], -0.5, 1.0));
    const double a = (n[i_] & 1) ? cr : sr;     // sin(r + n pi/2): sign flipped in quadrants 2 and 3
    res.v[i_] = ampl.v[i_] * __hiloint2double(__double2hiint(a) ^ ((n[i_] & 2) << 30), __double2loint(a));
  }
  return res;
}

// J1 on N arguments, all in [0, 8]
template <int N, class H>
__device__ __forceinline__ DV<N> j1_smallN(const DV<N>& x, const H& h)
{
  const double k32 = h.template at<H_MISC + 0>();
  DV<N> u;
  UPC_FOR_N u.v[i_] = fma(x.v[i_] * x.v[i_], k32, -1.);
  const DV<N> p = hornerN<N, H_J1P, UPC_J1_P_N>(u, h);
  DV<N> r;
  UPC_FOR_N r.v[i_] = x.v[i_] * p.v[i_];
  return r;
}

// J1 for three arbitrary positive arguments: each branch is evaluated for all three when any of them needs
// it, then selected per argument.  An argument outside a branch's domain gives a meaningless (possibly
// non-finite) value in that branch, which the selection drops: no clamping.  The class test is an integer
// comparison of the bit patterns (x >= 8 for positive x), off the FP64 pipe.
__device__ __forceinline__ bool j1_is_large(double x)
{
  // x >= 8 for positive x, on the high word alone (both branches of J1 are valid AT 8: the large-argument fit covers
  // u = 128 / x^2 - 1 in [-1, 1], closed).  One integer comparison; a 64-bit one makes ptxas form min/max chains, and
  // the strict "x > 8" needed a second comparison of the low word per argument.
  return (unsigned)__double2hiint(x) >= 0x40200000u;
}

__device__ __forceinline__ D3 j1_3(const D3& x)
{
  const HotConst h;
  const bool la = j1_is_large(x.v[0]), lb = j1_is_large(x.v[1]), lc = j1_is_large(x.v[2]);
  D3 r{{0., 0., 0.}};
  if (la || lb || lc) r = j1_largeN<3>(x, h);
  if (!(la && lb && lc)) {
    const D3 v = j1_smallN<3>(x, h);
    if (!la) r.v[0] = v.v[0];
    if (!lb) r.v[1] = v.v[1];
    if (!lc) r.v[2] = v.v[2];
  }
  return r;
}

}  // namespace upc
