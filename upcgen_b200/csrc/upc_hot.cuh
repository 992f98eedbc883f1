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
  H_SC = H_J1T + UPC_J1_T_N,         // 16: sincos (see kSinCosC)
  H_EPS = H_SC + 16,                 // 8 : sin/cos(eps) series
  H_MISC = H_EPS + 8,                // 8 : 1/32, 2/pi, 1/sqrt2, Q2min, 1/dQ2, dQ2, 64, pad
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
  // eps series: sin: 1/9!, -1/7!, 1/5!, -1/3!; cos: 1/8!, -1/6!, 1/4!, (pad)
  1. / 362880., -1. / 5040., 1. / 120., -1. / 6., 1. / 40320., -1. / 720., 1. / 24., 0.,
  // misc
  1. / 32., 0.63661977236758134308, 0.70710678118654752440, upc::kQ2min, 1. / upc::kDQ2, upc::kDQ2, 64., 0.};
}

namespace upc {

template <int I>
__device__ __forceinline__ double hot()
{
  double v;
  asm volatile("ld.const.f64 %0, [upc_hot+%1];" : "=d"(v) : "n"(I * 8));
  return v;
}

struct D3 {
  double a, b, c;
};

// r = r*u + K[I] on three lanes-in-a-thread, descending I from I0+N-2 to I0
template <int I0, int I>
struct Horner3 {
  static __device__ __forceinline__ void run(D3& r, const D3& u)
  {
    const double k = hot<I0 + I>();
    r.a = fma(r.a, u.a, k);
    r.b = fma(r.b, u.b, k);
    r.c = fma(r.c, u.c, k);
    Horner3<I0, I - 1>::run(r, u);
  }
};
template <int I0>
struct Horner3<I0, -1> {
  static __device__ __forceinline__ void run(D3&, const D3&) {}
};
template <int I0, int N>
__device__ __forceinline__ D3 horner3(const D3& u)
{
  const double top = hot<I0 + N - 1>();
  D3 r{top, top, top};
  Horner3<I0, N - 2>::run(r, u);
  return r;
}

// sin/cos of three moderate arguments (see sincos_mid)
__device__ __forceinline__ void sincos3(const D3& x, D3& s, D3& c)
{
  const double kMagic = 6755399441055744.0;
  const double two_over_pi = hot<H_SC + 0>();
  const double qa = fma(x.a, two_over_pi, kMagic), qb = fma(x.b, two_over_pi, kMagic), qc = fma(x.c, two_over_pi, kMagic);
  const int na = __double2loint(qa), nb = __double2loint(qb), nc = __double2loint(qc);
  const double fa = qa - kMagic, fb = qb - kMagic, fc = qc - kMagic;
  double ra, rb, rc;
  {
    const double p = hot<H_SC + 1>();
    ra = fma(-fa, p, x.a); rb = fma(-fb, p, x.b); rc = fma(-fc, p, x.c);
  }
  {
    const double p = hot<H_SC + 2>();
    ra = fma(-fa, p, ra); rb = fma(-fb, p, rb); rc = fma(-fc, p, rc);
  }
  {
    const double p = hot<H_SC + 3>();
    ra = fma(-fa, p, ra); rb = fma(-fb, p, rb); rc = fma(-fc, p, rc);
  }
  const D3 z{ra * ra, rb * rb, rc * rc};
  // the table stores S6..S1 and C6..C1 in evaluation order (ascending index)
  D3 sp, cp;
  {
    const double k6 = hot<H_SC + 4>(), k5 = hot<H_SC + 5>();
    sp.a = fma(z.a, k6, k5); sp.b = fma(z.b, k6, k5); sp.c = fma(z.c, k6, k5);
    const double k4 = hot<H_SC + 6>();
    sp.a = fma(z.a, sp.a, k4); sp.b = fma(z.b, sp.b, k4); sp.c = fma(z.c, sp.c, k4);
    const double k3 = hot<H_SC + 7>();
    sp.a = fma(z.a, sp.a, k3); sp.b = fma(z.b, sp.b, k3); sp.c = fma(z.c, sp.c, k3);
    const double k2 = hot<H_SC + 8>();
    sp.a = fma(z.a, sp.a, k2); sp.b = fma(z.b, sp.b, k2); sp.c = fma(z.c, sp.c, k2);
    const double k1 = hot<H_SC + 9>();
    sp.a = fma(z.a, sp.a, k1); sp.b = fma(z.b, sp.b, k1); sp.c = fma(z.c, sp.c, k1);
  }
  {
    const double k6 = hot<H_SC + 10>(), k5 = hot<H_SC + 11>();
    cp.a = fma(z.a, k6, k5); cp.b = fma(z.b, k6, k5); cp.c = fma(z.c, k6, k5);
    const double k4 = hot<H_SC + 12>();
    cp.a = fma(z.a, cp.a, k4); cp.b = fma(z.b, cp.b, k4); cp.c = fma(z.c, cp.c, k4);
    const double k3 = hot<H_SC + 13>();
    cp.a = fma(z.a, cp.a, k3); cp.b = fma(z.b, cp.b, k3); cp.c = fma(z.c, cp.c, k3);
    const double k2 = hot<H_SC + 14>();
    cp.a = fma(z.a, cp.a, k2); cp.b = fma(z.b, cp.b, k2); cp.c = fma(z.c, cp.c, k2);
    const double k1 = hot<H_SC + 15>();
    cp.a = fma(z.a, cp.a, k1); cp.b = fma(z.b, cp.b, k1); cp.c = fma(z.c, cp.c, k1);
  }
  const double sra = fma(z.a * ra, sp.a, ra), srb = fma(z.b * rb, sp.b, rb), src = fma(z.c * rc, sp.c, rc);
  const double cra = fma(z.a * z.a, cp.a, fma(z.a, -0.5, 1.0)), crb = fma(z.b * z.b, cp.b, fma(z.b, -0.5, 1.0)),
               crc = fma(z.c * z.c, cp.c, fma(z.c, -0.5, 1.0));
  auto quad = [](int n, double sr, double cr, double& so, double& co) {
    const double a = (n & 1) ? cr : sr;
    const double b = (n & 1) ? sr : cr;
    so = (n & 2) ? -a : a;
    co = ((n + 1) & 2) ? -b : b;
  };
  quad(na, sra, cra, s.a, c.a);
  quad(nb, srb, crb, s.b, c.b);
  quad(nc, src, crc, s.c, c.c);
}

// J1 on three arguments, all > 8 (modulus/phase form; see bessel_j1)
__device__ __forceinline__ D3 j1_large3(const D3& x)
{
  const D3 rx{1. / x.a, 1. / x.b, 1. / x.c};
  const double k64 = 64.;
  const D3 u{fma(2. * k64 * rx.a, rx.a, -1.), fma(2. * k64 * rx.b, rx.b, -1.), fma(2. * k64 * rx.c, rx.c, -1.)};
  const D3 m = horner3<H_J1M, UPC_J1_M_N>(u);
  const D3 t = horner3<H_J1T, UPC_J1_T_N>(u);
  const double two_over_pi = hot<H_MISC + 1>();
  const D3 ampl{m.a * sqrt(two_over_pi * rx.a), m.b * sqrt(two_over_pi * rx.b), m.c * sqrt(two_over_pi * rx.c)};
  const D3 eps{t.a * rx.a, t.b * rx.b, t.c * rx.c};
  D3 sy, cy;
  sincos3(x, sy, cy);
  const D3 e2{eps.a * eps.a, eps.b * eps.b, eps.c * eps.c};
  D3 se, ce;
  {
    const double k9 = hot<H_EPS + 0>(), k7 = hot<H_EPS + 1>();
    se.a = fma(e2.a, k9, k7); se.b = fma(e2.b, k9, k7); se.c = fma(e2.c, k9, k7);
    const double k5 = hot<H_EPS + 2>();
    se.a = fma(e2.a, se.a, k5); se.b = fma(e2.b, se.b, k5); se.c = fma(e2.c, se.c, k5);
    const double k3 = hot<H_EPS + 3>();
    se.a = fma(e2.a, se.a, k3); se.b = fma(e2.b, se.b, k3); se.c = fma(e2.c, se.c, k3);
    se.a = eps.a * fma(e2.a, se.a, 1.); se.b = eps.b * fma(e2.b, se.b, 1.); se.c = eps.c * fma(e2.c, se.c, 1.);
    const double k8 = hot<H_EPS + 4>(), k6 = hot<H_EPS + 5>();
    ce.a = fma(e2.a, k8, k6); ce.b = fma(e2.b, k8, k6); ce.c = fma(e2.c, k8, k6);
    const double k4 = hot<H_EPS + 6>();
    ce.a = fma(e2.a, ce.a, k4); ce.b = fma(e2.b, ce.b, k4); ce.c = fma(e2.c, ce.c, k4);
    ce.a = fma(e2.a, fma(e2.a, ce.a, -0.5), 1.); ce.b = fma(e2.b, fma(e2.b, ce.b, -0.5), 1.);
    ce.c = fma(e2.c, fma(e2.c, ce.c, -0.5), 1.);
  }
  const double inv_sqrt2 = hot<H_MISC + 2>();
  D3 r;
  r.a = ampl.a * fma(ce.a, sy.a - cy.a, se.a * (sy.a + cy.a)) * inv_sqrt2;
  r.b = ampl.b * fma(ce.b, sy.b - cy.b, se.b * (sy.b + cy.b)) * inv_sqrt2;
  r.c = ampl.c * fma(ce.c, sy.c - cy.c, se.c * (sy.c + cy.c)) * inv_sqrt2;
  return r;
}

// J1 on three arguments, all in [0, 8]
__device__ __forceinline__ D3 j1_small3(const D3& x)
{
  const double k32 = hot<H_MISC + 0>();
  const D3 u{fma(x.a * x.a, k32, -1.), fma(x.b * x.b, k32, -1.), fma(x.c * x.c, k32, -1.)};
  const D3 p = horner3<H_J1P, UPC_J1_P_N>(u);
  return D3{x.a * p.a, x.b * p.b, x.c * p.c};
}

// J1 for three arbitrary non-negative arguments: each branch is evaluated for all three when any
// of them needs it (arguments clamped into the branch's domain), then selected per argument.
__device__ __forceinline__ D3 j1_3(const D3& x)
{
  const bool la = x.a > 8., lb = x.b > 8., lc = x.c > 8.;
  D3 r{0., 0., 0.};
  if (la | lb | lc) {
    const D3 v = j1_large3(D3{fmax(x.a, 8.), fmax(x.b, 8.), fmax(x.c, 8.)});
    r = v;
  }
  if (!(la & lb & lc)) {
    const D3 v = j1_small3(D3{fmin(x.a, 8.), fmin(x.b, 8.), fmin(x.c, 8.)});
    if (!la) r.a = v.a;
    if (!lb) r.b = v.b;
    if (!lc) r.c = v.c;
  }
  return r;
}

}  // namespace upc
