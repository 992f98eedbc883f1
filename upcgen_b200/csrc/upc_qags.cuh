// upc_qags.cuh -- device QAGS: a per-lane state machine that follows gsl_integration_qags
// (QUADPACK dqagse: GK21 rule, bisection of the worst interval, Wynn epsilon extrapolation)
// decision for decision, because the reference's form-factor flux IS the un-converged result
// of that algorithm at epsabs = epsrel = 1e-4 (src/UpcCrossSection.cpp:206-212).
//
// The adaptive loop is cut into   pre() -> 1 or 2 GK21 evaluations -> post()   so that a warp
// whose lanes are at different stages of different integrals still executes the expensive
// part (the integrand evaluations) convergently; finished lanes pull the next integral from a
// global queue (see flux_rows_kernel in upc_kernels.cu).
#pragma once
#include <float.h>

#include "upc_math.cuh"

namespace upc {

// GK21 nodes/weights (QUADPACK dqk21).  Stored pair-wise in the order GSL's qk() visits them:
// first the 5 Gauss nodes (xgk[1],xgk[3],..,xgk[9]) then the 5 Kronrod-only nodes
// (xgk[0],xgk[2],..,xgk[8]), so that the partial sums are formed in the same order; entry 10 is
// the centre (abscissa 0).  Gauss weights are 0 for the Kronrod-only nodes.
__constant__ double kGkX[11] = {
  0.973906528517171720077964012084452, 0.865063366688984510732096688423493,
  0.679409568299024406234327365114874, 0.433395394129247190799265943165784,
  0.148874338981631210884826001129720,
  0.995657163025808080735527280689003, 0.930157491355708226001207180059508,
  0.780817726586416897063717578345042, 0.562757134668604683339000099272694,
  0.294392862701460198131126603103866, 0.0};
__constant__ double kGkWk[11] = {
  0.032558162307964727478818972459390, 0.075039674810919952767043140916190,
  0.109387158802297641899210590325805, 0.134709217311473325928054001771707,
  0.147739104901338491374841515972068,
  0.011694638867371874278064396062192, 0.054755896574351996031381300244580,
  0.093125454583697605535065465083366, 0.123491976262065851077958109585166,
  0.142775938577060080797094273138717, 0.149445554002916905664936468389821};
__constant__ double kGkWg[10] = {
  0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
  0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
  0.295524224714752870173815619188769, 0., 0., 0., 0., 0.};
constexpr double kGkWkC = 0.149445554002916905664936468389821;

struct GkOut {
  double result, abserr, resabs, resasc;
};

// signed abscissas of the 21 nodes in storage order: node 2p = -x_p, node 2p+1 = +x_p (pairs p in
// the order of kGkX), node 20 = centre
__constant__ double kGkNode[21] = {
  -0.973906528517171720077964012084452, 0.973906528517171720077964012084452,
  -0.865063366688984510732096688423493, 0.865063366688984510732096688423493,
  -0.679409568299024406234327365114874, 0.679409568299024406234327365114874,
  -0.433395394129247190799265943165784, 0.433395394129247190799265943165784,
  -0.148874338981631210884826001129720, 0.148874338981631210884826001129720,
  -0.995657163025808080735527280689003, 0.995657163025808080735527280689003,
  -0.930157491355708226001207180059508, 0.930157491355708226001207180059508,
  -0.780817726586416897063717578345042, 0.780817726586416897063717578345042,
  -0.562757134668604683339000099272694, 0.562757134668604683339000099272694,
  -0.294392862701460198131126603103866, 0.294392862701460198131126603103866, 0.0};

// gsl_integration_qk (integration/qk.c) specialised to the 21-point rule, with the integrand evaluated three nodes
// at a time (F::tri) -- 7 trips,
// no redundant evaluation -- into fv[0..20] (node order of kGkNode), followed by the weighted
// sums in GSL's order.  fv: per-thread scratch of 21 doubles at fv[n * fv_stride].
template <class F>
__device__ __forceinline__ GkOut gk21_tri(const F& f, double a, double b, double* fv, int fv_stride)
{
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
#pragma unroll 1
  for (int it = 0; it < 7; ++it) {
    const int n = 3 * it;
    double f0, f1, f2;
    f.tri(fma(half_length, kGkNode[n], center), fma(half_length, kGkNode[n + 1], center),
          fma(half_length, kGkNode[n + 2], center), f0, f1, f2);
    fv[n * fv_stride] = f0;
    fv[(n + 1) * fv_stride] = f1;
    fv[(n + 2) * fv_stride] = f2;
  }
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
    double scale = s * sqrt(s);  // pow(s, 1.5)
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

// The same sums fully unrolled (weights and indices are compile-time), values at fv[n * STRIDE].  The
// Gauss sum runs over the 5 Gauss pairs only, as in qk.c (the zero weights above add +0.0: the same
// value).
template <int STRIDE>
__device__ __forceinline__ GkOut gk21_sums_strided(const double* fv, double half_length)
{
  double f[21];
#pragma unroll
  for (int n = 0; n < 21; ++n) f[n] = fv[n * STRIDE];
  const double f_center = f[20];
  double result_gauss = 0;
  double result_kronrod = f_center * kGkWkC;
  double result_abs = fabs(result_kronrod);
#pragma unroll
  for (int p = 0; p < 10; ++p) {
    const double fsum = f[2 * p] + f[2 * p + 1];
    if (p < 5) result_gauss += kGkWg[p] * fsum;
    result_kronrod += kGkWk[p] * fsum;
    result_abs += kGkWk[p] * (fabs(f[2 * p]) + fabs(f[2 * p + 1]));
  }
  const double mean = result_kronrod * 0.5;
  double result_asc = kGkWkC * fabs(f_center - mean);
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    const int p = (j & 1) ? (j >> 1) : (5 + (j >> 1));
    result_asc += kGkWk[p] * (fabs(f[2 * p] - mean) + fabs(f[2 * p + 1] - mean));
  }
  const double abs_half_length = fabs(half_length);
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

// Interval-list / epsilon-table storage.  The QAGS logic below is written against this small
// interface so that the same code runs on in-thread arrays (test hooks), on a global-memory
// workspace of the reference's full size (overflow pass) and on the strided shared-memory state of
// the row-cooperative kernel (upc_qags_rows.cuh):
//   R(k), E(k)          result / error estimate of interval k          (rlist, elist)
//   ord(k), set_ord     the error-sorted permutation                   (order)
//   lvl(k)              bisection level of interval k                  (level)
//   get_iv / set_iv     bounds of interval k; set_iv returns false when the interval cannot be
//                       represented (treated like running out of capacity)
//   eps(k)              Wynn epsilon table entries                     (rlist2)
//   sc(k), k < 11       the FP64 scalars of the driver (area, errsum, res_ext, ...)
//   cap, eps_cap        capacities; exceeding either sets `overflow`
template <int CAP>
struct QagsLocalStore {
  double alist_[CAP], blist_[CAP], rlist_[CAP], elist_[CAP];
  short order_[CAP], level_[CAP];
  double eps_[52];
  double sc_[11];
  static constexpr int cap = CAP;
  static constexpr int eps_cap = 52;
  __device__ __forceinline__ double& R(int k) { return rlist_[k]; }
  __device__ __forceinline__ double& E(int k) { return elist_[k]; }
  __device__ __forceinline__ int ord(int k) const { return order_[k]; }
  __device__ __forceinline__ void set_ord(int k, int v) { order_[k] = (short)v; }
  __device__ __forceinline__ int lvl(int k) const { return level_[k]; }
  __device__ __forceinline__ void get_iv(int k, double& a, double& b) const { a = alist_[k]; b = blist_[k]; }
  __device__ __forceinline__ bool set_iv(int k, double a, double b, int level)
  {
    alist_[k] = a; blist_[k] = b; level_[k] = (short)level;
    return true;
  }
  __device__ __forceinline__ double& eps(int k) { return eps_[k]; }
  __device__ __forceinline__ double& sc(int k) { return sc_[k]; }
};
// caller-provided global-memory workspace of the reference's full size (overflow pass):
// 4 x 1000 doubles, 2 x 1000 shorts, 52 doubles
struct QagsGlobalStore {
  double *alist_, *blist_, *rlist_, *elist_, *eps_;
  short *order_, *level_;
  double sc_[11];
  static constexpr int cap = 1000;
  static constexpr int eps_cap = 52;
  __device__ __forceinline__ double& R(int k) { return rlist_[k]; }
  __device__ __forceinline__ double& E(int k) { return elist_[k]; }
  __device__ __forceinline__ int ord(int k) const { return order_[k]; }
  __device__ __forceinline__ void set_ord(int k, int v) { order_[k] = (short)v; }
  __device__ __forceinline__ int lvl(int k) const { return level_[k]; }
  __device__ __forceinline__ void get_iv(int k, double& a, double& b) const { a = alist_[k]; b = blist_[k]; }
  __device__ __forceinline__ bool set_iv(int k, double a, double b, int level)
  {
    alist_[k] = a; blist_[k] = b; level_[k] = (short)level;
    return true;
  }
  __device__ __forceinline__ double& eps(int k) { return eps_[k]; }
  __device__ __forceinline__ double& sc(int k) { return sc_[k]; }
};

// Per-lane QAGS state (integration/qags.c + qpsrt.c, qelg.c, util.c of GSL; QUADPACK dqagse).
// `kLimit` keeps the reference's value (1000) in all the places where the algorithm's decisions
// depend on it (qpsrt's `top`, increase_nrmax's `jupbnd`, the iteration cap); running out of the
// store's capacity sets `overflow` and the integral is redone by a pass with the full workspace.
template <class Store>
struct Qags : Store {
  static constexpr int kLimit = 1000;
  using Store::R; using Store::E; using Store::ord; using Store::set_ord; using Store::lvl;
  using Store::get_iv; using Store::set_iv; using Store::eps; using Store::sc; using Store::cap; using Store::eps_cap;
  int size, nrmax, i, maximum_level;
  // driver state (integration/qags.c)
  // the reference's call: gsl_integration_qags(..., epsabs = 1e-4, epsrel = 1e-4, limit = 1000, ...),
  // src/UpcCrossSection.cpp:209
  static constexpr double epsabs = 1e-4, epsrel = 1e-4;
  // the FP64 scalars of the driver live in the store (sc(0..10)): in the row-cooperative kernel
  // that is shared memory, so that they do not occupy registers across the evaluation phases
  __device__ __forceinline__ double& area() { return sc(0); }
  __device__ __forceinline__ double& errsum() { return sc(1); }
  __device__ __forceinline__ double& res_ext() { return sc(2); }
  __device__ __forceinline__ double& err_ext() { return sc(3); }
  __device__ __forceinline__ double& resabs0() { return sc(4); }
  __device__ __forceinline__ double& ertest() { return sc(5); }
  __device__ __forceinline__ double& error_over_large_intervals() { return sc(6); }
  __device__ __forceinline__ double& correc() { return sc(7); }
  __device__ __forceinline__ double& res3la0() { return sc(8); }
  __device__ __forceinline__ double& res3la1() { return sc(9); }
  __device__ __forceinline__ double& res3la2() { return sc(10); }
  int ktmin, roundoff_type1, roundoff_type2, roundoff_type3, error_type, error_type2, iteration;
  bool positive_integrand, extrapolate, disallow_extrapolation, overflow;
  // Wynn epsilon table (integration/qelg.c): entries in the store, bookkeeping here
  int tab_n, tab_nres;
  // outputs
  double result, abserr;
  int ier, neval;

  __device__ __forceinline__ void begin(double a, double b)
  {
    size = 0; nrmax = 0; i = 0; maximum_level = 0;
    set_iv(0, a, b, 0); R(0) = 0; E(0) = 0; set_ord(0, 0);
    ertest() = 0; error_over_large_intervals() = 0; correc() = 0;
    ktmin = 0; roundoff_type1 = roundoff_type2 = roundoff_type3 = 0;
    error_type = 0; error_type2 = 0; iteration = 0;
    positive_integrand = false; extrapolate = false; disallow_extrapolation = false; overflow = false;
    tab_n = 0; tab_nres = 0; res3la0() = res3la1() = res3la2() = 0;
    area() = errsum() = res_ext() = err_ext() = resabs0() = 0;
    result = 0; abserr = 0; ier = 0; neval = 0;
  }

  // append to the epsilon table; false when the store cannot hold what qelg will touch
  __device__ __forceinline__ bool eps_push(double v)
  {
    if (tab_n + 2 >= eps_cap) return false;  // qelg writes entry n + 2 (n = index of the new one)
    eps(tab_n++) = v;
    return true;
  }

  // integration/qelg.c
  __device__ __forceinline__ void qelg(double& result_, double& abserr_)
  {
    const int n = tab_n - 1;
    const double current = eps(n);
    double absolute = DBL_MAX;
    double relative = 5 * DBL_EPSILON * fabs(current);
    const int newelm = n / 2;
    const int n_orig = n;
    int n_final = n;
    const int nres_orig = tab_nres;
    result_ = current;
    abserr_ = DBL_MAX;
    if (n < 2) {
      result_ = current;
      abserr_ = fmax(absolute, relative);
      return;
    }
    eps(n + 2) = eps(n);
    eps(n) = DBL_MAX;
    for (int ii = 0; ii < newelm; ii++) {
      double res = eps(n - 2 * ii + 2);
      double e0 = eps(n - 2 * ii - 2);
      double e1 = eps(n - 2 * ii - 1);
      double e2 = res;
      double e1abs = fabs(e1);
      double delta2 = e2 - e1;
      double err2 = fabs(delta2);
      double tol2 = fmax(fabs(e2), e1abs) * DBL_EPSILON;
      double delta3 = e1 - e0;
      double err3 = fabs(delta3);
      double tol3 = fmax(e1abs, fabs(e0)) * DBL_EPSILON;
      if (err2 <= tol2 && err3 <= tol3) {
        result_ = res;
        absolute = err2 + err3;
        relative = 5 * DBL_EPSILON * fabs(res);
        abserr_ = fmax(absolute, relative);
        return;
      }
      double e3 = eps(n - 2 * ii);
      eps(n - 2 * ii) = e1;
      double delta1 = e1 - e3;
      double err1 = fabs(delta1);
      double tol1 = fmax(e1abs, fabs(e3)) * DBL_EPSILON;
      if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
        n_final = 2 * ii;
        break;
      }
      double ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
      if (fabs(ss * e1) <= 0.0001) {
        n_final = 2 * ii;
        break;
      }
      res = e1 + 1 / ss;
      eps(n - 2 * ii) = res;
      {
        const double error = err2 + fabs(res - e2) + err3;
        if (error <= abserr_) {
          abserr_ = error;
          result_ = res;
        }
      }
    }
    {
      const int limexp = 50 - 1;
      if (n_final == limexp) n_final = 2 * (limexp / 2);
    }
    if (n_orig % 2 == 1) {
      for (int ii = 0; ii <= newelm; ii++) eps(1 + ii * 2) = eps(ii * 2 + 3);
    } else {
      for (int ii = 0; ii <= newelm; ii++) eps(ii * 2) = eps(ii * 2 + 2);
    }
    if (n_orig != n_final) {
      for (int ii = 0; ii <= n_final; ii++) eps(ii) = eps(n_orig - n_final + ii);
    }
    tab_n = n_final + 1;
    if (nres_orig < 3) {
      if (nres_orig == 0) res3la0() = result_;
      else if (nres_orig == 1) res3la1() = result_;
      else res3la2() = result_;
      abserr_ = DBL_MAX;
    } else {
      abserr_ = (fabs(result_ - res3la2()) + fabs(result_ - res3la1()) + fabs(result_ - res3la0()));
      res3la0() = res3la1();
      res3la1() = res3la2();
      res3la2() = result_;
    }
    tab_nres = nres_orig + 1;
    abserr_ = fmax(abserr_, 5 * DBL_EPSILON * fabs(result_));
  }

  // after the first GK21 on [a0,b0]; returns true when the integral is finished
  __device__ __forceinline__ bool post_first(const GkOut& g)
  {
    neval += 21;
    size = 1; R(0) = g.result; E(0) = g.abserr;
    resabs0() = g.resabs;
    const double tolerance = fmax(epsabs, epsrel * fabs(g.result));
    if (g.abserr <= 100 * DBL_EPSILON * g.resabs && g.abserr > tolerance) {
      result = g.result; abserr = g.abserr; ier = 18;  // GSL_EROUND
      return true;
    } else if ((g.abserr <= tolerance && g.abserr != g.resasc) || g.abserr == 0.0) {
      result = g.result; abserr = g.abserr; ier = 0;
      return true;
    }
    tab_n = 0; tab_nres = 0;
    eps_push(g.result);
    area() = g.result;
    errsum() = g.abserr;
    res_ext() = g.result;
    err_ext() = DBL_MAX;
    positive_integrand = (fabs(g.result) >= (1 - 50 * DBL_EPSILON) * g.resabs);
    iteration = 1;
    return false;
  }

  // choose the interval to bisect (top of the do-loop in qags()): [a1,b1] and [a2,b2] are its
  // halves, `level` their bisection level
  __device__ __forceinline__ void pre_step(double& a1, double& b1, double& a2, double& b2, int& level)
  {
    double a_i, b_i;
    get_iv(i, a_i, b_i);
    level = lvl(i) + 1;
    a1 = a_i; b1 = 0.5 * (a_i + b_i); a2 = b1; b2 = b_i;
    iteration++;
  }

  // integration/qpsrt.c
  __device__ __forceinline__ void qpsrt()
  {
    const int last = size - 1;
    const int limit = kLimit;
    int i_nrmax = nrmax;
    int i_maxerr = ord(i_nrmax);
    if (last < 2) {
      set_ord(0, 0); set_ord(1, 1);
      i = i_maxerr;
      return;
    }
    double errmax = E(i_maxerr);
    while (i_nrmax > 0 && errmax > E(ord(i_nrmax - 1))) {
      set_ord(i_nrmax, ord(i_nrmax - 1));
      i_nrmax--;
    }
    int top = (last < (limit / 2 + 2)) ? last : limit - last + 1;
    int ii = i_nrmax + 1;
    while (ii < top && errmax < E(ord(ii))) {
      set_ord(ii - 1, ord(ii));
      ii++;
    }
    set_ord(ii - 1, i_maxerr);
    double errmin = E(last);
    int k = top - 1;
    while (k > ii - 2 && errmin >= E(ord(k))) {
      set_ord(k + 1, ord(k));
      k--;
    }
    set_ord(k + 1, last);
    i_maxerr = ord(i_nrmax);
    i = i_maxerr;
    nrmax = i_nrmax;
  }

  // integration/util.c update(): the half with the larger error stays at i_max
  __device__ __forceinline__ void update(double a1, double b1, double a2, double b2, int current_level, double area1,
                                         double error1, double area2, double error2)
  {
    const int i_max = i;
    const int i_new = size;
    const int new_level = current_level;  // = level[i_max] + 1
    bool ok;
    if (error2 > error1) {
      ok = set_iv(i_max, a2, b2, new_level);
      R(i_max) = area2; E(i_max) = error2;
      ok &= set_iv(i_new, a1, b1, new_level);
      R(i_new) = area1; E(i_new) = error1;
    } else {
      ok = set_iv(i_max, a1, b1, new_level);
      R(i_max) = area1; E(i_max) = error1;
      ok &= set_iv(i_new, a2, b2, new_level);
      R(i_new) = area2; E(i_new) = error2;
    }
    if (!ok) overflow = true;
    size++;
    if (new_level > maximum_level) maximum_level = new_level;
    qpsrt();
  }

  __device__ __forceinline__ bool increase_nrmax()
  {
    int id = nrmax;
    int last = size - 1;
    int jupbnd = (last > (1 + kLimit / 2)) ? kLimit + 1 - last : last;
    for (int k = id; k <= jupbnd; k++) {
      int i_max = ord(nrmax);
      i = i_max;
      if (lvl(i_max) < maximum_level) return true;
      nrmax++;
    }
    return false;
  }

  // body of the do-loop after the two GK21 evaluations; returns true when finished
  // (result/abserr/ier set)
  __device__ __forceinline__ bool post_step(const GkOut& g1, const GkOut& g2)
  {
    const int act = step_core(g1, g2);
    if (act == 0) return false;
    return finish(act == 1);  // the only call site: everything stays in registers
  }

  // returns 0: continue, 1: finish with the plain sum, 2: finish with the extrapolated value
  __device__ __forceinline__ int step_core(const GkOut& g1, const GkOut& g2)
  {
    neval += 42;
    // the interval being bisected (chosen by pre_step; nothing has touched the list since)
    double a1, b2;
    get_iv(i, a1, b2);
    const double b1 = 0.5 * (a1 + b2), a2 = b1;
    const double r_i = R(i), e_i = E(i);
    const int current_level = lvl(i) + 1;
    const double area1 = g1.result, area2 = g2.result;
    const double error1 = g1.abserr, error2 = g2.abserr;
    const double area12 = area1 + area2;
    const double error12 = error1 + error2;
    const double last_e_i = e_i;
    errsum() = errsum() + error12 - e_i;
    area() = area() + area12 - r_i;
    const double tolerance = fmax(epsabs, epsrel * fabs(area()));
    if (g1.resasc != error1 && g2.resasc != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) {
        if (!extrapolate) roundoff_type1++; else roundoff_type2++;
      }
      if (iteration > 10 && error12 > e_i) roundoff_type3++;
    }
    if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error_type = 2;
    if (roundoff_type2 >= 5) error_type2 = 1;
    {
      double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
      if (fabs(a1) <= tmp && fabs(b2) <= tmp) error_type = 4;
    }
    update(a1, b1, a2, b2, current_level, area1, error1, area2, error2);

    if (errsum() <= tolerance) return 1;
    if (error_type) return 2;
    if (iteration >= kLimit - 1) { error_type = 1; return 2; }
    if (overflow || size >= cap) { overflow = true; error_type = 1; return 2; }
    if (iteration == 2) {
      error_over_large_intervals() = errsum();
      ertest() = tolerance;
      if (!eps_push(area())) { overflow = true; error_type = 1; return 2; }
      return 0;
    }
    if (disallow_extrapolation) return 0;
    error_over_large_intervals() += -last_e_i;
    if (current_level < maximum_level) error_over_large_intervals() += error12;
    if (!extrapolate) {
      if (lvl(i) < maximum_level) return 0;  // large_interval()
      extrapolate = true;
      nrmax = 1;
    }
    if (!error_type2 && error_over_large_intervals() > ertest()) {
      if (increase_nrmax()) return 0;
    }
    if (!eps_push(area())) { overflow = true; error_type = 1; return 2; }
    double reseps, abseps;
    qelg(reseps, abseps);
    ktmin++;
    if (ktmin > 5 && err_ext() < 0.001 * errsum()) error_type = 5;
    if (abseps < err_ext()) {
      ktmin = 0;
      err_ext() = abseps;
      res_ext() = reseps;
      correc() = error_over_large_intervals();
      ertest() = fmax(epsabs, epsrel * fabs(reseps));
      if (err_ext() <= ertest()) return 2;
    }
    if (tab_n == 1) disallow_extrapolation = true;
    if (error_type == 5) return 2;
    nrmax = 0; i = ord(0);  // reset_nrmax
    extrapolate = false;
    error_over_large_intervals() = errsum();
    return 0;
  }

  // tail of qags(): choice between the extrapolated value and the plain sum
  __device__ __forceinline__ bool finish(bool direct_sum)
  {
    bool compute = direct_sum;
    bool ret_err = false;
    if (!compute) {
      result = res_ext();
      abserr = err_ext();
      if (err_ext() == DBL_MAX) {
        compute = true;
      } else {
        if (error_type || error_type2) {
          if (error_type2) err_ext() += correc();
          if (error_type == 0) error_type = 3;
          if (res_ext() != 0.0 && area() != 0.0) {
            if (err_ext() / fabs(res_ext()) > errsum() / fabs(area())) compute = true;
          } else if (err_ext() > errsum()) {
            compute = true;
          } else if (area() == 0.0) {
            ret_err = true;
          }
        }
        if (!compute && !ret_err) {
          double max_area = fmax(fabs(res_ext()), fabs(area()));
          if (!positive_integrand && max_area < 0.01 * resabs0()) {
            ret_err = true;
          } else {
            double ratio = res_ext() / area();
            if (ratio < 0.01 || ratio > 100.0 || errsum() > fabs(area())) error_type = 6;
          }
        }
      }
    }
    if (compute) {
      double s = 0;
      for (int k = 0; k < size; k++) s += R(k);
      result = s;
      abserr = errsum();
    }
    if (error_type > 2) error_type--;
    ier = error_type;
    return true;
  }
};

}  // namespace upc
