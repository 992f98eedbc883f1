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

// gsl_integration_qk (integration/qk.c) specialised to the 21-point rule.
//   F provides  void pair(double x1, double x2, double& f1, double& f2)  -- two integrand
//   evaluations through ONE inlined call site (two interleaved instruction streams for ILP, and
//   a kernel whose hot loop stays small enough for the instruction cache).
//   fv: per-thread scratch of 20 doubles at fv[j * fv_stride] (shared memory, lane-interleaved).
// Summation order: the centre term first, then the pairs in GSL's order; result_asc in GSL's
// index order.  The centre is evaluated by the same pair() site (its second slot re-evaluates
// the centre: one redundant evaluation in 22).
template <class F>
__device__ __forceinline__ GkOut gk21(const F& f, double a, double b, double* fv, int fv_stride)
{
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  double f_center = 0, result_gauss = 0, result_kronrod = 0, result_abs = 0;
#pragma unroll 1
  for (int it = 0; it < 11; ++it) {
    const int p = it == 0 ? 10 : it - 1;  // visiting order: centre (10), then pairs 0..9
    const double abscissa = half_length * kGkX[p];
    double fval1, fval2;
    f.pair(center - abscissa, center + abscissa, fval1, fval2);
    if (p == 10) {
      f_center = fval1;
      result_kronrod = f_center * kGkWkC;
      result_abs = fabs(result_kronrod);
    } else {
      const double fsum = fval1 + fval2;
      fv[(2 * p) * fv_stride] = fval1;
      fv[(2 * p + 1) * fv_stride] = fval2;
      result_gauss += kGkWg[p] * fsum;
      result_kronrod += kGkWk[p] * fsum;
      result_abs += kGkWk[p] * (fabs(fval1) + fabs(fval2));
    }
  }
  const double mean = result_kronrod * 0.5;
  double result_asc = kGkWkC * fabs(f_center - mean);
  // GSL sums j = 0..9 in xgk index order: j even -> pair 5 + j/2, j odd -> pair (j-1)/2
#pragma unroll 1
  for (int j = 0; j < 10; ++j) {
    const int p = (j & 1) ? (j >> 1) : (5 + (j >> 1));
    result_asc += kGkWk[p] * (fabs(fv[(2 * p) * fv_stride] - mean) + fabs(fv[(2 * p + 1) * fv_stride] - mean));
  }
  double err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  // rescale_error
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

// The same 21-point rule with the integrand evaluated three nodes at a time (F::tri) -- 7 trips,
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

// Wynn epsilon table (integration/qelg.c)
struct EpsTable {
  int n;
  int nres;
  double rlist2[52];
  double res3la[3];
};

__device__ __noinline__ void qelg(EpsTable& table, double& result, double& abserr)
{
  double* epstab = table.rlist2;
  double* res3la = table.res3la;
  const int n = table.n - 1;
  const double current = epstab[n];
  double absolute = DBL_MAX;
  double relative = 5 * DBL_EPSILON * fabs(current);
  const int newelm = n / 2;
  const int n_orig = n;
  int n_final = n;
  const int nres_orig = table.nres;
  result = current;
  abserr = DBL_MAX;
  if (n < 2) {
    result = current;
    abserr = fmax(absolute, relative);
    return;
  }
  epstab[n + 2] = epstab[n];
  epstab[n] = DBL_MAX;
  for (int i = 0; i < newelm; i++) {
    double res = epstab[n - 2 * i + 2];
    double e0 = epstab[n - 2 * i - 2];
    double e1 = epstab[n - 2 * i - 1];
    double e2 = res;
    double e1abs = fabs(e1);
    double delta2 = e2 - e1;
    double err2 = fabs(delta2);
    double tol2 = fmax(fabs(e2), e1abs) * DBL_EPSILON;
    double delta3 = e1 - e0;
    double err3 = fabs(delta3);
    double tol3 = fmax(e1abs, fabs(e0)) * DBL_EPSILON;
    if (err2 <= tol2 && err3 <= tol3) {
      result = res;
      absolute = err2 + err3;
      relative = 5 * DBL_EPSILON * fabs(res);
      abserr = fmax(absolute, relative);
      return;
    }
    double e3 = epstab[n - 2 * i];
    epstab[n - 2 * i] = e1;
    double delta1 = e1 - e3;
    double err1 = fabs(delta1);
    double tol1 = fmax(e1abs, fabs(e3)) * DBL_EPSILON;
    if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
      n_final = 2 * i;
      break;
    }
    double ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
    if (fabs(ss * e1) <= 0.0001) {
      n_final = 2 * i;
      break;
    }
    res = e1 + 1 / ss;
    epstab[n - 2 * i] = res;
    {
      const double error = err2 + fabs(res - e2) + err3;
      if (error <= abserr) {
        abserr = error;
        result = res;
      }
    }
  }
  {
    const int limexp = 50 - 1;
    if (n_final == limexp) n_final = 2 * (limexp / 2);
  }
  if (n_orig % 2 == 1) {
    for (int i = 0; i <= newelm; i++) epstab[1 + i * 2] = epstab[i * 2 + 3];
  } else {
    for (int i = 0; i <= newelm; i++) epstab[i * 2] = epstab[i * 2 + 2];
  }
  if (n_orig != n_final) {
    for (int i = 0; i <= n_final; i++) epstab[i] = epstab[n_orig - n_final + i];
  }
  table.n = n_final + 1;
  if (nres_orig < 3) {
    res3la[nres_orig] = result;
    abserr = DBL_MAX;
  } else {
    abserr = (fabs(result - res3la[2]) + fabs(result - res3la[1]) + fabs(result - res3la[0]));
    res3la[0] = res3la[1];
    res3la[1] = res3la[2];
    res3la[2] = result;
  }
  table.nres = nres_orig + 1;
  abserr = fmax(abserr, 5 * DBL_EPSILON * fabs(result));
}

// Per-lane QAGS state.  cap = capacity of the interval list (the reference allows 1000; every
// integral met on this path needs < 20).  `limit` keeps the reference's value
// (1000) in all the places where the algorithm's decisions depend on it (qpsrt's `top`,
// increase_nrmax's `jupbnd`, the iteration cap); running out of CAP sets `overflow` and the
// integral is redone by a second pass with cap = 1000.
// interval-list storage: in-thread arrays (main pass) ...
template <int CAP>
struct QagsLocalStore {
  double alist[CAP], blist[CAP], rlist[CAP], elist[CAP];
  short order[CAP], level[CAP];
  static constexpr int cap = CAP;
};
// ... or a caller-provided global-memory workspace of the reference's full size (overflow pass)
struct QagsGlobalStore {
  double *alist, *blist, *rlist, *elist;
  short *order, *level;
  static constexpr int cap = 1000;
};

template <class Store>
struct Qags : Store {
  static constexpr int kLimit = 1000;
  using Store::alist; using Store::blist; using Store::rlist; using Store::elist;
  using Store::order; using Store::level; using Store::cap;
  int size, nrmax, i, maximum_level;
  // driver state (integration/qags.c)
  double a0, b0, epsabs, epsrel;
  double area, errsum, res_ext, err_ext, resabs0, tolerance, ertest, error_over_large_intervals;
  double reseps, abseps, correc;
  int ktmin, roundoff_type1, roundoff_type2, roundoff_type3, error_type, error_type2, iteration;
  bool positive_integrand, extrapolate, disallow_extrapolation, overflow;
  EpsTable table;
  // pending bisection
  double a1, b1, a2, b2, r_i, e_i;
  int current_level;
  // outputs
  double result, abserr;
  int ier, neval;

  __device__ void begin(double a, double b, double ea, double er)
  {
    a0 = a; b0 = b; epsabs = ea; epsrel = er;
    size = 0; nrmax = 0; i = 0; maximum_level = 0;
    alist[0] = a; blist[0] = b; rlist[0] = 0; elist[0] = 0; order[0] = 0; level[0] = 0;
    ertest = 0; error_over_large_intervals = 0; reseps = 0; abseps = 0; correc = 0;
    ktmin = 0; roundoff_type1 = roundoff_type2 = roundoff_type3 = 0;
    error_type = 0; error_type2 = 0; iteration = 0;
    positive_integrand = false; extrapolate = false; disallow_extrapolation = false; overflow = false;
    result = 0; abserr = 0; ier = 0; neval = 0;
  }

  // after the first GK21 on [a0,b0]; returns true when the integral is finished
  __device__ bool post_first(const GkOut& g)
  {
    neval += 21;
    size = 1; rlist[0] = g.result; elist[0] = g.abserr;
    resabs0 = g.resabs;
    tolerance = fmax(epsabs, epsrel * fabs(g.result));
    if (g.abserr <= 100 * DBL_EPSILON * g.resabs && g.abserr > tolerance) {
      result = g.result; abserr = g.abserr; ier = 18;  // GSL_EROUND
      return true;
    } else if ((g.abserr <= tolerance && g.abserr != g.resasc) || g.abserr == 0.0) {
      result = g.result; abserr = g.abserr; ier = 0;
      return true;
    }
    table.n = 0; table.nres = 0;
    table.rlist2[table.n++] = g.result;
    area = g.result;
    errsum = g.abserr;
    res_ext = g.result;
    err_ext = DBL_MAX;
    positive_integrand = (fabs(g.result) >= (1 - 50 * DBL_EPSILON) * g.resabs);
    iteration = 1;
    return false;
  }

  // choose the interval to bisect (top of the do-loop in qags())
  __device__ void pre_step()
  {
    double a_i = alist[i], b_i = blist[i];
    r_i = rlist[i]; e_i = elist[i];
    current_level = level[i] + 1;
    a1 = a_i; b1 = 0.5 * (a_i + b_i); a2 = b1; b2 = b_i;
    iteration++;
  }

  __device__ void qpsrt()
  {
    const int last = size - 1;
    const int limit = kLimit;
    int i_nrmax = nrmax;
    int i_maxerr = order[i_nrmax];
    if (last < 2) {
      order[0] = 0; order[1] = 1;
      i = i_maxerr;
      return;
    }
    double errmax = elist[i_maxerr];
    while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
      order[i_nrmax] = order[i_nrmax - 1];
      i_nrmax--;
    }
    int top = (last < (limit / 2 + 2)) ? last : limit - last + 1;
    int ii = i_nrmax + 1;
    while (ii < top && errmax < elist[order[ii]]) {
      order[ii - 1] = order[ii];
      ii++;
    }
    order[ii - 1] = (short)i_maxerr;
    double errmin = elist[last];
    int k = top - 1;
    while (k > ii - 2 && errmin >= elist[order[k]]) {
      order[k + 1] = order[k];
      k--;
    }
    order[k + 1] = (short)last;
    i_maxerr = order[i_nrmax];
    i = i_maxerr;
    nrmax = i_nrmax;
  }

  __device__ void update(double area1, double error1, double area2, double error2)
  {
    const int i_max = i;
    const int i_new = size;
    const int new_level = level[i_max] + 1;
    if (error2 > error1) {
      alist[i_max] = a2;
      rlist[i_max] = area2; elist[i_max] = error2; level[i_max] = (short)new_level;
      alist[i_new] = a1; blist[i_new] = b1; rlist[i_new] = area1; elist[i_new] = error1;
      level[i_new] = (short)new_level;
    } else {
      blist[i_max] = b1;
      rlist[i_max] = area1; elist[i_max] = error1; level[i_max] = (short)new_level;
      alist[i_new] = a2; blist[i_new] = b2; rlist[i_new] = area2; elist[i_new] = error2;
      level[i_new] = (short)new_level;
    }
    size++;
    if (new_level > maximum_level) maximum_level = new_level;
    qpsrt();
  }

  __device__ bool increase_nrmax()
  {
    int id = nrmax;
    int last = size - 1;
    int jupbnd = (last > (1 + kLimit / 2)) ? kLimit + 1 - last : last;
    for (int k = id; k <= jupbnd; k++) {
      int i_max = order[nrmax];
      i = i_max;
      if (level[i_max] < maximum_level) return true;
      nrmax++;
    }
    return false;
  }

  // body of the do-loop after the two GK21 evaluations; returns true when finished
  // (result/abserr/ier set)
  __device__ bool post_step(const GkOut& g1, const GkOut& g2)
  {
    neval += 42;
    const double area1 = g1.result, area2 = g2.result;
    const double error1 = g1.abserr, error2 = g2.abserr;
    const double area12 = area1 + area2;
    const double error12 = error1 + error2;
    const double last_e_i = e_i;
    errsum = errsum + error12 - e_i;
    area = area + area12 - r_i;
    tolerance = fmax(epsabs, epsrel * fabs(area));
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
    update(area1, error1, area2, error2);

    if (errsum <= tolerance) return finish(true);
    if (error_type) return finish(false);
    if (iteration >= kLimit - 1) { error_type = 1; return finish(false); }
    if (size >= cap) { overflow = true; error_type = 1; return finish(false); }
    if (iteration == 2) {
      error_over_large_intervals = errsum;
      ertest = tolerance;
      table.rlist2[table.n++] = area;
      return false;
    }
    if (disallow_extrapolation) return false;
    error_over_large_intervals += -last_e_i;
    if (current_level < maximum_level) error_over_large_intervals += error12;
    if (!extrapolate) {
      if (level[i] < maximum_level) return false;  // large_interval()
      extrapolate = true;
      nrmax = 1;
    }
    if (!error_type2 && error_over_large_intervals > ertest) {
      if (increase_nrmax()) return false;
    }
    table.rlist2[table.n++] = area;
    qelg(table, reseps, abseps);
    ktmin++;
    if (ktmin > 5 && err_ext < 0.001 * errsum) error_type = 5;
    if (abseps < err_ext) {
      ktmin = 0;
      err_ext = abseps;
      res_ext = reseps;
      correc = error_over_large_intervals;
      ertest = fmax(epsabs, epsrel * fabs(reseps));
      if (err_ext <= ertest) return finish(false);
    }
    if (table.n == 1) disallow_extrapolation = true;
    if (error_type == 5) return finish(false);
    nrmax = 0; i = order[0];  // reset_nrmax
    extrapolate = false;
    error_over_large_intervals = errsum;
    return false;
  }

  // tail of qags(): choice between the extrapolated value and the plain sum
  __device__ __noinline__ bool finish(bool direct_sum)
  {
    bool compute = direct_sum;
    bool ret_err = false;
    if (!compute) {
      result = res_ext;
      abserr = err_ext;
      if (err_ext == DBL_MAX) {
        compute = true;
      } else {
        if (error_type || error_type2) {
          if (error_type2) err_ext += correc;
          if (error_type == 0) error_type = 3;
          if (res_ext != 0.0 && area != 0.0) {
            if (err_ext / fabs(res_ext) > errsum / fabs(area)) compute = true;
          } else if (err_ext > errsum) {
            compute = true;
          } else if (area == 0.0) {
            ret_err = true;
          }
        }
        if (!compute && !ret_err) {
          double max_area = fmax(fabs(res_ext), fabs(area));
          if (!positive_integrand && max_area < 0.01 * resabs0) {
            ret_err = true;
          } else {
            double ratio = res_ext / area;
            if (ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) error_type = 6;
          }
        }
      }
    }
    if (compute) {
      double s = 0;
      for (int k = 0; k < size; k++) s += rlist[k];
      result = s;
      abserr = errsum;
    }
    if (error_type > 2) error_type--;
    ier = error_type;
    return true;
  }
};

}  // namespace upc
