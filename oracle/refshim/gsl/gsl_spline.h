// gsl_spline / gsl_interp_cspline / gsl_interp_accel shim over the oracle's restatement
// (oracle/upc_oracle.c: upco_cspline_init / upco_cspline_eval).  Out-of-range evaluation aborts,
// as GSL's default error handler does.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
extern "C" {
void upco_cspline_init(const double* x, const double* y, int n, double* c);
double upco_cspline_eval(const double* x, const double* y, const double* c, int n, double xv);
}
struct gsl_interp_accel { size_t cache; };
struct gsl_interp_type { int id; };
static const gsl_interp_type gsl_interp_cspline_obj{1};
static const gsl_interp_type* const gsl_interp_cspline = &gsl_interp_cspline_obj;
struct gsl_spline { std::vector<double> x, y, c; size_t size; };
inline gsl_interp_accel* gsl_interp_accel_alloc() { return new gsl_interp_accel{0}; }
inline void gsl_interp_accel_free(gsl_interp_accel* a) { delete a; }
inline gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t n) { auto* s = new gsl_spline; s->size = n; return s; }
inline void gsl_spline_free(gsl_spline* s) { delete s; }
inline int gsl_spline_init(gsl_spline* s, const double* xa, const double* ya, size_t n)
{
  s->x.assign(xa, xa + n); s->y.assign(ya, ya + n); s->c.assign(n, 0.); s->size = n;
  upco_cspline_init(s->x.data(), s->y.data(), (int)n, s->c.data());
  return 0;
}
// gsl_interp_accel_find + cspline_eval: the cached interval is tried first, then a bisection on the
// side of the miss (same interval as a plain bisection; same cost profile as GSL's).  The reference
// shares one accelerator between its OpenMP threads; here each thread keeps its own cache slot.
inline size_t gsl_shim_bsearch(const double* xa, double x, size_t lo, size_t hi)
{
  while (hi > lo + 1) {
    size_t i = (hi + lo) / 2;
    if (xa[i] > x) hi = i; else lo = i;
  }
  return lo;
}
inline double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel* a)
{
  if (x < s->x.front() || x > s->x.back()) {
    std::fprintf(stderr, "gsl: interp.c: ERROR: interpolation error (x=%.17g outside [%.17g, %.17g])\n", x, s->x.front(), s->x.back());
    std::abort();
  }
  static thread_local const gsl_spline* owner[4] = {nullptr, nullptr, nullptr, nullptr};
  static thread_local size_t cache[4] = {0, 0, 0, 0};
  int slot = 0;
  for (; slot < 4; slot++) {
    if (owner[slot] == s) break;
    if (!owner[slot]) { owner[slot] = s; cache[slot] = 0; break; }
  }
  if (slot == 4) { slot = 0; owner[0] = s; cache[0] = 0; }
  (void)a;
  const double* xa = s->x.data();
  const size_t n = s->size;
  size_t idx = cache[slot];
  if (idx > n - 2) idx = 0;
  if (x < xa[idx]) idx = gsl_shim_bsearch(xa, x, 0, idx);
  else if (x >= xa[idx + 1]) idx = gsl_shim_bsearch(xa, x, idx, n - 1);
  cache[slot] = idx;
  // cspline_eval on the found interval
  const double x_lo = xa[idx], x_hi = xa[idx + 1];
  const double dx = x_hi - x_lo;
  const double y_lo = s->y[idx], y_hi = s->y[idx + 1];
  const double dy = y_hi - y_lo;
  const double delx = x - x_lo;
  const double c_i = s->c[idx], c_ip1 = s->c[idx + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}
