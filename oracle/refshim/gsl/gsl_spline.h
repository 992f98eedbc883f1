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
inline double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel*)
{
  if (x < s->x.front() || x > s->x.back()) {
    std::fprintf(stderr, "gsl: interp.c: ERROR: interpolation error (x=%.17g outside [%.17g, %.17g])\n", x, s->x.front(), s->x.back());
    std::abort();
  }
  return upco_cspline_eval(s->x.data(), s->y.data(), s->c.data(), (int)s->size, x);
}
