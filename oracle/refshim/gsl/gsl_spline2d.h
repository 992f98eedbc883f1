// gsl_spline2d with gsl_interp2d_bicubic, restated from GSL 2.x's published algorithm (interpolation/bicubic.c):
//   init: z_x at the grid points = derivative of the natural cubic spline (gsl_interp_cspline) along x through every
//         y row; z_y the same along y through every x column; z_xy = derivative along x of the spline through z_y;
//   eval: the cell is found by bisection in x and y, (t, u) are the unit-square coordinates, the derivatives are
//         scaled by the cell sizes and the patch is the sum over the 16 power-basis coefficients c_kl t^k u^l of
//         the bicubic that matches z, z_x, z_y, z_xy at the four corners (t^0 u^0 first, u fastest).
// Out-of-range evaluation aborts, as GSL's default error handler does.  Test infrastructure (oracle/_ref).
#pragma once
#include "gsl_interp2d.h"
#include <cstdio>
struct gsl_spline2d {
  size_t nx, ny;
  std::vector<double> x, y, z, zx, zy, zxy;   // z[j * nx + i] = z(x_i, y_j)
};
inline gsl_spline2d* gsl_spline2d_alloc(const gsl_interp2d_type*, size_t nx, size_t ny)
{
  auto* s = new gsl_spline2d; s->nx = nx; s->ny = ny; return s;
}
inline void gsl_spline2d_free(gsl_spline2d* s) { delete s; }
// derivative of the natural spline through (xa, ya) at each of its own knots: cspline_eval_deriv on the interval the
// accelerator returns for x = xa[i] (interval i, offset 0; the last knot is the right end of interval n - 2)
inline void gsl_shim_knot_derivs(const std::vector<double>& xa, const std::vector<double>& ya, std::vector<double>& out)
{
  const size_t n = xa.size();
  std::vector<double> c(n, 0.);
  upco_cspline_init(xa.data(), ya.data(), (int)n, c.data());
  out.resize(n);
  for (size_t i = 0; i < n; i++) {
    const size_t idx = i < n - 1 ? i : n - 2;
    const double dx = xa[idx + 1] - xa[idx], dy = ya[idx + 1] - ya[idx], delx = xa[i] - xa[idx];
    const double b_i = (dy / dx) - dx * (c[idx + 1] + 2.0 * c[idx]) / 3.0;
    const double d_i = (c[idx + 1] - c[idx]) / (3.0 * dx);
    out[i] = b_i + delx * (2.0 * c[idx] + 3.0 * d_i * delx);
  }
}
inline int gsl_spline2d_init(gsl_spline2d* s, const double* xa, const double* ya, const double* za, size_t nx, size_t ny)
{
  s->nx = nx; s->ny = ny;
  s->x.assign(xa, xa + nx); s->y.assign(ya, ya + ny); s->z.assign(za, za + nx * ny);
  s->zx.assign(nx * ny, 0.); s->zy.assign(nx * ny, 0.); s->zxy.assign(nx * ny, 0.);
  std::vector<double> line, d;
  for (size_t j = 0; j < ny; j++) {
    line.assign(s->z.begin() + j * nx, s->z.begin() + (j + 1) * nx);
    gsl_shim_knot_derivs(s->x, line, d);
    for (size_t i = 0; i < nx; i++) s->zx[j * nx + i] = d[i];
  }
  for (size_t i = 0; i < nx; i++) {
    line.resize(ny);
    for (size_t j = 0; j < ny; j++) line[j] = s->z[j * nx + i];
    gsl_shim_knot_derivs(s->y, line, d);
    for (size_t j = 0; j < ny; j++) s->zy[j * nx + i] = d[j];
  }
  for (size_t j = 0; j < ny; j++) {
    line.assign(s->zy.begin() + j * nx, s->zy.begin() + (j + 1) * nx);
    gsl_shim_knot_derivs(s->x, line, d);
    for (size_t i = 0; i < nx; i++) s->zxy[j * nx + i] = d[i];
  }
  return 0;
}
inline double gsl_spline2d_eval(const gsl_spline2d* s, double x, double y, gsl_interp_accel*, gsl_interp_accel*)
{
  const size_t nx = s->nx, ny = s->ny;
  if (x < s->x.front() || x > s->x.back() || y < s->y.front() || y > s->y.back()) {
    std::fprintf(stderr, "gsl: interp2d.c: ERROR: interpolation error (x=%.17g y=%.17g outside the grid)\n", x, y);
    std::abort();
  }
  const size_t xi = gsl_shim_bsearch(s->x.data(), x, 0, nx - 1), yi = gsl_shim_bsearch(s->y.data(), y, 0, ny - 1);
  const double dx = s->x[xi + 1] - s->x[xi], dy = s->y[yi + 1] - s->y[yi];
  const double t = (x - s->x[xi]) / dx, u = (y - s->y[yi]) / dy;
  const double dt = 1. / dx, du = 1. / dy;
  // corners: 00 = (xi, yi), 10 = (xi + 1, yi), 01 = (xi, yi + 1), 11 = (xi + 1, yi + 1)
  const size_t k00 = yi * nx + xi, k10 = k00 + 1, k01 = k00 + nx, k11 = k01 + 1;
  const double f00 = s->z[k00], f10 = s->z[k10], f01 = s->z[k01], f11 = s->z[k11];
  const double p00 = s->zx[k00] / dt, p10 = s->zx[k10] / dt, p01 = s->zx[k01] / dt, p11 = s->zx[k11] / dt;
  const double q00 = s->zy[k00] / du, q10 = s->zy[k10] / du, q01 = s->zy[k01] / du, q11 = s->zy[k11] / du;
  const double r00 = s->zxy[k00] / (dt * du), r10 = s->zxy[k10] / (dt * du), r01 = s->zxy[k01] / (dt * du),
               r11 = s->zxy[k11] / (dt * du);
  const double t2 = t * t, t3 = t * t2, u2 = u * u, u3 = u * u2;
  double z = 0, v;
  v = f00;                                                         z += v;
  v = q00;                                                         z += v * u;
  v = -3 * f00 + 3 * f01 - 2 * q00 - q01;                          z += v * u2;
  v = 2 * f00 - 2 * f01 + q00 + q01;                               z += v * u3;
  v = p00;                                                         z += v * t;
  v = r00;                                                         z += v * t * u;
  v = -3 * p00 + 3 * p01 - 2 * r00 - r01;                          z += v * t * u2;
  v = 2 * p00 - 2 * p01 + r00 + r01;                               z += v * t * u3;
  v = -3 * f00 + 3 * f10 - 2 * p00 - p10;                          z += v * t2;
  v = -3 * q00 + 3 * q10 - 2 * r00 - r10;                          z += v * t2 * u;
  v = 9 * f00 - 9 * f10 + 9 * f11 - 9 * f01 + 6 * p00 + 3 * p10 - 3 * p11 - 6 * p01 + 6 * q00 - 6 * q10 - 3 * q11 +
      3 * q01 + 4 * r00 + 2 * r10 + r11 + 2 * r01;                 z += v * t2 * u2;
  v = -6 * f00 + 6 * f10 - 6 * f11 + 6 * f01 - 4 * p00 - 2 * p10 + 2 * p11 + 4 * p01 - 3 * q00 + 3 * q10 + 3 * q11 -
      3 * q01 - 2 * r00 - r10 - r11 - 2 * r01;                     z += v * t2 * u3;
  v = 2 * f00 - 2 * f10 + p00 + p10;                               z += v * t3;
  v = 2 * q00 - 2 * q10 + r00 + r10;                               z += v * t3 * u;
  v = -6 * f00 + 6 * f10 - 6 * f11 + 6 * f01 - 3 * p00 - 3 * p10 + 3 * p11 + 3 * p01 - 4 * q00 + 4 * q10 + 2 * q11 -
      2 * q01 - 2 * r00 - 2 * r10 - r11 - r01;                     z += v * t3 * u2;
  v = 4 * f00 - 4 * f10 + 4 * f11 - 4 * f01 + 2 * p00 + 2 * p10 - 2 * p11 - 2 * p01 + 2 * q00 - 2 * q10 - 2 * q11 +
      2 * q01 + r00 + r10 + r11 + r01;                             z += v * t3 * u3;
  return z;
}
