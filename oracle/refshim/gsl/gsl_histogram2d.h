// GSL shim (oracle/refshim): gsl_histogram2d + gsl_histogram2d_pdf as include/UpcSampler.h:81-121 uses them;
// arithmetic = the oracle's restatement of GSL's histogram/pdf2d.c (upco_pdf_init, upco_sample2d).  TEST INFRASTRUCTURE.
#pragma once
#include "gsl_histogram.h"

extern "C" void upco_sample2d(const double* sum, int nx, int ny, const double* xe, const double* ye, double r1, double r2,
                              long long* kout, double* x, double* y);

struct gsl_histogram2d { size_t nx, ny; double* xrange; double* yrange; double* bin; };
struct gsl_histogram2d_pdf { size_t nx, ny; double* xrange; double* yrange; double* sum; };

inline gsl_histogram2d* gsl_histogram2d_alloc(size_t nx, size_t ny)
{
  return new gsl_histogram2d{nx, ny, new double[nx + 1](), new double[ny + 1](), new double[nx * ny]()};
}
inline void gsl_histogram2d_free(gsl_histogram2d* h) { if (h) { delete[] h->xrange; delete[] h->yrange; delete[] h->bin; delete h; } }
inline int gsl_histogram2d_set_ranges(gsl_histogram2d* h, const double xr[], size_t xs, const double yr[], size_t ys)
{
  if (xs != h->nx + 1 || ys != h->ny + 1) return GSL_EDOM;
  std::memcpy(h->xrange, xr, xs * sizeof(double));
  std::memcpy(h->yrange, yr, ys * sizeof(double));
  for (size_t i = 0; i < h->nx * h->ny; i++) h->bin[i] = 0;
  return GSL_SUCCESS;
}
inline gsl_histogram2d_pdf* gsl_histogram2d_pdf_alloc(size_t nx, size_t ny)
{
  return new gsl_histogram2d_pdf{nx, ny, new double[nx + 1](), new double[ny + 1](), new double[nx * ny + 1]()};
}
inline void gsl_histogram2d_pdf_free(gsl_histogram2d_pdf* p) { if (p) { delete[] p->xrange; delete[] p->yrange; delete[] p->sum; delete p; } }
inline int gsl_histogram2d_pdf_init(gsl_histogram2d_pdf* p, const gsl_histogram2d* h)
{
  if (p->nx != h->nx || p->ny != h->ny) return GSL_EDOM;
  const size_t n = h->nx * h->ny;
  for (size_t i = 0; i < n; i++)
    if (h->bin[i] < 0) return GSL_EDOM;
  std::memcpy(p->xrange, h->xrange, (h->nx + 1) * sizeof(double));
  std::memcpy(p->yrange, h->yrange, (h->ny + 1) * sizeof(double));
  upco_pdf_init(h->bin, n, p->sum);
  return GSL_SUCCESS;
}
inline int gsl_histogram2d_pdf_sample(const gsl_histogram2d_pdf* p, double r1, double r2, double* x, double* y)
{
  long long k;
  upco_sample2d(p->sum, (int)p->nx, (int)p->ny, p->xrange, p->yrange, r1, r2, &k, x, y);
  return k < 0 ? GSL_EDOM : GSL_SUCCESS;
}
