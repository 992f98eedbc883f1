// GSL shim (oracle/refshim): gsl_histogram + gsl_histogram_pdf as include/UpcSampler.h:40-71 uses them.  The
// arithmetic behind pdf_init / pdf_sample is the oracle's restatement of GSL's histogram/pdf.c and find.c
// (upco_pdf_init, upco_sample1d).  TEST INFRASTRUCTURE.
#pragma once
#include <cstddef>
#include <cstring>

#ifndef GSL_EDOM
#define GSL_SUCCESS 0
#define GSL_EDOM 1
#endif

extern "C" void upco_pdf_init(const double* bin, size_t n, double* sum);
extern "C" double upco_sample1d(const double* sum, int n, const double* edges, double r);

struct gsl_histogram { size_t n; double* range; double* bin; };
struct gsl_histogram_pdf { size_t n; double* range; double* sum; };

inline gsl_histogram* gsl_histogram_alloc(size_t n)
{
  auto* h = new gsl_histogram{n, new double[n + 1](), new double[n]()};
  return h;
}
inline void gsl_histogram_free(gsl_histogram* h) { if (h) { delete[] h->range; delete[] h->bin; delete h; } }
inline int gsl_histogram_set_ranges(gsl_histogram* h, const double range[], size_t size)
{
  if (size != h->n + 1) return GSL_EDOM;
  std::memcpy(h->range, range, size * sizeof(double));
  for (size_t i = 0; i < h->n; i++) h->bin[i] = 0;
  return GSL_SUCCESS;
}
inline gsl_histogram_pdf* gsl_histogram_pdf_alloc(size_t n) { return new gsl_histogram_pdf{n, new double[n + 1](), new double[n + 1]()}; }
inline void gsl_histogram_pdf_free(gsl_histogram_pdf* p) { if (p) { delete[] p->range; delete[] p->sum; delete p; } }
inline int gsl_histogram_pdf_init(gsl_histogram_pdf* p, const gsl_histogram* h)
{
  if (p->n != h->n) return GSL_EDOM;
  for (size_t i = 0; i < h->n; i++)
    if (h->bin[i] < 0) return GSL_EDOM;  // "histogram bins must be non-negative to compute a probability distribution"
  std::memcpy(p->range, h->range, (h->n + 1) * sizeof(double));
  upco_pdf_init(h->bin, h->n, p->sum);
  return GSL_SUCCESS;
}
inline double gsl_histogram_pdf_sample(const gsl_histogram_pdf* p, double r) { return upco_sample1d(p->sum, (int)p->n, p->range, r); }
