// gsl_integration_qags shim over the oracle's QAGS restatement (upco_qags)
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
struct gsl_function { double (*function)(double x, void* params); void* params; };
struct gsl_integration_workspace { size_t limit; };
extern "C" int upco_qags(double (*f)(double, void*), void* par, double a, double b, double epsabs, double epsrel,
                         size_t limit, double* result, double* abserr);
inline gsl_integration_workspace* gsl_integration_workspace_alloc(size_t n) { return new gsl_integration_workspace{n}; }
inline void gsl_integration_workspace_free(gsl_integration_workspace* w) { delete w; }
inline int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                                gsl_integration_workspace*, double* result, double* abserr)
{
  int rc = upco_qags(f->function, f->params, a, b, epsabs, epsrel, limit, result, abserr);
  if (rc) { std::fprintf(stderr, "gsl: qags.c: ERROR %d (default handler aborts)\n", rc); std::abort(); }
  return 0;
}
