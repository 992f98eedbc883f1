#pragma once
#include "gsl_interp2d.h"
struct gsl_spline2d { int unused; };
inline gsl_spline2d* gsl_spline2d_alloc(const gsl_interp2d_type*, size_t, size_t) { std::abort(); }
inline int gsl_spline2d_init(gsl_spline2d*, const double*, const double*, const double*, size_t, size_t) { std::abort(); }
inline double gsl_spline2d_eval(const gsl_spline2d*, double, double, gsl_interp_accel*, gsl_interp_accel*) { std::abort(); }
