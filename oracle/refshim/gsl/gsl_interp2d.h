#pragma once
#include "gsl_spline.h"
// gsl_interp2d / gsl_spline2d: src/UpcPhotoNuclearVM.cpp reaches them with SHADOWING 2 and 3 only (FGS10 grids), which
// the tests do not run; declared so that the file compiles, aborting if ever called.
#include <cstdlib>
struct gsl_interp2d_type { int unused; };
static const gsl_interp2d_type gsl_interp2d_bicubic_obj = {0};
static const gsl_interp2d_type* const gsl_interp2d_bicubic = &gsl_interp2d_bicubic_obj;
