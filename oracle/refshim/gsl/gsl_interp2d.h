#pragma once
#include "gsl_spline.h"
// gsl_interp2d types: src/UpcPhotoNuclearVM.cpp names gsl_interp2d_bicubic for the FGS10 grids (SHADOWING 2 and 3);
// the interpolation itself is in gsl_spline2d.h.
#include <cstdlib>
struct gsl_interp2d_type { int id; };
static const gsl_interp2d_type gsl_interp2d_bicubic_obj = {2};
static const gsl_interp2d_type* const gsl_interp2d_bicubic = &gsl_interp2d_bicubic_obj;
