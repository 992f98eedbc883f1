#pragma once
#include "gsl_spline.h"
struct gsl_spline2d { int unused; };
