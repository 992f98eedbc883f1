#pragma once
#include "gsl_spline.h"
