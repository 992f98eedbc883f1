#pragma once
extern "C" {
double upco_bessel_K0(double x);
double upco_bessel_K1(double x);
double upco_bessel_J1(double x);
}
inline double gsl_sf_bessel_K0(double x) { return upco_bessel_K0(x); }
inline double gsl_sf_bessel_K1(double x) { return upco_bessel_K1(x); }
inline double gsl_sf_bessel_J1(double x) { return upco_bessel_J1(x); }
