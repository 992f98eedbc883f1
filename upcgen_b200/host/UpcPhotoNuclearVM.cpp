// UpcPhotoNuclearVM.cpp -- see UpcPhotoNuclearVM.h.  Counterpart of the reference's src/UpcPhotoNuclearVM.cpp.
#include "UpcPhotoNuclearVM.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

#include "UpcCrossSection.h"
#include "UpcPhysConstants.h"

UpcPhotoNuclearVM::UpcPhotoNuclearVM(int partPDG_, int shadowingOpt, int dghtPDG_)
{
  partPDG = partPDG_;
  dghtPDG = dghtPDG_;
  fShadowing = shadowingOpt;
  isCharged = false;
  // src/UpcPhotoNuclearVM.cpp:62-97
  fPw = 0.4;
  if (partPDG == 443) { mPart = 3.0969; fC0 = 342.; fMu2 = 3.; }
  else if (partPDG == 100443) { mPart = 3.6861; fC0 = 56.8; fMu2 = 4.; }
  else if (partPDG == 553) { mPart = 9.3987; fC0 = 0.902; fPw = 0.447; fMu2 = 22.4; }
  else throw std::invalid_argument("Unsupported particle");
  if (dghtPDG == 11) mDght = phys_consts::mEl;
  else if (dghtPDG == 13) mDght = phys_consts::mMu;
  else if (dghtPDG == 2212) mDght = phys_consts::mProt;
  else throw std::invalid_argument("Unsupported decay mode");
  fMmin = phys_consts::mProt + mPart;
  if (fShadowing != 0 && fShadowing != 2 && fShadowing != 3 && fShadowing != 4) {
    ok = false;
    error = "SHADOWING " + std::to_string(fShadowing) +
            " is not available in this build (0: impulse approximation, 2/3: FGS10 grids, 4: LTA tables)";
  }
}

// :41-52
double UpcPhotoNuclearVM::dsdt(double Wgp) const
{
  const double Wgp2 = Wgp * Wgp;
  if (!(Wgp > fMmin)) return 0.;
  return fC0 * std::pow(1. - (fMmin * fMmin) / Wgp2, 1.5) * std::pow(Wgp2 * 1e-4, fPw);
}

namespace
{
double ff2(double t)
{
  const double f = UpcCrossSection::calcFormFac(t);
  return f * f;
}
// 15-point Gauss-Kronrod on [a, b]; err = |K15 - G7|
double gk15(double a, double b, double& err)
{
  static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
  static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
  static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                               0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  const double fc = ff2(c);
  double rk = fc * wgk[7], rg = fc * wg[3];
  for (int j = 0; j < 7; ++j) {
    const double dx = h * xgk[j];
    const double s = ff2(c - dx) + ff2(c + dx);
    rk += wgk[j] * s;
    if (j & 1) rg += wg[j / 2] * s;
  }
  err = std::fabs((rk - rg) * h);
  return rk * h;
}
double adapt(double a, double b, double whole, double tol, int depth)
{
  const double m = 0.5 * (a + b);
  double e1, e2;
  const double l = gk15(a, m, e1), r = gk15(m, b, e2);
  if (depth <= 0 || e1 + e2 <= tol * std::fabs(l + r) || std::fabs(l + r - whole) <= 1e-15 * std::fabs(whole)) return l + r;
  return adapt(a, m, l, tol, depth - 1) + adapt(m, b, r, tol, depth - 1);
}
} // namespace

// Phi_A: the reference calls TF1::Integral(tmin, tmax) on the squared form factor (:355-356, default relative
// tolerance 1e-12); the form factor has diffractive zeros inside the range, hence the bisection
double UpcPhotoNuclearVM::integrateFormFactorSq(double tmin, double tmax)
{
  double e;
  const double whole = gk15(tmin, tmax, e);
  return adapt(tmin, tmax, whole, 1e-12, 40);
}

// :303-337: the LTA shadowing of Guzey and Zhalov, a TGraph of 37 points with a TSpline3 through them
double UpcPhotoNuclearVM::getRgLtaVG(double x)
{
  if (!fLtaInit) {
    const char* dir = std::getenv("UPCGEN_CROSS_SEC_DIR");
#ifdef CROSS_SEC_DIR
    if (!dir) dir = CROSS_SEC_DIR;
#endif
    if (!dir) { ok = false; error = "UPCGEN_CROSS_SEC_DIR is not set (vm/lta tables)"; return 1.; }
    std::string fname;
    if (partPDG == 443) fname = std::string(dir) + "/vm/lta/LT2013_pb208_cteq6l1_m12_Q2_3.dat";
    else if (partPDG == 100443) fname = std::string(dir) + "/vm/lta/LT2013_pb208_cteq6l1_m12_Q2_4.dat";
    else { ok = false; error = "PDG " + std::to_string(partPDG) + " is not available for chosen shadowing option"; return 1.; }
    std::ifstream ifs(fname);
    if (!ifs) { ok = false; error = "Missing file: " + fname; return 1.; }
    const int n = 37;
    fX.resize(n); fY.resize(n);
    double rltas;
    for (int i = 0; i < n; ++i) ifs >> fX[i] >> rltas >> fY[i];
    // not-a-knot cubic spline (TSpline3 with default end conditions): S_i(x) = y_i + b_i d + c_i d^2 + d_i d^3
    std::vector<double> h(n - 1), dl(n - 1);
    for (int i = 0; i < n - 1; ++i) { h[i] = fX[i + 1] - fX[i]; dl[i] = (fY[i + 1] - fY[i]) / h[i]; }
    // unknowns: second derivatives m_i; interior equations + not-a-knot at both ends, dense Gauss elimination (n = 37)
    std::vector<std::vector<double>> A(n, std::vector<double>(n + 1, 0.));
    for (int i = 1; i < n - 1; ++i) {
      A[i][i - 1] = h[i - 1]; A[i][i] = 2 * (h[i - 1] + h[i]); A[i][i + 1] = h[i];
      A[i][n] = 6 * (dl[i] - dl[i - 1]);
    }
    A[0][0] = h[1]; A[0][1] = -(h[0] + h[1]); A[0][2] = h[0];                               // S''' continuous at x_1
    A[n - 1][n - 3] = h[n - 2]; A[n - 1][n - 2] = -(h[n - 3] + h[n - 2]); A[n - 1][n - 1] = h[n - 3];  // ... and at x_{n-2}
    for (int col = 0; col < n; ++col) {
      int piv = col;
      for (int r = col + 1; r < n; ++r) if (std::fabs(A[r][col]) > std::fabs(A[piv][col])) piv = r;
      std::swap(A[col], A[piv]);
      for (int r = col + 1; r < n; ++r) {
        const double f = A[r][col] / A[col][col];
        if (f != 0) for (int k = col; k <= n; ++k) A[r][k] -= f * A[col][k];
      }
    }
    std::vector<double> m(n);
    for (int r = n - 1; r >= 0; --r) {
      double s = A[r][n];
      for (int k = r + 1; k < n; ++k) s -= A[r][k] * m[k];
      m[r] = s / A[r][r];
    }
    fB.resize(n - 1); fC.resize(n - 1); fD.resize(n - 1);
    for (int i = 0; i < n - 1; ++i) {
      fB[i] = dl[i] - h[i] * (2 * m[i] + m[i + 1]) / 6;
      fC[i] = m[i] / 2;
      fD[i] = (m[i + 1] - m[i]) / (6 * h[i]);
    }
    fLtaInit = true;
  }
  const int n = (int)fX.size();
  if (x > 1e-5 && x < 1e-1) {
    // TSpline3::Eval: the segment that holds x (clamped to the ends)
    int i = 0;
    if (x <= fX[0]) i = 0;
    else if (x >= fX[n - 1]) i = n - 2;
    else { int lo = 0, hi = n - 1; while (hi - lo > 1) { int mid = (lo + hi) / 2; if (fX[mid] <= x) lo = mid; else hi = mid; } i = lo; }
    const double d = x - fX[i];
    return fY[i] + d * (fB[i] + d * (fC[i] + d * fD[i]));
  }
  // TGraph::Eval(x, 0, ""): linear interpolation, linear extrapolation from the two end points
  int lo, up;
  if (x <= fX[0]) { lo = 0; up = 1; }
  else if (x >= fX[n - 1]) { lo = n - 2; up = n - 1; }
  else { int a = 0, b = n - 1; while (b - a > 1) { int mid = (a + b) / 2; if (fX[mid] <= x) a = mid; else b = mid; } lo = a; up = a + 1; }
  return fY[up] + (x - fX[up]) * (fY[lo] - fY[up]) / (fX[lo] - fX[up]);
}

namespace
{
// first derivative of the natural cubic spline through (x_i, y_i) at its own knots (what GSL's bicubic initialisation
// takes from gsl_spline_eval_deriv: interp2d/bicubic.c); second-derivative form, Thomas elimination
std::vector<double> naturalSplineSlopes(const std::vector<double>& x, const std::vector<double>& y)
{
  const int n = (int)x.size();
  std::vector<double> h(n - 1), s(n - 1), m(n, 0.), slope(n);
  for (int i = 0; i < n - 1; ++i) { h[i] = x[i + 1] - x[i]; s[i] = (y[i + 1] - y[i]) / h[i]; }
  if (n > 2) {
    // h_{i-1} m_{i-1} + 2 (h_{i-1} + h_i) m_i + h_i m_{i+1} = 6 (s_i - s_{i-1}),  m_0 = m_{n-1} = 0
    std::vector<double> diag(n), rhs(n);
    for (int i = 1; i < n - 1; ++i) { diag[i] = 2 * (h[i - 1] + h[i]); rhs[i] = 6 * (s[i] - s[i - 1]); }
    for (int i = 2; i < n - 1; ++i) {
      const double f = h[i - 1] / diag[i - 1];
      diag[i] -= f * h[i - 1];
      rhs[i] -= f * rhs[i - 1];
    }
    for (int i = n - 2; i >= 1; --i) m[i] = (rhs[i] - (i + 1 < n - 1 ? h[i] * m[i + 1] : 0.)) / diag[i];
  }
  for (int i = 0; i < n - 1; ++i) slope[i] = s[i] - h[i] * (2 * m[i] + m[i + 1]) / 6;
  slope[n - 1] = s[n - 2] + h[n - 2] * (m[n - 2] + 2 * m[n - 1]) / 6;
  return slope;
}
} // namespace

// :249-297: gluon shadowing of Frankfurt, Guzey and Strikman (Phys. Rept. 512 (2012) 255), models 2 (weak, type 0) and
// 1 (strong, type 1): the glue column of QCDEvolution_pb<A>proton_2009_model<k>.dat on its 90 x 7 (x, Q^2) grid,
// interpolated at (x, mu^2) as gsl_spline2d with gsl_interp2d_bicubic does -- a bicubic Hermite patch whose
// derivatives z_x, z_y, z_xy at the grid points are those of natural cubic splines along the grid lines.
double UpcPhotoNuclearVM::getRgLta(int type, double x)
{
  if (partPDG == 443) {
    // :283-287 dereference graphs the reference never fills; unreachable there too (calcCrossSectionY asks for
    // mu^2 >= 4) -- kept as an error instead of a crash
    ok = false; error = "FGS10 shadowing is not available for J/psi"; return 1.;
  }
  if (fFgsType != type) {
    const char* dir = std::getenv("UPCGEN_CROSS_SEC_DIR");
#ifdef CROSS_SEC_DIR
    if (!dir) dir = CROSS_SEC_DIR;
#endif
    if (!dir) { ok = false; error = "UPCGEN_CROSS_SEC_DIR is not set (vm/lta grids)"; return 1.; }
    const std::string fname = std::string(dir) + "/vm/lta/QCDEvolution_pb" + std::to_string(UpcCrossSection::A) +
                              "proton_2009_model" + std::to_string(type == 0 ? 2 : 1) + ".dat";
    std::ifstream ifs(fname);
    if (!ifs) { ok = false; error = "Missing file: " + fname; return 1.; }
    const int nx = 90, nq = 7;
    fGx.assign(nx, 0.); fGq.assign(nq, 0.);
    fGz.assign((size_t)nx * nq, 0.);
    double xx, dv, uv, ubar, dbar, sbar, cbar, glue, f2;
    for (int i = 0; i < nq; ++i) {
      ifs >> fGq[i];
      for (int j = 0; j < nx; ++j) {
        ifs >> xx >> dv >> uv >> ubar >> dbar >> sbar >> cbar >> glue >> f2;
        fGx[j] = xx;
        fGz[(size_t)i * nx + j] = glue;
      }
    }
    if (!ifs) { ok = false; error = "Truncated file: " + fname; return 1.; }
    fGzx.assign(fGz.size(), 0.); fGzy.assign(fGz.size(), 0.); fGzxy.assign(fGz.size(), 0.);
    std::vector<double> line;
    for (int i = 0; i < nq; ++i) {            // along x, one Q^2 row at a time
      line.assign(fGz.begin() + (size_t)i * nx, fGz.begin() + (size_t)(i + 1) * nx);
      const std::vector<double> d = naturalSplineSlopes(fGx, line);
      for (int j = 0; j < nx; ++j) fGzx[(size_t)i * nx + j] = d[j];
    }
    for (int j = 0; j < nx; ++j) {            // along Q^2, one x column at a time
      line.resize(nq);
      for (int i = 0; i < nq; ++i) line[i] = fGz[(size_t)i * nx + j];
      const std::vector<double> d = naturalSplineSlopes(fGq, line);
      for (int i = 0; i < nq; ++i) fGzy[(size_t)i * nx + j] = d[i];
    }
    for (int i = 0; i < nq; ++i) {            // the cross derivative: along x of z_y (GSL's order)
      line.assign(fGzy.begin() + (size_t)i * nx, fGzy.begin() + (size_t)(i + 1) * nx);
      const std::vector<double> d = naturalSplineSlopes(fGx, line);
      for (int j = 0; j < nx; ++j) fGzxy[(size_t)i * nx + j] = d[j];
    }
    fFgsType = type;
  }
  // grid limitations (:291-295)
  if (x < 9.99999975e-6) x = 9.99999975e-6;
  if (x > 0.95) x = 0.95;
  const int nx = (int)fGx.size(), nq = (int)fGq.size();
  const double q2 = fMu2;
  if (x < fGx[0] || x > fGx[nx - 1] || q2 < fGq[0] || q2 > fGq[nq - 1]) {
    ok = false; error = "FGS10 grid: (x, mu^2) outside the table (gsl: interpolation error)"; return 1.;
  }
  auto cell = [](const std::vector<double>& g, double v) {
    int lo = 0, hi = (int)g.size() - 1;
    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (g[mid] > v) hi = mid; else lo = mid; }
    return lo;
  };
  const int ix = cell(fGx, x), iq = cell(fGq, q2);
  const double dx = fGx[ix + 1] - fGx[ix], dq = fGq[iq + 1] - fGq[iq];
  const double t = (x - fGx[ix]) / dx, u = (q2 - fGq[iq]) / dq;
  // cubic Hermite basis on the unit interval: value at 0 / at 1, slope at 0 / at 1
  const double ht[4] = {(2 * t - 3) * t * t + 1, (3 - 2 * t) * t * t, ((t - 2) * t + 1) * t, (t - 1) * t * t};
  const double hu[4] = {(2 * u - 3) * u * u + 1, (3 - 2 * u) * u * u, ((u - 2) * u + 1) * u, (u - 1) * u * u};
  double r = 0;
  for (int b = 0; b < 2; ++b)
    for (int a = 0; a < 2; ++a) {
      const size_t k = (size_t)(iq + b) * nx + (ix + a);
      r += fGz[k] * ht[a] * hu[b] + dx * fGzx[k] * ht[2 + a] * hu[b] + dq * fGzy[k] * ht[a] * hu[2 + b] +
           dx * dq * fGzxy[k] * ht[2 + a] * hu[2 + b];
    }
  return r;
}

// :340-381
double UpcPhotoNuclearVM::calcCrossSectionY(double y)
{
  const double w = mPart / 2. * std::exp(y);           // photon energy
  const double beamE = 0.5 * UpcCrossSection::sqrts;   // beam energy in lab cms
  const double Wgp2 = 4. * w * beamE;                  // photon energy in nucleon cms
  const double Wgp = std::sqrt(Wgp2);
  const double csGammaP = dsdt(Wgp);
  const double m2 = mPart * mPart;
  const double x = m2 / Wgp2;
  const double tmin = x * x * UpcCrossSection::mNucl * UpcCrossSection::mNucl;
  const double tmax = tmin + 1.;
  const double PhiA = integrateFormFactorSq(tmin, tmax);
  double cAcP2 = 1.;
  double Rg = 1.;
  // FGS10 parametrization works for psi(2s) and heavier (:366-374); for J/psi these two options leave the impulse
  // approximation in place, as the reference does
  if ((fShadowing == 2 || fShadowing == 3) && fMu2 > 4. - 1e-6) {
    cAcP2 = 0.9 * 0.9;
    Rg = getRgLta(fShadowing == 2 ? 0 : 1, x);
  }
  if (fShadowing == 4) {
    cAcP2 = 0.9 * 0.9;
    Rg = getRgLtaVG(x);
  }
  return cAcP2 * csGammaP * Rg * Rg * PhiA * 1e-6;
}
