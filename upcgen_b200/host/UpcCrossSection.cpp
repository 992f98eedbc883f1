// UpcCrossSection over the CUDA C-ABI.  Method-by-method counterpart of the reference's
// src/UpcCrossSection.cpp; every numeric kernel of that file runs on the GPU here.
#include "UpcCrossSection.h"
#include "UpcRootFile.h"
#include "UpcRootHist.h"
#include "UpcPhotoNuclearVM.h"
#include "UpcTwoPhotonTabulated.h"

#include <fcntl.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>

namespace
{
std::mt19937_64& hostRng()
{
  static std::mt19937_64 rng(0x5eedULL); // stands in for gRandom (TRandomMT64 is std::mt19937_64)
  return rng;
}
double uniform(double a, double b) { return a + (b - a) * std::generate_canonical<double, 53>(hostRng()); }
} // namespace

UpcCrossSection::UpcCrossSection()
{
  // the reference fixes gtot here, before parameters.in is read (src/UpcCrossSection.cpp:51-54)
  gtot = std::cosh((std::acosh(g1) + std::acosh(g2)) / 2.);
}

UpcCrossSection::~UpcCrossSection()
{
  delete elemProcess;
  if (ctx) upcgpu_destroy(ctx);
}

void UpcCrossSection::fail(const char* what, int rc)
{
  // error convention of the reference: log and std::_Exit(-1) (src/UpcCrossSection.cpp:110-111)
  PLOG_FATAL << what << " failed (" << rc << "): " << upcgpu_last_error(ctx);
  std::_Exit(-1);
}

void UpcCrossSection::setElemProcess(int procID)
{
  switch (procID) {
    case 11:
    case 13:
    case 15:
      elemProcess = new UpcTwoPhotonDilep(procID);
      break;
    case 51:
      elemProcess = new UpcTwoPhotonALP(alpMass, alpWidth);
      break;
    case 22:
    case 111: {
      // cross sections from the reference's ROOT files (cross_sections/lbyl, cross_sections/pi0pi0)
      auto* tab = procID == 22 ? (UpcTwoPhotonTabulated*)new UpcTwoPhotonLbyL(doMassCut, lowMCut, hiMCut)
                               : (UpcTwoPhotonTabulated*)new UpcTwoPhotonDipion(doMassCut, lowMCut, hiMCut);
      if (!tab->ok) {
        PLOG_FATAL << "Cannot read the elementary cross sections: " << tab->error
                   << " (set UPCGEN_CROSS_SEC_DIR to the reference's cross_sections directory). Exiting...";
        std::_Exit(-1);
      }
      elemProcess = tab;
      break;
    }
    case 443:
    case 100443:
    case 553: {
      UpcPhotoNuclearVM* vm = nullptr;
      try {
        vm = new UpcPhotoNuclearVM(procID, shadowingOption, dghtPDG);
      } catch (const std::exception& e) {
        PLOG_FATAL << "Vector-meson process " << procID << ": " << e.what() << " (DECAY_PDG " << dghtPDG << "). Exiting...";
        std::_Exit(-1);
      }
      if (!vm->ok) {
        PLOG_FATAL << vm->error << ". Exiting...";
        std::_Exit(-1);
      }
      elemProcess = vm;
      break;
    }
    default:
      PLOG_FATAL << "Unknown process ID! Check manual and enter a correct ID! Exiting...";
      std::_Exit(-1);
  }
}

upcgpu_params UpcCrossSection::makeParams() const
{
  upcgpu_params p{};
  p.Z = Z; p.A = A; p.R = R; p.a = a; p.sqrts = sqrts; p.g1 = g1; p.g2 = g2; p.gtot = gtot;
  p.is_point = isPoint; p.breakup_mode = breakupMode; p.use_pol = usePolarizedCS; p.nonzero_gam_pt = useNonzeroGamPt;
  p.nm = nm; p.ny = ny; p.nz = nz;
  p.mmin = mmin; p.mmax = mmax; p.ymin = ymin; p.ymax = ymax; p.zmin = zmin; p.zmax = zmax;
  p.nb1 = nb1; p.nb2 = nb2;
  if (elemProcess) {
    p.part_pdg = elemProcess->partPDG; p.m_part = elemProcess->mPart; p.is_charged = elemProcess->isCharged;
  }
  p.is_pair = evIsPair; p.is_single = evIsSingle; p.ignore_csz = evIgnoreCSZ; p.decay_uniform_pdg = evDecayUniformPDG;
  p.do_pt_cut = evDoPtCut; p.do_eta_cut = evDoEtaCut; p.pt_min = evMinPt; p.eta_min = evMinEta; p.eta_max = evMaxEta;
  return p;
}

void UpcCrossSection::ensureContext()
{
  if (ctx) return;
  upcgpu_params p = makeParams();
  int rc;
  if (numGpus > 1) {
    // the reference's -nthreads OpenMP team becomes -ngpus devices behind one handle: m rows dealt to the devices,
    // NCCL all-gather of the table, events split by candidate ranges (include/upcgpu.h: upcgpu_create_multi)
    std::vector<int> devs(numGpus);
    for (int i = 0; i < numGpus; ++i) devs[i] = device + i;
    rc = upcgpu_create_multi(&p, numGpus, devs.data(), &ctx);
    if (rc == UPCGPU_OK) {
      char buf[256];
      upcgpu_group_describe(ctx, buf, sizeof(buf));
      PLOG_INFO << "GPU group: " << buf;
    }
  } else {
    rc = upcgpu_create(&p, device, &ctx);
  }
  if (rc != UPCGPU_OK) {
    PLOG_FATAL << "upcgpu_create failed (" << rc << "): " << upcgpu_last_error(nullptr);
    std::_Exit(-1);
  }
}

void UpcCrossSection::ensureTables()
{
  ensureContext();
  if (tablesReady) return;
  int rc = upcgpu_prepare_tables(ctx);
  if (rc) fail("upcgpu_prepare_tables", rc);
  upcgpu_table_info info;
  upcgpu_get_table_info(ctx, &info);
  rho0 = info.rho0;
  tablesReady = true;
}

void UpcCrossSection::init()
{
  PLOG_INFO << "Initializing caches ...";
  factor = Z * Z * phys_consts::alpha / M_PI / M_PI / phys_consts::hc / phys_consts::hc;
  mNucl = (Z * phys_consts::mProt + (A - Z) * phys_consts::mNeut) / A;
  rho0 = calcWSRho();
  prepareGAA();
  prepareFormFac();
  if (breakupMode > 1) prepareBreakupProb();
  if (elemProcess->partPDG != 443 && elemProcess->partPDG != 100443 && elemProcess->partPDG != 553)
    prepareTwoPhotonLumi(); // two-photon luminosity (the vector-meson path needs the photon flux only, :127-131)
}

template <typename ArrayType>
double UpcCrossSection::simpson(int n, ArrayType* v, double h)
{
  double sum = v[0] + v[n - 1];
  for (int i = 1; i < n - 1; i += 2) sum += 4. * v[i];
  for (int i = 2; i < n - 1; i += 2) sum += 2. * v[i];
  return sum * h / 3.;
}
template double UpcCrossSection::simpson<double>(int, double*, double);

// rho0, G_AA, the form-factor spline and the breakup spline are built together on the device
double UpcCrossSection::calcWSRho() { ensureTables(); return rho0; }
void UpcCrossSection::prepareGAA() { ensureTables(); }
void UpcCrossSection::prepareFormFac() { ensureTables(); }
void UpcCrossSection::prepareBreakupProb() { ensureTables(); }

double UpcCrossSection::fluxPoint(const double b, const double k)
{
  ensureTables();
  double out = 0;
  int rc = upcgpu_flux_point(ctx, &b, &k, 1, &out);
  if (rc) fail("upcgpu_flux_point", rc);
  return out;
}

double UpcCrossSection::fluxForm(const double b, const double k)
{
  ensureTables();
  double out = 0;
  int rc = upcgpu_flux_form(ctx, &b, &k, 1, &out, nullptr);
  if (rc) fail("upcgpu_flux_form", rc);
  return out;
}

// analytic Woods-Saxon form factor; static in the reference (it is only called directly by the
// vector-meson path), so it is a plain host function here as well
double UpcCrossSection::calcFormFac(double Q2)
{
  double Q = std::sqrt(Q2) / phys_consts::hc;
  double coshVal = std::cosh(M_PI * Q * a);
  double sinhVal = std::sinh(M_PI * Q * a);
  double ff = 4 * M_PI * M_PI * rho0 * a * a * a / (Q * a * Q * a * sinhVal * sinhVal) *
              (M_PI * Q * a * coshVal * std::sin(Q * R) - Q * R * std::cos(Q * R) * sinhVal);
  ff += 8 * M_PI * rho0 * a * a * a * std::exp(-R / a) / (1 + Q * Q * a * a) / (1 + Q * Q * a * a);
  return ff;
}

double UpcCrossSection::calcTwoPhotonLumi(double M, double Y)
{
  ensureTables();
  double out = 0;
  int rc = upcgpu_lumi_cells(ctx, &M, &Y, 1, &out, nullptr, nullptr);
  if (rc) fail("upcgpu_lumi_cells", rc);
  return out;
}

void UpcCrossSection::calcTwoPhotonLumiPol(double& ns, double& np, double M, double Y)
{
  ensureTables();
  int rc = upcgpu_lumi_cells(ctx, &M, &Y, 1, nullptr, &ns, &np);
  if (rc) fail("upcgpu_lumi_cells", rc);
}

void UpcCrossSection::fillCrossSectionZM(std::vector<std::vector<double>>& crossSectionZM, double zmin, double zmax,
                                         int nz, double mmin, double mmax, int nm, int flag)
{
  constexpr double scalingFactor = phys_consts::hc * phys_consts::hc * 1e7; // to [nb]
  double dm = (mmax - mmin) / nm;
  double dz = (zmax - zmin) / nz;
  double cs = 0;
  for (int im = 0; im < nm; ++im) {
    double m = mmin + dm * im;
    for (int iz = 0; iz < nz; ++iz) {
      double z = zmin + dz * iz;
      if (flag == 0) cs = elemProcess->calcCrossSectionZM(z, m);
      if (flag == 1) cs = elemProcess->calcCrossSectionZMPolS(z, m);
      if (flag == 2) cs = elemProcess->calcCrossSectionZMPolPS(z, m);
      crossSectionZM[im][iz] = cs * scalingFactor / dm;
    }
  }
}

namespace
{
// The lock file of the reference (src/UpcCrossSection.cpp:465-478, :587-591): <dir>/.lumiIsCalculated[Pol], created with
// O_CREAT | O_EXCL before the cache is looked for, removed after it is read or written; a generator that finds it waits
// in one-second steps.  Kept so that this build and the reference -- or several jobs of either -- can share one cache
// directory.  Two deviations: the descriptor is closed (the reference leaks it and treats descriptor 0 as a failure),
// and an error other than "exists" (a read-only directory, say) is a warning and no lock, not an endless wait.
struct LumiLock {
  std::string path;
  bool owned{false};
  explicit LumiLock(const std::string& p) : path(p)
  {
    int waited = 0;
    for (;;) {
      const int fd = ::open(path.c_str(), O_CREAT | O_EXCL, 0644);
      if (fd >= 0) { ::close(fd); owned = true; return; }
      if (errno != EEXIST) {
        PLOG_WARNING << "Cannot create the lock file " << path << " (" << std::strerror(errno) << "): continuing without it";
        return;
      }
      if (waited % 30 == 0) PLOG_INFO << "Another generator holds " << path << ": waiting";
      ::sleep(1);
      ++waited;
    }
  }
  ~LumiLock()
  {
    if (owned && std::remove(path.c_str()) != 0) PLOG_WARNING << "Lock file " << path << " was not properly removed!";
  }
};
} // namespace

void UpcCrossSection::prepareTwoPhotonLumi()
{
  ensureTables();
  LumiLock lock(std::string(lumiFileDirectory) + "/.lumiIsCalculated" + (usePolarizedCS ? "Pol" : ""));
  const size_t n = (size_t)nm * ny;
  // The cache file, as in the reference (src/UpcCrossSection.cpp:481-491, :578-585): twoPhotonLumi[Pol].root with the
  // TH2D hD2LDMDY (or hD2LDMDY_s / hD2LDMDY_p), bin (im + 1, iy + 1) = table[im][iy].  A file the REFERENCE wrote is
  // picked up here, and one written here is picked up by the reference (UpcRootHist / UpcRootFile: no ROOT needed).
  // Like the reference, an existing file is used without looking at the parameters it was made with -- except that a
  // file whose grid has other dimensions cannot be used and is recomputed.
  const std::string fname = std::string(lumiFileDirectory) + "/twoPhotonLumi" + (usePolarizedCS ? "Pol.root" : ".root");
  {
    std::ifstream probe(fname, std::ios::binary);
    if (probe.good()) {
      probe.close();
      const char* names[2] = {usePolarizedCS ? "hD2LDMDY_s" : "hD2LDMDY", "hD2LDMDY_p"};
      std::vector<double>* tabs[2] = {usePolarizedCS ? &lumiS : &lumi, &lumiPs};
      bool ok = true;
      std::string why;
      for (int t = 0; t < (usePolarizedCS ? 2 : 1) && ok; ++t) {
        UpcRootHist h;
        std::string err;
        if (!h.Read(fname, names[t], err)) { ok = false; why = err; break; }
        if (h.dim != 2 || h.GetNbinsX() != nm || h.GetNbinsY() != ny) {
          ok = false;
          why = "its grid is " + std::to_string(h.GetNbinsX()) + " x " + std::to_string(h.GetNbinsY()) + ", not " +
                std::to_string(nm) + " x " + std::to_string(ny);
          break;
        }
        if (h.fXaxis.fXmin != mmin || h.fXaxis.fXmax != mmax || h.fYaxis.fXmin != ymin || h.fYaxis.fXmax != ymax)
          PLOG_WARNING << fname << ": axis ranges differ from the current MMIN/MMAX/YMIN/YMAX (used as it is, like the reference does)";
        tabs[t]->resize(n);
        for (int im = 0; im < nm; ++im)
          for (int iy = 0; iy < ny; ++iy) (*tabs[t])[(size_t)im * ny + iy] = h.GetBinContent(im + 1, iy + 1);
      }
      if (ok) {
        PLOG_INFO << "Found pre-calculated " << (usePolarizedCS ? "polarized" : "unpolarized") << " 2D luminosity";
        int rc = usePolarizedCS ? upcgpu_lumi_upload(ctx, 1, lumiS.data()) : upcgpu_lumi_upload(ctx, 0, lumi.data());
        if (!rc && usePolarizedCS) rc = upcgpu_lumi_upload(ctx, 2, lumiPs.data());
        if (rc) fail("upcgpu_lumi_upload", rc);
        return;
      }
      PLOG_WARNING << "Cannot use " << fname << " (" << why << "): computing the luminosity again";
    }
  }
  PLOG_INFO << "Precalculated 2D luminosity is not found. Starting all over...";
  int rc;
  if (usePolarizedCS) {
    lumiS.assign(n, 0.); lumiPs.assign(n, 0.);
    rc = upcgpu_fill_lumi(ctx, nullptr, lumiS.data(), lumiPs.data());
  } else {
    lumi.assign(n, 0.);
    rc = upcgpu_fill_lumi(ctx, lumi.data(), nullptr, nullptr);
  }
  if (rc) fail("upcgpu_fill_lumi", rc);
  upcgpu_fill_stats st;
  upcgpu_get_fill_stats(ctx, &st);
  PLOG_INFO << "Two-photon luminosity: " << n << " cells in " << st.ms_total << " ms on " << upcgpu_group_size(ctx) << " GPU(s)";
  {
    UpcRootFileWriter w;
    // the reference writes this file with ROOT's default compression, zlib level 1 (setting 101); with the same name,
    // axes and cells the compressed TH2D record written here is byte-identical to ROOT's (tests/test_root_file.py)
    if (UpcRootFileDefaultCompression() == 0) w.SetCompression(101);
    auto cells_of = [&](const std::vector<double>& t) {
      std::vector<double> c((size_t)(nm + 2) * (ny + 2), 0.);
      for (int im = 0; im < nm; ++im)
        for (int iy = 0; iy < ny; ++iy) c[(size_t)(iy + 1) * (nm + 2) + (im + 1)] = t[(size_t)im * ny + iy];
      return c;
    };
    if (usePolarizedCS) {
      w.AddTH2D("hD2LDMDY_s", "", nm, mmin, mmax, ny, ymin, ymax, cells_of(lumiS), (double)n);
      w.AddTH2D("hD2LDMDY_p", "", nm, mmin, mmax, ny, ymin, ymax, cells_of(lumiPs), (double)n);
    } else {
      w.AddTH2D("hD2LDMDY", "", nm, mmin, mmax, ny, ymin, ymax, cells_of(lumi), (double)n);
    }
    std::string err;
    if (w.Write(fname, err)) PLOG_INFO << "Two-photon luminosity was written to " << fname;
    else PLOG_WARNING << "Two-photon luminosity was NOT cached: " << err;
  }
}

void UpcCrossSection::calcNucCrossSectionYM(std::vector<std::vector<double>>& crossSectionYM,
                                            std::vector<std::vector<double>>& polCSRatio, double& totCS)
{
  PLOG_INFO << "Calculating nuclear cross section...";
  ensureTables();
  const double dm = (mmax - mmin) / nm;
  std::vector<double> sig0(nm), sig1(nm), cs((size_t)nm * ny), ratio;
  for (int im = 0; im < nm; ++im) {
    double m = mmin + dm * im;
    if (!usePolarizedCS) {
      sig0[im] = elemProcess->calcCrossSectionM(m);
    } else {
      sig0[im] = elemProcess->calcCrossSectionMPolS(m);
      sig1[im] = elemProcess->calcCrossSectionMPolPS(m);
    }
  }
  double tot = 0;
  int rc;
  if (usePolarizedCS) {
    ratio.resize((size_t)nm * ny);
    rc = upcgpu_fold_sigma(ctx, nullptr, sig0.data(), sig1.data(), cs.data(), ratio.data(), &tot);
  } else {
    rc = upcgpu_fold_sigma(ctx, sig0.data(), nullptr, nullptr, cs.data(), nullptr, &tot);
  }
  if (rc) fail("upcgpu_fold_sigma", rc);
  for (int iy = 0; iy < ny; ++iy)
    for (int im = 0; im < nm; ++im) {
      crossSectionYM[iy][im] = cs[(size_t)iy * nm + im];
      if (usePolarizedCS && (int)polCSRatio.size() > iy && (int)polCSRatio[iy].size() > im)
        polCSRatio[iy][im] = ratio[(size_t)iy * nm + im];
    }
  totCS = totCS * 1e-6 + tot; // [mb]; the reference sums into the caller's variable, then scales it by 1e-6
  PLOG_INFO << "Calculating nuclear cross section...Done!";
  PLOG_INFO << "Total nuclear cross section = " << std::fixed << totCS << " mb";
}

double UpcCrossSection::calcBreakupProb(const double impactparameter, const int mode)
{
  ensureTables();
  double out = 0;
  int rc = upcgpu_breakup_raw(ctx, &impactparameter, mode, 1, &out);
  if (rc) fail("upcgpu_breakup_raw", rc);
  return out;
}

// per-call accessor kept for interface compatibility (the event loop of UpcGenerator uses the
// batched upcgpu_generate instead): pdf tabulated on the GPU, cached per integer-MeV key as the
// reference's photPtDistrMap, inverted with one uniform like TH1::GetRandom
double UpcCrossSection::getPhotonPt(double ePhot)
{
  ensureTables();
  int key = ePhot * 1e3;
  auto it = photPtCdf.find(key);
  if (it == photPtCdf.end()) {
    std::vector<double> cdf(5001);
    int rc = upcgpu_photon_pt_cdf(ctx, ePhot, cdf.data());
    if (rc) fail("upcgpu_photon_pt_cdf", rc);
    it = photPtCdf.emplace(key, std::move(cdf)).first;
  }
  const std::vector<double>& cdf = it->second;
  if (cdf[5000] == 0) return 0;
  double r1 = uniform(0., 1.);
  int lo = 0, hi = 5000;
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (cdf[mid] <= r1) lo = mid; else hi = mid;
  }
  double bw = 6. * phys_consts::hc / R / 5000;
  double x = lo * bw;
  if (r1 > cdf[lo]) x += bw * (r1 - cdf[lo]) / (cdf[lo + 1] - cdf[lo]);
  if (photPtCdf.size() > 30000) photPtCdf.clear();
  return x;
}

void UpcCrossSection::getPairMomentum(double mPair, double yPair, TLorentzVector& pPair)
{
  if (!useNonzeroGamPt) {
    double mtPair = mPair;
    pPair.SetPxPyPzE(0., 0., mtPair * std::sinh(yPair), mtPair * std::cosh(yPair));
    return;
  }
  double k1 = mPair / 2 * std::exp(yPair);
  double k2 = mPair / 2 * std::exp(-yPair);
  double angle1 = uniform(0, 2 * M_PI);
  double angle2 = uniform(0, 2 * M_PI);
  double pt1 = getPhotonPt(k1);
  double pt2 = getPhotonPt(k2);
  double px = pt1 * std::cos(angle1) + pt2 * std::cos(angle2);
  double py = pt1 * std::sin(angle1) + pt2 * std::sin(angle2);
  double pt = std::sqrt(px * px + py * py);
  double mtPair = std::sqrt(mPair * mPair + pt * pt);
  pPair.SetPxPyPzE(px, py, mtPair * std::sinh(yPair), mtPair * std::cosh(yPair));
}

// ---- vector-meson photoproduction path (src/UpcCrossSection.cpp:700-748, :1076-1104) ----
// the b-integrated photon flux, on the GPU (flux rows of the luminosity stage + one thread per photon energy)
double UpcCrossSection::calcPhotonFlux(double M, double Y)
{
  ensureTables();
  double out = 0;
  int rc = upcgpu_photon_flux(ctx, &M, &Y, 1, &out, nullptr);
  if (rc) fail("upcgpu_photon_flux", rc);
  return out;
}

// :724-748.  All 2 ny fluxes come from one GPU call; the fold with the plug-in's sigma(+-y) is the reference's loop.
void UpcCrossSection::calcNucCrossSectionY(std::vector<std::vector<double>>& crossSectionY,
                                           std::vector<std::vector<double>>& csYRatio, double& totCS)
{
  PLOG_INFO << "Calculating nuclear cross section...";
  ensureTables();
  const double dy = (ymax - ymin) / ny;
  std::vector<double> mm(ny, elemProcess->mPart), yy(ny), f1(ny), f2(ny);
  for (int iy = 0; iy < ny; iy++) yy[iy] = ymin + dy * iy;
  int rc = upcgpu_photon_flux(ctx, mm.data(), yy.data(), ny, f1.data(), f2.data());
  if (rc) fail("upcgpu_photon_flux", rc);
  for (int iy = 0; iy < ny; iy++) {
    const double y = yy[iy];
    const double cs1 = elemProcess->calcCrossSectionY(y);
    const double cs2 = elemProcess->calcCrossSectionY(-y);
    const double upcCs1 = f1[iy] * cs1;
    const double upcCs2 = f2[iy] * cs2;
    crossSectionY[iy][0] = upcCs1 + upcCs2;
    totCS += upcCs1 + upcCs2;
    csYRatio[iy][0] = upcCs1 / upcCs2;
  }
  totCS *= dy; // already in [mb] for VM
  PLOG_INFO << "Total nuclear cross section = " << std::fixed << totCS << " mb";
}

// :1076-1104: photon pT from the tabulated pdf, pomeron pT by rejection from the squared form factor
void UpcCrossSection::getMomentumVM(double m, double y, int target, TLorentzVector& p)
{
  double sign = target ? 1 : -1;
  double ePhot = m / 2 * std::exp(sign * y);
  double ePom = m / 2 * std::exp(-sign * y);
  double ptPhot = getPhotonPt(ePhot);
  double phi1 = uniform(0, 2 * M_PI);
  double phi2 = uniform(0, 2 * M_PI);
  double tmin = (ePom * ePom) / (g1 * g1);
  double ptPom;
  while (true) {
    ptPom = 32. * uniform(0., 1.) * phys_consts::hc * R;
    double t2 = tmin + ptPom * ptPom;
    double ff = calcFormFac(t2) / A;
    double test = uniform(0., 1.);
    if (test < ff * ff * ptPom) break;
  }
  double px = ptPhot * std::cos(phi1) + ptPom * std::cos(phi2);
  double py = ptPhot * std::sin(phi1) + ptPom * std::sin(phi2);
  double pt = std::sqrt(px * px + py * py);
  double mtPair = std::sqrt(m * m + pt * pt);
  p.SetPxPyPzE(px, py, mtPair * std::sinh(y), mtPair * std::cosh(y));
}
