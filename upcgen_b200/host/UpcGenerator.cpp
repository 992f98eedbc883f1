// UpcGenerator over the GPU path.  Counterpart of the reference's src/UpcGenerator.cpp for the
// two-photon processes with closed-form elementary cross sections (dileptons, ALP).
#include "UpcGenerator.h"
#include "UpcRootFile.h"

#include <algorithm>
#include <cstdlib>
#include <ctime>
#include <random>
#include <sstream>

int UpcGenerator::debug = 0;

namespace upc_host
{
static upcgpu_ctx* g_samplerCtx = nullptr;
static upcgpu_ctx* g_ownCtx = nullptr;
void registerSamplerContext(upcgpu_ctx* ctx) { g_samplerCtx = ctx; }
upcgpu_ctx* samplerContext()
{
  if (g_samplerCtx) return g_samplerCtx;
  if (!g_ownCtx) {
    UpcCrossSection defaults;
    defaults.setElemProcess(11);
    upcgpu_params p = defaults.makeParams();
    if (upcgpu_create(&p, 0, &g_ownCtx) != UPCGPU_OK) {
      PLOG_FATAL << "UpcSampler: " << upcgpu_last_error(nullptr);
      std::_Exit(-1);
    }
  }
  return g_ownCtx;
}
// gRandom stand-in of the host-side event code (vector mesons): std::mt19937_64, as TRandomMT64
static std::mt19937_64& genRng()
{
  static std::mt19937_64 rng(0x5eedULL);
  return rng;
}
void seedHost(uint64_t s) { genRng().seed(s); }
double hostUniform(double a, double b) { return a + (b - a) * std::generate_canonical<double, 53>(genRng()); }
} // namespace upc_host

UpcGenerator::UpcGenerator()
{
  nucProcessCS = new UpcCrossSection();
  totCS = 0.0;
  fidCS = 0.0;
  setCollisionSystem(5020., 82, 208);
}

UpcGenerator::~UpcGenerator()
{
  delete samplerCsYM;
  upc_host::registerSamplerContext(nullptr);
  delete nucProcessCS; // (the reference leaks it)
}

void UpcGenerator::setCollisionSystem(float sqrts, int nucl_z, int nucl_a)
{
  UpcCrossSection::sqrts = sqrts;
  UpcCrossSection::g1 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
  UpcCrossSection::g2 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
  UpcCrossSection::Z = nucl_z;
  UpcCrossSection::A = nucl_a;
  UpcCrossSection::mNucl = (nucl_z * phys_consts::mProt + (nucl_a - nucl_z) * phys_consts::mNeut) / nucl_a;
}

// KEY value pairs of parameters.in; unknown keys are ignored (src/UpcGenerator.cpp:183-305)
void UpcGenerator::setParameterValue(const std::string& parameter, const std::string& parValue)
{
  using std::stod;
  using std::stoi;
  using std::stol;
  auto* cs = nucProcessCS;
  if (parameter == "NEVENTS") nEvents = stol(parValue);
  if (parameter == "SQRTS") {
    UpcCrossSection::sqrts = stod(parValue);
    UpcCrossSection::g1 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
    UpcCrossSection::g2 = UpcCrossSection::sqrts / (2. * phys_consts::mProt);
  }
  if (parameter == "PROC_ID") procID = stoi(parValue);
  if (parameter == "LEP_A") aLep = stod(parValue);
  if (parameter == "ALP_MASS") cs->alpMass = stod(parValue);
  if (parameter == "ALP_WIDTH") cs->alpWidth = stod(parValue);
  if (parameter == "DO_PT_CUT") doPtCut = stoi(parValue);
  if (parameter == "PT_MIN") minPt = stod(parValue);
  if (parameter == "DO_ETA_CUT") doEtaCut = stoi(parValue);
  if (parameter == "ETA_MIN") minEta = stod(parValue);
  if (parameter == "ETA_MAX") maxEta = stod(parValue);
  if (parameter == "ZMIN") cs->zmin = stod(parValue);
  if (parameter == "ZMAX") cs->zmax = stod(parValue);
  if (parameter == "MMIN") cs->mmin = stod(parValue);
  if (parameter == "MMAX") cs->mmax = stod(parValue);
  if (parameter == "YMIN") cs->ymin = stod(parValue);
  if (parameter == "YMAX") cs->ymax = stod(parValue);
  if (parameter == "BINS_Z") cs->nz = stoi(parValue);
  if (parameter == "BINS_M") cs->nm = stoi(parValue);
  if (parameter == "BINS_Y") cs->ny = stoi(parValue);
  if (parameter == "WS_R") UpcCrossSection::R = stod(parValue);
  if (parameter == "WS_A") UpcCrossSection::a = stod(parValue);
  if (parameter == "NUCLEUS_Z") UpcCrossSection::Z = stoi(parValue);
  if (parameter == "NUCLEUS_A") UpcCrossSection::A = stoi(parValue);
  if (parameter == "FLUX_POINT") cs->isPoint = stoi(parValue);
  if (parameter == "BREAKUP_MODE") cs->breakupMode = stoi(parValue);
  if (parameter == "PYTHIA_VERSION") pythiaVersion = stoi(parValue);
  if (parameter == "PYTHIA8_FSR") doFSR = stoi(parValue);
  if (parameter == "PYTHIA8_DECAYS") doDecays = stoi(parValue);
  if (parameter == "NON_ZERO_GAM_PT") cs->useNonzeroGamPt = stoi(parValue);
  if (parameter == "USE_POLARIZED_CS") {
    cs->usePolarizedCS = stoi(parValue);
    usePolarizedCS = stoi(parValue);
  }
  if (parameter == "SEED") seed = stol(parValue);
  if (parameter == "USE_ROOT_OUTPUT") useROOTOut = stoi(parValue);
  if (parameter == "USE_HEPMC_OUTPUT") useHepMCOut = stoi(parValue);
  if (parameter == "DO_M_CUT") cs->doMassCut = stoi(parValue);
  if (parameter == "LOW_M_CUT") cs->lowMCut = stod(parValue);
  if (parameter == "HIGH_M_CUT") cs->hiMCut = stod(parValue);
  if (parameter == "SHADOWING") cs->shadowingOption = stoi(parValue);
  if (parameter == "DECAY_PDG") cs->dghtPDG = stoi(parValue);
}

void UpcGenerator::configGeneratorFromFile()
{
  std::ifstream fInputs(parFileName);
  if (fInputs) {
    PLOG_INFO << "Reading parameters from " << parFileName << " ...";
    std::string line, parameter, parValue;
    while (getline(fInputs, line)) {
      std::istringstream iss(line);
      if (line[0] == '#') continue;
      iss >> parameter >> parValue; // not reset between lines, as the reference (a blank line re-applies the last pair)
      setParameterValue(parameter, parValue);
    }
    if (!useROOTOut && !useHepMCOut)
      PLOG_WARNING << "Output format not set! Choose ROOT or/and HepMC via flags USE_ROOT_OUTPUT and USE_HEPMC_OUTPUT!";
  } else {
    PLOG_WARNING << "Input file not found! Using default parameters...";
  }
}

void UpcGenerator::printParameters()
{
  auto* cs = nucProcessCS;
  PLOG_INFO << "OMP_NTHREADS " << numThreads << " (unused: GPU build)";
  PLOG_INFO << "NUCLEUS_Z " << UpcCrossSection::Z;
  PLOG_INFO << "NUCLEUS_A " << UpcCrossSection::A;
  PLOG_INFO << "WS_R " << UpcCrossSection::R;
  PLOG_INFO << "WS_A " << UpcCrossSection::a;
  PLOG_INFO << "SQRTS " << UpcCrossSection::sqrts;
  PLOG_INFO << "PROC_ID " << procID;
  PLOG_INFO << "LEP_A " << aLep;
  PLOG_INFO << "NEVENTS " << nEvents;
  PLOG_INFO << "DO_PT_CUT " << doPtCut << " PT_MIN " << minPt << " DO_ETA_CUT " << doEtaCut << " ETA_MIN " << minEta
            << " ETA_MAX " << maxEta;
  PLOG_INFO << "ZMIN " << cs->zmin << " ZMAX " << cs->zmax << " MMIN " << cs->mmin << " MMAX " << cs->mmax << " YMIN "
            << cs->ymin << " YMAX " << cs->ymax;
  PLOG_INFO << "BINS_Z " << cs->nz << " BINS_M " << cs->nm << " BINS_Y " << cs->ny;
  PLOG_INFO << "FLUX_POINT " << cs->isPoint << " BREAKUP_MODE " << cs->breakupMode << " NON_ZERO_GAM_PT "
            << cs->useNonzeroGamPt << " USE_POLARIZED_CS " << usePolarizedCS;
  PLOG_INFO << "SEED " << seed;
}

void UpcGenerator::init()
{
  if (nucProcessCS == nullptr) {
    PLOG_FATAL << "UpcCrossSection was not initialized! Exiting...";
    std::_Exit(-1);
  }
  ignoreCSZ = false;
  auto* cs = nucProcessCS;
  // process-specific set-up, src/UpcGenerator.cpp:69-140
  if (procID == 22 || procID == 111) {
    // the elementary cross sections come from files with a fixed (z, m) grid: the grid is set explicitly (:69-103)
    isPairProduction = true;
    PLOG_WARNING << "For this process grid sizes along Z and M are fixed -- see parameters check";
    if (procID == 22) {
      cs->zmin = -0.99; cs->zmax = 0.99; cs->nz = 198;
      cs->mmin = 0.05; cs->mmax = 50.; cs->nm = 1000;
    } else {
      cs->zmin = -1; cs->zmax = 1; cs->nz = 100;
      cs->mmin = 0.275; cs->mmax = 5.; cs->nm = 91;
    }
    if (usePolarizedCS) {
      // The reference clears only the generator's flag here and leaves the cross-section object's set, which makes it
      // fold the polarised luminosities with cross sections that are identically 0 and write a ratio table that was never
      // sized.  Both flags are cleared here: the process is run unpolarised, as the warning says.
      PLOG_WARNING << "For this process polarized cross section is not available";
      usePolarizedCS = false;
      cs->usePolarizedCS = false;
    }
  }
  if (procID == 51) {
    ignoreCSZ = true; // cos(theta) uniform for the ALP
    isSingleProduction = true;
    cs->mmin = cs->alpMass - 4. * cs->alpWidth;
    cs->mmax = cs->alpMass + 4. * cs->alpWidth;
    PLOG_WARNING << "For ALP production angular distribution is ignored!";
  }
  cs->setElemProcess(procID);
  if (procID == 443 || procID == 100443 || procID == 553) {  // :120-130
    isSingleProduction = true; // one particle
    isPairProductionVM = true; // two-part decay
    ignoreCSZ = true;
    double mPart = cs->elemProcess->mPart;
    cs->mmin = mPart - 1e-6;
    cs->mmax = mPart + 1e-6;
    cs->nm = 1;
  }
  if (procID >= 11 && procID <= 15) {
    isPairProduction = true;
    auto* proc = (UpcTwoPhotonDilep*)cs->elemProcess;
    proc->aLep = aLep;
    if (cs->mmin < proc->mPart * 2.) {
      PLOG_WARNING << "MMIN is lower than 2 lepton masses! Setting MMIN to 2 lepton masses...";
      cs->mmin = proc->mPart * 2.;
    }
  }
  cs->numThreads = numThreads;
  cs->evIsPair = isPairProduction;
  cs->evIsSingle = isSingleProduction;
  cs->evIgnoreCSZ = ignoreCSZ;
  cs->evDecayUniformPDG = (procID == 51 || procID == 111) ? 22 : 0;  // :799-808: the ALP, or both pi0 of the pair
  cs->evDoPtCut = doPtCut; cs->evMinPt = minPt;
  cs->evDoEtaCut = doEtaCut; cs->evMinEta = minEta; cs->evMaxEta = maxEta;

  PLOG_WARNING << "Check inputs:";
  printParameters();
  cs->init(); // tables + two-photon luminosity on the GPU
  upc_host::registerSamplerContext(cs->gpu());
  if (seed == 0) seed = time(nullptr); // the reference seeds gRandom with the wall clock for SEED 0
  upc_host::seedHost((uint64_t)seed);
  if ((doFSR || doDecays) && pythiaVersion > 0)
    PLOG_WARNING << "Decays with Pythia are not used! (Pythia is not part of the GPU build)";
  computeNuclXsection();
}

void UpcGenerator::computeNuclXsection()
{
  auto* cs = nucProcessCS;
  const int nm = cs->nm, nz = cs->nz, ny = cs->ny;
  const double dm = (cs->mmax - cs->mmin) / nm, dz = (cs->zmax - cs->zmin) / nz, dy = (cs->ymax - cs->ymin) / ny;
  if (procID == 443 || procID == 100443 || procID == 553) {
    // vector mesons: a 1-D table in y (src/UpcGenerator.cpp:701-712)
    nucCSYM.assign(ny, std::vector<double>(1, 0.));
    nucTargRatioCSYM.assign(ny, std::vector<double>(1, 0.));
    totCS = 0;
    cs->calcNucCrossSectionY(nucCSYM, nucTargRatioCSYM, totCS);
    binEdgesM.assign(2, 0.);
    binEdgesM[0] = cs->mmin;
    binEdgesM[1] = cs->mmax;
    binEdgesY.resize(ny + 1);
    for (int i = 0; i < ny + 1; ++i) binEdgesY[i] = cs->ymin + dy * i;
    delete samplerCsYM;
    samplerCsYM = new UpcSampler2D(nucCSYM, binEdgesY, binEdgesM, seed);
    return;
  }
  std::vector<std::vector<double>> csZM, csZMS, csZMPS;
  if (usePolarizedCS) {
    csZMS.resize(nm, std::vector<double>(nz, 0.));
    csZMPS.resize(nm, std::vector<double>(nz, 0.));
    cs->fillCrossSectionZM(csZMS, cs->zmin, cs->zmax, nz, cs->mmin, cs->mmax, nm, 1);
    cs->fillCrossSectionZM(csZMPS, cs->zmin, cs->zmax, nz, cs->mmin, cs->mmax, nm, 2);
  } else {
    csZM.resize(nm, std::vector<double>(nz, 0.));
    cs->fillCrossSectionZM(csZM, cs->zmin, cs->zmax, nz, cs->mmin, cs->mmax, nm, 0);
  }
  nucCSYM.assign(ny, std::vector<double>(nm, 0.));
  if (usePolarizedCS) polCSRatio.assign(ny, std::vector<double>(nm));
  totCS = 0;
  cs->calcNucCrossSectionYM(nucCSYM, polCSRatio, totCS);

  binEdgesM.resize(nm + 1);
  binEdgesZ.resize(nz + 1);
  binEdgesY.resize(ny + 1);
  for (int i = 0; i < nm + 1; ++i) binEdgesM[i] = cs->mmin + dm * i;
  for (int i = 0; i < nz + 1; ++i) binEdgesZ[i] = cs->zmin + dz * i;
  for (int i = 0; i < ny + 1; ++i) binEdgesY[i] = cs->ymin + dy * i;

  // samplers: the 2-D (y, m) CDF from the folded table already on the device, the nm z-CDFs from the
  // plug-in's d sigma/dz table (src/UpcGenerator.cpp:684-700)
  auto flat = [&](const std::vector<std::vector<double>>& t) {
    std::vector<double> f((size_t)nm * nz);
    for (int i = 0; i < nm; ++i) std::copy(t[i].begin(), t[i].end(), f.begin() + (size_t)i * nz);
    return f;
  };
  int rc;
  if (ignoreCSZ) {
    rc = upcgpu_sampler_build(cs->gpu(), nullptr, nullptr, nullptr, nullptr);
  } else if (usePolarizedCS) {
    auto fs = flat(csZMS), fp = flat(csZMPS);
    rc = upcgpu_sampler_build(cs->gpu(), nullptr, nullptr, fs.data(), fp.data());
  } else {
    auto f0 = flat(csZM);
    rc = upcgpu_sampler_build(cs->gpu(), nullptr, f0.data(), nullptr, nullptr);
  }
  if (rc) {
    PLOG_FATAL << "upcgpu_sampler_build failed: " << upcgpu_last_error(cs->gpu());
    std::_Exit(-1);
  }
}

void UpcGenerator::refillBlock()
{
  const size_t n = 1 << 16;
  // the particle arrays carry as many slots per candidate as an event of this process has particles
  const int ps = upcgpu_particles_per_event(nucProcessCS->gpu());
  block.stride = ps;
  block.npart.resize(n);
  block.pdg.resize(n * ps);
  block.status.resize(n * ps);
  block.mother.resize(n * ps);
  block.p4.resize(n * ps * 4);
  uint64_t nacc = 0;
  int rc = upcgpu_generate_packed(nucProcessCS->gpu(), (uint64_t)seed, nextCandidate, n, ps, block.npart.data(),
                                  block.pdg.data(), block.status.data(), block.mother.data(), block.p4.data(), nullptr, &nacc);
  if (rc) {
    PLOG_FATAL << "upcgpu_generate_packed failed: " << upcgpu_last_error(nucProcessCS->gpu());
    std::_Exit(-1);
  }
  nextCandidate += n;
  block.pos = 0;
  block.n = n;
}

// one candidate per call; returns 1 when it passes the kinematic cuts, else 0 with empty vectors
// vector-meson event, host code following src/UpcGenerator.cpp:715-832 for isVM
long int UpcGenerator::generateEventVM(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                                       std::vector<TLorentzVector>& particles)
{
  auto* cs = nucProcessCS;
  double mPair, yPair;
  (*samplerCsYM)(yPair, mPair);
  int yPairBin = samplerCsYM->getBinX(yPair);
  int mPairBin = samplerCsYM->getBinY(mPair);
  if (yPairBin < 0) yPairBin = 0;
  if (yPairBin >= (int)nucTargRatioCSYM.size()) yPairBin = (int)nucTargRatioCSYM.size() - 1;
  if (mPairBin != 0) mPairBin = 0;  // one mass bin (the reference's getBinY can leave it through its integer arithmetic)
  (void)upc_host::hostUniform(-1., 1.);  // cos(theta) of ignoreCSZ processes: drawn, then unused (:756)
  TLorentzVector pPair;
  double ratio = nucTargRatioCSYM[yPairBin][mPairBin];
  bool target = upc_host::hostUniform(0, 1) < ratio;
  cs->getMomentumVM(mPair, yPair, target, pPair);
  // singleProduction :474-485
  particles.emplace_back(pPair);
  pdgs.emplace_back(cs->elemProcess->partPDG);
  mothers.emplace_back(0);
  statuses.emplace_back(23);
  if (!checkKinCuts(particles)) {
    pdgs.clear(); statuses.clear(); mothers.clear(); particles.clear();
    return 0;
  }
  if (isPairProductionVM) twoPartDecayVM(pdgs, statuses, mothers, particles, 1);
  genParticles.clear();
  for (size_t ii = 0; ii < particles.size(); ++ii)
    genParticles.emplace_back(pdgs[ii], statuses[ii], mothers[ii], mothers[ii], -1, -1, particles[ii].Px(), particles[ii].Py(),
                              particles[ii].Pz(), particles[ii].E(), 0.0, 0.0, 0.0, 0.0);
  return 1;
}

// :563-587
bool UpcGenerator::checkKinCuts(std::vector<TLorentzVector>& particles)
{
  for (const auto& tlvec : particles) {
    if (doPtCut && tlvec.Pt() < minPt) return false;
    if (doEtaCut) {
      double eta = tlvec.Eta();
      if (eta < minEta || eta > maxEta) return false;
    }
  }
  return true;
}

// :425-472: decay angle by rejection from 1 + cos^2 (leptons) or 1 + 0.605 cos^2 (protons), J. Breitweg et al.
void UpcGenerator::twoPartDecayVM(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                                  std::vector<TLorentzVector>& particles, int id)
{
  auto* cs = nucProcessCS;
  int decayProdPDG = cs->elemProcess->dghtPDG;
  double theta, dndtheta = 0;
  while (true) {
    theta = M_PI * upc_host::hostUniform(0., 1.);
    double test = upc_host::hostUniform(0., 1.);
    if (decayProdPDG == 11 || decayProdPDG == 13) dndtheta = std::sin(theta) * (1. + (std::cos(theta) * std::cos(theta)));
    else if (decayProdPDG == 2212) dndtheta = std::sin(theta) * (1. + (0.605 * std::cos(theta) * std::cos(theta)));
    if (test < dndtheta) break;
  }
  int sign1 = upc_host::hostUniform(-1, 1) > 0 ? 1 : -1;
  int sign2 = -sign1;
  const int status = 33;
  const TLorentzVector particle = particles[id - 1];
  double mDecay = cs->elemProcess->mDght;
  double pMag = std::sqrt(particle.Mag2() / 4. - mDecay * mDecay);
  double phi = upc_host::hostUniform(0., 2. * M_PI);
  const double amag = std::fabs(pMag);
  double v[3] = {amag * std::sin(theta) * std::cos(phi), amag * std::sin(theta) * std::sin(phi), amag * std::cos(theta)};
  const double bx = particle.Px() / particle.E(), by = particle.Py() / particle.E(), bz = particle.Pz() / particle.E();
  const double pm = particle.P();
  const double tot = pm > 0 ? 1.0 / pm : 1.0;
  const double u1 = particle.Px() * tot, u2 = particle.Py() * tot, u3 = particle.Pz() * tot;
  for (int k = 0; k < 2; ++k) {
    const double sgn = k == 0 ? -1. : 1.;
    double x = sgn * v[0], y = sgn * v[1], z = sgn * v[2];
    const double e0 = std::sqrt(x * x + y * y + z * z + mDecay * mDecay);
    // TVector3::RotateUz
    double up = u1 * u1 + u2 * u2;
    if (up) {
      up = std::sqrt(up);
      const double px = x, py = y, pz = z;
      x = (u1 * u3 * px - u2 * py + u1 * up * pz) / up;
      y = (u2 * u3 * px + u1 * py + u2 * up * pz) / up;
      z = (u3 * u3 * px - px + u3 * up * pz) / up;
    } else if (u3 < 0.) {
      x = -x; z = -z;
    }
    // TLorentzVector::Boost
    const double b2 = bx * bx + by * by + bz * bz;
    const double gamma = 1.0 / std::sqrt(1.0 - b2);
    const double bp = bx * x + by * y + bz * z;
    const double gamma2 = b2 > 0 ? (gamma - 1.0) / b2 : 0.0;
    TLorentzVector d(x + gamma2 * bp * bx + gamma * bx * e0, y + gamma2 * bp * by + gamma * by * e0,
                     z + gamma2 * bp * bz + gamma * bz * e0, gamma * (e0 + bp));
    pdgs.emplace_back((k == 0 ? sign1 : sign2) * decayProdPDG);
    statuses.emplace_back(status);
    mothers.emplace_back(id);
    particles.emplace_back(d);
  }
}

long int UpcGenerator::generateEvent(std::vector<int>& pdgs, std::vector<int>& statuses, std::vector<int>& mothers,
                                     std::vector<TLorentzVector>& particles)
{
  pdgs.clear();
  statuses.clear();
  mothers.clear();
  particles.clear();
  if (procID == 443 || procID == 100443 || procID == 553) return generateEventVM(pdgs, statuses, mothers, particles);
  if (block.pos == block.n) refillBlock();
  const size_t i = block.pos++;
  const int np = block.npart[i];
  if (np == 0) return 0;
  genParticles.clear();
  for (int j = 0; j < np; ++j) {
    const size_t o = i * block.stride + j;
    pdgs.emplace_back(block.pdg[o]);
    statuses.emplace_back(block.status[o]);
    mothers.emplace_back(block.mother[o]);
    particles.emplace_back(block.p4[o * 4 + 0], block.p4[o * 4 + 1], block.p4[o * 4 + 2], block.p4[o * 4 + 3]);
    genParticles.emplace_back(block.pdg[o], block.status[o], block.mother[o], block.mother[o], -1, -1,
                              block.p4[o * 4 + 0], block.p4[o * 4 + 1], block.p4[o * 4 + 2], block.p4[o * 4 + 3], 0.0,
                              0.0, 0.0, 0.0);
  }
  return 1;
}

void UpcGenerator::writeEvent(long int evt, const std::vector<int>& pdgs, const std::vector<int>& statuses,
                              const std::vector<int>& mothers, const std::vector<TLorentzVector>& particles)
{
  if (useROOTOut) {
    // the rows of the tree "particles" (src/UpcGenerator.cpp:595-608); written by generateEvents when the loop ends
    for (size_t i = 0; i < particles.size(); i++) {
      treeCols[0].push_back((double)(int)evt);  // eventNumber/I: the low 32 bits of the long it is bound to (Q10)
      treeCols[1].push_back(pdgs[i]);
      treeCols[2].push_back((double)(i + 1));
      treeCols[3].push_back(statuses[i]);
      treeCols[4].push_back(mothers[i]);
      treeCols[5].push_back(particles[i].Px());
      treeCols[6].push_back(particles[i].Py());
      treeCols[7].push_back(particles[i].Pz());
      treeCols[8].push_back(particles[i].E());
    }
  }
  if (!writerHepMC) return;
  int nVertices = -1;
  int lastMotherId = -1;
  for (auto mother : mothers) {
    if (mother != lastMotherId) {
      nVertices++;
      lastMotherId = mother;
    }
  }
  writerHepMC->writeEventInfo(evt, static_cast<int>(particles.size()), nVertices);
  for (size_t i = 0; i < particles.size(); ++i)
    writerHepMC->writeParticleInfo((int)i + 1, mothers[i], pdgs[i], particles[i].Px(), particles[i].Py(),
                                   particles[i].Pz(), particles[i].E(), particles[i].M(), statuses[i]);
}

void UpcGenerator::generateEvents()
{
  std::vector<int> pdgs, statuses, mothers;
  std::vector<TLorentzVector> particles;
  if (useROOTOut) {
    // events.root, tree "particles" with the reference's nine branches (src/UpcGenerator.cpp:842-857), written without
    // ROOT by UpcRootFile.cpp when the loop ends (uncompressed unless UPCGEN_ROOT_COMPRESSION=409 asks for the LZ4 records of the reference, :843)
    PLOG_WARNING << "Using ROOT tree for output!";
    PLOG_INFO << "Events will be written to events.root";
    treeCols.assign(9, std::vector<double>());
  }
  if (useHepMCOut) writerHepMC = new WriterHepMC("events.hepmc");
  PLOG_INFO << "Generating " << nEvents << " events...";
  long int rejected = 0;
  long int evt = 0;
  while (evt < nEvents) {
    if (debug > 1) PLOG_DEBUG << "Event number: " << evt + 1;  // :868-870
    if (debug <= 1 && ((evt + 1) % 100000 == 0)) PLOG_INFO << "Event number: " << evt + 1;
    if (generateEvent(pdgs, statuses, mothers, particles) == 1) {
      writeEvent(evt, pdgs, statuses, mothers, particles);
      evt++;
    } else {
      rejected++;
    }
  }
  fidCS = totCS * (double)nEvents / (double)(nEvents + rejected);
  PLOG_INFO << "Event generation is finished!";
  if ((doPtCut || doEtaCut) && nEvents > 0) {
    PLOG_INFO << "Kinematic cuts were used";
    PLOG_INFO << "Number of rejected events = " << rejected;
    PLOG_INFO << std::fixed << std::setprecision(6) << "Cross section with cuts = " << fidCS << " mb";
  }
  if (useROOTOut) {
    static const char* names[9] = {"eventNumber", "pdgCode", "particleID", "statusID", "motherID", "px", "py", "pz", "e"};
    std::vector<UpcRootFileWriter::Column> cols(9);
    for (int i = 0; i < 9; i++) {
      cols[i].name = names[i];
      cols[i].type = i < 5 ? 'I' : 'D';
      cols[i].values.swap(treeCols[i]);
    }
    UpcRootFileWriter w;
    std::string err;
    if (debug > 0) UpcAddSigmaHists(w, binEdgesY, binEdgesM, nucCSYM);  // :900-917: the cross section table and its projections
    w.AddTree("particles", "Generated particles", cols);
    if (!w.Write("events.root", err)) {
      PLOG_FATAL << "cannot write events.root: " << err;
      std::_Exit(-1);
    }
  }
  delete writerHepMC;
  writerHepMC = nullptr;
}
