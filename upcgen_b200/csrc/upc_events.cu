// upc_events.cu -- event stage E1-E5: counter-based (Philox4x32-10) sampling of (y, m), cos(theta),
// photon pT, pair kinematics, uniform two-body decay and kinematic cuts.
// Reference: src/UpcGenerator.cpp:715-832 (generateEvent), :388-423 (pairProduction), :474-485,
// :526-587; src/UpcCrossSection.cpp:1021-1074 (getPhotonPt, getPairMomentum); ROOT
// TLorentzVector/TVector3/TH1::GetRandom arithmetic restated.
//
// Uniform slot map: Philox counter = candidate index, block j:
//   0 -> (r1, r2) of the 2-D sampler      1 -> (scalar/pseudoscalar pick, z uniform)
//   2 -> (angle1, angle2)                 3 -> (u_pT1, u_pT2)
//   4 -> (phi, charge sign)               5 -> (phi_decay, cos(theta)_decay)
// Slots are consumed by position, so results do not depend on batch size or GPU count.
//
// Photon pT: the reference caches one 5000-bin pdf per integer-MeV photon energy in a host map
// (Q9).  Here a batch's photon energies are reduced to their distinct MeV keys on the device
// (radix sort + unique), one CTA tabulates the cumulative pdf of each key, and the event kernel
// inverts it; the pdf of a key is evaluated at the key's centre (key + 0.5) MeV.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>

#include "upc_ctx.h"
#include "upc_internal.h"
#include "upc_sampler.cuh"

namespace upc {

constexpr int kPtBins = 5000;

struct EvScratch {
  size_t cap = 0;        // candidates
  size_t cap_keys = 0;   // pT tables
  double *y = nullptr, *m = nullptr, *cost = nullptr;
  int *keys = nullptr, *keys_sorted = nullptr, *uniq = nullptr, *n_uniq = nullptr;
  double* cdf = nullptr;  // [cap_keys][kPtBins+1]
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  int *npart = nullptr, *pdg = nullptr, *status = nullptr, *mother = nullptr;
  double *p4 = nullptr, *aux = nullptr;
  unsigned long long* n_acc = nullptr;
  int* err = nullptr;
};

void free_event_scratch(upcgpu_ctx* c)
{
  EvScratch* s = (EvScratch*)c->ev;
  if (!s) return;
  cudaFree(s->y); cudaFree(s->m); cudaFree(s->cost); cudaFree(s->keys); cudaFree(s->keys_sorted); cudaFree(s->uniq);
  cudaFree(s->n_uniq); cudaFree(s->cdf); cudaFree(s->cub_tmp); cudaFree(s->npart); cudaFree(s->pdg);
  cudaFree(s->status); cudaFree(s->mother); cudaFree(s->p4); cudaFree(s->aux); cudaFree(s->n_acc); cudaFree(s->err);
  delete s;
  c->ev = nullptr;
}

struct EvParams {
  int ny, nm, nz;
  int pol, ignore_csz, nonzero_gam_pt;
  int is_pair, is_single, is_charged, part_pdg, decay_pdg;
  int do_pt_cut, do_eta_cut;
  double pt_min, eta_min, eta_max;
  double m_part, gtot, R;
  const double *sum2d, *sumz, *sumz_ps, *ratio, *ye, *me, *ze;
};

__device__ __forceinline__ int energy_key(double e) { return (int)(e * 1e3); }  // UpcCrossSection.cpp:1025

// phase 1: (y, m), bins, cos(theta), photon-energy keys
__global__ void k_ev_sample(EvParams P, uint64_t seed, uint64_t first, size_t n, double* __restrict__ y,
                            double* __restrict__ m, double* __restrict__ cost, int* __restrict__ keys,
                            int* __restrict__ err)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t cand = first + t;
  double r1, r2, u2, u3;
  philox4x32_10(seed, cand, 0, r1, r2);
  philox4x32_10(seed, cand, 1, u2, u3);
  long long k; double yy, mm;
  sample_ym_dev(P.sum2d, P.ny, P.nm, P.ye, P.me, r1, r2, k, yy, mm);  // UpcGenerator.cpp:737
  if (k < 0) {  // GSL: "cannot find r1 in cumulative pdf" -> abort in the reference
    atomicAdd(err, 1);
    y[t] = nan(""); m[t] = nan(""); cost[t] = 0;
    if (keys) { keys[2 * t] = 0; keys[2 * t + 1] = 0; }
    return;
  }
  const int ybin = get_bin(P.ny, yy, P.ye[0], P.ye[P.ny]);  // :738
  const int mbin = get_bin(P.nm, mm, P.me[0], P.me[P.nm]);  // :739
  double cz;
  if (!P.ignore_csz) {                                      // :743-757
    const double* tab = P.sumz;
    if (P.pol) {
      const double frac = P.ratio[(size_t)ybin * P.nm + mbin];
      tab = (u2 < frac) ? P.sumz : P.sumz_ps;
    }
    cz = sample_1d_dev(tab + (size_t)mbin * (P.nz + 1), P.nz, P.ze, u3);
  } else {
    cz = -1. + 2. * u3;  // gRandom->Uniform(-1., 1.)
  }
  y[t] = yy; m[t] = mm; cost[t] = cz;
  if (keys) {
    keys[2 * t] = energy_key(mm / 2 * exp(yy));       // UpcCrossSection.cpp:1060-1061
    keys[2 * t + 1] = energy_key(mm / 2 * exp(-yy));
  }
}

// phase 3: cumulative pT pdf of one photon-energy key per CTA (getPhotonPt, :1021-1038, and
// TH1::ComputeIntegral): cdf[0] = 0, cdf[i] = sum_{j<=i} p_j / total
__device__ __forceinline__ double ff_lookup(const SplineSeg* __restrict__ ff, double t)
{
  // the reference calls gsl_spline_eval without a clamp (GSL domain error beyond the last
  // knot); semantics here and in the oracle: clamp to the last knot
  const double tmax = kQ2min + (kNQ2 - 1) * kDQ2;
  t = fmin(t, tmax);
  int idx = (int)((t - kQ2min) * (1. / kDQ2));
  idx = max(0, min(idx, kNQ2 - 2));
  return seg_eval(ff[idx], t - fma((double)idx, kDQ2, kQ2min));
}

constexpr int kPtPerThread = (kPtBins + 255) / 256;  // 20 bins per thread
constexpr int kPtRunPad = kPtPerThread + 1;          // runs padded to 21 doubles in shared memory (2-way conflicts)

__global__ void __launch_bounds__(256) k_pt_tables(int n_keys, const int* __restrict__ keys, const double* e_direct,
                                                   const SplineSeg* __restrict__ ff, double gtot, double R,
                                                   double* __restrict__ cdf)
{
  // The pdf is evaluated and the cdf stored with bin = base + thread (neighbouring threads gather neighbouring
  // form-factor segments and write neighbouring doubles); the prefix sum goes through shared memory: each thread
  // sums a RUN of 20 consecutive bins serially and ONE block scan ranks the runs (a block scan per 256 bins -- 20
  // of them, two barriers each -- was 60 % of this kernel's instructions).
  typedef cub::BlockScan<double, 256> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double sp[256 * kPtRunPad];
  const int kidx = blockIdx.x;
  if (kidx >= n_keys) return;
  const int tid = threadIdx.x;
  const double e = e_direct ? e_direct[kidx] : (keys[kidx] + 0.5) * 1e-3;
  const double ereds = (e * e) / (gtot * gtot);
  const double pi2x4 = 4 * kPi * kPi;
  double* out = cdf + (size_t)kidx * (kPtBins + 1);
#pragma unroll 4
  for (int it = 0; it < kPtPerThread; ++it) {
    const int b0 = it * 256 + tid;           // 0-based bin
    double prob = 0;
    if (b0 < kPtBins) {
      const double pt = 6. * kHc / R / kPtBins * (b0 + 1);  // upper bin edge, :1033
      const double arg = pt * pt + ereds;
      const double f = ff_lookup(ff, arg);
      prob = (f * f) * pt * pt * pt / (pi2x4 * arg * arg);
    }
    sp[(b0 / kPtPerThread) * kPtRunPad + b0 % kPtPerThread] = prob;
  }
  __syncthreads();
  double acc = 0;
  double* run = sp + tid * kPtRunPad;
#pragma unroll
  for (int j = 0; j < kPtPerThread; ++j) {
    acc += run[j];
    run[j] = acc;
  }
  double before, tot;
  Scan(tmp).ExclusiveSum(acc, before, tot);
  run[kPtPerThread] = before;                // the pad slot carries the run's offset
  __syncthreads();
  if (tid == 0) out[0] = 0;
#pragma unroll 4
  for (int it = 0; it < kPtPerThread; ++it) {
    const int b0 = it * 256 + tid;
    if (b0 < kPtBins) {
      const int r = b0 / kPtPerThread;
      const double v = sp[r * kPtRunPad + kPtPerThread] + sp[r * kPtRunPad + b0 % kPtPerThread];
      out[b0 + 1] = tot != 0 ? v / tot : v;
    }
  }
}

// TH1::GetRandom on a tabulated integral
__device__ __forceinline__ double pt_sample(const double* __restrict__ cdf, double r1, double R)
{
  if (cdf[kPtBins] == 0) return 0;
  int lo = 0, hi = kPtBins;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= r1) lo = mid; else hi = mid;
  }
  const double bw = (6. * kHc / R) / kPtBins;
  double x = lo * bw;
  if (r1 > cdf[lo]) x += bw * (r1 - cdf[lo]) / (cdf[lo + 1] - cdf[lo]);
  return x;
}

struct LV { double x, y, z, t; };

__device__ __forceinline__ void lv_boost(LV& v, double bx, double by, double bz)
{
  const double b2 = bx * bx + by * by + bz * bz;
  const double gamma = 1.0 / sqrt(1.0 - b2);
  const double bp = bx * v.x + by * v.y + bz * v.z;
  const double gamma2 = b2 > 0 ? (gamma - 1.0) / b2 : 0.0;
  v.x = v.x + gamma2 * bp * bx + gamma * bx * v.t;
  v.y = v.y + gamma2 * bp * by + gamma * by * v.t;
  v.z = v.z + gamma2 * bp * bz + gamma * bz * v.t;
  v.t = gamma * (v.t + bp);
}
__device__ __forceinline__ LV lv_vect_m(double x, double y, double z, double m)
{
  LV v{x, y, z, 0};
  v.t = sqrt(x * x + y * y + z * z + m * m);
  return v;
}
__device__ __forceinline__ void rotate_uz(double& x, double& y, double& z, double u1, double u2, double u3)
{
  double up = u1 * u1 + u2 * u2;
  if (up) {
    up = sqrt(up);
    const double px = x, py = y, pz = z;
    x = (u1 * u3 * px - u2 * py + u1 * up * pz) / up;
    y = (u2 * u3 * px + u1 * py + u2 * up * pz) / up;
    z = (u3 * u3 * px - px + u3 * up * pz) / up;
  } else if (u3 < 0.) {
    x = -x;
    z = -z;
  }
}
__device__ __forceinline__ double lv_eta(const LV& v)
{
  const double ptot = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  const double cosTheta = ptot == 0.0 ? 1.0 : v.z / ptot;
  if (cosTheta * cosTheta < 1) return -0.5 * log((1.0 - cosTheta) / (1.0 + cosTheta));
  if (v.z == 0) return 0;
  return v.z > 0 ? 10e10 : -10e10;
}

// phase 4: kinematics
__global__ void k_ev_kin(EvParams P, uint64_t seed, uint64_t first, size_t n, const double* __restrict__ y,
                         const double* __restrict__ m, const double* __restrict__ cost, const int* __restrict__ uniq,
                         int n_uniq, const double* __restrict__ cdf, int* __restrict__ npart, int* __restrict__ pdg,
                         int* __restrict__ status, int* __restrict__ mother, double* __restrict__ p4,
                         double* __restrict__ aux, unsigned long long* __restrict__ n_acc)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t cand = first + t;
  const double yPair = y[t], mPair = m[t], cz = cost[t];
  int np = 0;
  LV parts[3];
  int ppdg[3] = {0, 0, 0}, pst[3] = {0, 0, 0}, pmo[3] = {0, 0, 0};
  double pt1 = 0, pt2 = 0;
  bool ok = !(mPair != mPair);
  if (ok) {
    // getPairMomentum, UpcCrossSection.cpp:1053-1074
    LV pPair;
    if (!P.nonzero_gam_pt) {
      pPair = LV{0., 0., mPair * sinh(yPair), mPair * cosh(yPair)};
    } else {
      double a1, a2, u6, u7;
      philox4x32_10(seed, cand, 2, a1, a2);
      philox4x32_10(seed, cand, 3, u6, u7);
      const double k1 = mPair / 2 * exp(yPair), k2 = mPair / 2 * exp(-yPair);
      const double angle1 = 2 * kPi * a1, angle2 = 2 * kPi * a2;
      const int keyv[2] = {energy_key(k1), energy_key(k2)};
      double pts[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        int lo = 0, hi = n_uniq;  // position of the key among the distinct keys
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (uniq[mid] <= keyv[s]) lo = mid; else hi = mid;
        }
        pts[s] = pt_sample(cdf + (size_t)lo * (kPtBins + 1), s ? u7 : u6, P.R);
      }
      pt1 = pts[0]; pt2 = pts[1];
      double s1, c1, s2, c2;
      sincos(angle1, &s1, &c1);
      sincos(angle2, &s2, &c2);
      const double px = pt1 * c1 + pt2 * c2;
      const double py = pt1 * s1 + pt2 * s2;
      const double pt = sqrt(px * px + py * py);
      const double mt = sqrt(mPair * mPair + pt * pt);
      pPair = LV{px, py, mt * sinh(yPair), mt * cosh(yPair)};
    }
    double u8, u9;
    philox4x32_10(seed, cand, 4, u8, u9);
    if (P.is_pair) {  // UpcGenerator.cpp:769-776, pairProduction :388-423
      const double mag2 = pPair.t * pPair.t - (pPair.x * pPair.x + pPair.y * pPair.y + pPair.z * pPair.z);
      const double pMag = sqrt(mag2 / 4 - P.m_part * P.m_part);
      const double theta = acos(cz);
      const double phi = 2. * kPi * u8;
      const double amag = fabs(pMag);
      double st, ct, sp, cp;
      sincos(theta, &st, &ct);
      sincos(phi, &sp, &cp);
      const double vx = amag * st * cp, vy = amag * st * sp, vz = amag * ct;
      parts[0] = lv_vect_m(vx, vy, vz, P.m_part);
      parts[1] = lv_vect_m(-vx, -vy, -vz, P.m_part);
      const double bx = pPair.x / pPair.t, by = pPair.y / pPair.t, bz = pPair.z / pPair.t;
      lv_boost(parts[0], bx, by, bz);
      lv_boost(parts[1], bx, by, bz);
      int sign1 = 1, sign2 = 1;
      if (P.is_charged) { sign1 = (-1. + 2. * u9) > 0 ? 1 : -1; sign2 = -sign1; }
      ppdg[0] = sign1 * P.part_pdg; ppdg[1] = sign2 * P.part_pdg;
      pst[0] = pst[1] = 23;
      np = 2;
    }
    if (P.is_single) {  // singleProduction :474-485
      parts[0] = pPair; ppdg[0] = P.part_pdg; pst[0] = 23; pmo[0] = 0;
      np = 1;
    }
    // checkKinCuts :563-587
    for (int i = 0; i < np; i++) {
      if (P.do_pt_cut && sqrt(parts[i].x * parts[i].x + parts[i].y * parts[i].y) < P.pt_min) { ok = false; break; }
      if (P.do_eta_cut) {
        const double eta = lv_eta(parts[i]);
        if (eta < P.eta_min || eta > P.eta_max) { ok = false; break; }
      }
    }
    // twoPartDecayUniform(id = 1, mass 0) :526-561
    if (ok && P.decay_pdg != 0 && np >= 1) {
      double u10, u11;
      philox4x32_10(seed, cand, 5, u10, u11);
      const LV part = parts[0];
      const double mm2 = part.t * part.t - (part.x * part.x + part.y * part.y + part.z * part.z);
      const double mag = mm2 < 0 ? -sqrt(-mm2) : sqrt(mm2);
      const double ePhot1 = mag / 2.;
      const double pPhot1 = sqrt(ePhot1 * ePhot1);
      const double phi1 = 2. * kPi * u10;
      const double cost1 = -1. + 2. * u11;
      const double theta1 = acos(cost1);
      double st, ct, sp, cp;
      sincos(theta1, &st, &ct);
      sincos(phi1, &sp, &cp);
      const double vx = pPhot1 * st * cp, vy = pPhot1 * st * sp, vz = pPhot1 * ct;
      LV d0 = lv_vect_m(-vx, -vy, -vz, 0.), d1 = lv_vect_m(vx, vy, vz, 0.);
      const double bx = part.x / part.t, by = part.y / part.t, bz = part.z / part.t;
      const double pm = sqrt(part.x * part.x + part.y * part.y + part.z * part.z);
      double ux = part.x, uy = part.y, uz = part.z;
      if (pm > 0) { ux /= pm; uy /= pm; uz /= pm; }
      rotate_uz(d0.x, d0.y, d0.z, ux, uy, uz);
      rotate_uz(d1.x, d1.y, d1.z, ux, uy, uz);
      lv_boost(d0, bx, by, bz);
      lv_boost(d1, bx, by, bz);
      parts[np] = d0; parts[np + 1] = d1;
      ppdg[np] = ppdg[np + 1] = P.decay_pdg;
      pst[np] = pst[np + 1] = 33;
      pmo[np] = pmo[np + 1] = 1;
      np += 2;
    }
  }
  if (!ok) np = 0;
  npart[t] = np;
  for (int i = 0; i < UPCGPU_MAX_PART; i++) {
    const size_t o = t * UPCGPU_MAX_PART + i;
    const bool v = i < np && i < 3;
    pdg[o] = v ? ppdg[i] : 0;
    status[o] = v ? pst[i] : 0;
    mother[o] = v ? pmo[i] : 0;
    p4[o * 4 + 0] = v ? parts[i].x : 0.;
    p4[o * 4 + 1] = v ? parts[i].y : 0.;
    p4[o * 4 + 2] = v ? parts[i].z : 0.;
    p4[o * 4 + 3] = v ? parts[i].t : 0.;
  }
  if (aux) {
    aux[t * 5 + 0] = yPair; aux[t * 5 + 1] = mPair; aux[t * 5 + 2] = cz; aux[t * 5 + 3] = pt1; aux[t * 5 + 4] = pt2;
  }
  if (np > 0) atomicAdd(n_acc, 1ull);
}

static int ensure_scratch(upcgpu_ctx* c, size_t n, size_t n_keys)
{
  EvScratch* s = (EvScratch*)c->ev;
  if (!s) { s = new EvScratch(); c->ev = s; }
  if (n > s->cap) {
    cudaFree(s->y); cudaFree(s->m); cudaFree(s->cost); cudaFree(s->keys); cudaFree(s->keys_sorted); cudaFree(s->uniq);
    cudaFree(s->npart); cudaFree(s->pdg); cudaFree(s->status); cudaFree(s->mother); cudaFree(s->p4); cudaFree(s->aux);
    cudaFree(s->cub_tmp);
    UPC_CUDA(c, cudaMalloc(&s->y, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->m, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->cost, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->keys, 2 * n * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->keys_sorted, 2 * n * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->uniq, 2 * n * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->npart, n * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->pdg, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->status, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->mother, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->p4, n * UPCGPU_MAX_PART * 4 * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->aux, n * 5 * sizeof(double)));
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b1, (int*)nullptr, (int*)nullptr, (int)(2 * n), 0, 32, c->stream);
    cub::DeviceSelect::Unique(nullptr, b2, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)(2 * n), c->stream);
    s->cub_bytes = std::max(b1, b2);
    UPC_CUDA(c, cudaMalloc(&s->cub_tmp, s->cub_bytes + 16));
    s->cap = n;
  }
  if (!s->n_uniq) {
    UPC_CUDA(c, cudaMalloc(&s->n_uniq, sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->n_acc, sizeof(unsigned long long)));
    UPC_CUDA(c, cudaMalloc(&s->err, sizeof(int)));
  }
  if (n_keys > s->cap_keys) {
    cudaFree(s->cdf);
    s->cdf = nullptr;
    UPC_CUDA(c, cudaMalloc(&s->cdf, n_keys * (size_t)(kPtBins + 1) * sizeof(double)));
    s->cap_keys = n_keys;
  }
  return UPCGPU_OK;
}

static EvParams make_evp(const upcgpu_ctx* c)
{
  const upcgpu_params& p = c->p;
  EvParams P;
  P.ny = p.ny; P.nm = p.nm; P.nz = p.nz;
  P.pol = p.use_pol && c->ratio && c->sumz_ps; P.ignore_csz = p.ignore_csz; P.nonzero_gam_pt = p.nonzero_gam_pt;
  P.is_pair = p.is_pair; P.is_single = p.is_single; P.is_charged = p.is_charged; P.part_pdg = p.part_pdg;
  P.decay_pdg = p.decay_uniform_pdg;
  P.do_pt_cut = p.do_pt_cut; P.do_eta_cut = p.do_eta_cut; P.pt_min = p.pt_min; P.eta_min = p.eta_min; P.eta_max = p.eta_max;
  P.m_part = p.m_part; P.gtot = p.gtot; P.R = p.R;
  P.sum2d = c->sum2d; P.sumz = c->sumz; P.sumz_ps = c->sumz_ps; P.ratio = c->ratio;
  P.ye = c->edges_y; P.me = c->edges_m; P.ze = c->edges_z;
  return P;
}

int generate(upcgpu_ctx* c, uint64_t seed, uint64_t first, size_t n, int* npart, int* pdg, int* status, int* mother,
             double* p4, double* aux, uint64_t* n_acc_out, bool device_only)
{
  if (!c->sampler_ready) { c->err = "generate: samplers not built"; return UPCGPU_EINVAL; }
  if (c->p.nonzero_gam_pt && !c->tables_ready) { c->err = "generate: form-factor table not prepared"; return UPCGPU_EINVAL; }
  if (!c->p.ignore_csz && !c->sumz) { c->err = "generate: z samplers missing"; return UPCGPU_EINVAL; }
  cudaStream_t st = c->stream;
  const size_t kChunk = (size_t)1 << 18;  // bounds the pT-table scratch: <= 2*kChunk keys * 40 KB = 21 GB worst case
  const EvParams P = make_evp(c);
  uint64_t total_acc = 0;
  for (size_t off = 0; off < n; off += kChunk) {
    const size_t cn = std::min(kChunk, n - off);
    int rc = ensure_scratch(c, std::min(kChunk, n), 0);
    if (rc) return rc;
    EvScratch* s = (EvScratch*)c->ev;
    UPC_CUDA(c, cudaMemsetAsync(s->n_acc, 0, sizeof(unsigned long long), st));
    UPC_CUDA(c, cudaMemsetAsync(s->err, 0, sizeof(int), st));
    const unsigned g = (unsigned)((cn + 127) / 128);
    UPC_K(c), k_ev_sample<<<g, 128, 0, st>>>(P, seed, first + off, cn, s->y, s->m, s->cost, P.nonzero_gam_pt ? s->keys : nullptr,
                                   s->err);
    int n_uniq = 0;
    if (P.nonzero_gam_pt) {
      size_t tb = s->cub_bytes;
      cub::DeviceRadixSort::SortKeys(s->cub_tmp, tb, s->keys, s->keys_sorted, (int)(2 * cn), 0, 32, st);
      tb = s->cub_bytes;
      cub::DeviceSelect::Unique(s->cub_tmp, tb, s->keys_sorted, s->uniq, s->n_uniq, (int)(2 * cn), st);
      UPC_CUDA(c, cudaMemcpyAsync(&n_uniq, s->n_uniq, sizeof(int), cudaMemcpyDeviceToHost, st));
      UPC_CUDA(c, cudaStreamSynchronize(st));
      rc = ensure_scratch(c, 0, (size_t)n_uniq);
      if (rc) return rc;
      s = (EvScratch*)c->ev;
      UPC_K(c), k_pt_tables<<<n_uniq, 256, 0, st>>>(n_uniq, s->uniq, nullptr, c->ff_seg, c->p.gtot, c->p.R, s->cdf);
    }
    UPC_K(c), k_ev_kin<<<g, 128, 0, st>>>(P, seed, first + off, cn, s->y, s->m, s->cost, s->uniq, n_uniq, s->cdf, s->npart, s->pdg,
                                s->status, s->mother, s->p4, s->aux, s->n_acc);
    unsigned long long acc = 0;
    int herr = 0;
    UPC_CUDA(c, cudaMemcpyAsync(&acc, s->n_acc, sizeof(acc), cudaMemcpyDeviceToHost, st));
    UPC_CUDA(c, cudaMemcpyAsync(&herr, s->err, sizeof(herr), cudaMemcpyDeviceToHost, st));
    if (!device_only) {
      if (npart) UPC_CUDA(c, cudaMemcpyAsync(npart + off, s->npart, cn * sizeof(int), cudaMemcpyDeviceToHost, st));
      const size_t o4 = off * UPCGPU_MAX_PART;
      if (pdg) UPC_CUDA(c, cudaMemcpyAsync(pdg + o4, s->pdg, cn * UPCGPU_MAX_PART * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (status) UPC_CUDA(c, cudaMemcpyAsync(status + o4, s->status, cn * UPCGPU_MAX_PART * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (mother) UPC_CUDA(c, cudaMemcpyAsync(mother + o4, s->mother, cn * UPCGPU_MAX_PART * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (p4) UPC_CUDA(c, cudaMemcpyAsync(p4 + o4 * 4, s->p4, cn * UPCGPU_MAX_PART * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (aux) UPC_CUDA(c, cudaMemcpyAsync(aux + off * 5, s->aux, cn * 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    UPC_CUDA(c, cudaStreamSynchronize(st));
    UPC_CUDA(c, cudaGetLastError());
    if (herr) {
      c->err = "generate: " + std::to_string(herr) + " uniforms fell outside the cumulative pdf (GSL: cannot find r1)";
      return UPCGPU_ERANGE;
    }
    total_acc += acc;
  }
  if (n_acc_out) *n_acc_out = total_acc;
  return UPCGPU_OK;
}

int photon_pt_cdf(upcgpu_ctx* c, double e, double* cdf)
{
  if (!c->tables_ready) { c->err = "photon_pt_cdf: tables not prepared"; return UPCGPU_EINVAL; }
  double *de = nullptr, *dc = nullptr;
  UPC_CUDA(c, cudaMalloc(&de, sizeof(double)));
  UPC_CUDA(c, cudaMalloc(&dc, (kPtBins + 1) * sizeof(double)));
  UPC_CUDA(c, cudaMemcpy(de, &e, sizeof(double), cudaMemcpyHostToDevice));
  UPC_K(c), k_pt_tables<<<1, 256, 0, c->stream>>>(1, nullptr, de, c->ff_seg, c->p.gtot, c->p.R, dc);
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  UPC_CUDA(c, cudaGetLastError());
  UPC_CUDA(c, cudaMemcpy(cdf, dc, (kPtBins + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(de); cudaFree(dc);
  return UPCGPU_OK;
}

}  // namespace upc
