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
// (Q9).  Here the photons of a batch (two per candidate) are sorted by their MeV key on the device
// (radix sort of (key, photon) pairs), and one CTA per DISTINCT key builds that key's cumulative pdf
// in SHARED memory and serves all photons of the key from it -- the pdf of a key is evaluated at the
// key's centre (key + 0.5) MeV.  Nothing but the photon pT (8 B per photon) reaches HBM: writing the
// 5001-entry tables out and searching them from another kernel (round 1) moved 40 KB per key, ~300x
// the bytes of the events themselves.  Batches are large (8 M candidates) because the cost of the
// stage is the number of distinct keys per batch x 5000 form-factor evaluations.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>

#include "upc_ctx.h"
#include "upc_internal.h"
#include "upc_sampler.cuh"

namespace upc {

constexpr int kPtBins = 5000;
constexpr size_t kEvChunk = (size_t)1 << 23;  // candidates per batch (~300 B of scratch each: 2.5 GB)

struct EvReport {  // pinned host mirror of the per-batch counters
  unsigned long long n_acc;
  int err;
  int n_uniq;
};

struct EvScratch {
  size_t cap = 0;        // candidates
  double *y = nullptr, *m = nullptr, *cost = nullptr;
  unsigned *keys = nullptr, *keys_sorted = nullptr, *pid = nullptr, *pid_sorted = nullptr;  // [2 cap] photons
  unsigned* seg_off = nullptr;  // [2 cap]: first sorted photon of every distinct key
  int* n_uniq = nullptr;
  unsigned* next_tile = nullptr;  // work counter of k_pt_serve
  double* pt = nullptr;         // [2 cap] photon pT, indexed by photon id 2 t + side
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  int *npart = nullptr, *pdg = nullptr, *status = nullptr, *mother = nullptr;
  double *p4 = nullptr, *aux = nullptr;
  unsigned long long* n_acc = nullptr;
  int* err = nullptr;
  EvReport* report = nullptr;
  cudaStream_t copy_stream = nullptr;          // device -> host copies of a chunk's events, beside the next chunk's kernels
  cudaEvent_t kin_done = nullptr, copy_done = nullptr;
  void release()
  {
    cudaFree(y); cudaFree(m); cudaFree(cost); cudaFree(keys); cudaFree(keys_sorted); cudaFree(pid); cudaFree(pid_sorted);
    cudaFree(seg_off); cudaFree(pt); cudaFree(cub_tmp); cudaFree(npart); cudaFree(pdg); cudaFree(status); cudaFree(mother);
    cudaFree(p4); cudaFree(aux);
    y = m = cost = pt = p4 = aux = nullptr;
    keys = keys_sorted = pid = pid_sorted = seg_off = nullptr;
    cub_tmp = nullptr;
    npart = pdg = status = mother = nullptr;
    cap = 0;
    cub_bytes = 0;
  }
};

void free_event_scratch(upcgpu_ctx* c)
{
  EvScratch* s = (EvScratch*)c->ev;
  if (!s) return;
  s->release();
  cudaFree(s->n_uniq); cudaFree(s->n_acc); cudaFree(s->err); cudaFree(s->next_tile);
  if (s->report) cudaFreeHost(s->report);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  if (s->kin_done) cudaEventDestroy(s->kin_done);
  if (s->copy_done) cudaEventDestroy(s->copy_done);
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
                            double* __restrict__ m, double* __restrict__ cost, unsigned* __restrict__ keys,
                            unsigned* __restrict__ pid, int* __restrict__ err)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t cand = first + t;
  double r1, r2, u2, u3;
  philox4x32_10(seed, cand, 0, r1, r2);
  philox4x32_10(seed, cand, 1, u2, u3);
  long long k; double yy, mm;
  sample_ym_dev(P.sum2d, P.ny, P.nm, P.ye, P.me, r1, r2, k, yy, mm);  // UpcGenerator.cpp:737
  if (keys) { pid[2 * t] = (unsigned)(2 * t); pid[2 * t + 1] = (unsigned)(2 * t + 1); }
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
    keys[2 * t] = (unsigned)energy_key(mm / 2 * exp(yy));       // UpcCrossSection.cpp:1060-1061
    keys[2 * t + 1] = (unsigned)energy_key(mm / 2 * exp(-yy));
  }
}

// first sorted photon of every distinct key
struct KeyHead {
  const unsigned* ks;
  __device__ __forceinline__ bool operator()(unsigned i) const { return i == 0 || ks[i] != ks[i - 1]; }
};

// phase 3: cumulative pT pdf of one photon-energy key (getPhotonPt, :1021-1038, and
// TH1::ComputeIntegral): cdf[0] = 0, cdf[i] = sum_{j<=i} p_j / total
__device__ __forceinline__ double ff_lookup(const SplineSeg* __restrict__ ff, double t)
{
  // the reference calls gsl_spline_eval without a clamp (GSL domain error beyond the last
  // knot); semantics here and in the oracle: clamp to the last knot
  const double tmax = kQ2min + (kNQ2 - 1) * kDQ2;
  t = fmin(t, tmax);
  int idx = (int)((t - kQ2min) * (1. / kDQ2));
  idx = max(0, min(idx, kNQ2 - 2));
  return seg_eval(ld_seg(ff + idx), t - fma((double)idx, kDQ2, kQ2min));
}

constexpr int kPtThreads = 256;
constexpr int kPtPerThread = (kPtBins + kPtThreads - 1) / kPtThreads;  // 20 bins per thread
constexpr int kPtRunPad = kPtPerThread + 1;  // runs padded to 21 doubles in shared memory (2-way conflicts)

struct PtShared {
  typename cub::BlockScan<double, kPtThreads>::TempStorage scan;
  double sp[kPtThreads * kPtRunPad];  // after pt_build: sp[(i / 20) * 21 + i % 20] = cdf[i + 1]
};

// cdf[i], i = 0 .. 5000, of the table pt_build left in shared memory
__device__ __forceinline__ double pt_cdf_at(const double* sp, int i)
{
  if (i == 0) return 0.;
  const int b = i - 1;
  return sp[(b / kPtPerThread) * kPtRunPad + b % kPtPerThread];
}

// The pdf is evaluated with bin = base + thread (neighbouring threads gather neighbouring form-factor segments);
// the prefix sum goes through shared memory: each thread sums a RUN of 20 consecutive bins serially and ONE block
// scan ranks the runs; the thread then normalises its run in place.  Ends with a barrier.
__device__ __forceinline__ void pt_build(PtShared& S, double e, const SplineSeg* __restrict__ ff, double gtot, double R)
{
  typedef cub::BlockScan<double, kPtThreads> Scan;
  const int tid = threadIdx.x;
  const double ereds = (e * e) / (gtot * gtot);
  const double pi2x4 = 4 * kPi * kPi;
#pragma unroll 4
  for (int it = 0; it < kPtPerThread; ++it) {
    const int b0 = it * kPtThreads + tid;           // 0-based bin
    double prob = 0;
    if (b0 < kPtBins) {
      const double pt = 6. * kHc / R / kPtBins * (b0 + 1);  // upper bin edge, :1033
      const double arg = pt * pt + ereds;
      const double f = ff_lookup(ff, arg);
      // (the division of :1036 as a multiplication by a Newton-refined reciprocal: ~1e-16 relative on a pdf whose
      // draws are compared statistically; the table is the FP64-bound part of the event stage)
      prob = (f * f) * pt * pt * pt * rcp_fast(pi2x4 * arg * arg);
    }
    S.sp[(b0 / kPtPerThread) * kPtRunPad + b0 % kPtPerThread] = prob;
  }
  __syncthreads();
  double acc = 0;
  double* run = S.sp + tid * kPtRunPad;
#pragma unroll
  for (int j = 0; j < kPtPerThread; ++j) {
    acc += run[j];
    run[j] = acc;
  }
  double before, tot;
  Scan(S.scan).ExclusiveSum(acc, before, tot);
#pragma unroll
  for (int j = 0; j < kPtPerThread; ++j) {
    const double v = before + run[j];
    run[j] = tot != 0 ? v / tot : v;           // TH1::ComputeIntegral: integral[i] /= integral[n]
  }
  __syncthreads();
}

// TH1::GetRandom on the table in shared memory
__device__ __forceinline__ double pt_sample(const double* sp, double r1, double R)
{
  if (pt_cdf_at(sp, kPtBins) == 0) return 0;
  int lo = 0, hi = kPtBins;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pt_cdf_at(sp, mid) <= r1) lo = mid; else hi = mid;
  }
  const double bw = (6. * kHc / R) / kPtBins;
  double x = lo * bw;
  const double c0 = pt_cdf_at(sp, lo);
  if (r1 > c0) x += bw * (r1 - c0) / (pt_cdf_at(sp, lo + 1) - c0);
  return x;
}

// Two kernels serve the photons of a batch from per-key tables in shared memory; the keys are very unequal -- at low
// photon energies one integer-MeV key holds 10^5-10^6 photons of a 4 M-candidate batch, at high energies every photon
// has a key of its own -- and each regime gets the work decomposition that suits it:
//   k_pt_serve_keys   one table per DISTINCT key, keys dealt to a persistent grid round-robin (neighbouring CTAs work
//                     on neighbouring keys, i.e. on the same part of the form-factor table at the same time); the CTA
//                     serves the first kPtHead photons of the key;
//   k_pt_serve_tiles  the photons beyond the first kPtHead of their key, in tiles of kPtThreads sorted photons handed
//                     out by an atomic counter: a heavy key is shared by many CTAs (each rebuilds its table: one
//                     extra table per 256 photons), a tile without such photons costs one look at its run lengths.
// Photon id = 2 t + side; its uniform is slot `side` of Philox block 3 of candidate first + t (the slot map above),
// recomputed here rather than stored.
constexpr unsigned kPtHead = 2048;

__device__ __forceinline__ void pt_serve_one(const PtShared& S, unsigned q, const unsigned* __restrict__ pid_sorted, uint64_t seed,
                                             uint64_t first, double R, double* __restrict__ pt_out)
{
  const unsigned id = pid_sorted[q];
  double u6, u7;
  philox4x32_10(seed, first + (id >> 1), 3, u6, u7);
  pt_out[id] = pt_sample(S.sp, (id & 1) ? u7 : u6, R);
}

__global__ void __launch_bounds__(kPtThreads) k_pt_serve_keys(const int* __restrict__ n_uniq_p, const unsigned* __restrict__ seg_off,
                                                              unsigned n_photons, const unsigned* __restrict__ keys_sorted,
                                                              const unsigned* __restrict__ pid_sorted, uint64_t seed,
                                                              uint64_t first, const SplineSeg* __restrict__ ff, double gtot,
                                                              double R, double* __restrict__ pt_out)
{
  extern __shared__ __align__(16) unsigned char pt_smem[];
  PtShared& S = *reinterpret_cast<PtShared*>(pt_smem);
  const int n_uniq = *n_uniq_p;
  for (int k = blockIdx.x; k < n_uniq; k += gridDim.x) {
    const unsigned q0 = seg_off[k], q1 = (k + 1 < n_uniq) ? seg_off[k + 1] : n_photons;
    pt_build(S, (keys_sorted[q0] + 0.5) * 1e-3, ff, gtot, R);
    const unsigned qe = min(q1, q0 + kPtHead);
    for (unsigned q = q0 + threadIdx.x; q < qe; q += kPtThreads) pt_serve_one(S, q, pid_sorted, seed, first, R, pt_out);
    __syncthreads();  // the table is rebuilt for the next key
  }
}

__global__ void __launch_bounds__(kPtThreads) k_pt_serve_tiles(const int* __restrict__ n_uniq_p, const unsigned* __restrict__ seg_off,
                                                               unsigned n_photons, const unsigned* __restrict__ keys_sorted,
                                                               const unsigned* __restrict__ pid_sorted, uint64_t seed,
                                                               uint64_t first, const SplineSeg* __restrict__ ff, double gtot,
                                                               double R, double* __restrict__ pt_out, unsigned* __restrict__ next_tile)
{
  extern __shared__ __align__(16) unsigned char pt_smem[];
  PtShared& S = *reinterpret_cast<PtShared*>(pt_smem);
  __shared__ unsigned s_tile;
  const int n_uniq = *n_uniq_p;
  const unsigned n_tiles = (n_photons + kPtThreads - 1) / kPtThreads;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(next_tile, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= n_tiles) break;
    const unsigned t0 = tile * kPtThreads, t1 = min(t0 + kPtThreads, n_photons);
    // the run that holds photon t0: the last k with seg_off[k] <= t0
    int lo = 0, hi = n_uniq;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (seg_off[mid] <= t0) lo = mid; else hi = mid;
    }
    // Only the run that holds t0 can have photons of rank >= kPtHead inside this tile beyond its start ... and any
    // later run that starts inside the tile has ranks < kPtThreads <= kPtHead there: at most ONE run needs serving.
    const unsigned q0 = seg_off[lo];
    const unsigned q1 = (lo + 1 < n_uniq) ? seg_off[lo + 1] : n_photons;
    const unsigned from = max(t0, q0 + kPtHead), to = min(q1, t1);
    if (from < to) {  // uniform over the CTA
      pt_build(S, (keys_sorted[q0] + 0.5) * 1e-3, ff, gtot, R);
      const unsigned q = from + threadIdx.x;
      if (q < to) pt_serve_one(S, q, pid_sorted, seed, first, R, pt_out);
    }
    __syncthreads();  // s_tile and the table are reused
  }
}

// test hook: the table of one photon energy, written out
__global__ void __launch_bounds__(kPtThreads) k_pt_table_one(const double* __restrict__ e_direct, const SplineSeg* __restrict__ ff,
                                                             double gtot, double R, double* __restrict__ cdf)
{
  extern __shared__ __align__(16) unsigned char pt_smem[];
  PtShared& S = *reinterpret_cast<PtShared*>(pt_smem);
  pt_build(S, e_direct[0], ff, gtot, R);
  for (int i = threadIdx.x; i <= kPtBins; i += kPtThreads) cdf[i] = pt_cdf_at(S.sp, i);
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
                         const double* __restrict__ m, const double* __restrict__ cost, const double* __restrict__ pt_ph,
                         int* __restrict__ npart, int* __restrict__ pdg, int* __restrict__ status,
                         int* __restrict__ mother, double* __restrict__ p4, double* __restrict__ aux,
                         unsigned long long* __restrict__ n_acc, int part_stride)
{
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t cand = first + t;
  const double yPair = y[t], mPair = m[t], cz = cost[t];
  int np = 0;
  LV parts[UPCGPU_MAX_PART];
  int ppdg[UPCGPU_MAX_PART] = {}, pst[UPCGPU_MAX_PART] = {}, pmo[UPCGPU_MAX_PART] = {};
  double pt1 = 0, pt2 = 0;
  bool ok = !(mPair != mPair);
  if (ok) {
    // getPairMomentum, UpcCrossSection.cpp:1053-1074
    LV pPair;
    if (!P.nonzero_gam_pt) {
      pPair = LV{0., 0., mPair * sinh(yPair), mPair * cosh(yPair)};
    } else {
      double a1, a2;
      philox4x32_10(seed, cand, 2, a1, a2);
      const double angle1 = 2 * kPi * a1, angle2 = 2 * kPi * a2;
      pt1 = pt_ph[2 * t];        // getPhotonPt(k1), getPhotonPt(k2): drawn by k_pt_serve from the keys' tables
      pt2 = pt_ph[2 * t + 1];
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
    // twoPartDecayUniform(id, mass 0) :526-561 of the single particle (ALP, :806-808) or of both particles of a pair
    // (pi0 pi0, :799-803: id = 1, then id = 2); decay d takes its two uniforms from Philox block 5 + d
    const int n_dec = (ok && P.decay_pdg != 0) ? (P.is_pair ? 2 : 1) : 0;
    for (int dcy = 0; dcy < n_dec; ++dcy) {
      double u10, u11;
      philox4x32_10(seed, cand, 5 + dcy, u10, u11);
      const LV part = parts[dcy];
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
      pmo[np] = pmo[np + 1] = dcy + 1;
      np += 2;
    }
  }
  if (!ok) np = 0;
  npart[t] = np;
  for (int i = 0; i < part_stride; i++) {  // part_stride >= the particles an event of this process can have
    const size_t o = t * part_stride + i;
    const bool v = i < np;
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

// bits of the largest photon-energy key the grid can produce (radix-sort passes)
static int key_bits(const upcgpu_params& p)
{
  const double kmax = p.mmax / 2. * exp(fmax(fabs(p.ymin), fabs(p.ymax))) * 1e3 + 2.;
  int bits = 1;
  while (bits < 32 && (double)(1ull << bits) <= kmax) ++bits;
  return bits;
}

static int ensure_scratch(upcgpu_ctx* c, size_t n)
{
  EvScratch* s = (EvScratch*)c->ev;
  if (!s) { s = new EvScratch(); c->ev = s; }
  if (!s->n_uniq) {
    UPC_CUDA(c, cudaMalloc(&s->n_uniq, sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->next_tile, sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->n_acc, sizeof(unsigned long long)));
    UPC_CUDA(c, cudaMalloc(&s->err, sizeof(int)));
    UPC_CUDA(c, cudaMallocHost(&s->report, sizeof(EvReport)));
  }
  if (n > s->cap) {
    s->release();  // frees and nulls: a failed allocation below leaves nothing dangling
    UPC_CUDA(c, cudaMalloc(&s->y, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->m, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->cost, n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->keys, 2 * n * sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->keys_sorted, 2 * n * sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->pid, 2 * n * sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->pid_sorted, 2 * n * sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->seg_off, 2 * n * sizeof(unsigned)));
    UPC_CUDA(c, cudaMalloc(&s->pt, 2 * n * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->npart, n * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->pdg, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->status, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->mother, n * UPCGPU_MAX_PART * sizeof(int)));
    UPC_CUDA(c, cudaMalloc(&s->p4, n * UPCGPU_MAX_PART * 4 * sizeof(double)));
    UPC_CUDA(c, cudaMalloc(&s->aux, n * 5 * sizeof(double)));
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr, (unsigned*)nullptr,
                                    (int)(2 * n), 0, 32, c->stream);
    cub::CountingInputIterator<unsigned> first(0u);
    cub::DeviceSelect::If(nullptr, b2, first, (unsigned*)nullptr, (int*)nullptr, (int)(2 * n), KeyHead{nullptr}, c->stream);
    s->cub_bytes = std::max(b1, b2);
    UPC_CUDA(c, cudaMalloc(&s->cub_tmp, s->cub_bytes + 16));
    s->cap = n;
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

// particles an accepted event of this context's process has: the pair or the single particle, plus two decay products
// of each when they decay (ALP: 3, pi0 pi0: 6)
int particles_per_event(const upcgpu_ctx* c)
{
  const EvParams P = make_evp(c);
  const int prim = P.is_pair ? 2 : 1;
  return prim + (P.decay_pdg != 0 ? 2 * prim : 0);
}

// Candidates are processed in chunks.  Device-resident runs take chunks of kEvChunk (the photon-pT stage costs one table
// per DISTINCT key of a chunk, so chunks are large).  Runs that deliver to host buffers take chunks of kEvChunkHost and
// pipeline them: the copies of chunk i run on a second stream while the kernels of chunk i + 1 sample, sort and serve;
// only that chunk's kinematics kernel, which overwrites the output arrays, waits for the copies.  One host wait at the end.
constexpr size_t kEvChunkHost = (size_t)1 << 21;

int generate(upcgpu_ctx* c, uint64_t seed, uint64_t first, size_t n, int part_stride, int* npart, int* pdg, int* status, int* mother,
             double* p4, double* aux, uint64_t* n_acc_out, bool device_only)
{
  if (!c->sampler_ready) { c->err = "generate: samplers not built"; return UPCGPU_EINVAL; }
  if (c->p.nonzero_gam_pt && !c->tables_ready) { c->err = "generate: form-factor table not prepared"; return UPCGPU_EINVAL; }
  if (!c->p.ignore_csz && !c->sumz) { c->err = "generate: z samplers missing"; return UPCGPU_EINVAL; }
  cudaStream_t st = c->stream;
  const EvParams P = make_evp(c);
  if (!c->ev_attr_set) {
    cudaFuncSetAttribute(k_pt_serve_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    cudaFuncSetAttribute(k_pt_serve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    cudaFuncSetAttribute(k_pt_table_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    c->ev_attr_set = true;
  }
  const int kbits = key_bits(c->p);
  const size_t chunk = device_only ? kEvChunk : kEvChunkHost;
  int rc = ensure_scratch(c, std::min(chunk, n));
  if (rc) return rc;
  EvScratch* s = (EvScratch*)c->ev;
  if (!device_only && !s->copy_stream) {
    UPC_CUDA(c, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    UPC_CUDA(c, cudaEventCreateWithFlags(&s->kin_done, cudaEventDisableTiming));
    UPC_CUDA(c, cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming));
  }
  cudaStream_t cs = s->copy_stream;
  // the accepted count and the error flag accumulate over the chunks on the device
  UPC_CUDA(c, cudaMemsetAsync(s->n_acc, 0, sizeof(unsigned long long), st));
  UPC_CUDA(c, cudaMemsetAsync(s->err, 0, sizeof(int), st));
  bool copies_in_flight = false;
  // whichever way this function returns, no copy into the caller's buffers is left running
  struct CopyGuard {
    cudaStream_t s;
    const bool& active;
    ~CopyGuard() { if (active && s) cudaStreamSynchronize(s); }
  } copy_guard{cs, copies_in_flight};
  for (size_t off = 0; off < n; off += chunk) {
    const size_t cn = std::min(chunk, n - off);
    const unsigned g = (unsigned)((cn + 127) / 128);
    UPC_K(c), k_ev_sample<<<g, 128, 0, st>>>(P, seed, first + off, cn, s->y, s->m, s->cost, P.nonzero_gam_pt ? s->keys : nullptr,
                                   s->pid, s->err);
    if (P.nonzero_gam_pt) {
      const int n_ph = (int)(2 * cn);
      size_t tb = s->cub_bytes;
      cub::DeviceRadixSort::SortPairs(s->cub_tmp, tb, s->keys, s->keys_sorted, s->pid, s->pid_sorted, n_ph, 0, kbits, st);
      tb = s->cub_bytes;
      cub::CountingInputIterator<unsigned> cnt0(0u);
      cub::DeviceSelect::If(s->cub_tmp, tb, cnt0, s->seg_off, s->n_uniq, n_ph, KeyHead{s->keys_sorted}, st);
      // persistent grids: the number of distinct keys stays on the device
      const int n_tiles = (n_ph + kPtThreads - 1) / kPtThreads;
      const int grid = std::min(n_ph, c->prop.multiProcessorCount * 5);
      UPC_K(c), k_pt_serve_keys<<<grid, kPtThreads, sizeof(PtShared), st>>>(s->n_uniq, s->seg_off, (unsigned)n_ph, s->keys_sorted,
                                                                  s->pid_sorted, seed, first + off, c->ff_seg, c->p.gtot, c->p.R, s->pt);
      UPC_CUDA(c, cudaMemsetAsync(s->next_tile, 0, sizeof(unsigned), st));
      UPC_K(c), k_pt_serve_tiles<<<std::min(n_tiles, grid), kPtThreads, sizeof(PtShared), st>>>(
          s->n_uniq, s->seg_off, (unsigned)n_ph, s->keys_sorted, s->pid_sorted, seed, first + off, c->ff_seg, c->p.gtot, c->p.R, s->pt,
          s->next_tile);
    }
    if (copies_in_flight) UPC_CUDA(c, cudaStreamWaitEvent(st, s->copy_done, 0));  // the previous chunk's events have left
    UPC_K(c), k_ev_kin<<<g, 128, 0, st>>>(P, seed, first + off, cn, s->y, s->m, s->cost, s->pt, s->npart, s->pdg, s->status, s->mother,
                                s->p4, s->aux, s->n_acc, part_stride);
    if (!device_only) {
      UPC_CUDA(c, cudaEventRecord(s->kin_done, st));
      UPC_CUDA(c, cudaStreamWaitEvent(cs, s->kin_done, 0));
      const size_t os = off * (size_t)part_stride, ns = cn * (size_t)part_stride;
      if (npart) UPC_CUDA(c, cudaMemcpyAsync(npart + off, s->npart, cn * sizeof(int), cudaMemcpyDeviceToHost, cs));
      if (pdg) UPC_CUDA(c, cudaMemcpyAsync(pdg + os, s->pdg, ns * sizeof(int), cudaMemcpyDeviceToHost, cs));
      if (status) UPC_CUDA(c, cudaMemcpyAsync(status + os, s->status, ns * sizeof(int), cudaMemcpyDeviceToHost, cs));
      if (mother) UPC_CUDA(c, cudaMemcpyAsync(mother + os, s->mother, ns * sizeof(int), cudaMemcpyDeviceToHost, cs));
      if (p4) UPC_CUDA(c, cudaMemcpyAsync(p4 + os * 4, s->p4, ns * 4 * sizeof(double), cudaMemcpyDeviceToHost, cs));
      if (aux) UPC_CUDA(c, cudaMemcpyAsync(aux + off * 5, s->aux, cn * 5 * sizeof(double), cudaMemcpyDeviceToHost, cs));
      UPC_CUDA(c, cudaEventRecord(s->copy_done, cs));
      copies_in_flight = true;
    }
  }
  UPC_CUDA(c, cudaMemcpyAsync(&s->report->n_acc, s->n_acc, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  UPC_CUDA(c, cudaMemcpyAsync(&s->report->err, s->err, sizeof(int), cudaMemcpyDeviceToHost, st));
  UPC_CUDA(c, cudaStreamSynchronize(st));
  if (copies_in_flight) UPC_CUDA(c, cudaStreamSynchronize(cs));
  UPC_CUDA(c, cudaGetLastError());
  if (s->report->err) {
    c->err = "generate: " + std::to_string(s->report->err) + " uniforms fell outside the cumulative pdf (GSL: cannot find r1)";
    return UPCGPU_ERANGE;
  }
  if (n_acc_out) *n_acc_out = s->report->n_acc;
  return UPCGPU_OK;
}

int photon_pt_cdf(upcgpu_ctx* c, double e, double* cdf)
{
  if (!c->tables_ready) { c->err = "photon_pt_cdf: tables not prepared"; return UPCGPU_EINVAL; }
  if (!c->ev_attr_set) {
    cudaFuncSetAttribute(k_pt_serve_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    cudaFuncSetAttribute(k_pt_serve_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    cudaFuncSetAttribute(k_pt_table_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtShared));
    c->ev_attr_set = true;
  }
  double *de = nullptr, *dc = nullptr;
  UPC_CUDA(c, cudaMalloc(&de, sizeof(double)));
  if (cudaMalloc(&dc, (kPtBins + 1) * sizeof(double)) != cudaSuccess) { cudaFree(de); c->err = "photon_pt_cdf: out of memory"; return UPCGPU_ECUDA; }
  int rc = UPCGPU_OK;
  if (cudaMemcpy(de, &e, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = UPCGPU_ECUDA;
  if (!rc) {
    UPC_K(c), k_pt_table_one<<<1, kPtThreads, sizeof(PtShared), c->stream>>>(de, c->ff_seg, c->p.gtot, c->p.R, dc);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess ||
        cudaMemcpy(cdf, dc, (kPtBins + 1) * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = UPCGPU_ECUDA;
  }
  cudaFree(de); cudaFree(dc);
  if (rc) c->err = "photon_pt_cdf: CUDA error";
  return rc;
}

}  // namespace upc
