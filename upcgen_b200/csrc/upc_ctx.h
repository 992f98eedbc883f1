// upc_ctx.h -- host-side context behind the C-ABI handle (include/upcgpu.h).
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <map>
#include <string>
#include <vector>

#include "../../include/upcgpu.h"
#include "upc_math.cuh"

namespace upc {

struct Group;  // upc_group.cu: the member contexts of a multi-GPU handle
constexpr int kMaxPeers = 16;

// device-resident lookup tables handed to the kernels by value
struct DevTables {
  // G_AA: 200 knots on [0,20]; seg[i], i<199, plus seg[199] = {1,0,0,0} for b >= 20
  const SplineSeg* gaa_seg;
  double gaa_inv_db;  // 199/20
  double gaa_db;
  int gaa_i0;         // first segment of the live window kept in shared memory by the cell kernel (segments below are cut)
  double b_in2;       // G_AA's spline is <= 1e-30 in magnitude on [0, sqrt(b_in2)]: (b1, b2) pairs that stay below contribute nothing
  // breakup: knots b_i = 1e-6 + db*i, i < nbk (covers [0, 20.2]); seg[nbk-1] = {P20,0,0,0}
  const SplineSeg* bk_seg;
  int bk_n;      // number of real segments (index clamp = bk_n)
  double p20;
  int use_breakup;
  // form factor: 1e6 knots on [1e-9, 2)
  const SplineSeg* ff_seg;
  double ff_last;  // spline value at Q2max - dQ2 (the clamp of fluxFormIntegrand)
};

struct upcgpu_ctx_impl {
  upcgpu_params p;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux[2] = {nullptr, nullptr};   // side streams: the three lookup tables are built concurrently
  cudaEvent_t aux_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  cudaDeviceProp prop;
  bool tables_ready = false;
  upcgpu_table_info info{};
  upcgpu_fill_stats stats{};
  long long launches = 0;  // kernels of this library launched so far (UPC_K)

  // raw spline tables (x, y, c) kept for the get_table test hooks
  double *gaa_x = nullptr, *gaa_y = nullptr, *gaa_c = nullptr, *ta_y = nullptr, *ta_c = nullptr;
  double *ff_y = nullptr, *ff_c = nullptr;
  double *bk_y = nullptr, *bk_c = nullptr;
  int bk_nknots = 0;
  void* bk_table = nullptr;  // photo-nuclear energy table of calcBreakupProb (device)
  double bk_table_g1 = -1.;  // beam gamma the energy table was built for (the reference builds it once per process:
                             // function-local statics, src/UpcCrossSection.cpp:853-969)
  SplineSeg *gaa_seg = nullptr, *ff_seg = nullptr, *bk_seg = nullptr;
  double* d_scal = nullptr;  // small device scratch for scalars
  double* h_scal = nullptr;  // its pinned host mirror (one asynchronous read-back per table stage)
  // The scalars are a pure function of the (immutable) parameter block: the first table stage waits for them, later
  // ones take them from this cache and leave the comparison with what the device produced to the next host wait
  bool scal_cached = false;
  double scal_cache[5] = {0, 0, 0, 0, 0};
  bool bk_deferred = false;     // the breakup chain of the queued table stage runs beside the flux stage: cells wait for aux_ev[3]
  bool tables_pending = false;  // a table stage is queued whose scalars (and stage time) have not been collected
  cudaEvent_t tab_ev[2] = {nullptr, nullptr};
  DevTables tab{};

  // luminosity tables, full [nm][ny], index 0 unpol, 1 scalar, 2 pseudoscalar
  double* lumi[3] = {nullptr, nullptr, nullptr};
  double* shard[3] = {nullptr, nullptr, nullptr};
  double* gather[3] = {nullptr, nullptr, nullptr};
  size_t shard_rows = 0;
  int shard_n = 0, gather_n = 0;
  bool lumi_ready = false;

  // fold / samplers
  double *cs = nullptr, *ratio = nullptr;
  double *sum2d = nullptr, *sumz = nullptr, *sumz_ps = nullptr;
  double *edges_y = nullptr, *edges_m = nullptr, *edges_z = nullptr;
  bool spec_attr_set = false;          // dynamic shared memory opt-in of k_seq_spec on this context's device
  struct SpecStats* spec_stats = nullptr;  // device counters of the speculate-and-verify recurrences (may stay null)
  double *samp_term = nullptr, *samp_mean = nullptr, *samp_dz = nullptr;  // sampler_build's scratch (kept: no malloc/free per build)
  bool fold_ready = false, sampler_ready = false;
  double* fold_ws = nullptr;  // scratch of fold_sigma: 3 x nm sigma values + the block sums of the total (kept: a
                              // cudaMalloc / cudaFree pair per call costs more than the fold itself)

  // event stage scratch (grown on demand)
  void* ev = nullptr;
  // flux-row scratch of the lumi fill (allocated once, reused by every fill)
  void* slab = nullptr;
  int slab_max_m = 0;
  unsigned* cell_counter = nullptr;  // work counter of the persistent cell kernel
  // A fill is queued on the stream and collected later (finish_fill): number of slabs whose reports are pending
  int fill_pending = 0;
  cudaEvent_t fill_ev[2] = {nullptr, nullptr};
  // integral count of a (shard, nshards, first local row, rows) slab: a pure function of the parameter block, read back
  // from the device the first time and reused to size the head kernel's grid afterwards
  std::map<std::array<int, 4>, long long> n_items_cache;
  // several GPUs behind one handle (upcgpu_create_multi): set on every member; the leader (rank 0) owns the group
  Group* group = nullptr;
  int group_rank = 0;
  // peer-store exchange: full lumi tables of every member, written by this member's cell kernel (set per fill)
  double* peer_lumi[kMaxPeers][3] = {};
  int n_peers = 0;
  // ... the same between PROCESSES (one rank per GPU): the other ranks' tables mapped through CUDA IPC
  // (upcgpu_lumi_ipc_export / _import); ipc_n = number of ranks, ipc_rank = this one
  double* ipc_lumi[kMaxPeers][3] = {};
  int ipc_n = 0, ipc_rank = -1;
  bool func_attrs_set = false;   // cudaFuncSetAttribute done on this context's device
  bool ev_attr_set = false;      // ... for the event kernels
  long long test_head_pool = 0;  // UPCGPU_TEST_HEAD_POOL: slots of the hand-over state pool (tests of the fallback path)
};

}  // namespace upc

struct upcgpu_ctx : upc::upcgpu_ctx_impl {};
