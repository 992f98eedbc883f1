// upc_capi.cu -- the extern "C" boundary declared in include/upcgpu.h.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "upc_ctx.h"
#include "upc_internal.h"

using namespace upc;

static thread_local std::string g_create_err;

#define CHECK_CTX(c)                 \
  if (!(c)) return UPCGPU_EINVAL;    \
  cudaSetDevice((c)->device);

// entry points that read results on the host (or free / reuse the fill's scratch) first collect a queued fill
#define CHECK_CTX_SYNC(c)                         \
  CHECK_CTX(c);                                   \
  if ((c)->fill_pending) {                        \
    int _rc = finish_fill(c);                     \
    if (_rc) return _rc;                          \
  }                                               \
  if ((c)->tables_pending) {                      \
    int _rc = finish_tables(c);                   \
    if (_rc) return _rc;                          \
  }

// The context's main stream (and the form-factor side stream) get the highest priority the device offers; the breakup
// table's side stream keeps the default, lowest one: its chain runs beside the flux stage and must not take SM slots
// from it (pending blocks of a higher-priority stream are placed first).
cudaError_t upc::create_stream(cudaStream_t* st, bool high_priority)
{
  int least = 0, greatest = 0;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, high_priority ? greatest : least);
}

extern "C" {

int upcgpu_abi_version(void) { return 1; }

const char* upcgpu_last_error(const upcgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int upcgpu_create(const upcgpu_params* params, int device, upcgpu_ctx** out)
{
  if (!params || !out) { g_create_err = "upcgpu_create: null argument"; return UPCGPU_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("upcgpu_create: no CUDA device (") + cudaGetErrorString(e) +
                   "); this library has no CPU fallback";
    return UPCGPU_ENODEV;
  }
  if (device < 0 || device >= ndev) { g_create_err = "upcgpu_create: bad device index"; return UPCGPU_EINVAL; }
  const upcgpu_params& p = *params;
  if (p.nm < 1 || p.ny < 1 || p.nz < 1 || p.nb1 < 2 || p.nb2 < 2 || p.A < 1 || p.Z < 1 || !(p.R > 0) || !(p.a > 0) ||
      !(p.g1 > 1) || !(p.g2 > 1) || p.breakup_mode < 1 || p.breakup_mode > 4 || !(p.mmax > p.mmin) ||
      !(p.ymax > p.ymin) || !(p.mmin > 0) || !(p.sqrts > 0) || !(p.gtot > 1) || !(p.zmax > p.zmin)) {
    g_create_err = "upcgpu_create: parameter block out of range";
    return UPCGPU_EINVAL;
  }
  upcgpu_ctx* c = new (std::nothrow) upcgpu_ctx();
  if (!c) { g_create_err = "upcgpu_create: out of memory"; return UPCGPU_EINVAL; }
  c->p = p;
  c->device = device;
  if (const char* e = std::getenv("UPCGPU_TEST_HEAD_POOL")) c->test_head_pool = std::atoll(e);
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&c->prop, device) != cudaSuccess ||
      create_stream(&c->stream, /*high_priority=*/true) != cudaSuccess) {
    g_create_err = std::string("upcgpu_create: ") + cudaGetErrorString(cudaGetLastError());
    delete c;
    return UPCGPU_ECUDA;
  }
  *out = c;
  return UPCGPU_OK;
}

int upcgpu_create_multi(const upcgpu_params* params, int n_gpus, const int* devices, upcgpu_ctx** out)
{
  if (!params || !out) { g_create_err = "upcgpu_create_multi: null argument"; return UPCGPU_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_err = "upcgpu_create_multi: no CUDA device; this library has no CPU fallback";
    return UPCGPU_ENODEV;
  }
  if (n_gpus < 1 || n_gpus > kMaxPeers) { g_create_err = "upcgpu_create_multi: n_gpus must be 1.." + std::to_string(kMaxPeers); return UPCGPU_EINVAL; }
  std::vector<int> devs(n_gpus);
  for (int r = 0; r < n_gpus; ++r) {
    devs[r] = devices ? devices[r] : r;
    if (devs[r] < 0 || devs[r] >= ndev) {
      g_create_err = "upcgpu_create_multi: device " + std::to_string(devs[r]) + " requested, " + std::to_string(ndev) + " present";
      return UPCGPU_EINVAL;
    }
    for (int q = 0; q < r; ++q)
      if (devs[q] == devs[r]) { g_create_err = "upcgpu_create_multi: a device is listed twice"; return UPCGPU_EINVAL; }
  }
  upcgpu_ctx* leader = nullptr;
  int rc = upcgpu_create(params, devs[0], &leader);
  if (rc) return rc;
  if (n_gpus > 1) {
    std::string err;
    rc = group_create(leader, n_gpus, devs.data(), err);
    if (rc) { g_create_err = "upcgpu_create_multi: " + err; upcgpu_destroy(leader); return rc; }
  }
  *out = leader;
  return UPCGPU_OK;
}

int upcgpu_group_size(const upcgpu_ctx* c) { return c ? group_size(c) : 0; }

int upcgpu_group_member(upcgpu_ctx* c, int rank, upcgpu_ctx** member)
{
  if (!c || !member) return UPCGPU_EINVAL;
  *member = group_member(c, rank);
  return *member ? UPCGPU_OK : UPCGPU_EINVAL;
}

int upcgpu_group_set_exchange(upcgpu_ctx* c, int mode)
{
  if (!c) return UPCGPU_EINVAL;
  return group_set_exchange(c, mode);
}

int upcgpu_group_describe(const upcgpu_ctx* c, char* buf, size_t cap)
{
  if (!c || !buf || cap == 0) return UPCGPU_EINVAL;
  if (!c->group) { std::snprintf(buf, cap, "1 device"); return UPCGPU_OK; }
  group_describe(c, buf, cap);
  return UPCGPU_OK;
}

void upcgpu_destroy(upcgpu_ctx* c)
{
  if (!c) return;
  if (c->group && c->group_rank == 0) group_destroy(c);  // the members, their threads, the NCCL communicators
  cudaSetDevice(c->device);
  for (int d = 0; d < c->ipc_n; ++d)
    for (int w = 0; w < 3; ++w)
      if (d != c->ipc_rank && c->ipc_lumi[d][w]) cudaIpcCloseMemHandle(c->ipc_lumi[d][w]);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->gaa_x); cudaFree(c->gaa_y); cudaFree(c->gaa_c); cudaFree(c->ta_y); cudaFree(c->ta_c);
  cudaFree(c->ff_y); cudaFree(c->ff_c); cudaFree(c->bk_y); cudaFree(c->bk_c);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  cudaFree(c->gaa_seg); cudaFree(c->ff_seg); cudaFree(c->bk_seg); cudaFree(c->d_scal); cudaFree(c->bk_table); cudaFree(c->fold_ws);
  for (int w = 0; w < 3; w++) { cudaFree(c->lumi[w]); cudaFree(c->shard[w]); cudaFree(c->gather[w]); }
  cudaFree(c->cs); cudaFree(c->ratio); cudaFree(c->sum2d); cudaFree(c->sumz); cudaFree(c->sumz_ps);
  cudaFree(c->edges_y); cudaFree(c->edges_m); cudaFree(c->edges_z);
  cudaFree(c->samp_term); cudaFree(c->samp_mean); cudaFree(c->samp_dz); cudaFree(c->spec_stats);
  free_event_scratch(c);
  free_lumi_scratch(c);
  cudaFree(c->cell_counter);
  for (int i = 0; i < 2; ++i) if (c->aux[i]) cudaStreamDestroy(c->aux[i]);
  for (int i = 0; i < 4; ++i) if (c->aux_ev[i]) cudaEventDestroy(c->aux_ev[i]);
  for (int i = 0; i < 2; ++i) if (c->tab_ev[i]) cudaEventDestroy(c->tab_ev[i]);
  for (int i = 0; i < 2; ++i) if (c->fill_ev[i]) cudaEventDestroy(c->fill_ev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int upcgpu_device_name(const upcgpu_ctx* c, char* buf, size_t cap)
{
  if (!c || !buf || cap == 0) return UPCGPU_EINVAL;
  std::snprintf(buf, cap, "%s (sm_%d%d, %d SMs)", c->prop.name, c->prop.major, c->prop.minor, c->prop.multiProcessorCount);
  return UPCGPU_OK;
}


int upcgpu_prepare_tables(upcgpu_ctx* c)
{
  CHECK_CTX(c);
  if (c->group && c->group_rank == 0) return group_prepare_tables(c);
  if (c->tables_ready) return UPCGPU_OK;
  return prepare_tables(c);
}

int upcgpu_stream_handle(const upcgpu_ctx* c, uint64_t* stream)
{
  if (!c || !stream) return UPCGPU_EINVAL;
  *stream = (uint64_t)(uintptr_t)c->stream;
  return UPCGPU_OK;
}

long long upcgpu_launch_count(const upcgpu_ctx* c) { return c ? c->launches : -1; }

int upcgpu_invalidate_tables(upcgpu_ctx* c)
{
  if (!c) return UPCGPU_EINVAL;
  c->tables_ready = false;
  if (c->group && c->group_rank == 0) group_invalidate_tables(c);
  return UPCGPU_OK;
}

int upcgpu_fp64_peak(upcgpu_ctx* c, int iters, double* tflops, double* ms)
{
  CHECK_CTX_SYNC(c);
  return fp64_peak(c, iters, tflops, ms);
}

int upcgpu_get_table_info(const upcgpu_ctx* c, upcgpu_table_info* info)
{
  if (!c || !info || !c->tables_ready) return UPCGPU_EINVAL;
  *info = c->info;
  return UPCGPU_OK;
}

int upcgpu_get_table(upcgpu_ctx* c, int which, size_t i0, size_t n, double* x, double* y, double* cc)
{
  CHECK_CTX_SYNC(c);
  if (!c->tables_ready) { c->err = "get_table: tables not prepared"; return UPCGPU_EINVAL; }
  const double *dy = nullptr, *dc = nullptr;
  size_t size = 0;
  double x0 = 0, dx = 0;
  switch (which) {
    case UPCGPU_TABLE_GAA: dy = c->gaa_y; dc = c->gaa_c; size = kNB; break;
    case UPCGPU_TABLE_TA: dy = c->ta_y; dc = c->ta_c; size = kNB; break;
    case UPCGPU_TABLE_FORMFAC: dy = c->ff_y; dc = c->ff_c; size = kNQ2; x0 = kQ2min; dx = kDQ2; break;
    case UPCGPU_TABLE_BREAKUP: dy = c->bk_y; dc = c->bk_c; size = c->bk_nknots; x0 = kBkBmin; dx = kBkDb; break;
    default: c->err = "get_table: unknown table"; return UPCGPU_EINVAL;
  }
  if (!dy || i0 + n > size) { c->err = "get_table: range/table unavailable"; return UPCGPU_EINVAL; }
  if (y) UPC_CUDA(c, cudaMemcpy(y, dy + i0, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (cc) UPC_CUDA(c, cudaMemcpy(cc, dc + i0, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (x) {
    if (which == UPCGPU_TABLE_GAA || which == UPCGPU_TABLE_TA) {
      UPC_CUDA(c, cudaMemcpy(x, c->gaa_x + i0, n * sizeof(double), cudaMemcpyDeviceToHost));
    } else {
      for (size_t i = 0; i < n; i++) {
        volatile double prod = (double)(i0 + i) * dx;  // two roundings, as the reference's knots
        x[i] = x0 + prod;
      }
    }
  }
  return UPCGPU_OK;
}

int upcgpu_eval_table(upcgpu_ctx* c, int which, const double* x, size_t n, double* out)
{
  CHECK_CTX_SYNC(c);
  if (!c->tables_ready || !x || !out) { c->err = "eval_table: bad state/argument"; return UPCGPU_EINVAL; }
  if (n == 0) return UPCGPU_OK;
  return eval_table(c, which, x, n, out);
}

int upcgpu_breakup_raw(upcgpu_ctx* c, const double* b, int mode, size_t n, double* out)
{
  CHECK_CTX_SYNC(c);
  if (!c->tables_ready || !c->bk_seg) { c->err = "breakup_raw: breakup table not prepared (BREAKUP_MODE 1?)"; return UPCGPU_EINVAL; }
  if (n == 0) return UPCGPU_OK;
  return breakup_raw(c, b, mode, n, out);
}

int upcgpu_flux_point(upcgpu_ctx* c, const double* b, const double* k, size_t n, double* out)
{
  CHECK_CTX_SYNC(c);
  if (!b || !k || !out) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return flux_points(c, b, k, n, 1, out, nullptr);
}

int upcgpu_flux_form(upcgpu_ctx* c, const double* b, const double* k, size_t n, double* out, int* neval)
{
  CHECK_CTX_SYNC(c);
  if (!b || !k || !out) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return flux_points(c, b, k, n, 0, out, neval);
}

int upcgpu_fill_lumi_shard(upcgpu_ctx* c, int shard, int nshards)
{
  CHECK_CTX(c);
  return fill_lumi_rows(c, shard, nshards, /*wait=*/false);
}

int upcgpu_lumi_download(upcgpu_ctx* c, int which, double* host)
{
  CHECK_CTX_SYNC(c);
  if (which < 0 || which > 2 || !host || !c->lumi[which]) { c->err = "lumi_download: table not available"; return UPCGPU_EINVAL; }
  // on the context's stream: upcgpu_lumi_unpack and the fill are queued there and do not wait
  UPC_CUDA(c, cudaMemcpyAsync(host, c->lumi[which], (size_t)c->p.nm * c->p.ny * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  return UPCGPU_OK;
}

int upcgpu_lumi_upload(upcgpu_ctx* c, int which, const double* host)
{
  CHECK_CTX_SYNC(c);
  if (which < 0 || which > 2 || !host) return UPCGPU_EINVAL;
  if ((which == 0) == (c->p.use_pol != 0)) { c->err = "lumi_upload: table kind does not match use_pol"; return UPCGPU_EINVAL; }
  int rc = ensure_lumi_buffers(c, 0);
  if (rc) return rc;
  UPC_CUDA(c, cudaMemcpyAsync(c->lumi[which], host, (size_t)c->p.nm * c->p.ny * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  UPC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->lumi_ready = true;
  return UPCGPU_OK;
}

int upcgpu_fill_lumi(upcgpu_ctx* c, double* lumi, double* lumi_s, double* lumi_p)
{
  CHECK_CTX_SYNC(c);
  int rc = upcgpu_prepare_tables(c);
  if (rc) return rc;
  // several GPUs behind this handle: the m rows are dealt to the devices, exchanged, and every device holds the table
  rc = (c->group && c->group_rank == 0) ? group_fill_lumi(c) : fill_lumi_rows(c, 0, 1, /*wait=*/true);
  if (rc) return rc;
  if (!c->p.use_pol) {
    if (lumi) rc = upcgpu_lumi_download(c, 0, lumi);
  } else {
    if (lumi_s) rc = upcgpu_lumi_download(c, 1, lumi_s);
    if (!rc && lumi_p) rc = upcgpu_lumi_download(c, 2, lumi_p);
  }
  return rc;
}

int upcgpu_lumi_cells(upcgpu_ctx* c, const double* M, const double* Y, size_t n, double* out, double* out_s, double* out_p)
{
  CHECK_CTX_SYNC(c);
  if (!M || !Y) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return lumi_cells(c, M, Y, n, out, out_s, out_p);
}

int upcgpu_photon_flux(upcgpu_ctx* c, const double* M, const double* Y, size_t n, double* flux_pos, double* flux_neg)
{
  CHECK_CTX_SYNC(c);
  if (!M || !Y || (!flux_pos && !flux_neg)) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return lumi_cells(c, M, Y, n, nullptr, nullptr, nullptr, flux_pos, flux_neg);
}

int upcgpu_get_fill_stats(upcgpu_ctx* c, upcgpu_fill_stats* st)
{
  if (!st) return UPCGPU_EINVAL;
  CHECK_CTX_SYNC(c);
  if (c->group && c->group_rank == 0) group_fill_stats(c, st);
  else *st = c->stats;
  return UPCGPU_OK;
}

int upcgpu_lumi_shard_buffer(upcgpu_ctx* c, int which, uint64_t* dev_ptr, size_t* n_doubles)
{
  CHECK_CTX(c);
  if (which < 0 || which > 2 || !c->shard[which]) { c->err = "lumi_shard_buffer: no shard computed"; return UPCGPU_EINVAL; }
  if (dev_ptr) *dev_ptr = (uint64_t)(uintptr_t)c->shard[which];
  if (n_doubles) *n_doubles = c->shard_rows * c->p.ny;
  return UPCGPU_OK;
}

int upcgpu_lumi_gather_buffer(upcgpu_ctx* c, int which, int nshards, uint64_t* dev_ptr, size_t* n_doubles)
{
  CHECK_CTX(c);
  if (which < 0 || which > 2) { c->err = "lumi_gather_buffer: bad table kind"; return UPCGPU_EINVAL; }
  const int rc = ensure_gather_buffers(c, nshards);
  if (rc) return rc;
  if (!c->gather[which]) { c->err = "lumi_gather_buffer: table kind does not match use_pol"; return UPCGPU_EINVAL; }
  if (dev_ptr) *dev_ptr = (uint64_t)(uintptr_t)c->gather[which];
  if (n_doubles) *n_doubles = c->shard_rows * c->p.ny * nshards;
  return UPCGPU_OK;
}

// ---- peer stores between processes (one rank per GPU): CUDA IPC handles of the full tables ----
int upcgpu_lumi_ipc_export(upcgpu_ctx* c, void* handles /* 3 x 64 bytes */)
{
  CHECK_CTX_SYNC(c);
  if (!handles) return UPCGPU_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  const int rc = ensure_lumi_buffers(c, 0);
  if (rc) return rc;
  std::memset(handles, 0, 3 * sizeof(cudaIpcMemHandle_t));
  for (int w = 0; w < 3; ++w)
    if (c->lumi[w]) UPC_CUDA(c, cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handles + w, c->lumi[w]));
  return UPCGPU_OK;
}

int upcgpu_lumi_ipc_import(upcgpu_ctx* c, int nshards, int rank, const void* handles /* nshards x 3 x 64 bytes */)
{
  CHECK_CTX_SYNC(c);
  if (!handles || nshards < 1 || nshards > kMaxPeers || rank < 0 || rank >= nshards) { c->err = "lumi_ipc_import: bad argument"; return UPCGPU_EINVAL; }
  int rc = ensure_lumi_buffers(c, 0);
  if (rc) return rc;
  for (int d = 0; d < c->ipc_n; ++d)
    for (int w = 0; w < 3; ++w)
      if (d != c->ipc_rank && c->ipc_lumi[d][w]) { cudaIpcCloseMemHandle(c->ipc_lumi[d][w]); c->ipc_lumi[d][w] = nullptr; }
  c->ipc_n = 0;
  const int w0 = c->p.use_pol ? 1 : 0, w1 = c->p.use_pol ? 2 : 0;
  for (int d = 0; d < nshards; ++d)
    for (int w = w0; w <= w1; ++w) {
      if (d == rank) { c->ipc_lumi[d][w] = c->lumi[w]; continue; }
      void* ptr = nullptr;
      UPC_CUDA(c, cudaIpcOpenMemHandle(&ptr, ((const cudaIpcMemHandle_t*)handles)[d * 3 + w], cudaIpcMemLazyEnablePeerAccess));
      c->ipc_lumi[d][w] = (double*)ptr;
    }
  c->ipc_n = nshards;
  c->ipc_rank = rank;
  return UPCGPU_OK;
}

// Sharded fill whose cell kernel stores every finished cell into the full table of EVERY rank (the tables imported
// with upcgpu_lumi_ipc_import): no gather buffer, no un-permute.  Queued like upcgpu_fill_lumi_shard; the caller
// orders the ranks behind it (any collective on upcgpu_stream_handle, e.g. a one-element all-reduce) before the fold.
int upcgpu_fill_lumi_shard_peers(upcgpu_ctx* c)
{
  CHECK_CTX(c);
  if (c->ipc_n < 1) { c->err = "fill_lumi_shard_peers: call upcgpu_lumi_ipc_import first"; return UPCGPU_EINVAL; }
  for (int d = 0; d < c->ipc_n; ++d)
    for (int w = 0; w < 3; ++w) c->peer_lumi[d][w] = c->ipc_lumi[d][w];
  c->n_peers = c->ipc_n;
  const int rc = fill_lumi_rows(c, c->ipc_rank, c->ipc_n, /*wait=*/false);
  c->n_peers = 0;
  c->lumi_ready = true;  // ... once the caller's collective has passed on this stream
  return rc;
}

int upcgpu_lumi_unpack(upcgpu_ctx* c, int nshards)
{
  CHECK_CTX(c);
  return lumi_unpack(c, nshards);
}

int upcgpu_fold_sigma(upcgpu_ctx* c, const double* sig_m, const double* sig_s, const double* sig_p, double* cs,
                      double* ratio, double* totcs_mb)
{
  CHECK_CTX(c);
  if (c->group && c->group_rank == 0) return group_fold_sigma(c, sig_m, sig_s, sig_p, cs, ratio, totcs_mb);
  return fold_sigma(c, sig_m, sig_s, sig_p, cs, ratio, totcs_mb);
}

int upcgpu_sampler_build(upcgpu_ctx* c, const double* cs, const double* cszm, const double* cszm_s, const double* cszm_ps)
{
  CHECK_CTX_SYNC(c);
  if (c->group && c->group_rank == 0) return group_sampler_build(c, cs, cszm, cszm_s, cszm_ps);
  return sampler_build(c, cs, cszm, cszm_s, cszm_ps);
}

int upcgpu_sampler_spec_stats(upcgpu_ctx* c, unsigned long long* out6)
{
  CHECK_CTX_SYNC(c);
  if (!out6) return UPCGPU_EINVAL;
  return sampler_spec_stats(c, out6);
}

int upcgpu_sampler_get_cdf(upcgpu_ctx* c, double* sum2d, double* sumz, double* sumz_ps)
{
  CHECK_CTX_SYNC(c);
  if (!c->sampler_ready) { c->err = "sampler_get_cdf: samplers not built"; return UPCGPU_EINVAL; }
  const size_t n = (size_t)c->p.nm * c->p.ny, nsz = (size_t)c->p.nm * (c->p.nz + 1);
  if (sum2d) UPC_CUDA(c, cudaMemcpy(sum2d, c->sum2d, (n + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  if (sumz && c->sumz) UPC_CUDA(c, cudaMemcpy(sumz, c->sumz, nsz * sizeof(double), cudaMemcpyDeviceToHost));
  if (sumz_ps && c->sumz_ps) UPC_CUDA(c, cudaMemcpy(sumz_ps, c->sumz_ps, nsz * sizeof(double), cudaMemcpyDeviceToHost));
  return UPCGPU_OK;
}

int upcgpu_sample_ym(upcgpu_ctx* c, const double* u, size_t n, long long* k, int* ybin, int* mbin, double* y, double* m)
{
  CHECK_CTX_SYNC(c);
  if (!u) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return sample_ym(c, u, n, k, ybin, mbin, y, m);
}

int upcgpu_sample_z(upcgpu_ctx* c, const int* mbin, const double* u, size_t n, int ps, double* z)
{
  CHECK_CTX_SYNC(c);
  if (!u || !mbin || !z) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return sample_z(c, mbin, u, n, ps, z);
}

int upcgpu_hist_pdf_init(upcgpu_ctx* c, const double* bins, size_t n, double* sum)
{
  CHECK_CTX_SYNC(c);
  if (!bins || !sum || n == 0) return UPCGPU_EINVAL;
  return hist_pdf_init(c, bins, n, sum);
}

int upcgpu_hist_sample2d(upcgpu_ctx* c, const double* sum, int nx, int ny, const double* xe, const double* ye,
                         const double* u, size_t n, long long* k, double* x, double* y)
{
  CHECK_CTX_SYNC(c);
  if (!sum || !xe || !ye || !u || !x || !y || nx < 1 || ny < 1) return UPCGPU_EINVAL;
  if (n == 0) return UPCGPU_OK;
  return hist_sample2d(c, sum, nx, ny, xe, ye, u, n, k, x, y);
}

int upcgpu_hist_sample1d(upcgpu_ctx* c, const double* sum, int n, const double* edges, const double* u, size_t nsamp,
                         double* x)
{
  CHECK_CTX_SYNC(c);
  if (!sum || !edges || !u || !x || n < 1) return UPCGPU_EINVAL;
  if (nsamp == 0) return UPCGPU_OK;
  return hist_sample1d(c, sum, n, edges, u, nsamp, x);
}

int upcgpu_particles_per_event(const upcgpu_ctx* c)
{
  if (!c) return UPCGPU_EINVAL;
  return particles_per_event(c);
}

int upcgpu_generate_packed(upcgpu_ctx* c, uint64_t seed, uint64_t first_candidate, size_t n_candidates, int part_stride, int* npart,
                           int* pdg, int* status, int* mother, double* p4, double* aux, uint64_t* n_accepted)
{
  CHECK_CTX_SYNC(c);
  if (part_stride < particles_per_event(c) || part_stride > UPCGPU_MAX_PART) {
    c->err = "generate: part_stride must be between upcgpu_particles_per_event() and UPCGPU_MAX_PART";
    return UPCGPU_EINVAL;
  }
  if (n_candidates == 0) { if (n_accepted) *n_accepted = 0; return UPCGPU_OK; }
  if (c->group && c->group_rank == 0)
    return group_generate(c, seed, first_candidate, n_candidates, part_stride, npart, pdg, status, mother, p4, aux, n_accepted, false);
  return generate(c, seed, first_candidate, n_candidates, part_stride, npart, pdg, status, mother, p4, aux, n_accepted, false);
}

int upcgpu_generate(upcgpu_ctx* c, uint64_t seed, uint64_t first_candidate, size_t n_candidates, int* npart, int* pdg,
                    int* status, int* mother, double* p4, double* aux, uint64_t* n_accepted)
{
  return upcgpu_generate_packed(c, seed, first_candidate, n_candidates, UPCGPU_MAX_PART, npart, pdg, status, mother, p4, aux,
                                n_accepted);
}

int upcgpu_generate_device(upcgpu_ctx* c, uint64_t seed, uint64_t first_candidate, size_t n_candidates, uint64_t* n_accepted)
{
  CHECK_CTX_SYNC(c);
  if (n_candidates == 0) { if (n_accepted) *n_accepted = 0; return UPCGPU_OK; }
  if (c->group && c->group_rank == 0)
    return group_generate(c, seed, first_candidate, n_candidates, UPCGPU_MAX_PART, nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, n_accepted, true);
  return generate(c, seed, first_candidate, n_candidates, UPCGPU_MAX_PART, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                  n_accepted, true);
}

int upcgpu_photon_pt_cdf(upcgpu_ctx* c, double e_phot, double* cdf)
{
  CHECK_CTX_SYNC(c);
  if (!cdf) return UPCGPU_EINVAL;
  return photon_pt_cdf(c, e_phot, cdf);
}

int upcgpu_philox(uint64_t seed, uint64_t ctr0, uint32_t block, size_t n, double* out)
{
  if (!out) return UPCGPU_EINVAL;
  for (size_t i = 0; i < n; i++) philox4x32_10(seed, ctr0 + i, block, out[2 * i], out[2 * i + 1]);
  return UPCGPU_OK;
}

}  // extern "C"
