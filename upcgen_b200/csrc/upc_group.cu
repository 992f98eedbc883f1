// upc_group.cu -- several GPUs behind ONE context handle, from one process.
//
// upcgpu_create_multi gives the caller a context like upcgpu_create does; behind it stands one member
// context per device, each driven by its own host thread.  The reference's grid driver
// (UpcCrossSection::prepareTwoPhotonLumi, src/UpcCrossSection.cpp:463-592) slabs the m rows over OpenMP
// threads; here the rows are dealt to the devices (blocks of 32 rows in a snake, see shard_of_row) and every
// device ends up with the full table, because every device samples events from it afterwards:
//
//   exchange 0: NCCL (the fallback where the devices cannot map each other).  ncclCommInitAll over the devices;
//     each member all-gathers its packed shard (ncclAllGather on the member's stream) and un-permutes the
//     gathered buffer (k_unpack).
//   exchange 1 (default with peer access): peer stores.  The cell kernel itself writes every finished cell into
//     the full table of EVERY device (8-byte stores over NVLink while the quadrature of the other cells goes
//     on); no gather buffer, no un-permute kernel -- the collective is folded into the kernel's epilogue.
//     The members then wait for each other's "cells done" events (cudaStreamWaitEvent across devices).
//
// Events shard by contiguous candidate ranges of the Philox counter: results do not depend on the number
// of devices, and there is no collective on that path.
//
// NCCL is loaded with dlopen at group creation (no link-time dependency: the library also loads where no
// NCCL is installed, and inside a PyTorch process it binds to the copy torch already loaded).
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "upc_ctx.h"
#include "upc_internal.h"

namespace upc {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err)
  {
    const char* names[] = {std::getenv("UPCGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load NCCL (libnccl.so.2; set UPCGPU_NCCL_LIB): ") + dlerror(); return false; }
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetVersion = (decltype(GetVersion))dlsym(lib, "ncclGetVersion");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !AllGather || !GetErrorString) { err = "NCCL library lacks the expected symbols"; return false; }
    return true;
  }
};

// one host thread per device: runs the jobs posted to it, in order
class Worker
{
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker()
  {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<int()> job)
  {
    {
      std::lock_guard<std::mutex> l(m_);
      job_ = std::move(job);
      has_ = true;
      done_ = false;
    }
    cv_.notify_all();
  }
  int wait()
  {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [this] { return done_; });
    return rc_;
  }

 private:
  void loop()
  {
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return has_ || stop_; });
        if (stop_) return;
        job = std::move(job_);
        has_ = false;
      }
      const int rc = job();
      {
        std::lock_guard<std::mutex> l(m_);
        rc_ = rc;
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_ = false, stop_ = false, done_ = true;
  int rc_ = 0;
  std::thread th_;
};

// reusable barrier of the member threads (C++17: no std::barrier)
class HostBarrier
{
 public:
  explicit HostBarrier(int n) : n_(n) {}
  void arrive_and_wait()
  {
    std::unique_lock<std::mutex> l(m_);
    const unsigned long gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(l, [&] { return gen_ != gen; });
    }
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  unsigned long gen_ = 0;
};

struct Group {
  int n = 0;
  std::vector<upcgpu_ctx*> ctx;       // ctx[0] is the leader (the handle the caller holds)
  std::vector<Worker*> workers;
  NcclApi nccl;
  std::vector<ncclComm_t> comms;
  bool have_nccl = false, have_peer = false;
  int exchange = 0;
  std::vector<cudaEvent_t> cells_done;  // one per member, recorded on its stream (peer exchange)
  HostBarrier* bar = nullptr;
  std::atomic<int> failed{0};           // a member failed before the exchange: nobody enters the collective
  std::string nccl_version;
};

// runs job(rank) on every member's thread; first failure wins (its message is copied to the leader)
static int run_all(Group* g, const std::function<int(int)>& job)
{
  for (int r = 0; r < g->n; ++r) g->workers[r]->post([&job, r] { return job(r); });
  int rc = UPCGPU_OK;
  for (int r = 0; r < g->n; ++r) {
    const int rr = g->workers[r]->wait();
    if (rr && !rc) {
      rc = rr;
      if (r != 0) g->ctx[0]->err = "device " + std::to_string(g->ctx[r]->device) + ": " + g->ctx[r]->err;
    }
  }
  return rc;
}

void group_destroy(upcgpu_ctx* leader)
{
  Group* g = leader->group;
  if (!g) return;
  for (Worker* w : g->workers) delete w;  // joins
  for (int r = 0; r < g->n; ++r) {
    cudaSetDevice(g->ctx[r]->device);
    if (g->have_nccl && g->comms[r]) g->nccl.CommDestroy(g->comms[r]);
    if (g->cells_done[r]) cudaEventDestroy(g->cells_done[r]);
  }
  for (int r = 1; r < g->n; ++r) {
    g->ctx[r]->group = nullptr;
    upcgpu_destroy(g->ctx[r]);
  }
  delete g->bar;
  delete g;
  leader->group = nullptr;
}

int group_create(upcgpu_ctx* leader, int n_gpus, const int* devices, std::string& err)
{
  Group* g = new Group();
  g->n = n_gpus;
  g->ctx.assign(n_gpus, nullptr);
  g->ctx[0] = leader;
  g->comms.assign(n_gpus, nullptr);
  g->cells_done.assign(n_gpus, nullptr);
  leader->group = g;
  for (int r = 1; r < n_gpus; ++r) {
    const int rc = upcgpu_create(&leader->p, devices[r], &g->ctx[r]);
    if (rc) { err = std::string("member ") + std::to_string(r) + ": " + upcgpu_last_error(nullptr); g->n = r; group_destroy(leader); return rc; }
    g->ctx[r]->group_rank = r;
  }
  leader->group_rank = 0;
  g->bar = new HostBarrier(n_gpus);
  // peer access between every pair (the peer-store exchange; NCCL uses its own mappings)
  g->have_peer = true;
  for (int a = 0; a < n_gpus; ++a) {
    cudaSetDevice(g->ctx[a]->device);
    for (int b = 0; b < n_gpus; ++b) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, g->ctx[a]->device, g->ctx[b]->device);
      if (!can) { g->have_peer = false; continue; }
      const cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[b]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) g->have_peer = false;
      cudaGetLastError();
    }
    cudaEventCreateWithFlags(&g->cells_done[a], cudaEventDisableTiming);
  }
  std::string nerr;
  if (g->nccl.load(nerr)) {
    std::vector<int> devs(n_gpus);
    for (int r = 0; r < n_gpus; ++r) devs[r] = g->ctx[r]->device;
    const ncclResult_t nr = g->nccl.CommInitAll(g->comms.data(), n_gpus, devs.data());
    if (nr == ncclSuccess) {
      g->have_nccl = true;
      int v = 0;
      if (g->nccl.GetVersion && g->nccl.GetVersion(&v) == ncclSuccess) g->nccl_version = std::to_string(v);
    } else {
      nerr = std::string("ncclCommInitAll: ") + g->nccl.GetErrorString(nr);
    }
  }
  if (!g->have_nccl && !g->have_peer) {
    err = "no exchange path between the devices: " + nerr + "; peer access unavailable";
    group_destroy(leader);
    return UPCGPU_ECUDA;
  }
  g->exchange = g->have_peer ? 1 : 0;  // peer stores where the devices can map each other, else the all-gather
  if (const char* e = std::getenv("UPCGPU_EXCHANGE")) {
    if (!std::strcmp(e, "peer") && g->have_peer) g->exchange = 1;
    if (!std::strcmp(e, "nccl") && g->have_nccl) g->exchange = 0;
  }
  for (int r = 0; r < n_gpus; ++r) g->workers.push_back(new Worker());
  return UPCGPU_OK;
}

int group_size(const upcgpu_ctx* c) { return c->group ? c->group->n : 1; }
upcgpu_ctx* group_member(upcgpu_ctx* c, int rank) { return (c->group && rank >= 0 && rank < c->group->n) ? c->group->ctx[rank] : (rank == 0 ? c : nullptr); }
int group_exchange(const upcgpu_ctx* c) { return c->group ? c->group->exchange : -1; }
int group_set_exchange(upcgpu_ctx* c, int mode)
{
  Group* g = c->group;
  if (!g) { c->err = "set_exchange: not a multi-GPU context"; return UPCGPU_EINVAL; }
  if (mode == 0 && !g->have_nccl) { c->err = "set_exchange: NCCL is not available"; return UPCGPU_EINVAL; }
  if (mode == 1 && !g->have_peer) { c->err = "set_exchange: peer access is not available between all devices"; return UPCGPU_EINVAL; }
  if (mode != 0 && mode != 1) { c->err = "set_exchange: 0 (NCCL all-gather) or 1 (peer stores)"; return UPCGPU_EINVAL; }
  g->exchange = mode;
  return UPCGPU_OK;
}

int group_prepare_tables(upcgpu_ctx* leader)
{
  Group* g = leader->group;
  return run_all(g, [g](int r) {
    upcgpu_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    return c->tables_ready ? UPCGPU_OK : prepare_tables(c);
  });
}

void group_invalidate_tables(upcgpu_ctx* leader)
{
  for (upcgpu_ctx* c : leader->group->ctx) c->tables_ready = false;
}

// tables -> sharded fill -> exchange -> full table on every device.  Every member synchronises once, at the end.
int group_fill_lumi(upcgpu_ctx* leader)
{
  Group* g = leader->group;
  const int n = g->n;
  const bool pol = leader->p.use_pol != 0;
  const int w0 = pol ? 1 : 0, w1 = pol ? 2 : 0;
  const int exchange = g->exchange;
  g->failed = 0;
  return run_all(g, [=](int r) {
    upcgpu_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    int rc = c->tables_ready ? UPCGPU_OK : prepare_tables(c);
    if (!rc) rc = ensure_lumi_buffers(c, n);
    if (!rc && exchange == 0) {
      c->shard_n = n;  // ensure_lumi_buffers sized the shard buffers for n shards
      rc = ensure_gather_buffers(c, n);
    }
    // every member's tables and buffers exist (or the step is called off) before any kernel stores into a peer
    // and before anyone enters a collective
    if (rc) g->failed = 1;
    g->bar->arrive_and_wait();
    if (g->failed) return rc ? rc : (c->err = "another device failed while preparing the fill", (int)UPCGPU_ECUDA);
    if (exchange == 1) {
      for (int d = 0; d < n; ++d)
        for (int w = 0; w < 3; ++w) c->peer_lumi[d][w] = g->ctx[d]->lumi[w];
      c->n_peers = n;
    } else {
      c->n_peers = 0;
    }
    rc = fill_lumi_rows(c, r, n, /*wait=*/false);
    if (rc) g->failed = 1;
    g->bar->arrive_and_wait();
    if (g->failed) { c->n_peers = 0; finish_fill(c); return rc ? rc : (c->err = "another device failed while queueing the fill", (int)UPCGPU_ECUDA); }
    if (exchange == 0) {
      if (!rc) {
        const size_t cnt = c->shard_rows * (size_t)c->p.ny;
        for (int w = w0; w <= w1 && !rc; ++w) {
          const ncclResult_t nr = g->nccl.AllGather(c->shard[w], c->gather[w], cnt, ncclDouble, g->comms[r], c->stream);
          if (nr != ncclSuccess) { c->err = std::string("ncclAllGather: ") + g->nccl.GetErrorString(nr); rc = UPCGPU_ECUDA; }
        }
      }
      if (!rc) rc = lumi_unpack(c, n);
    } else {
      cudaEventRecord(g->cells_done[r], c->stream);
      g->bar->arrive_and_wait();  // every member has recorded its event
      for (int d = 0; d < n; ++d)
        if (d != r) cudaStreamWaitEvent(c->stream, g->cells_done[d], 0);
      c->lumi_ready = true;
    }
    c->n_peers = 0;
    const int frc = finish_fill(c);  // the member's one host wait
    return rc ? rc : frc;
  });
}

int group_fold_sigma(upcgpu_ctx* leader, const double* sig_m, const double* sig_s, const double* sig_p, double* cs, double* ratio,
                     double* totcs_mb)
{
  Group* g = leader->group;
  return run_all(g, [=](int r) {
    upcgpu_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    // replicated (16 B per cell): every device samples events from its own copy; the leader hands the table back
    return fold_sigma(c, sig_m, sig_s, sig_p, r == 0 ? cs : nullptr, r == 0 ? ratio : nullptr, r == 0 ? totcs_mb : nullptr);
  });
}

int group_sampler_build(upcgpu_ctx* leader, const double* cs, const double* cszm, const double* cszm_s, const double* cszm_ps)
{
  Group* g = leader->group;
  return run_all(g, [=](int r) {
    upcgpu_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    return sampler_build(c, cs, cszm, cszm_s, cszm_ps);
  });
}

// candidates [first, first + n) in contiguous ranges, one per device; host arrays are filled in place
int group_generate(upcgpu_ctx* leader, uint64_t seed, uint64_t first, size_t n, int part_stride, int* npart, int* pdg, int* status,
                   int* mother, double* p4, double* aux, uint64_t* n_acc, bool device_only)
{
  Group* g = leader->group;
  const int G = g->n;
  std::vector<uint64_t> acc(G, 0);
  const size_t per = (n + G - 1) / G;
  const int rc = run_all(g, [&, per](int r) {
    upcgpu_ctx* c = g->ctx[r];
    cudaSetDevice(c->device);
    const size_t o = std::min(n, per * (size_t)r), cn = std::min(per, n - o);
    if (cn == 0) return (int)UPCGPU_OK;
    const size_t os = o * (size_t)part_stride;
    return generate(c, seed, first + o, cn, part_stride, npart ? npart + o : nullptr, pdg ? pdg + os : nullptr,
                    status ? status + os : nullptr, mother ? mother + os : nullptr, p4 ? p4 + os * 4 : nullptr,
                    aux ? aux + o * 5 : nullptr, &acc[r], device_only);
  });
  uint64_t tot = 0;
  for (uint64_t a : acc) tot += a;
  if (n_acc) *n_acc = tot;
  return rc;
}

// counters summed over the members, stage times = the slowest member's
void group_fill_stats(const upcgpu_ctx* leader, upcgpu_fill_stats* out)
{
  upcgpu_fill_stats s{};
  for (const upcgpu_ctx* c : leader->group->ctx) {
    const upcgpu_fill_stats& t = c->stats;
    s.qags_integrals += t.qags_integrals; s.qags_evals += t.qags_evals; s.qags_overflow += t.qags_overflow;
    s.qags_errors += t.qags_errors; s.flux_rows += t.flux_rows; s.band_pairs += t.band_pairs;
    s.qags_head_evals += t.qags_head_evals; s.qags_head_done += t.qags_head_done; s.qags_table_evals += t.qags_table_evals;
    s.cells_evaluated += t.cells_evaluated;
    s.ms_tables = std::max(s.ms_tables, t.ms_tables); s.ms_flux = std::max(s.ms_flux, t.ms_flux);
    s.ms_cells = std::max(s.ms_cells, t.ms_cells); s.ms_total = std::max(s.ms_total, t.ms_total);
    s.ms_qags = std::max(s.ms_qags, t.ms_qags); s.ms_qags_head = std::max(s.ms_qags_head, t.ms_qags_head);
  }
  *out = s;
}

const char* group_describe(const upcgpu_ctx* leader, char* buf, size_t cap)
{
  const Group* g = leader->group;
  std::snprintf(buf, cap, "%d devices, exchange = %s, NCCL %s, peer access %s", g->n, g->exchange ? "peer stores" : "NCCL all-gather",
                g->have_nccl ? g->nccl_version.c_str() : "unavailable", g->have_peer ? "yes" : "no");
  return buf;
}

}  // namespace upc
