// upc_internal.h -- declarations shared between the translation units of libupcgpu.so
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "upc_ctx.h"

#define UPC_CUDA(ctx, call)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +      \
                   std::to_string(__LINE__) + ")";                                               \
      return UPCGPU_ECUDA;                                                                       \
    }                                                                                            \
  } while (0)

// every kernel launch of this library goes through UPC_K so that callers can report how many
// of OUR kernels ran (bench.py: gpu_launches)
#define UPC_K(ctx) ((ctx)->launches++)

namespace upc {

// upc_tables.cu
cudaError_t create_stream(cudaStream_t* st, bool high_priority);
int prepare_tables(upcgpu_ctx* c);
int finish_tables(upcgpu_ctx* c);  // collects a queued table stage: waits, checks the scalars against the cache
int eval_table(upcgpu_ctx* c, int which, const double* x, size_t n, double* out);
int breakup_raw(upcgpu_ctx* c, const double* b, int mode, size_t n, double* out);

// upc_lumi.cu
int flux_points(upcgpu_ctx* c, const double* b, const double* k, size_t n, int force_point, double* out, int* neval);
int fill_lumi_rows(upcgpu_ctx* c, int shard, int nshards, bool wait);
int finish_fill(upcgpu_ctx* c);
int lumi_cells(upcgpu_ctx* c, const double* M, const double* Y, size_t n, double* out, double* out_s, double* out_p,
               double* flux_pos = nullptr, double* flux_neg = nullptr);
int lumi_unpack(upcgpu_ctx* c, int nshards);
int ensure_lumi_buffers(upcgpu_ctx* c, int nshards);
int ensure_gather_buffers(upcgpu_ctx* c, int nshards);
void free_lumi_scratch(upcgpu_ctx* c);
int fp64_peak(upcgpu_ctx* c, int iters, double* tflops, double* ms);

// upc_fold.cu
int fold_sigma(upcgpu_ctx* c, const double* sig_m, const double* sig_s, const double* sig_p, double* cs, double* ratio,
               double* totcs_mb);
int sampler_spec_stats(upcgpu_ctx* c, unsigned long long out[6]);
int sampler_build(upcgpu_ctx* c, const double* cs, const double* cszm, const double* cszm_s, const double* cszm_ps);
int sample_ym(upcgpu_ctx* c, const double* u, size_t n, long long* k, int* ybin, int* mbin, double* y, double* m);
int sample_z(upcgpu_ctx* c, const int* mbin, const double* u, size_t n, int ps, double* z);
int hist_pdf_init(upcgpu_ctx* c, const double* bins, size_t n, double* sum);
int hist_sample2d(upcgpu_ctx* c, const double* sum, int nx, int ny, const double* xe, const double* ye, const double* u,
                  size_t n, long long* k, double* x, double* y);
int hist_sample1d(upcgpu_ctx* c, const double* sum, int nb, const double* edges, const double* u, size_t n, double* x);

// upc_events.cu
int particles_per_event(const upcgpu_ctx* c);
int generate(upcgpu_ctx* c, uint64_t seed, uint64_t first, size_t n, int part_stride, int* npart, int* pdg, int* status, int* mother,
             double* p4, double* aux, uint64_t* n_acc, bool device_only);
int photon_pt_cdf(upcgpu_ctx* c, double e, double* cdf);
void free_event_scratch(upcgpu_ctx* c);


// upc_group.cu: several GPUs behind one handle
int group_create(upcgpu_ctx* leader, int n_gpus, const int* devices, std::string& err);
void group_destroy(upcgpu_ctx* leader);
int group_size(const upcgpu_ctx* c);
upcgpu_ctx* group_member(upcgpu_ctx* c, int rank);
int group_exchange(const upcgpu_ctx* c);
int group_set_exchange(upcgpu_ctx* c, int mode);
int group_prepare_tables(upcgpu_ctx* leader);
void group_invalidate_tables(upcgpu_ctx* leader);
int group_fill_lumi(upcgpu_ctx* leader);
int group_fold_sigma(upcgpu_ctx* leader, const double* sig_m, const double* sig_s, const double* sig_p, double* cs, double* ratio,
                     double* totcs_mb);
int group_sampler_build(upcgpu_ctx* leader, const double* cs, const double* cszm, const double* cszm_s, const double* cszm_ps);
int group_generate(upcgpu_ctx* leader, uint64_t seed, uint64_t first, size_t n, int part_stride, int* npart, int* pdg, int* status, int* mother,
                   double* p4, double* aux, uint64_t* n_acc, bool device_only);
void group_fill_stats(const upcgpu_ctx* leader, upcgpu_fill_stats* out);
const char* group_describe(const upcgpu_ctx* leader, char* buf, size_t cap);

}  // namespace upc
