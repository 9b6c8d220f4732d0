// internal.h -- host-side seams between the translation units of libaleo_b200.so.
// (Three kernel TUs so that nvcc/ptxas run in parallel; each TU owns a copy of the constant-bank
// moduli and uploads it through its *_upload_constants().)
#pragma once
#include "platform.cuh"

namespace aleo {
// A non-blocking stream owned by one host thread for its current device: created on first use, re-created when the
// thread moves to another device, destroyed when the thread ends (thread_local instances in capi.cu / msm_lib.cu).
struct ThreadStream {
  cudaStream_t s = nullptr;
  int dev = -1;
  ThreadStream() = default;
  ThreadStream(const ThreadStream&) = delete;
  ThreadStream& operator=(const ThreadStream&) = delete;
  ~ThreadStream() {
    if (s) cudaStreamDestroy(s);  // fails harmlessly once the runtime is unloading
  }
  cudaError_t get(cudaStream_t* out) {
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) return e;
    if (s == nullptr || dev != d) {
      if (s) cudaStreamDestroy(s);
      s = nullptr;
      e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
      dev = d;
    }
    *out = s;
    return cudaSuccess;
  }
};

// capi.cu -- workspace allocator: stream-ordered allocations from a PRIVATE memory pool per device (the process's default
// pool and its release threshold are left alone: torch's caching allocator and other users of the default pool share
// the process).  Freed blocks stay in the pool up to ALEO_B200_POOL_KEEP_MB (default 16384) so that the next call does
// not pay for mapping gigabytes again; aleo_b200_shutdown() trims it.  Free with cudaFreeAsync.
cudaError_t pool_malloc_async(void** p, size_t bytes, cudaStream_t s);
// ntt_lib.cu
cudaError_t ntt_upload_constants();
cudaError_t ntt_transform(int device, u32 log_n, size_t batch, bool inverse, bool coset, void* data_dev, cudaStream_t s,
                          float* pass_ms = nullptr /* >= 4 floats; synchronises the stream when given */);
cudaError_t ntt_twiddle_matrix(int device, u32 log_n_global, bool inverse, void* data_dev, u32 rows, u32 cols, u32 row0, u32 col0,
                               cudaStream_t s);
cudaError_t ntt_bitrev(u32 log_n, size_t batch, void* data_dev, cudaStream_t s);
int ntt_launches(u32 log_n);
int ntt_dist_layout(u32 log_n, int world, u32* log_r_first, u32* log_r_last, int* npass);
cudaError_t ntt_dist_create(u32 log_n, int rank, int world, void** ctx_out);
cudaError_t ntt_dist_handles(void* ctx, void* handles_out /* 128 bytes */);
cudaError_t ntt_dist_open(void* ctx, const void* all_handles /* world x 128 bytes */);
cudaError_t ntt_dist_stage1(void* ctx, const void* local_in_dev, bool inverse, bool coset, cudaStream_t s);
cudaError_t ntt_dist_stage2(void* ctx, void* local_out_dev, bool inverse, bool coset, cudaStream_t s);
cudaError_t ntt_dist_profile(void* ctx, const void* local_in_dev, void* local_out_dev, bool inverse, bool coset, cudaStream_t s,
                             float* stage_ms4);
void ntt_dist_destroy(void* ctx);
int ntt_max_log_n();
void ntt_clear_plans();
// msm_lib.cu
cudaError_t msm_upload_constants();
cudaError_t msm_run(const void* bases_dev, u32 stride, const void* scalars_dev, size_t n, void* out144_dev, cudaStream_t s,
                    bool dry, int* launches_out, float* phase_ms = nullptr /* 3 floats; synchronises when given */);
// host-pointer MSM: copies and accumulates point range by point range (overlapped); synchronises s
cudaError_t msm_run_host(const void* bases_host, u32 stride, const void* scalars_host, size_t n, void* out144_host, cudaStream_t s);
int msm_host_chunks(size_t n);
int msm_host_window_bits(size_t n);
cudaError_t srs_msm_host(const void* handle, const void* in_host, size_t n, bool montgomery_in, void* out_host, size_t out_bytes,
                         cudaStream_t s);
// host <-> device copies of caller memory that may be pageable (staged through pinned double buffers by helper threads)
cudaError_t feed_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t cs, cudaEvent_t ready);
cudaError_t feed_d2h_sync(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s);
bool msm_size_supported(size_t n);
int msm_window_bits(size_t n);
int msm_ba_levels(size_t n);
cudaError_t g1_sum(const void* points144_dev, u32 count, void* out144_dev, cudaStream_t s);
cudaError_t srs_create(const void* bases_dev, u32 stride, size_t n, cudaStream_t s, void** handle_out);
void srs_destroy(void* handle);
void srs_info(const void* handle, size_t* n, int* c, int* W, size_t* bytes);
int srs_device(const void* handle);  // the device the handle was created on
bool srs_size_supported(size_t n);   // n * windows must stay below 2^31
cudaError_t srs_msm(const void* handle, const void* scalars_dev, size_t n_used, void* out144_dev, cudaStream_t s, bool dry,
                    int* launches_out, float* phase_ms);
cudaError_t srs_msm_batch(const void* handle, const void* const* scalars_dev_ptrs, const size_t* n_each, size_t count,
                          bool montgomery_in, void* out_dev, bool compressed, cudaStream_t s);
cudaError_t fr_to_bigint(const void* in_dev, void* out_dev, size_t n, cudaStream_t s);
cudaError_t g1_compress(const void* jac144_dev, void* out48_dev, cudaStream_t s);
// poly_lib.cu
cudaError_t poly_upload_constants();
cudaError_t field_op(int field, int op, void* out_dev, const void* a_dev, const void* b_dev, size_t n, cudaStream_t s);
cudaError_t fr_lagrange_coeffs(void* out_dev, u32 log_n, const void* tau32, cudaStream_t s);
cudaError_t fr_axpy(void* y_dev, const void* x_dev, const void* a32, size_t n, cudaStream_t s);
cudaError_t fr_distribute_powers(void* inout_dev, size_t n, const void* g32, const void* k32, cudaStream_t s);
cudaError_t fr_divide_by_vanishing_on_coset(void* inout_dev, u32 log_m, u32 log_n, const void* g32, cudaStream_t s);
cudaError_t fr_poly_eval(void* out_dev, const void* coeffs_dev, size_t n, const void* z32, cudaStream_t s);
cudaError_t fr_divide_by_linear(void* quotient_dev, const void* coeffs_dev, size_t n, const void* z32, cudaStream_t s);
cudaError_t g1_decompress(const void* in48_dev, size_t n, void* out_affine_dev, u32 stride, cudaStream_t s, u32* bad_out, bool unchecked);
cudaError_t g1_compress_affine(const void* affine_dev, u32 stride, size_t n, void* out48_dev, cudaStream_t s);
// util_lib.cu
cudaError_t util_upload_constants();
cudaError_t gen_bases(void* bases_dev, size_t n, u32 stride, const void* s0_32, const void* d_32, u64 first, cudaStream_t s);
cudaError_t gen_scalars(void* scalars_dev, size_t n, u64 seed, u64 first, bool montgomery, cudaStream_t s);
cudaError_t dlog_dot(void* out_dev, const void* scalars_dev, size_t n, const void* s0_32, const void* d_32, u64 first, cudaStream_t s);
cudaError_t check_on_curve(const void* bases_dev, size_t n, u32 stride, cudaStream_t s, int* ok);
cudaError_t bench_imad(int kind, int iters, double* ms_out, double* ops_out);
cudaError_t fq_mul_fp64(void* out_dev, const void* a_dev, const void* b_dev, size_t n, bool square, cudaStream_t s);
}  // namespace aleo
