// capi.cu -- the C ABI of libaleo_b200.so (include/aleo_b200.h): argument checking, device
// readiness, per-thread streams and host<->device staging.  The kernels live in ntt_lib.cu,
// msm_lib.cu and util_lib.cu.
//
// No CPU fallback: if the calling thread's CUDA device is missing or is not sm_100, every entry
// point returns ALEO_B200_ENODEVICE.
#include "../../include/aleo_b200.h"
#include <cstdio>
#include <mutex>
#include <string>
#ifndef ALEO_EMU
#include <thread>
#endif
#include <vector>
#include "internal.h"

namespace {

thread_local std::string t_last_cuda_error;
thread_local cudaStream_t t_stream = nullptr;
thread_local int t_stream_dev = -1;

std::mutex g_dev_mu;
bool g_dev_ready[64] = {false};
bool g_dev_bad[64] = {false};

int fail_cuda(cudaError_t e) {
  t_last_cuda_error = cudaGetErrorString(e);
  if (e == cudaErrorMemoryAllocation) return ALEO_B200_ENOMEM;
  return ALEO_B200_ECUDA;
}

#define API_CK(expr)                                   \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) return fail_cuda(_e);       \
  } while (0)

// current device must be a B200-class part; constants uploaded once per device
int ensure_ready(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    t_last_cuda_error = (e != cudaSuccess) ? cudaGetErrorString(e) : "device index out of range";
    return ALEO_B200_ENODEVICE;
  }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev_bad[dev]) return ALEO_B200_ENODEVICE;
  if (!g_dev_ready[dev]) {
#ifndef ALEO_EMU
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
      t_last_cuda_error = cudaGetErrorString(e);
      return ALEO_B200_ENODEVICE;
    }
    if (prop.major != 10) {  // the library holds sm_100a code only
      g_dev_bad[dev] = true;
      t_last_cuda_error =
          std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", need sm_100 (B200)";
      return ALEO_B200_ENODEVICE;
    }
    // Workspaces come from the stream-ordered allocator.  Keep freed blocks in the pool: with the
    // default threshold (0) every synchronisation hands the memory back to the driver and the next
    // call pays for mapping gigabytes again (measured: ~100 ms per 2^24-point host-pointer MSM).
    {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
    }
#endif
    API_CK(aleo::ntt_upload_constants());
    API_CK(aleo::msm_upload_constants());
    API_CK(aleo::util_upload_constants());
    API_CK(aleo::poly_upload_constants());
    g_dev_ready[dev] = true;
  }
  if (dev_out) *dev_out = dev;
  return ALEO_B200_OK;
}

int thread_stream(int dev, cudaStream_t* out) {
  if (t_stream == nullptr || t_stream_dev != dev) {
    API_CK(cudaStreamCreateWithFlags(&t_stream, cudaStreamNonBlocking));
    t_stream_dev = dev;
  }
  *out = t_stream;
  return ALEO_B200_OK;
}

bool stride_ok(size_t s) { return s == 96 || s == 104; }

}  // namespace

extern "C" {

const char* aleo_b200_version(void) {
#ifdef ALEO_EMU
  return "aleo_b200 0.1.0 (EMULATOR BUILD - development tooling, not a product)";
#else
  return "aleo_b200 0.1.0 (sm_100a)";
#endif
}

const char* aleo_b200_strerror(int code) {
  switch (code) {
    case ALEO_B200_OK: return "success";
    case ALEO_B200_EINVAL: return "invalid argument";
    case ALEO_B200_ETOOLARGE: return "problem too large for one call";
    case ALEO_B200_ENODEVICE: return "no usable B200 (sm_100) device; this library has no CPU fallback";
    case ALEO_B200_ECUDA: return "CUDA failure";
    case ALEO_B200_ENOMEM: return "out of memory";
  }
  return "unknown error";
}

const char* aleo_b200_last_cuda_error(void) { return t_last_cuda_error.c_str(); }

int aleo_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int aleo_b200_init(int device) {
  if (cudaSetDevice(device) != cudaSuccess) {
    t_last_cuda_error = "cudaSetDevice failed";
    return ALEO_B200_ENODEVICE;
  }
  return ensure_ready(nullptr);
}

int aleo_b200_shutdown(void) {
  aleo::ntt_clear_plans();
#ifndef ALEO_EMU
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    cudaDeviceSynchronize();
    cudaMemPoolTrimTo(pool, 0);
  }
#endif
  return ALEO_B200_OK;
}

// ------------------------------------------------------------------------------------------ NTT
int aleo_b200_ntt_launches(uint32_t log_n) { return aleo::ntt_launches(log_n); }

int aleo_b200_ntt_fr_dev(void* inout_dev, uint32_t log_n, size_t batch, int direction, int kind, void* stream) {
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  if (batch == 0) return ALEO_B200_OK;
  if (inout_dev == nullptr) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  API_CK(aleo::ntt_transform(dev, log_n, batch, direction == ALEO_B200_NTT_INVERSE, kind == ALEO_B200_NTT_COSET, inout_dev,
                             (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_fr_ordered_dev(void* inout_dev, uint32_t log_n, size_t batch, int direction, int kind, int order, void* stream) {
  if (order != ALEO_B200_NTT_ORDER_II && order != ALEO_B200_NTT_ORDER_IO && order != ALEO_B200_NTT_ORDER_OI) return ALEO_B200_EINVAL;
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (batch == 0) return ALEO_B200_OK;
  if (inout_dev == nullptr) return ALEO_B200_EINVAL;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  int rc0 = ensure_ready(nullptr);
  if (rc0) return rc0;
  if (order == ALEO_B200_NTT_ORDER_OI) API_CK(aleo::ntt_bitrev(log_n, batch, inout_dev, (cudaStream_t)stream));
  int rc = aleo_b200_ntt_fr_dev(inout_dev, log_n, batch, direction, kind, stream);
  if (rc) return rc;
  if (order == ALEO_B200_NTT_ORDER_IO) API_CK(aleo::ntt_bitrev(log_n, batch, inout_dev, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_fr_dev_profile(void* inout_dev, uint32_t log_n, int direction, int kind, void* stream, float* pass_ms4) {
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  if (inout_dev == nullptr || pass_ms4 == nullptr || log_n == 0) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  API_CK(aleo::ntt_transform(dev, log_n, 1, direction == ALEO_B200_NTT_INVERSE, kind == ALEO_B200_NTT_COSET, inout_dev,
                             (cudaStream_t)stream, pass_ms4));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_twiddle_dev(void* data_dev, uint32_t log_n_global, int direction, uint32_t rows, uint32_t cols, uint32_t row0,
                              uint32_t col0, void* stream) {
  if (log_n_global > (uint32_t)aleo::ntt_max_log_n() || log_n_global < 12) return ALEO_B200_ETOOLARGE;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (data_dev == nullptr) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  API_CK(aleo::ntt_twiddle_matrix(dev, log_n_global, direction == ALEO_B200_NTT_INVERSE, data_dev, rows, cols, row0, col0,
                                  (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_layout(uint32_t log_n, int world, uint32_t* log_r_first_out, uint32_t* log_r_last_out, int* passes_out) {
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  return aleo::ntt_dist_layout(log_n, world, log_r_first_out, log_r_last_out, passes_out) == 0 ? ALEO_B200_OK : ALEO_B200_EINVAL;
}

int aleo_b200_ntt_dist_create(void** ctx_out, uint32_t log_n, int rank, int world) {
  if (ctx_out == nullptr) return ALEO_B200_EINVAL;
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (aleo::ntt_dist_layout(log_n, world, nullptr, nullptr, nullptr) != 0 || rank < 0 || rank >= world) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::ntt_dist_create(log_n, rank, world, ctx_out));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_handles(void* ctx, void* handles128_out) {
  if (ctx == nullptr || handles128_out == nullptr) return ALEO_B200_EINVAL;
  API_CK(aleo::ntt_dist_handles(ctx, handles128_out));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_open(void* ctx, const void* all_handles) {
  if (ctx == nullptr || all_handles == nullptr) return ALEO_B200_EINVAL;
  API_CK(aleo::ntt_dist_open(ctx, all_handles));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_stage1(void* ctx, const void* local_in_dev, int direction, int kind, void* stream) {
  if (ctx == nullptr || local_in_dev == nullptr) return ALEO_B200_EINVAL;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::ntt_dist_stage1(ctx, local_in_dev, direction == ALEO_B200_NTT_INVERSE, kind == ALEO_B200_NTT_COSET, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_stage2(void* ctx, void* local_out_dev, int direction, int kind, void* stream) {
  if (ctx == nullptr || local_out_dev == nullptr) return ALEO_B200_EINVAL;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::ntt_dist_stage2(ctx, local_out_dev, direction == ALEO_B200_NTT_INVERSE, kind == ALEO_B200_NTT_COSET, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_ntt_dist_destroy(void* ctx) {
  aleo::ntt_dist_destroy(ctx);
  return ALEO_B200_OK;
}

int aleo_b200_ntt_fr(void* inout_host, uint32_t log_n, int direction, int kind) {
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (inout_host == nullptr) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  const size_t bytes = (size_t)32 << log_n;
  void* d = nullptr;
  API_CK(cudaMallocAsync(&d, bytes, s));
  cudaEvent_t ready = nullptr;
  cudaError_t e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ready, s);
  if (e == cudaSuccess) e = aleo::feed_h2d(d, inout_host, bytes, s, ready);  // pageable callers (a Rust Vec) are staged
  if (e == cudaSuccess) {
    rc = aleo_b200_ntt_fr_dev(d, log_n, 1, direction, kind, (void*)s);
    if (rc == ALEO_B200_OK) e = aleo::feed_d2h_sync(inout_host, d, bytes, s);
  }
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  if (ready) cudaEventDestroy(ready);
  if (rc) return rc;
  if (e != cudaSuccess) return fail_cuda(e);
  if (e2 != cudaSuccess) return fail_cuda(e2);
  return ALEO_B200_OK;
}

// ------------------------------------------------------------------------------------------ MSM
int aleo_b200_msm_window_bits(size_t n) { return aleo::msm_window_bits(n); }

int aleo_b200_msm_launches(size_t n) {
  int launches = 0;
  aleo::msm_run(nullptr, 96, nullptr, n, nullptr, nullptr, true, &launches);
  return launches;
}

int aleo_b200_msm_host_plan(size_t n, int* ranges_out, int* window_bits_out) {
  if (!aleo::msm_size_supported(n)) return ALEO_B200_ETOOLARGE;
  if (ranges_out) *ranges_out = n ? aleo::msm_host_chunks(n) : 1;
  if (window_bits_out) *window_bits_out = aleo::msm_host_window_bits(n);
  return ALEO_B200_OK;
}

int aleo_b200_msm_g1_dev(void* out_projective_dev, const void* bases_dev, size_t n, const void* scalars_dev,
                         size_t affine_stride, void* stream) {
  if (!stride_ok(affine_stride) || out_projective_dev == nullptr) return ALEO_B200_EINVAL;
  if (n > 0 && (bases_dev == nullptr || scalars_dev == nullptr)) return ALEO_B200_EINVAL;
  if (!aleo::msm_size_supported(n)) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::msm_run(bases_dev, (u32)affine_stride, scalars_dev, n, out_projective_dev, (cudaStream_t)stream, false,
                       nullptr));
  return ALEO_B200_OK;
}

int aleo_b200_msm_g1(void* out_projective_host, const void* bases_host, size_t n, const void* scalars_host,
                     size_t affine_stride) {
  if (!stride_ok(affine_stride) || out_projective_host == nullptr) return ALEO_B200_EINVAL;
  if (n > 0 && (bases_host == nullptr || scalars_host == nullptr)) return ALEO_B200_EINVAL;
  if (!aleo::msm_size_supported(n)) return ALEO_B200_ETOOLARGE;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  API_CK(aleo::msm_run_host(bases_host, (u32)affine_stride, scalars_host, n, out_projective_host, s));
  return ALEO_B200_OK;
}

int aleo_b200_msm_g1_multi(void* out_projective_host, const void* bases_host, size_t n, const void* scalars_host,
                           size_t affine_stride, int n_devices) {
  if (n_devices < 1 || n_devices > 64) return ALEO_B200_EINVAL;
  if (n_devices == 1 || n < (size_t)n_devices * 1024) return aleo_b200_msm_g1(out_projective_host, bases_host, n, scalars_host, affine_stride);
  if (!stride_ok(affine_stride) || out_projective_host == nullptr || bases_host == nullptr || scalars_host == nullptr) return ALEO_B200_EINVAL;
  if (n_devices > aleo_b200_device_count()) return ALEO_B200_ENODEVICE;
#ifdef ALEO_EMU
  return ALEO_B200_ENODEVICE;
#else
  int home = 0;
  cudaGetDevice(&home);
  std::vector<unsigned char> partials((size_t)n_devices * 144);
  std::vector<int> rcs(n_devices, ALEO_B200_OK);
  std::vector<std::thread> workers;
  const size_t base = n / n_devices, rem = n % n_devices;
  size_t first = 0;
  for (int d = 0; d < n_devices; d++) {
    const size_t cnt = base + ((size_t)d < rem ? 1 : 0);
    const unsigned char* b = (const unsigned char*)bases_host + first * affine_stride;
    const unsigned char* sc = (const unsigned char*)scalars_host + first * 32;
    unsigned char* out = partials.data() + (size_t)d * 144;
    workers.emplace_back([=, &rcs]() {
      int rc = aleo_b200_init(d);  // selects device d for this thread and prepares it
      if (rc == ALEO_B200_OK) rc = aleo_b200_msm_g1(out, b, cnt, sc, affine_stride);
      rcs[d] = rc;
    });
    first += cnt;
  }
  for (auto& w : workers) w.join();
  cudaSetDevice(home);
  for (int rc : rcs)
    if (rc) return rc;
  // single final combine on the calling thread's device
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  unsigned char* dbuf = nullptr;
  API_CK(cudaMallocAsync((void**)&dbuf, partials.size() + 256, s));
  cudaError_t e = cudaMemcpyAsync(dbuf, partials.data(), partials.size(), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = aleo::g1_sum(dbuf, (u32)n_devices, dbuf + partials.size(), s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_projective_host, dbuf + partials.size(), 144, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(dbuf, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail_cuda(e);
  if (e2 != cudaSuccess) return fail_cuda(e2);
  return ALEO_B200_OK;
#endif
}

int aleo_b200_msm_g1_dev_profile(void* out_projective_dev, const void* bases_dev, size_t n, const void* scalars_dev,
                                 size_t affine_stride, void* stream, float* phase_ms3) {
  if (!stride_ok(affine_stride) || out_projective_dev == nullptr || phase_ms3 == nullptr || n == 0) return ALEO_B200_EINVAL;
  if (bases_dev == nullptr || scalars_dev == nullptr) return ALEO_B200_EINVAL;
  if (!aleo::msm_size_supported(n)) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::msm_run(bases_dev, (u32)affine_stride, scalars_dev, n, out_projective_dev, (cudaStream_t)stream, false,
                       nullptr, phase_ms3));
  return ALEO_B200_OK;
}

int aleo_b200_g1_sum_dev(void* out_projective_dev, const void* points_dev, size_t count, void* stream) {
  if (out_projective_dev == nullptr || (count && points_dev == nullptr) || count >= ((size_t)1 << 32))
    return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::g1_sum(points_dev, (u32)count, out_projective_dev, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

// ------------------------------------------------------------------- generators / checks / bench
int aleo_b200_gen_bases_dev(void* bases_dev, size_t n, size_t affine_stride, const void* s0_host, const void* d_host,
                            uint64_t first_index, void* stream) {
  if (!stride_ok(affine_stride) || s0_host == nullptr || d_host == nullptr) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  if (bases_dev == nullptr || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::gen_bases(bases_dev, n, (u32)affine_stride, s0_host, d_host, first_index, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_gen_scalars_dev(void* scalars_dev, size_t n, uint64_t seed, uint64_t first_index, int montgomery,
                              void* stream) {
  if (n == 0) return ALEO_B200_OK;
  if (scalars_dev == nullptr || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::gen_scalars(scalars_dev, n, seed, first_index, montgomery != 0, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_dlog_dot_dev(void* out_scalar_dev, const void* scalars_dev, size_t n, const void* s0_host,
                           const void* d_host, uint64_t first_index, void* stream) {
  if (out_scalar_dev == nullptr || (n && scalars_dev == nullptr) || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  if (s0_host == nullptr || d_host == nullptr) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::dlog_dot(out_scalar_dev, scalars_dev, n, s0_host, d_host, first_index, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_check_on_curve_dev(const void* bases_dev, size_t n, size_t affine_stride, void* stream) {
  if (!stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n == 0) return 1;
  if (bases_dev == nullptr || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  int ok = 0;
  API_CK(aleo::check_on_curve(bases_dev, n, (u32)affine_stride, (cudaStream_t)stream, &ok));
  return ok;
}

int aleo_b200_field_op_dev(int field, int op, void* out_dev, const void* a_dev, const void* b_dev, size_t n, void* stream) {
  if (field != ALEO_B200_FIELD_FR && field != ALEO_B200_FIELD_FQ) return ALEO_B200_EINVAL;
  if (op < ALEO_B200_OP_ADD || op > ALEO_B200_OP_NEG) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  const bool binary = op == ALEO_B200_OP_ADD || op == ALEO_B200_OP_SUB || op == ALEO_B200_OP_MUL;
  if (out_dev == nullptr || a_dev == nullptr || (binary && b_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::field_op(field, op, out_dev, a_dev, b_dev, n, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_lagrange_coeffs_dev(void* out_dev, uint32_t log_n, const void* tau_host, void* stream) {
  if (out_dev == nullptr || tau_host == nullptr) return ALEO_B200_EINVAL;
  if (log_n > 31) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_lagrange_coeffs(out_dev, log_n, tau_host, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_axpy_dev(void* y_inout_dev, const void* x_dev, const void* a_host, size_t n, void* stream) {
  if (a_host == nullptr) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  if (y_inout_dev == nullptr || x_dev == nullptr) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_axpy(y_inout_dev, x_dev, a_host, n, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_divide_by_vanishing_on_coset_dev(void* evals_inout_dev, uint32_t log_m, uint32_t log_n, const void* g_host,
                                                  void* stream) {
  if (evals_inout_dev == nullptr || g_host == nullptr) return ALEO_B200_EINVAL;
  if (log_n > log_m) return ALEO_B200_EINVAL;
  if (log_m > 31) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_divide_by_vanishing_on_coset(evals_inout_dev, log_m, log_n, g_host, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_distribute_powers_dev(void* inout_dev, size_t n, const void* g_host, const void* k_host, void* stream) {
  if (g_host == nullptr) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  if (inout_dev == nullptr) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_distribute_powers(inout_dev, n, g_host, k_host, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_poly_eval_dev(void* out_dev, const void* coeffs_dev, size_t n, const void* z_host, void* stream) {
  if (out_dev == nullptr || z_host == nullptr || (n && coeffs_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_poly_eval(out_dev, coeffs_dev, n, z_host, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_fr_divide_by_linear_dev(void* quotient_dev, const void* coeffs_dev, size_t n, const void* z_host, void* stream) {
  if (z_host == nullptr) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  if (quotient_dev == nullptr || coeffs_dev == nullptr || quotient_dev == coeffs_dev) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fr_divide_by_linear(quotient_dev, coeffs_dev, n, z_host, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_kzg_open_dev(const void* handle, void* out_compressed48_dev, const void* coeffs_montgomery_dev, size_t n_coeffs,
                           const void* z_host, void* stream) {
  if (handle == nullptr || out_compressed48_dev == nullptr || z_host == nullptr || (n_coeffs && coeffs_montgomery_dev == nullptr))
    return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_coeffs <= 1) return aleo_b200_kzg_commit_dev(handle, out_compressed48_dev, coeffs_montgomery_dev, 0, stream);  // q = 0
  unsigned char* q = nullptr;
  API_CK(cudaMallocAsync((void**)&q, n_coeffs * 32, s));
  cudaError_t e = aleo::fr_divide_by_linear(q, coeffs_montgomery_dev, n_coeffs, z_host, s);
  int rc2 = ALEO_B200_OK;
  if (e == cudaSuccess) rc2 = aleo_b200_kzg_commit_dev(handle, out_compressed48_dev, q, n_coeffs - 1, stream);
  cudaFreeAsync(q, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return rc2;
}

static int g1_decompress_any(void* out_affine_dev, size_t affine_stride, const void* in48_dev, size_t n, void* stream, bool unchecked) {
  if (!stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n == 0) return 0;
  if (out_affine_dev == nullptr || in48_dev == nullptr || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  u32 bad = 0;
  API_CK(aleo::g1_decompress(in48_dev, n, out_affine_dev, (u32)affine_stride, (cudaStream_t)stream, &bad, unchecked));
  return bad > 0x7fffffffu ? 0x7fffffff : (int)bad;
}

int aleo_b200_g1_decompress_dev(void* out_affine_dev, size_t affine_stride, const void* in48_dev, size_t n, void* stream) {
  return g1_decompress_any(out_affine_dev, affine_stride, in48_dev, n, stream, false);
}

int aleo_b200_g1_decompress_unchecked_dev(void* out_affine_dev, size_t affine_stride, const void* in48_dev, size_t n, void* stream) {
  return g1_decompress_any(out_affine_dev, affine_stride, in48_dev, n, stream, true);
}

int aleo_b200_g1_compress_dev(void* out48_dev, const void* affine_dev, size_t affine_stride, size_t n, void* stream) {
  if (!stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n == 0) return ALEO_B200_OK;
  if (out48_dev == nullptr || affine_dev == nullptr || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::g1_compress_affine(affine_dev, (u32)affine_stride, n, out48_dev, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_bench_imad(int kind, int iters, double* ms_out, double* ops_out) {
  if (kind < 0 || kind > 5 || iters <= 0 || ms_out == nullptr || ops_out == nullptr) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::bench_imad(kind, iters, ms_out, ops_out));
  return ALEO_B200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------- resident SRS / KZG commit
extern "C" {

int aleo_b200_srs_create_dev(void** handle_out, const void* bases_dev, size_t n, size_t affine_stride, void* stream) {
  if (handle_out == nullptr || bases_dev == nullptr || n == 0 || !stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n >= ((size_t)1 << 28)) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::srs_create(bases_dev, (u32)affine_stride, n, (cudaStream_t)stream, handle_out));
  return ALEO_B200_OK;
}

int aleo_b200_srs_create(void** handle_out, const void* bases_host, size_t n, size_t affine_stride) {
  if (handle_out == nullptr || bases_host == nullptr || n == 0 || !stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n >= ((size_t)1 << 28)) return ALEO_B200_ETOOLARGE;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  void* d = nullptr;
  API_CK(cudaMallocAsync(&d, n * affine_stride, s));
  cudaError_t e = cudaMemcpyAsync(d, bases_host, n * affine_stride, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = aleo::srs_create(d, (u32)affine_stride, n, s, handle_out);
  cudaFreeAsync(d, s);
  cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
}

int aleo_b200_srs_destroy(void* handle) {
  aleo::srs_destroy(handle);
  return ALEO_B200_OK;
}

int aleo_b200_srs_info(const void* handle, size_t* n_out, int* window_bits_out, int* windows_out, size_t* bytes_out) {
  if (handle == nullptr) return ALEO_B200_EINVAL;
  aleo::srs_info(handle, n_out, window_bits_out, windows_out, bytes_out);
  return ALEO_B200_OK;
}

int aleo_b200_srs_msm_launches(const void* handle, size_t n_used) {
  if (handle == nullptr) return ALEO_B200_EINVAL;
  int launches = 0;
  if (aleo::srs_msm(handle, nullptr, n_used, nullptr, nullptr, true, &launches, nullptr) != cudaSuccess) return ALEO_B200_EINVAL;
  return launches;
}

int aleo_b200_srs_msm_dev(const void* handle, void* out_projective_dev, const void* scalars_dev, size_t n_used, void* stream) {
  if (handle == nullptr || out_projective_dev == nullptr || (n_used && scalars_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::srs_msm(handle, scalars_dev, n_used, out_projective_dev, (cudaStream_t)stream, false, nullptr, nullptr));
  return ALEO_B200_OK;
}

int aleo_b200_srs_msm_dev_profile(const void* handle, void* out_projective_dev, const void* scalars_dev, size_t n_used,
                                  void* stream, float* phase_ms3) {
  if (handle == nullptr || out_projective_dev == nullptr || scalars_dev == nullptr || n_used == 0 || phase_ms3 == nullptr)
    return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::srs_msm(handle, scalars_dev, n_used, out_projective_dev, (cudaStream_t)stream, false, nullptr, phase_ms3));
  return ALEO_B200_OK;
}

// host scalars -> device, run `body` on the staged copy, copy `out_bytes` of result back
static int srs_host_call(const void* handle, void* out_host, size_t out_bytes, const void* in_host, size_t n, bool montgomery_in) {
  if (handle == nullptr || out_host == nullptr || (n && in_host == nullptr)) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  API_CK(aleo::srs_msm_host(handle, in_host, n, montgomery_in, out_host, out_bytes, s));
  return ALEO_B200_OK;
}

int aleo_b200_srs_msm(const void* handle, void* out_projective_host, const void* scalars_host, size_t n_used) {
  return srs_host_call(handle, out_projective_host, 144, scalars_host, n_used, false);
}

int aleo_b200_kzg_commit(const void* handle, void* out_compressed48_host, const void* coeffs_montgomery_host, size_t n_coeffs) {
  return srs_host_call(handle, out_compressed48_host, 48, coeffs_montgomery_host, n_coeffs, true);
}

int aleo_b200_kzg_commit_hiding_dev(const void* handle_beta, const void* handle_beta_gamma, void* out_compressed48_dev,
                                    const void* coeffs_montgomery_dev, size_t n_coeffs, const void* random_coeffs_montgomery_dev,
                                    size_t n_random, void* stream) {
  if (handle_beta == nullptr || handle_beta_gamma == nullptr || out_compressed48_dev == nullptr) return ALEO_B200_EINVAL;
  if ((n_coeffs && coeffs_montgomery_dev == nullptr) || (n_random && random_coeffs_montgomery_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* d = nullptr;
  const size_t sb = ((n_coeffs > n_random ? n_coeffs : n_random) * 32 + 255) & ~(size_t)255;
  API_CK(cudaMallocAsync((void**)&d, sb + 512, s));
  unsigned char* parts = d + sb;  // two 144-byte partial results, then the sum
  cudaError_t e = aleo::fr_to_bigint(coeffs_montgomery_dev, d, n_coeffs, s);
  if (e == cudaSuccess) e = aleo::srs_msm(handle_beta, d, n_coeffs, parts, s, false, nullptr, nullptr);
  if (e == cudaSuccess) e = aleo::fr_to_bigint(random_coeffs_montgomery_dev, d, n_random, s);
  if (e == cudaSuccess) e = aleo::srs_msm(handle_beta_gamma, d, n_random, parts + 144, s, false, nullptr, nullptr);
  if (e == cudaSuccess) e = aleo::g1_sum(parts, 2, parts + 288, s);
  if (e == cudaSuccess) e = aleo::g1_compress(parts + 288, out_compressed48_dev, s);
  cudaFreeAsync(d, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
}

int aleo_b200_kzg_commit_batch_dev(const void* handle, void* out_compressed48_dev, const void* const* coeffs_dev_ptrs_host,
                                   const size_t* n_coeffs_host, size_t count, void* stream) {
  if (handle == nullptr) return ALEO_B200_EINVAL;
  if (count == 0) return ALEO_B200_OK;
  if (out_compressed48_dev == nullptr || coeffs_dev_ptrs_host == nullptr || n_coeffs_host == nullptr || count > 64) return ALEO_B200_EINVAL;
  for (size_t m = 0; m < count; m++)
    if (n_coeffs_host[m] && coeffs_dev_ptrs_host[m] == nullptr) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::srs_msm_batch(handle, coeffs_dev_ptrs_host, n_coeffs_host, count, true, out_compressed48_dev, true, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_kzg_commit_dev(const void* handle, void* out_compressed48_dev, const void* coeffs_montgomery_dev, size_t n_coeffs,
                             void* stream) {
  if (handle == nullptr || out_compressed48_dev == nullptr || (n_coeffs && coeffs_montgomery_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* d = nullptr;
  const size_t sb = (n_coeffs * 32 + 255) & ~(size_t)255;
  API_CK(cudaMallocAsync((void**)&d, sb + 256, s));
  cudaError_t e = aleo::fr_to_bigint(coeffs_montgomery_dev, d, n_coeffs, s);
  if (e == cudaSuccess) e = aleo::srs_msm(handle, d, n_coeffs, d + sb, s, false, nullptr, nullptr);
  if (e == cudaSuccess) e = aleo::g1_compress(d + sb, out_compressed48_dev, s);
  cudaFreeAsync(d, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
}

}  // extern "C"
