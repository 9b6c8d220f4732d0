// capi.cu -- the C ABI of libaleo_b200.so (include/aleo_b200.h): argument checking, device
// readiness, per-thread streams and host<->device staging.  The kernels live in ntt_lib.cu,
// msm_lib.cu and util_lib.cu.
//
// No CPU fallback: if the calling thread's CUDA device is missing or is not sm_100, every entry
// point returns ALEO_B200_ENODEVICE.
#include "../../include/aleo_b200.h"
#include <cstdlib>
#ifndef ALEO_EMU
#include <condition_variable>
#include <functional>
#endif
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
thread_local aleo::ThreadStream t_stream;  // released when the calling thread ends

std::mutex g_dev_mu;
bool g_dev_ready[64] = {false};
bool g_dev_bad[64] = {false};

#ifndef ALEO_EMU
cudaMemPool_t g_pool[64] = {nullptr};
std::mutex g_pool_mu;
#endif

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
  (void)dev;  // the calling thread's current device (ensure_ready has just read it)
  API_CK(t_stream.get(out));
  return ALEO_B200_OK;
}

bool stride_ok(size_t s) { return s == 96 || s == 104; }

#ifndef ALEO_EMU
// One persistent host thread per device for the single-process multi-GPU entry point.  Created on first use and never
// destroyed (a leaked singleton: no ordering problems against the CUDA runtime's own teardown at process exit); calls
// are serialised.
class DeviceWorkers {
 public:
  static DeviceWorkers& get() {
    static DeviceWorkers* p = new DeviceWorkers();
    return *p;
  }
  // fn(d) for d = 0 .. nd - 1, each on device d's worker thread; returns when all of them are done
  void run(int nd, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> call(call_mu_);
    std::unique_lock<std::mutex> lk(mu_);
    while ((int)threads_.size() < nd) {
      const int d = (int)threads_.size();
      threads_.emplace_back([this, d]() { loop(d); });
      threads_.back().detach();
    }
    fn_ = &fn;
    active_ = nd;
    pending_ = nd;
    generation_++;
    cv_.notify_all();
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int d) {
    unsigned long long seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (d >= active_) continue;
        fn = fn_;
      }
      (*fn)(d);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void(int)>* fn_ = nullptr;
  int active_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
};
#endif

// a resident-SRS handle lives on the device it was created on: using it from a thread whose current device differs
// would fault on the device instead of failing here
int srs_ready(const void* handle) {
  if (handle == nullptr) return ALEO_B200_EINVAL;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  if (aleo::srs_device(handle) != dev) {
    t_last_cuda_error = "SRS handle belongs to device " + std::to_string(aleo::srs_device(handle)) + ", current device is " +
                        std::to_string(dev);
    return ALEO_B200_EINVAL;
  }
  return ALEO_B200_OK;
}

}  // namespace

namespace aleo {
cudaError_t pool_malloc_async(void** p, size_t bytes, cudaStream_t s) {
#ifdef ALEO_EMU
  return cudaMalloc(p, bytes);
#else
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool[dev] == nullptr) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      e = cudaMemPoolCreate(&g_pool[dev], &props);
      if (e != cudaSuccess) {
        g_pool[dev] = nullptr;
        return e;
      }
      const char* env = getenv("ALEO_B200_POOL_KEEP_MB");
      unsigned long long keep = (env && atoll(env) >= 0 ? (unsigned long long)atoll(env) : 65536ull) << 20;  // above the largest MSM workspace (sorted list + buckets + 40 GB of batch-affine levels at 2^26): a pool trimmed below what every call needs re-maps gigabytes per call (measured: 146 instead of 80 ms per 2^24 MSM)
      cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = g_pool[dev];
  }
  return cudaMallocFromPoolAsync(p, bytes, pool, s);
#endif
}
}  // namespace aleo

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
  // Must not race with calls in flight on the same device (a transform holds its plan while it enqueues work).
  aleo::ntt_clear_plans();
#ifndef ALEO_EMU
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool[dev]) cudaMemPoolTrimTo(g_pool[dev], 0);
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

int aleo_b200_ntt_dist_transform(void* ctx, const void* local_in_dev, void* local_out_dev, int direction, int kind, void* stream) {
  int rc = aleo_b200_ntt_dist_stage1(ctx, local_in_dev, direction, kind, stream);
  if (rc) return rc;
  return aleo_b200_ntt_dist_stage2(ctx, local_out_dev, direction, kind, stream);
}

int aleo_b200_ntt_dist_profile(void* ctx, const void* local_in_dev, void* local_out_dev, int direction, int kind, void* stream,
                               float* stage_ms4) {
  if (ctx == nullptr || local_in_dev == nullptr || local_out_dev == nullptr || stage_ms4 == nullptr) return ALEO_B200_EINVAL;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::ntt_dist_profile(ctx, local_in_dev, local_out_dev, direction == ALEO_B200_NTT_INVERSE, kind == ALEO_B200_NTT_COSET,
                                (cudaStream_t)stream, stage_ms4));
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
  API_CK(aleo::pool_malloc_async(&d, bytes, s));
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

// host-pointer transform with upstream's `order` argument (SURVEY.md App. E snarkvm_ntt): one H2D, the transform with
// its bit-reversal pass, one D2H
int aleo_b200_ntt_fr_ordered(void* inout_host, uint32_t log_n, int direction, int kind, int order) {
  if (order != ALEO_B200_NTT_ORDER_II && order != ALEO_B200_NTT_ORDER_IO && order != ALEO_B200_NTT_ORDER_OI) return ALEO_B200_EINVAL;
  if (direction != ALEO_B200_NTT_FORWARD && direction != ALEO_B200_NTT_INVERSE) return ALEO_B200_EINVAL;
  if (kind != ALEO_B200_NTT_STANDARD && kind != ALEO_B200_NTT_COSET) return ALEO_B200_EINVAL;
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
  API_CK(aleo::pool_malloc_async(&d, bytes, s));
  cudaEvent_t ready = nullptr;
  cudaError_t e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ready, s);
  if (e == cudaSuccess) e = aleo::feed_h2d(d, inout_host, bytes, s, ready);
  if (e == cudaSuccess) {
    rc = aleo_b200_ntt_fr_ordered_dev(d, log_n, 1, direction, kind, order, (void*)s);
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

// ---- polymul: out = ifft( prod_i fft(p_i) * prod_j e_j ) over the domain of size 2^log_n -----------------------------
// work: (pcount ? pcount : 1) * n elements of scratch (one batched forward transform for all polynomials)
static int polymul_check(const void* out, size_t pcount, const void* const* polys, const size_t* plens, size_t ecount,
                         const void* const* evals, const size_t* elens, uint32_t log_n) {
  if (log_n > (uint32_t)aleo::ntt_max_log_n()) return ALEO_B200_ETOOLARGE;
  if (out == nullptr || pcount + ecount == 0 || pcount > 64 || ecount > 64) return ALEO_B200_EINVAL;
  if ((pcount && (polys == nullptr || plens == nullptr)) || (ecount && (evals == nullptr || elens == nullptr))) return ALEO_B200_EINVAL;
  const size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < pcount; i++)
    if (plens[i] > n || (plens[i] && polys[i] == nullptr)) return ALEO_B200_EINVAL;
  for (size_t j = 0; j < ecount; j++)
    if (elens[j] != n || evals[j] == nullptr) return ALEO_B200_EINVAL;  // evaluations live on the whole domain
  return ALEO_B200_OK;
}

// polynomials already staged in `work` (pcount slots of n elements, zero padded); evals on the device
static cudaError_t polymul_core(int dev, unsigned char* out, unsigned char* work, size_t pcount, size_t ecount,
                                const void* const* evals_dev, uint32_t log_n, cudaStream_t s) {
  const size_t n = (size_t)1 << log_n, nb = n * 32;
  cudaError_t e = cudaSuccess;
  const unsigned char* acc = nullptr;
  if (pcount) {
    e = aleo::ntt_transform(dev, log_n, pcount, false, false, work, s);
    acc = work;
    for (size_t i = 1; i < pcount && e == cudaSuccess; i++) {
      e = aleo::field_op(ALEO_B200_FIELD_FR, ALEO_B200_OP_MUL, i + 1 == pcount && ecount == 0 ? out : work, acc, work + i * nb, n, s);
      acc = (i + 1 == pcount && ecount == 0) ? out : work;
    }
  }
  for (size_t j = 0; j < ecount && e == cudaSuccess; j++) {
    if (acc == nullptr) {
      if (ecount == 1) {
        e = cudaMemcpyAsync(out, evals_dev[0], nb, cudaMemcpyDeviceToDevice, s);
      } else {
        e = aleo::field_op(ALEO_B200_FIELD_FR, ALEO_B200_OP_MUL, out, evals_dev[0], evals_dev[1], n, s);
        j = 1;
      }
      acc = out;
      continue;
    }
    e = aleo::field_op(ALEO_B200_FIELD_FR, ALEO_B200_OP_MUL, out, acc, evals_dev[j], n, s);
    acc = out;
  }
  if (e == cudaSuccess && acc != out) e = cudaMemcpyAsync(out, acc, nb, cudaMemcpyDeviceToDevice, s);  // pcount == 1, ecount == 0
  if (e == cudaSuccess) e = aleo::ntt_transform(dev, log_n, 1, true, false, out, s);
  return e;
}

int aleo_b200_polymul_dev(void* out_dev, size_t pcount, const void* const* polynomials_dev_ptrs_host, const size_t* plens_host,
                          size_t ecount, const void* const* evaluations_dev_ptrs_host, const size_t* elens_host, uint32_t log_n,
                          void* stream) {
  int rc = polymul_check(out_dev, pcount, polynomials_dev_ptrs_host, plens_host, ecount, evaluations_dev_ptrs_host, elens_host, log_n);
  if (rc) return rc;
  int dev = 0;
  rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)1 << log_n, nb = n * 32;
  unsigned char* work = nullptr;
  if (pcount) API_CK(aleo::pool_malloc_async((void**)&work, pcount * nb, s));
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < pcount && e == cudaSuccess; i++) {
    if (plens_host[i]) e = cudaMemcpyAsync(work + i * nb, polynomials_dev_ptrs_host[i], plens_host[i] * 32, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && plens_host[i] < n) e = cudaMemsetAsync(work + i * nb + plens_host[i] * 32, 0, (n - plens_host[i]) * 32, s);
  }
  if (e == cudaSuccess) e = polymul_core(dev, (unsigned char*)out_dev, work, pcount, ecount, evaluations_dev_ptrs_host, log_n, s);
  if (work) cudaFreeAsync(work, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
}

int aleo_b200_polymul(void* out_host, size_t pcount, const void* const* polynomials_host, const size_t* plens, size_t ecount,
                      const void* const* evaluations_host, const size_t* elens, uint32_t log_n) {
  int rc = polymul_check(out_host, pcount, polynomials_host, plens, ecount, evaluations_host, elens, log_n);
  if (rc) return rc;
  int dev = 0;
  rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  const size_t n = (size_t)1 << log_n, nb = n * 32;
  unsigned char* d = nullptr;  // [out][pcount polynomial slots][ecount evaluation vectors]
  API_CK(aleo::pool_malloc_async((void**)&d, (1 + pcount + ecount) * nb, s));
  unsigned char *work = d + nb, *ev = d + (1 + pcount) * nb;
  cudaEvent_t ready = nullptr;
  cudaError_t e = cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ready, s);
  std::vector<const void*> ev_ptrs(ecount);
  for (size_t i = 0; i < pcount && e == cudaSuccess; i++) {
    e = aleo::feed_h2d(work + i * nb, polynomials_host[i], plens[i] * 32, s, ready);
    if (e == cudaSuccess && plens[i] < n) e = cudaMemsetAsync(work + i * nb + plens[i] * 32, 0, (n - plens[i]) * 32, s);
  }
  for (size_t j = 0; j < ecount && e == cudaSuccess; j++) {
    ev_ptrs[j] = ev + j * nb;
    e = aleo::feed_h2d(ev + j * nb, evaluations_host[j], nb, s, ready);
  }
  if (e == cudaSuccess) e = polymul_core(dev, d, work, pcount, ecount, ev_ptrs.data(), log_n, s);
  if (e == cudaSuccess) e = aleo::feed_d2h_sync(out_host, d, nb, s);
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  if (ready) cudaEventDestroy(ready);
  if (e != cudaSuccess) return fail_cuda(e);
  if (e2 != cudaSuccess) return fail_cuda(e2);
  return ALEO_B200_OK;
}

// ------------------------------------------------------------------------------------------ MSM
int aleo_b200_msm_window_bits(size_t n) { return aleo::msm_window_bits(n); }
int aleo_b200_msm_ba_levels(size_t n) { return aleo::msm_ba_levels(n); }

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
  const size_t base = n / n_devices, rem = n % n_devices;
  // one PERSISTENT worker thread per device (DeviceWorkers): its stream, staging buffers and helper threads are created
  // once and reused by every later call instead of being rebuilt (tens of MB of pinned memory per device) per call
  DeviceWorkers::get().run(n_devices, [&](int d) {
    size_t first = 0;
    for (int k = 0; k < d; k++) first += base + ((size_t)k < rem ? 1 : 0);
    const size_t cnt = base + ((size_t)d < rem ? 1 : 0);
    int rc = aleo_b200_init(d);  // selects device d for the worker thread and prepares it
    if (rc == ALEO_B200_OK)
      rc = aleo_b200_msm_g1(partials.data() + (size_t)d * 144, (const unsigned char*)bases_host + first * affine_stride, cnt,
                            (const unsigned char*)scalars_host + first * 32, affine_stride);
    rcs[d] = rc;
  });
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
  API_CK(aleo::pool_malloc_async((void**)&dbuf, partials.size() + 256, s));
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
  int rc = srs_ready(handle);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_coeffs <= 1) return aleo_b200_kzg_commit_dev(handle, out_compressed48_dev, coeffs_montgomery_dev, 0, stream);  // q = 0
  unsigned char* q = nullptr;
  API_CK(aleo::pool_malloc_async((void**)&q, n_coeffs * 32, s));
  cudaError_t e = aleo::fr_divide_by_linear(q, coeffs_montgomery_dev, n_coeffs, z_host, s);
  int rc2 = ALEO_B200_OK;
  if (e == cudaSuccess) rc2 = aleo_b200_kzg_commit_dev(handle, out_compressed48_dev, q, n_coeffs - 1, stream);
  cudaFreeAsync(q, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return rc2;
}

// SonicKZG10::open_combinations / kzg10 batch_open shape: m openings, opening k of the linear combination
// p_k = sum_i lc[k][i] * poly_i at the point z_k.  All m witness polynomials are committed in ONE launch sequence.
int aleo_b200_kzg_open_combinations_dev(const void* handle, void* out_compressed48_dev, const void* const* polys_dev_ptrs_host,
                                        const size_t* n_coeffs_host, size_t n_polys, const void* lc_coeffs_host,
                                        const void* points_host, size_t m, void* stream) {
  if (handle == nullptr) return ALEO_B200_EINVAL;
  if (m == 0) return ALEO_B200_OK;
  if (out_compressed48_dev == nullptr || lc_coeffs_host == nullptr || points_host == nullptr || m > 64 || n_polys == 0 || n_polys > 1024)
    return ALEO_B200_EINVAL;
  if (polys_dev_ptrs_host == nullptr || n_coeffs_host == nullptr) return ALEO_B200_EINVAL;
  size_t max_len = 0;
  for (size_t i = 0; i < n_polys; i++) {
    if (n_coeffs_host[i] && polys_dev_ptrs_host[i] == nullptr) return ALEO_B200_EINVAL;
    if (n_coeffs_host[i] > max_len) max_len = n_coeffs_host[i];
  }
  int rc = srs_ready(handle);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (max_len <= 1) {  // every combination is a constant: all witness polynomials are zero
    std::vector<const void*> none(m, nullptr);
    std::vector<size_t> zero(m, 0);
    API_CK(aleo::srs_msm_batch(handle, none.data(), zero.data(), m, true, out_compressed48_dev, true, s));
    return ALEO_B200_OK;
  }
  const size_t nb = (max_len * 32 + 255) & ~(size_t)255;
  unsigned char* d = nullptr;  // [combination][m witness polynomials]
  API_CK(aleo::pool_malloc_async((void**)&d, (1 + m) * nb, s));
  const unsigned char* lc = (const unsigned char*)lc_coeffs_host;
  const unsigned char* zs = (const unsigned char*)points_host;
  std::vector<const void*> wit(m);
  std::vector<size_t> wlen(m);
  cudaError_t e = cudaSuccess;
  for (size_t k = 0; k < m && e == cudaSuccess; k++) {
    e = cudaMemsetAsync(d, 0, max_len * 32, s);
    size_t len = 0;
    for (size_t i = 0; i < n_polys && e == cudaSuccess; i++) {
      const unsigned char* a = lc + (k * n_polys + i) * 32;
      bool is_zero = true;
      for (int b = 0; b < 32; b++) is_zero = is_zero && a[b] == 0;
      if (is_zero || n_coeffs_host[i] == 0) continue;
      e = aleo::fr_axpy(d, polys_dev_ptrs_host[i], a, n_coeffs_host[i], s);
      if (n_coeffs_host[i] > len) len = n_coeffs_host[i];
    }
    unsigned char* q = d + (1 + k) * nb;
    wit[k] = q;
    wlen[k] = len > 1 ? len - 1 : 0;
    if (e == cudaSuccess && len > 1) e = aleo::fr_divide_by_linear(q, d, len, zs + k * 32, s);
  }
  if (e == cudaSuccess) e = aleo::srs_msm_batch(handle, wit.data(), wlen.data(), m, true, out_compressed48_dev, true, s);
  cudaFreeAsync(d, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
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

int aleo_b200_fq_mul_fp64_dev(void* out_dev, const void* a_dev, const void* b_dev, size_t n, int square, void* stream) {
  if (n == 0) return ALEO_B200_OK;
  if (out_dev == nullptr || a_dev == nullptr || (!square && b_dev == nullptr) || n >= ((size_t)1 << 32)) return ALEO_B200_EINVAL;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::fq_mul_fp64(out_dev, a_dev, b_dev, n, square != 0, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_bench_imad(int kind, int iters, double* ms_out, double* ops_out) {
  if (kind < 0 || kind > 8 || iters <= 0 || ms_out == nullptr || ops_out == nullptr) return ALEO_B200_EINVAL;
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
  if (n >= ((size_t)1 << 28) || !aleo::srs_size_supported(n)) return ALEO_B200_ETOOLARGE;
  int rc = ensure_ready(nullptr);
  if (rc) return rc;
  API_CK(aleo::srs_create(bases_dev, (u32)affine_stride, n, (cudaStream_t)stream, handle_out));
  return ALEO_B200_OK;
}

int aleo_b200_srs_create(void** handle_out, const void* bases_host, size_t n, size_t affine_stride) {
  if (handle_out == nullptr || bases_host == nullptr || n == 0 || !stride_ok(affine_stride)) return ALEO_B200_EINVAL;
  if (n >= ((size_t)1 << 28) || !aleo::srs_size_supported(n)) return ALEO_B200_ETOOLARGE;
  int dev = 0;
  int rc = ensure_ready(&dev);
  if (rc) return rc;
  cudaStream_t s;
  rc = thread_stream(dev, &s);
  if (rc) return rc;
  void* d = nullptr;
  API_CK(aleo::pool_malloc_async(&d, n * affine_stride, s));
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
  int rc = srs_ready(handle);
  if (rc) return rc;
  API_CK(aleo::srs_msm(handle, scalars_dev, n_used, out_projective_dev, (cudaStream_t)stream, false, nullptr, nullptr));
  return ALEO_B200_OK;
}

int aleo_b200_srs_msm_dev_profile(const void* handle, void* out_projective_dev, const void* scalars_dev, size_t n_used,
                                  void* stream, float* phase_ms3) {
  if (handle == nullptr || out_projective_dev == nullptr || scalars_dev == nullptr || n_used == 0 || phase_ms3 == nullptr)
    return ALEO_B200_EINVAL;
  int rc = srs_ready(handle);
  if (rc) return rc;
  API_CK(aleo::srs_msm(handle, scalars_dev, n_used, out_projective_dev, (cudaStream_t)stream, false, nullptr, phase_ms3));
  return ALEO_B200_OK;
}

// host scalars -> device, run `body` on the staged copy, copy `out_bytes` of result back
static int srs_host_call(const void* handle, void* out_host, size_t out_bytes, const void* in_host, size_t n, bool montgomery_in) {
  if (handle == nullptr || out_host == nullptr || (n && in_host == nullptr)) return ALEO_B200_EINVAL;
  int rc = srs_ready(handle);
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
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
  int rc = srs_ready(handle_beta);
  if (rc) return rc;
  rc = srs_ready(handle_beta_gamma);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* d = nullptr;
  const size_t sb = ((n_coeffs > n_random ? n_coeffs : n_random) * 32 + 255) & ~(size_t)255;
  API_CK(aleo::pool_malloc_async((void**)&d, sb + 512, s));
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
  int rc = srs_ready(handle);
  if (rc) return rc;
  API_CK(aleo::srs_msm_batch(handle, coeffs_dev_ptrs_host, n_coeffs_host, count, true, out_compressed48_dev, true, (cudaStream_t)stream));
  return ALEO_B200_OK;
}

int aleo_b200_kzg_commit_dev(const void* handle, void* out_compressed48_dev, const void* coeffs_montgomery_dev, size_t n_coeffs,
                             void* stream) {
  if (handle == nullptr || out_compressed48_dev == nullptr || (n_coeffs && coeffs_montgomery_dev == nullptr)) return ALEO_B200_EINVAL;
  int rc = srs_ready(handle);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* d = nullptr;
  const size_t sb = (n_coeffs * 32 + 255) & ~(size_t)255;
  API_CK(aleo::pool_malloc_async((void**)&d, sb + 256, s));
  cudaError_t e = aleo::fr_to_bigint(coeffs_montgomery_dev, d, n_coeffs, s);
  if (e == cudaSuccess) e = aleo::srs_msm(handle, d, n_coeffs, d + sb, s, false, nullptr, nullptr);
  if (e == cudaSuccess) e = aleo::g1_compress(d + sb, out_compressed48_dev, s);
  cudaFreeAsync(d, s);
  if (e != cudaSuccess) return fail_cuda(e);
  return ALEO_B200_OK;
}

}  // extern "C"
