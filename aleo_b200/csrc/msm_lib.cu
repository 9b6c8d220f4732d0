// msm_lib.cu -- translation unit holding the G1 MSM kernels.
#include "internal.h"
#include "msm_host.cuh"

namespace aleo {
cudaError_t msm_upload_constants() { return aleo_upload_field_constants(); }
bool msm_size_supported(size_t n) { return msm::size_supported(n); }
int msm_window_bits(size_t n) { return (int)msm::make_params(n).c; }

// ---- host-pointer calls: the MSM arrives in point ranges so that the copy of range k + 1 overlaps the
// accumulation of range k (PCIe moves 136 bytes per point about 2.5x faster than the integer pipe adds it) ----
namespace {
thread_local cudaStream_t t_copy_stream = nullptr;
thread_local int t_copy_dev = -1;

cudaError_t copy_stream(cudaStream_t* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (t_copy_stream == nullptr || t_copy_dev != dev) {
    e = cudaStreamCreateWithFlags(&t_copy_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    t_copy_dev = dev;
  }
  *out = t_copy_stream;
  return cudaSuccess;
}

// Point ranges of a host-pointer MSM.  The first range is small (its copy is the only one not hidden) and the
// later ones grow: a range's accumulation must outlast the copy of the next one.
int chunk_schedule(size_t n, size_t* sizes) {
  const char* env = getenv("ALEO_B200_MSM_CHUNKS");  // read per call: tests and sweeps switch it
  const long k_env = env ? atol(env) : 0L;
  int k = n < ((size_t)1 << 19) ? 1 : (n < ((size_t)1 << 22) ? 2 : 3);
  if (k_env >= 1 && k_env <= 3 && n >= 8) k = (int)k_env;
  if (k == 1) {
    sizes[0] = n;
  } else if (k == 2) {
    sizes[0] = n / 4;
    sizes[1] = n - sizes[0];
  } else {
    sizes[0] = n / 8;
    sizes[1] = (n * 3) / 8;
    sizes[2] = n - sizes[0] - sizes[1];
  }
  return k;
}

struct EventSet {
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaError_t init() {
    for (auto& e : ev) {
      cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (r != cudaSuccess) return r;
    }
    return cudaSuccess;
  }
  ~EventSet() {
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};
}  // namespace

int msm_host_chunks(size_t n) {
  size_t sizes[4];
  return chunk_schedule(n, sizes);
}

int msm_host_window_bits(size_t n) { return (int)msm::make_params(n, nullptr, (u32)msm_host_chunks(n)).c; }

// out_host: 144 bytes.  Synchronises `s` before returning.
cudaError_t msm_run_host(const void* bases_host, u32 stride, const void* scalars_host, size_t n, void* out_host, cudaStream_t s) {
  unsigned char* d = nullptr;
  const size_t bb = n * stride, sb = n * 32;
  const size_t o_s = (bb + 255) & ~(size_t)255, o_out = o_s + ((sb + 255) & ~(size_t)255);
  MSM_CK(cudaMallocAsync((void**)&d, o_out + 256, s));
  cudaError_t e = cudaSuccess;
  if (n == 0) {
    e = msm::run(nullptr, stride, nullptr, 0, d + o_out, s, false, nullptr);
  } else {
    size_t sizes[4];
    const int k = chunk_schedule(n, sizes);
    cudaStream_t cs = nullptr;
    EventSet evs;
    e = copy_stream(&cs);
    if (e == cudaSuccess) e = evs.init();
    msm::Session ss;
    if (e == cudaSuccess) e = cudaEventRecord(evs.ev[4], s);          // the allocation is ordered on s
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, evs.ev[4], 0);
    size_t first = 0;
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      const size_t m = sizes[i];
      e = cudaMemcpyAsync(d + o_s + first * 32, (const unsigned char*)scalars_host + first * 32, m * 32, cudaMemcpyHostToDevice, cs);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(d + first * stride, (const unsigned char*)bases_host + first * stride, m * stride,
                            cudaMemcpyHostToDevice, cs);
      if (e == cudaSuccess) e = cudaEventRecord(evs.ev[i], cs);
      first += m;
    }
    size_t max_chunk = 0;
    for (int i = 0; i < k; i++) max_chunk = sizes[i] > max_chunk ? sizes[i] : max_chunk;
    if (e == cudaSuccess) e = ss.begin(n, max_chunk, (u32)k, nullptr, s, false);
    first = 0;
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      e = cudaStreamWaitEvent(s, evs.ev[i], 0);
      if (e == cudaSuccess)
        e = ss.add_chunk(d + first * stride, stride, (const u32*)(d + o_s + first * 32), sizes[i], first, s);
      first += sizes[i];
    }
    if (e == cudaSuccess) e = ss.finish(d + o_out, s);
    else ss.release(s);
    if (e != cudaSuccess) cudaStreamSynchronize(cs);  // nothing may still be writing into d when it is freed
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d + o_out, 144, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

cudaError_t msm_run(const void* bases_dev, u32 stride, const void* scalars_dev, size_t n, void* out144_dev, cudaStream_t s,
                    bool dry, int* launches_out, float* phase_ms) {
  if (!phase_ms || dry || n == 0)
    return msm::run((const unsigned char*)bases_dev, stride, (const u32*)scalars_dev, n, (unsigned char*)out144_dev, s, dry,
                    launches_out);
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]);
  cudaError_t e = msm::run((const unsigned char*)bases_dev, stride, (const u32*)scalars_dev, n, (unsigned char*)out144_dev,
                           s, false, launches_out, ev);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  for (int i = 0; i < 3; i++) {
    phase_ms[i] = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  return e;
}

// ---- resident SRS ---------------------------------------------------------------------------------
struct Srs {
  int device;
  unsigned char* pre;   // [W][n] packed 96-byte affine, pre[w][i] = 2^(c w) * P_i
  size_t n;
  u32 c, W;
};

cudaError_t srs_create(const void* bases_dev, u32 stride, size_t n, cudaStream_t s, void** handle_out) {
  if (n == 0 || n >= ((size_t)1 << 31)) return cudaErrorInvalidValue;
  Srs* h = new Srs();
  cudaGetDevice(&h->device);
  h->n = n;
  h->c = msm::choose_window_srs(n);
  h->W = msm::windows_for_srs(h->c);
  if (h->W > msm::SRS_MAX_WINDOWS || (unsigned long long)n * h->W >= (1ull << 31)) {
    delete h;
    return cudaErrorInvalidValue;
  }
  cudaError_t e = cudaMalloc((void**)&h->pre, n * h->W * 96);
  if (e != cudaSuccess) {
    delete h;
    return e;
  }
  LAUNCH_NOSYNC(msm::srs_expand_kernel, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)bases_dev, stride,
                (u32)n, h->c, h->W, h->pre);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    cudaFree(h->pre);
    delete h;
    return e;
  }
  *handle_out = h;
  return cudaSuccess;
}

void srs_destroy(void* handle) {
  Srs* h = (Srs*)handle;
  if (!h) return;
  cudaFree(h->pre);
  delete h;
}

void srs_info(const void* handle, size_t* n, int* c, int* W, size_t* bytes) {
  const Srs* h = (const Srs*)handle;
  if (n) *n = h->n;
  if (c) *c = (int)h->c;
  if (W) *W = (int)h->W;
  if (bytes) *bytes = h->n * h->W * 96;
}

cudaError_t srs_msm(const void* handle, const void* scalars_dev, size_t n_used, void* out144_dev, cudaStream_t s, bool dry,
                    int* launches_out, float* phase_ms) {
  const Srs* h = (const Srs*)handle;
  if (n_used > h->n) return cudaErrorInvalidValue;
  msm::SrsView v{h->pre, (u32)h->n, h->c, h->W};
  if (!phase_ms || dry || n_used == 0)
    return msm::run(nullptr, 96, (const u32*)scalars_dev, n_used, (unsigned char*)out144_dev, s, dry, launches_out, nullptr, &v);
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]);
  cudaError_t e = msm::run(nullptr, 96, (const u32*)scalars_dev, n_used, (unsigned char*)out144_dev, s, false, launches_out, ev, &v);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  for (int i = 0; i < 3; i++) {
    phase_ms[i] = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  return e;
}

// Host scalars against a resident SRS, copied in point ranges like msm_run_host.  montgomery_in: the scalars
// are Fr in Montgomery form (KZG10::commit's polynomial) and are converted range by range on the device.
// out_bytes = 144 (normalised Jacobian) or 48 (compressed commitment).  Synchronises `s`.
cudaError_t srs_msm_host(const void* handle, const void* in_host, size_t n, bool montgomery_in, void* out_host, size_t out_bytes,
                         cudaStream_t s) {
  const Srs* h = (const Srs*)handle;
  if (n > h->n) return cudaErrorInvalidValue;
  msm::SrsView v{h->pre, (u32)h->n, h->c, h->W};
  unsigned char* d = nullptr;
  const size_t sb = (n * 32 + 255) & ~(size_t)255;
  MSM_CK(cudaMallocAsync((void**)&d, sb + 512, s));
  cudaError_t e = cudaSuccess;
  if (n == 0) {
    e = msm::run(nullptr, 96, nullptr, 0, d + sb, s, false, nullptr);
  } else {
    size_t sizes[4];
    const int k = chunk_schedule(n, sizes);
    cudaStream_t cs = nullptr;
    EventSet evs;
    e = copy_stream(&cs);
    if (e == cudaSuccess) e = evs.init();
    msm::Session ss;
    if (e == cudaSuccess) e = cudaEventRecord(evs.ev[4], s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, evs.ev[4], 0);
    size_t first = 0, max_chunk = 0;
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      e = cudaMemcpyAsync(d + first * 32, (const unsigned char*)in_host + first * 32, sizes[i] * 32, cudaMemcpyHostToDevice, cs);
      if (e == cudaSuccess) e = cudaEventRecord(evs.ev[i], cs);
      first += sizes[i];
      max_chunk = sizes[i] > max_chunk ? sizes[i] : max_chunk;
    }
    if (e == cudaSuccess) e = ss.begin(n, max_chunk, (u32)k, &v, s, false);
    first = 0;
    for (int i = 0; i < k && e == cudaSuccess; i++) {
      e = cudaStreamWaitEvent(s, evs.ev[i], 0);
      if (e == cudaSuccess && montgomery_in) e = fr_to_bigint(d + first * 32, d + first * 32, sizes[i], s);
      if (e == cudaSuccess) e = ss.add_chunk(nullptr, 96, (const u32*)(d + first * 32), sizes[i], first, s);
      first += sizes[i];
    }
    if (e == cudaSuccess) e = ss.finish(d + sb, s);
    else ss.release(s);
    if (e != cudaSuccess) cudaStreamSynchronize(cs);
  }
  if (e == cudaSuccess && out_bytes == 48) {
    e = g1_compress(d + sb, d + sb + 256, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d + sb + 256, 48, cudaMemcpyDeviceToHost, s);
  } else if (e == cudaSuccess) {
    e = cudaMemcpyAsync(out_host, d + sb, 144, cudaMemcpyDeviceToHost, s);
  }
  cudaFreeAsync(d, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : e2;
}

cudaError_t fr_to_bigint(const void* in_dev, void* out_dev, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  LAUNCH_NOSYNC(msm::fr_to_bigint_kernel, dim3((u32)((n + 255) / 256)), dim3(256), 0, s, (const Fr*)in_dev, (Fr*)out_dev, (u32)n);
  return cudaGetLastError();
}

cudaError_t g1_compress(const void* jac144_dev, void* out48_dev, cudaStream_t s) {
  LAUNCH_NOSYNC(msm::g1_compress_kernel, dim3(1), dim3(1), 0, s, (const unsigned char*)jac144_dev, (unsigned char*)out48_dev);
  return cudaGetLastError();
}

cudaError_t g1_sum(const void* points144_dev, u32 count, void* out144_dev, cudaStream_t s) {
  LAUNCH_NOSYNC(msm::g1_sum_kernel, dim3(1), dim3(1), 0, s, (const unsigned char*)points144_dev, count,
                (unsigned char*)out144_dev);
  return cudaGetLastError();
}
}  // namespace aleo
