// msm_lib.cu -- translation unit holding the G1 MSM kernels.
#include "internal.h"
#include "msm_host.cuh"

namespace aleo {
cudaError_t msm_upload_constants() { return aleo_upload_field_constants(); }
bool msm_size_supported(size_t n) { return msm::size_supported(n); }
int msm_window_bits(size_t n) { return (int)msm::make_params(n).c; }

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
  h->W = msm::SCALAR_BITS / h->c + 1;
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
