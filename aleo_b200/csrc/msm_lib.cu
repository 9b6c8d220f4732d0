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

cudaError_t g1_sum(const void* points144_dev, u32 count, void* out144_dev, cudaStream_t s) {
  LAUNCH_NOSYNC(msm::g1_sum_kernel, dim3(1), dim3(1), 0, s, (const unsigned char*)points144_dev, count,
                (unsigned char*)out144_dev);
  return cudaGetLastError();
}
}  // namespace aleo
