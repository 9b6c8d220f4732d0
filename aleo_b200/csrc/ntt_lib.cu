// ntt_lib.cu -- translation unit holding the Fr NTT kernels and their plan cache.
#include "internal.h"
#include "ntt_host.cuh"

namespace {
ntt::PlanCache g_plans;
}

namespace aleo {
cudaError_t ntt_upload_constants() { return aleo_upload_field_constants(); }
int ntt_max_log_n() { return ntt::MAX_LOG_N; }
void ntt_clear_plans() { g_plans.clear(); }

int ntt_launches(u32 log_n) {
  if (log_n == 0) return 0;
  if (log_n <= (u32)ntt::SMALL_MAX_LOG) return 1;
  int K[4];
  return ntt::split_passes((int)log_n, K);
}

cudaError_t ntt_transform(int device, u32 log_n, size_t batch, bool inverse, bool coset, void* data_dev, cudaStream_t s,
                          float* pass_ms) {
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(device, log_n, inverse, coset, s, &plan);
  if (e != cudaSuccess) return e;
  Fr* scratch = nullptr;
  if (log_n > (u32)ntt::SMALL_MAX_LOG) {
    e = cudaMallocAsync((void**)&scratch, (sizeof(Fr) << log_n) * ntt::scratch_batch(log_n, batch), s);
    if (e != cudaSuccess) return e;
  }
  cudaEvent_t ev[5];
  const int nev = ntt_launches(log_n) + 1;
  if (pass_ms)
    for (int i = 0; i < nev; i++) cudaEventCreate(&ev[i]);
  e = ntt::run(*plan, (Fr*)data_dev, batch, scratch, s, pass_ms ? ev : nullptr);
  if (scratch) cudaFreeAsync(scratch, s);
  if (pass_ms) {
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    for (int i = 0; i < 4; i++) pass_ms[i] = 0.f;
    for (int i = 0; i + 1 < nev && e == cudaSuccess; i++) cudaEventElapsedTime(&pass_ms[i], ev[i], ev[i + 1]);
    for (int i = 0; i < nev; i++) cudaEventDestroy(ev[i]);
  }
  return e;
}

// in-place bit-reversal permutation of `batch` vectors of 2^log_n elements
cudaError_t ntt_bitrev(u32 log_n, size_t batch, void* data_dev, cudaStream_t s) {
  if (log_n < 2 || batch == 0) return cudaSuccess;  // sizes 1 and 2 are their own reversal
  const u64 n = (u64)1 << log_n;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  for (size_t b0 = 0; b0 < batch; b0 += 65535) {
    const u32 nb = (u32)(batch - b0 < 65535 ? batch - b0 : 65535);
    LAUNCH_NOSYNC(ntt::bitrev_permute_kernel, dim3((u32)blocks, nb), dim3(256), 0, s, (Fr*)data_dev + (b0 << log_n), log_n);
  }
  return cudaGetLastError();
}

// data[r][c] *= w_N^(+-(r + row0)(c + col0)), N = 2^log_n_global (multi-GPU four-step twiddle step)
cudaError_t ntt_twiddle_matrix(int device, u32 log_n_global, bool inverse, void* data_dev, u32 rows, u32 cols, u32 row0, u32 col0,
                               cudaStream_t s) {
  if (log_n_global <= (u32)ntt::SMALL_MAX_LOG) return cudaErrorInvalidValue;
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(device, log_n_global, inverse, false, s, &plan);
  if (e != cudaSuccess) return e;
  const u64 total = (u64)rows * cols;
  if (total == 0) return cudaSuccess;
  LAUNCH_NOSYNC(ntt::twiddle_matrix_kernel, dim3((u32)((total + 255) / 256)), dim3(256), 0, s, (Fr*)data_dev, rows, cols, row0, col0,
                (ntt::PowTable{plan->tw_lo, plan->tw_hi, plan->lo_bits}), log_n_global);
  return cudaGetLastError();
}
}  // namespace aleo
