// util_lib.cu -- translation unit holding workload generators, on-device checks and the
// integer-pipe microbenchmark (bench / tests support).
#include "internal.h"
#include "util_kernels.cuh"

#define UT_CK(expr)                       \
  do {                                    \
    cudaError_t _e = (expr);              \
    if (_e != cudaSuccess) return _e;     \
  } while (0)

namespace aleo {
cudaError_t util_upload_constants() { return aleo_upload_field_constants(); }

static Fr load_fr(const void* p) {
  Fr v;
  std::memcpy(v.l, p, 32);
  return v;
}

cudaError_t gen_bases(void* bases_dev, size_t n, u32 stride, const void* s0_32, const void* d_32, u64 first, cudaStream_t s) {
  const Fr s0 = load_fr(s0_32), d = load_fr(d_32);
  const size_t slab = (size_t)1 << 22;
  Fq* qaff = nullptr;
  G1Xyzz* scratch = nullptr;
  UT_CK(aleo::pool_malloc_async((void**)&qaff, 2 * sizeof(Fq), s));
  UT_CK(aleo::pool_malloc_async((void**)&scratch, (n < slab ? n : slab) * sizeof(G1Xyzz), s));
  LAUNCH_NOSYNC(util::gen_step_point_kernel, dim3(1), dim3(1), 0, s, qaff, d);
  cudaError_t e = cudaGetLastError();
  for (size_t done = 0; done < n && e == cudaSuccess; done += slab) {
    const u32 cnt = (u32)((n - done < slab) ? (n - done) : slab);
    const u32 threads = (cnt + util::GEN_CHUNK - 1) / util::GEN_CHUNK;
    LAUNCH_NOSYNC(util::gen_bases_xyzz_kernel, dim3((threads + 127) / 128), dim3(128), 0, s, scratch, cnt, s0, d,
                  (u64)(first + done), (const Fq*)qaff);
    LAUNCH_NOSYNC(util::gen_bases_normalise_kernel, dim3((threads + 127) / 128), dim3(128), 0, s,
                  (unsigned char*)bases_dev, stride, (const G1Xyzz*)scratch, cnt, (u64)done);
    e = cudaGetLastError();
  }
  cudaFreeAsync(scratch, s);
  cudaFreeAsync(qaff, s);
  return e;
}

cudaError_t gen_scalars(void* scalars_dev, size_t n, u64 seed, u64 first, bool montgomery, cudaStream_t s) {
  LAUNCH_NOSYNC(util::gen_scalars_kernel, dim3((u32)((n + 255) / 256)), dim3(256), 0, s, (Fr*)scalars_dev, (u32)n, seed,
                first, (u32)(montgomery ? 1 : 0));
  return cudaGetLastError();
}

cudaError_t dlog_dot(void* out_dev, const void* scalars_dev, size_t n, const void* s0_32, const void* d_32, u64 first,
                     cudaStream_t s) {
  const Fr s0 = load_fr(s0_32), d = load_fr(d_32);
  const u32 nblocks = dev_props().sms * 4;
  Fr* partials = nullptr;
  UT_CK(aleo::pool_malloc_async((void**)&partials, nblocks * sizeof(Fr), s));
  LAUNCH(util::dlog_dot_kernel, dim3(nblocks), dim3(256), 256 * sizeof(Fr), s, (const Fr*)scalars_dev, (u32)n, s0, d, first,
         partials);
  LAUNCH_NOSYNC(util::dlog_dot_final_kernel, dim3(1), dim3(1), 0, s, (const Fr*)partials, nblocks, (Fr*)out_dev);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partials, s);
  return e;
}

cudaError_t check_on_curve(const void* bases_dev, size_t n, u32 stride, cudaStream_t s, int* ok) {
  u32* bad = nullptr;
  UT_CK(aleo::pool_malloc_async((void**)&bad, 4, s));
  cudaMemsetAsync(bad, 0, 4, s);
  LAUNCH_NOSYNC(util::check_on_curve_kernel, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)bases_dev,
                stride, (u32)n, bad);
  u32 h = 1;
  cudaError_t e = cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(bad, s);
  *ok = (h == 0) ? 1 : 0;
  return e;
}

cudaError_t fq_mul_fp64(void* out_dev, const void* a_dev, const void* b_dev, size_t n, bool square, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  LAUNCH_NOSYNC(util::fq_mul_fp64_kernel, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (Fq*)out_dev, (const Fq*)a_dev, (const Fq*)b_dev,
                (u32)n, (u32)(square ? 1 : 0));
  return cudaGetLastError();
}

cudaError_t bench_imad(int kind, int iters, double* ms_out, double* ops_out) {
#ifdef ALEO_EMU
  (void)kind; (void)iters;
  *ms_out = 0.0;
  *ops_out = 0.0;
  return cudaSuccess;
#else
  u32* sink = nullptr;
  UT_CK(cudaMalloc((void**)&sink, 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const u32 grid = dev_props().sms * 8, block = 256;
  const double per_thread[9] = {64.0, 64.0, 64.0, 2.0 * 276.0, 64.0, 64.0, 2.0 * 276.0, 2.0 * 276.0, 2.0 * 276.0};  // kinds 6-8: counted as 276-MAC products  // kind 5: 64 DFMA + 64 IMAD.WIDE, counted as 64 pairs
  for (int rep = 0; rep < 2; rep++) {  // first launch warms up
    cudaEventRecord(e0, 0);
    switch (kind) {
      case 0: util::imad_bench_kernel<0><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 1: util::imad_bench_kernel<1><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 2: util::imad_bench_kernel<2><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 4: util::imad_bench_kernel<4><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 5: util::imad_bench_kernel<5><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 6: util::imad_bench_kernel<6><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 7: util::imad_bench_kernel<7><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      case 8: util::imad_bench_kernel<8><<<grid, block>>>(sink, (u32)iters, 12345u); break;
      default: util::imad_bench_kernel<3><<<grid, block>>>(sink, (u32)iters, 12345u); break;
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *ms_out = ms;
  *ops_out = per_thread[(kind >= 0 && kind < 9) ? kind : 3] * (double)iters * (double)grid * (double)block;
  return e;
#endif
}
}  // namespace aleo
