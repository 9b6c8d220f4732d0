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
  const int forced = ntt::forced_split((int)log_n, K);
  return forced ? forced : ntt::split_passes((int)log_n, K);
}

cudaError_t ntt_transform(int device, u32 log_n, size_t batch, bool inverse, bool coset, void* data_dev, cudaStream_t s,
                          float* pass_ms) {
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(device, log_n, inverse, coset, s, &plan);
  if (e != cudaSuccess) return e;
  Fr* scratch = nullptr;
  if (log_n > (u32)ntt::SMALL_MAX_LOG) {
    e = aleo::pool_malloc_async((void**)&scratch, (sizeof(Fr) << log_n) * ntt::scratch_batch(log_n, batch), s);
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
  if (blocks > dev_props().sms * 8) blocks = dev_props().sms * 8;
  for (size_t b0 = 0; b0 < batch; b0 += 65535) {
    const u32 nb = (u32)(batch - b0 < 65535 ? batch - b0 : 65535);
    LAUNCH_NOSYNC(ntt::bitrev_permute_kernel, dim3((u32)blocks, nb), dim3(256), 0, s, (Fr*)data_dev + (b0 << log_n), log_n);
  }
  return cudaGetLastError();
}

// ---- one transform over the GPUs of a node, exchange by stores into peer memory (NVLink) -------------------------------
struct NttDist {
  int device = 0, rank = 0, world = 1, lg = 0;
  u32 log_n = 0;
  Fr* recv[2] = {nullptr, nullptr};          // double-buffered receive buffers (N / world elements each), cudaMalloc'ed
  Fr* peers[2][8] = {{nullptr}, {nullptr}};  // recv[b] of every rank, mapped into this process
  bool opened[2][8] = {{false}, {false}};
  bool have_peers = false;
  unsigned long long calls = 0;              // parity selects the buffer; every rank issues the same sequence of calls
  size_t flag_off() const { return (sizeof(Fr) << log_n) / (size_t)world; }  // bytes: 8 flag slots follow the buffer
  // flags of the transform about to run on buffer b: signal slots in every rank's array, own slots to wait on
  ntt::DistFlags flags(int b) const {
    ntt::DistFlags f;
    for (int r = 0; r < world; r++) f.signal[r] = (u32*)((unsigned char*)peers[b][r] + flag_off()) + rank;
    f.local = (const u32*)((const unsigned char*)recv[b] + flag_off());
    f.epoch = (u32)(calls + 1);            // every rank issues the same call sequence: the epochs agree
    return f;
  }
};

int ntt_dist_layout(u32 log_n, int world, u32* log_r_first, u32* log_r_last, int* npass) {
  int lg = 0;
  while ((1 << lg) < world) lg++;
  if ((1 << lg) != world || world > 8 || world < 1) return -1;
  int K[4];
  const int P = ntt::split_passes_dist((int)log_n, lg, K);
  if (P == 0 || (int)log_n - 2 * lg < 0) return -1;
  if (log_r_first) *log_r_first = (u32)K[0];
  if (log_r_last) *log_r_last = (u32)K[P - 1];
  if (npass) *npass = P;
  return 0;
}

cudaError_t ntt_dist_create(u32 log_n, int rank, int world, void** ctx_out) {
  if (ntt_dist_layout(log_n, world, nullptr, nullptr, nullptr) != 0 || rank < 0 || rank >= world) return cudaErrorInvalidValue;
  NttDist* c = new NttDist();
  cudaGetDevice(&c->device);
  c->rank = rank;
  c->world = world;
  while ((1 << c->lg) < world) c->lg++;
  c->log_n = log_n;
  const size_t bytes = (sizeof(Fr) << log_n) / (size_t)world;
  for (int b = 0; b < 2; b++) {
    cudaError_t e = cudaMalloc((void**)&c->recv[b], bytes + 256);  // + the cross-rank flag slots (one u32 per source rank)
    if (e == cudaSuccess) e = cudaMemset((unsigned char*)c->recv[b] + bytes, 0, 256);
    if (e != cudaSuccess) {
      if (b == 1) cudaFree(c->recv[0]);
      delete c;
      return e;
    }
  }
  {
    cudaError_t e = cudaDeviceSynchronize();  // the flag slots are zero before any peer can map them
    if (e != cudaSuccess) {
      cudaFree(c->recv[0]);
      cudaFree(c->recv[1]);
      delete c;
      return e;
    }
  }
  *ctx_out = c;
  return cudaSuccess;
}

// handles_out: 2 x 64 bytes describing this rank's two receive buffers for the other processes
cudaError_t ntt_dist_handles(void* ctx, void* handles_out) {
  NttDist* c = (NttDist*)ctx;
  std::memset(handles_out, 0, 128);
  for (int b = 0; b < 2; b++) {
#ifndef ALEO_EMU
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->recv[b]);
    if (e != cudaSuccess) return e;
    std::memcpy((unsigned char*)handles_out + 64 * b, &h, 64);
#else
    std::memcpy((unsigned char*)handles_out + 64 * b, &c->recv[b], sizeof(Fr*));  // emulator: ranks share one address space
#endif
  }
  return cudaSuccess;
}

// all_handles: world x (2 x 64 bytes), rank-major, as gathered from every rank's ntt_dist_handles
cudaError_t ntt_dist_open(void* ctx, const void* all_handles) {
  NttDist* c = (NttDist*)ctx;
  for (int r = 0; r < c->world; r++)
    for (int b = 0; b < 2; b++) {
      const unsigned char* h = (const unsigned char*)all_handles + ((size_t)r * 2 + b) * 64;
      if (r == c->rank) {
        c->peers[b][r] = c->recv[b];
        continue;
      }
#ifndef ALEO_EMU
      cudaIpcMemHandle_t ih;
      std::memcpy(&ih, h, 64);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return e;
      c->peers[b][r] = (Fr*)ptr;
      c->opened[b][r] = true;
#else
      Fr* ptr = nullptr;
      std::memcpy(&ptr, h, sizeof(Fr*));
      c->peers[b][r] = ptr;
#endif
    }
  c->have_peers = true;
  return cudaSuccess;
}

// passes 0 .. P-2 of the local part; the last of them writes into the peers' receive buffers
cudaError_t ntt_dist_stage1(void* ctx, const void* local_in_dev, bool inverse, bool coset, cudaStream_t s) {
  NttDist* c = (NttDist*)ctx;
  if (!c->have_peers) return cudaErrorInvalidValue;
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(c->device, c->log_n, inverse, coset, s, &plan, c->lg);
  if (e != cudaSuccess) return e;
  Fr* scratch = nullptr;
  if (plan->npass > 2) {
    e = aleo::pool_malloc_async((void**)&scratch, (sizeof(Fr) << c->log_n) / (size_t)c->world, s);
    if (e != cudaSuccess) return e;
  }
  const ntt::DistFlags fl = c->flags((int)(c->calls & 1));
  e = ntt::run_dist_stage1(*plan, c->lg, c->rank, (const Fr*)local_in_dev, scratch, c->peers[c->calls & 1], s, &fl);
  if (scratch) cudaFreeAsync(scratch, s);
  return e;
}

// the last pass from this rank's receive buffer (call after a cross-rank barrier ordered on the same stream)
cudaError_t ntt_dist_stage2(void* ctx, void* local_out_dev, bool inverse, bool coset, cudaStream_t s) {
  NttDist* c = (NttDist*)ctx;
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(c->device, c->log_n, inverse, coset, s, &plan, c->lg);
  if (e != cudaSuccess) return e;
  const ntt::DistFlags fl = c->flags((int)(c->calls & 1));
  e = ntt::run_dist_stage2(*plan, c->lg, c->rank, c->recv[c->calls & 1], (Fr*)local_out_dev, s, &fl);
  c->calls++;
  return e;
}

// the whole transform with per-stage events: stage_ms4 = {local passes before the exchange pass, exchange pass (stores
// into the peers), wait for every peer's flag, last pass}.  Synchronises s.
cudaError_t ntt_dist_profile(void* ctx, const void* local_in_dev, void* local_out_dev, bool inverse, bool coset, cudaStream_t s,
                             float* stage_ms4) {
  NttDist* c = (NttDist*)ctx;
  if (!c->have_peers) return cudaErrorInvalidValue;
  const ntt::Plan* plan = nullptr;
  cudaError_t e = g_plans.get(c->device, c->log_n, inverse, coset, s, &plan, c->lg);
  if (e != cudaSuccess) return e;
  Fr* scratch = nullptr;
  if (plan->npass > 2) {
    e = aleo::pool_malloc_async((void**)&scratch, (sizeof(Fr) << c->log_n) / (size_t)c->world, s);
    if (e != cudaSuccess) return e;
  }
  cudaEvent_t ev[5];
  for (auto& x : ev) cudaEventCreate(&x);
  const int b = (int)(c->calls & 1);
  const ntt::DistFlags fl = c->flags(b);
  e = ntt::run_dist_stage1(*plan, c->lg, c->rank, (const Fr*)local_in_dev, scratch, c->peers[b], s, &fl, ev);
  cudaEventRecord(ev[2], s);
  LAUNCH_NOSYNC(ntt::dist_wait_kernel, dim3(1), dim3(32), 0, s, fl.local, (u32)c->world, fl.epoch);
  cudaEventRecord(ev[3], s);
  if (e == cudaSuccess) e = ntt::run_dist_stage2(*plan, c->lg, c->rank, c->recv[b], (Fr*)local_out_dev, s, &fl);
  cudaEventRecord(ev[4], s);
  c->calls++;
  if (scratch) cudaFreeAsync(scratch, s);
  cudaError_t e2 = cudaStreamSynchronize(s);
  for (int i = 0; i < 4; i++) {
    stage_ms4[i] = 0.f;
    if (e == cudaSuccess && e2 == cudaSuccess) cudaEventElapsedTime(&stage_ms4[i], ev[i], ev[i + 1]);
  }
  for (auto& x : ev) cudaEventDestroy(x);
  return e != cudaSuccess ? e : e2;
}

void ntt_dist_destroy(void* ctx) {
  NttDist* c = (NttDist*)ctx;
  if (!c) return;
#ifndef ALEO_EMU
  for (int b = 0; b < 2; b++)
    for (int r = 0; r < 8; r++)
      if (c->opened[b][r]) cudaIpcCloseMemHandle(c->peers[b][r]);
#endif
  cudaFree(c->recv[0]);
  cudaFree(c->recv[1]);
  delete c;
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
