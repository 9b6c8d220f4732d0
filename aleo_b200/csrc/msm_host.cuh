// msm_host.cuh -- host-side driver of msm.cuh: window choice, workspace carving, launch sequence.
#pragma once
#include <vector>
#include "msm.cuh"

namespace msm {

#define MSM_CK(expr)                      \
  do {                                    \
    cudaError_t _e = (expr);              \
    if (_e != cudaSuccess) return _e;     \
  } while (0)

static inline u32 floor_log2(size_t n) {
  u32 l = 0;
  while ((n >> (l + 1)) != 0) l++;
  return l;
}

// window bits: about n / 2^(c-1) >= 32 entries per bucket, capped so that the bucket reduction
// (2 * 2^(c-1) full additions per window) stays a small fraction of the n mixed additions per window
static inline u32 choose_window(size_t n) {
  if (n < 2) return 4;
  long c = (long)floor_log2(n) - 4;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  return (u32)c;
}

static inline Params make_params(size_t n) {
  Params p;
  p.c = choose_window(n);
  p.W = SCALAR_BITS / p.c + 1;
  p.B = 1u << (p.c - 1);
  size_t t = ((size_t)n * p.W + 149999) / 150000;  // ~1000 tasks per SM keeps 148 SMs busy
  if (t < 16) t = 16;
  if (t > 1024) t = 1024;
  p.T = (u32)t;
  return p;
}

// largest n one call accepts: sorted positions are 32-bit
static inline bool size_supported(size_t n) {
  if (n >= ((size_t)1 << 31)) return false;
  Params p = make_params(n);
  return (unsigned long long)n * p.W < (1ull << 32);
}

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  }
};

static inline cudaError_t exclusive_scan(const u32* in, u32 n, u32* block_sums, u32* out, u32* out_copy, u32* total,
                                         cudaStream_t s, int& launches) {
  const u32 nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  LAUNCH(scan_block_sums_kernel, dim3(nblocks), dim3(SCAN_TPB), 0, s, in, n, block_sums);
  LAUNCH(scan_sums_kernel, dim3(1), dim3(SCAN_TPB), 0, s, block_sums, nblocks, total);
  LAUNCH(scan_apply_kernel, dim3(nblocks), dim3(SCAN_TPB), 0, s, in, n, (const u32*)block_sums, out, out_copy);
  launches += 3;
  return cudaGetLastError();
}

struct RedLevel {
  u32 m, T;  // input length per window, chunks per window
};

// Runs (or, with dry = true, only counts the launches of) one MSM.  bases / scalars / out144 are
// device pointers; everything is asynchronous on `s` except the workspace allocation call itself.
// phase_ev (optional, 4 events): recorded before the sort phase, before / after the bucket
// accumulation kernel, and after the final kernel -- bench.py's per-kernel timing.
static inline cudaError_t run(const unsigned char* bases, u32 stride, const u32* scalars, size_t n_sz,
                              unsigned char* out144, cudaStream_t s, bool dry, int* launches_out,
                              cudaEvent_t* phase_ev = nullptr) {
  int launches = 0;
  if (n_sz == 0) {
    if (!dry) LAUNCH_NOSYNC(write_identity_kernel, dim3(1), dim3(1), 0, s, out144);
    launches = 1;
    if (launches_out) *launches_out = launches;
    return dry ? cudaSuccess : cudaGetLastError();
  }
  const u32 n = (u32)n_sz;
  const Params prm = make_params(n_sz);
  const u32 NB = prm.W * prm.B;
  const size_t entries_ub = (size_t)n * prm.W;
  const u32 task_ub = (u32)(NB + entries_ub / prm.T + 1);
  const u32 cap_small = (u32)(entries_ub / prm.T + 1);
  const u32 cap_large = (u32)(entries_ub / ((size_t)prm.T * SMALL_SPLIT_MAX) + 1);
  const u32 scan_blocks = (NB + SCAN_BLOCK - 1) / SCAN_BLOCK;

  // reduction level structure
  std::vector<RedLevel> lv;
  {
    u32 m = prm.B;
    while (m > 0 && (int)lv.size() < MAX_RED_LEVELS) {
      RedLevel l;
      l.m = m;
      l.T = (m + RED_KC - 1) / RED_KC;
      lv.push_back(l);
      m = l.T - 1;
    }
  }

  Carver cv;
  const size_t o_counts = cv.take((size_t)NB * 4), o_starts = cv.take((size_t)NB * 4), o_ends = cv.take((size_t)NB * 4);
  const size_t o_ntasks = cv.take((size_t)NB * 4), o_taskoff = cv.take((size_t)NB * 4);
  const size_t o_bsums = cv.take((size_t)(scan_blocks + 1) * 4), o_meta = cv.take(64);
  const size_t o_sorted = cv.take(entries_ub * 4);
  const size_t o_small = cv.take((size_t)cap_small * 4), o_large = cv.take((size_t)cap_large * 4);
  const size_t o_buckets = cv.take((size_t)NB * sizeof(G1Xyzz));
  const size_t o_partials = cv.take((size_t)task_ub * sizeof(G1Xyzz));
  std::vector<size_t> o_R(lv.size()), o_WS(lv.size());
  for (size_t l = 0; l < lv.size(); l++) {
    o_R[l] = cv.take((size_t)prm.W * lv[l].T * sizeof(G1Xyzz));
    o_WS[l] = cv.take((size_t)prm.W * lv[l].T * sizeof(G1Xyzz));
  }
  // ping-pong buffers for the plain sums of WS (largest needed: W * ceil(T0 / Kc))
  const size_t ps_elems = (size_t)prm.W * ((lv[0].T + RED_KC - 1) / RED_KC);
  std::vector<size_t> o_PSfinal(lv.size());
  const size_t o_psA = cv.take(ps_elems * sizeof(G1Xyzz)), o_psB = cv.take(ps_elems * sizeof(G1Xyzz));
  for (size_t l = 0; l < lv.size(); l++) o_PSfinal[l] = cv.take((size_t)prm.W * sizeof(G1Xyzz));
  const size_t o_D = cv.take((size_t)prm.W * sizeof(G1Xyzz));

  unsigned char* ws = nullptr;
  if (!dry) MSM_CK(cudaMallocAsync((void**)&ws, cv.off, s));
#define WSP(type, off) reinterpret_cast<type*>(ws + (off))
  u32* counts = WSP(u32, o_counts);
  u32* starts = WSP(u32, o_starts);
  u32* ends = WSP(u32, o_ends);
  u32* ntasks = WSP(u32, o_ntasks);
  u32* task_off = WSP(u32, o_taskoff);
  u32* bsums = WSP(u32, o_bsums);
  u32* meta = WSP(u32, o_meta);
  u32* sorted = WSP(u32, o_sorted);
  u32* small_list = WSP(u32, o_small);
  u32* large_list = WSP(u32, o_large);
  G1Xyzz* buckets = WSP(G1Xyzz, o_buckets);
  G1Xyzz* partials = WSP(G1Xyzz, o_partials);

  cudaError_t err = cudaSuccess;
#define STEP(stmt)                                       \
  do {                                                   \
    if (!dry && err == cudaSuccess) {                    \
      stmt;                                              \
      err = cudaGetLastError();                          \
    }                                                    \
  } while (0)

  if (phase_ev && !dry) cudaEventRecord(phase_ev[0], s);
  STEP(cudaMemsetAsync(counts, 0, (size_t)NB * 4, s));
  STEP(cudaMemsetAsync(meta, 0, 64, s));
  STEP(cudaMemsetAsync(buckets, 0, (size_t)NB * sizeof(G1Xyzz), s));
  const u32 g_n = (n + 255) / 256;
  STEP(LAUNCH_NOSYNC(count_kernel, dim3(g_n), dim3(256), 0, s, scalars, n, prm, counts));
  launches++;
  if (!dry && err == cudaSuccess) err = exclusive_scan(counts, NB, bsums, starts, ends, meta + 0, s, launches);
  else launches += 3;
  STEP(LAUNCH_NOSYNC(scatter_kernel, dim3(g_n), dim3(256), 0, s, scalars, n, prm, ends, sorted));
  launches++;
  STEP(LAUNCH_NOSYNC(plan_tasks_kernel, dim3((NB + 255) / 256), dim3(256), 0, s, (const u32*)starts, (const u32*)ends, NB,
                     prm.T, ntasks, small_list, large_list, cap_small, cap_large, meta));
  launches++;
  if (!dry && err == cudaSuccess) err = exclusive_scan(ntasks, NB, bsums, task_off, nullptr, meta + 1, s, launches);
  else launches += 3;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[1], s);
  STEP(LAUNCH_NOSYNC(accumulate_kernel, dim3((task_ub + 127) / 128), dim3(128), 0, s, bases, stride, (const u32*)sorted,
                     (const u32*)starts, (const u32*)ends, (const u32*)ntasks, (const u32*)task_off, NB, prm.T,
                     (const u32*)meta, buckets, partials));
  launches++;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[2], s);
  STEP(LAUNCH_NOSYNC(combine_small_kernel, dim3((cap_small + 127) / 128), dim3(128), 0, s, (const u32*)small_list,
                     (const u32*)ntasks, (const u32*)task_off, (const u32*)meta, (const G1Xyzz*)partials, buckets));
  launches++;
  {
    const u32 g = cap_large < 592 ? cap_large : 592;  // 4 CTAs per SM; the kernel strides over the list
    STEP(LAUNCH(combine_large_kernel, dim3(g), dim3(COMBINE_TPB), COMBINE_TPB * sizeof(G1Xyzz), s, (const u32*)large_list,
                (const u32*)ntasks, (const u32*)task_off, (const u32*)meta, (const G1Xyzz*)partials, buckets));
    launches++;
  }
  // bucket reduction
  FinalArgs fa;
  fa.levels = (int)lv.size();
  fa.W = prm.W;
  fa.c = prm.c;
  for (size_t l = 0; l < lv.size(); l++) {
    const G1Xyzz* X = (l == 0) ? buckets : (WSP(G1Xyzz, o_R[l - 1]) + 1);
    const u32 xs = (l == 0) ? prm.B : lv[l - 1].T;
    G1Xyzz* R = WSP(G1Xyzz, o_R[l]);
    G1Xyzz* WS = WSP(G1Xyzz, o_WS[l]);
    const u32 threads = prm.W * lv[l].T;
    STEP(LAUNCH_NOSYNC(wsum_kernel, dim3((threads + 127) / 128), dim3(128), 0, s, X, xs, lv[l].m, prm.W, lv[l].T, R, WS));
    launches++;
    // plain sum of WS[l] (T entries per window) down to one entry per window
    const G1Xyzz* cur = WS;
    u32 cur_n = lv[l].T;
    int flip = 0;
    while (cur_n > 1) {
      const u32 t = (cur_n + RED_KC - 1) / RED_KC;
      G1Xyzz* dst = (t == 1) ? WSP(G1Xyzz, o_PSfinal[l]) : WSP(G1Xyzz, flip ? o_psB : o_psA);
      const u32 th = prm.W * t;
      STEP(LAUNCH_NOSYNC(psum_kernel, dim3((th + 127) / 128), dim3(128), 0, s, cur, cur_n, cur_n, prm.W, t, dst));
      launches++;
      cur = dst;
      cur_n = t;
      flip ^= 1;
    }
    fa.f[l] = cur;
    fa.stride[l] = 1;
  }
  G1Xyzz* D = WSP(G1Xyzz, o_D);
  STEP(LAUNCH_NOSYNC(window_weigh_kernel, dim3((prm.W + 31) / 32), dim3(32), 0, s, fa, D));
  launches++;
  STEP(LAUNCH_NOSYNC(final_kernel, dim3(1), dim3(1), 0, s, (const G1Xyzz*)D, prm.W, out144));
  launches++;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[3], s);
  if (ws) cudaFreeAsync(ws, s);
#undef STEP
#undef WSP
  if (launches_out) *launches_out = launches;
  return err;
}

}  // namespace msm
