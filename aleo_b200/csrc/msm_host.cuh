// msm_host.cuh -- host-side driver of msm.cuh: window choice, workspace carving, launch sequence.
#pragma once
#include <cstdlib>
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

// a resident SRS expanded by srs_expand_kernel: pre[w][i] = 2^(c w) * P_i, packed 96-byte affine
struct SrsView {
  const unsigned char* pre;
  u32 n_total;  // points per window block
  u32 c, W;
};

// window bits of a resident SRS of n points: the doubling tail and the per-window buckets are gone, so
// the optimum moves up (n * W(c) mixed additions against 2 * 2^(c-1) reduction additions)
static inline u32 choose_window_srs(size_t n) {
  if (n < 2) return 8;
  long c = (long)floor_log2(n) - 2;
  if (c < 8) c = 8;
  if (c > 22) c = 22;
  return (u32)c;
}

static inline Params make_params(size_t n, const SrsView* srs = nullptr) {
  Params p;
  p.c = srs ? srs->c : choose_window(n);
  p.W = SCALAR_BITS / p.c + 1;
  p.B = 1u << (p.c - 1);
  p.n_stride = srs ? srs->n_total : 0;
  // accumulation threads: 16 waves of (148 SMs x 3 CTAs x 128 threads) -- measured best on B200: 4 waves
  // 96.1 ms, 16 waves 93.4 ms at n = 2^24 -- but at least ~32 entries per run
  size_t lanes = ((size_t)n * p.W + 31) / 32;
  static const long waves_env = []() { const char* e = getenv("ALEO_B200_MSM_WAVES"); return e ? atol(e) : 0L; }();
  const size_t full = (size_t)148 * 384 * (waves_env > 0 ? (size_t)waves_env : 16);
  if (lanes > full) lanes = full;
  lanes = (lanes + 127) / 128 * 128;
  p.nlanes = (u32)lanes;
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
                              cudaEvent_t* phase_ev = nullptr, const SrsView* srs = nullptr) {
  int launches = 0;
  if (n_sz == 0) {
    if (!dry) LAUNCH_NOSYNC(write_identity_kernel, dim3(1), dim3(1), 0, s, out144);
    launches = 1;
    if (launches_out) *launches_out = launches;
    return dry ? cudaSuccess : cudaGetLastError();
  }
  const u32 n = (u32)n_sz;
  const Params prm = make_params(n_sz, srs);
  if (srs) {
    bases = srs->pre;
    stride = 96;
  }
  const u32 nwin = srs ? 1u : prm.W;  // bucket sets (the resident SRS shares one across all windows)
  const u32 NB = nwin * prm.B;
  const size_t entries_ub = (size_t)n * prm.W;
  const u32 cap_small = prm.nlanes + 1;                   // every run boundary cuts at most one bucket
  const u32 cap_large = prm.nlanes / SMALL_SPLIT_MAX + 1;
  const u32 scan_blocks = (NB + SCAN_BLOCK - 1) / SCAN_BLOCK;

  // reduction level structure
  std::vector<RedLevel> lv;
  {
    // level l: weighted input of length m (buckets, then R[1..) of the level below), T chunks out;
    // the plain input P of the level below has T_prev entries and needs ceil(T_prev / Kc) chunks too
    u32 m = prm.B, t_prev = 0;
    for (;;) {
      RedLevel l;
      l.m = m;
      const u32 need = t_prev > m ? t_prev : m;
      l.T = (need + RED_KC - 1) / RED_KC;
      if (l.T == 0) l.T = 1;
      lv.push_back(l);
      if (l.T == 1) break;
      t_prev = l.T;
      m = l.T - 1;
    }
  }

  Carver cv;
  const size_t o_counts = cv.take((size_t)NB * 4), o_starts = cv.take((size_t)NB * 4), o_ends = cv.take((size_t)NB * 4);
  const size_t o_piece_bucket = cv.take((size_t)prm.nlanes * 2 * 4);
  const size_t o_bsums = cv.take((size_t)(scan_blocks + 1) * 4), o_meta = cv.take(64);
  const size_t o_sorted = cv.take(entries_ub * 4);
  const size_t o_small = cv.take((size_t)cap_small * 4), o_large = cv.take((size_t)cap_large * 4);
  const size_t o_buckets = cv.take((size_t)NB * sizeof(G1Xyzz));
  const size_t o_pieces = cv.take((size_t)prm.nlanes * 2 * sizeof(G1Xyzz));
  std::vector<size_t> o_R(lv.size()), o_P(lv.size());
  for (size_t l = 0; l < lv.size(); l++) {
    o_R[l] = cv.take((size_t)nwin * lv[l].T * sizeof(G1Xyzz));
    o_P[l] = cv.take((size_t)nwin * lv[l].T * sizeof(G1Xyzz));
  }
  const size_t o_D = cv.take((size_t)nwin * sizeof(G1Xyzz));

  unsigned char* ws = nullptr;
  if (!dry) MSM_CK(cudaMallocAsync((void**)&ws, cv.off, s));
#define WSP(type, off) reinterpret_cast<type*>(ws + (off))
  u32* counts = WSP(u32, o_counts);
  u32* starts = WSP(u32, o_starts);
  u32* ends = WSP(u32, o_ends);
  u32* piece_bucket = WSP(u32, o_piece_bucket);
  u32* bsums = WSP(u32, o_bsums);
  u32* meta = WSP(u32, o_meta);
  u32* sorted = WSP(u32, o_sorted);
  u32* small_list = WSP(u32, o_small);
  u32* large_list = WSP(u32, o_large);
  G1Xyzz* buckets = WSP(G1Xyzz, o_buckets);
  G1Xyzz* pieces = WSP(G1Xyzz, o_pieces);

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
  STEP(LAUNCH_NOSYNC(plan_pieces_kernel, dim3((NB + 255) / 256), dim3(256), 0, s, (const u32*)starts, (const u32*)ends, NB,
                     prm.nlanes, small_list, large_list, cap_small, cap_large, meta));
  launches++;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[1], s);
  static const bool acc_inline = []() { const char* e = getenv("ALEO_B200_MSM_ACC"); return e && e[0] == 'i'; }();
  if (acc_inline)
    STEP(LAUNCH_NOSYNC(accumulate_kernel<false>, dim3(prm.nlanes / 128), dim3(128), 0, s, bases, stride, (const u32*)sorted,
                       (const u32*)starts, (const u32*)ends, NB, prm.nlanes, (const u32*)meta, buckets, pieces, piece_bucket));
  else
    STEP(LAUNCH_NOSYNC(accumulate_kernel<true>, dim3(prm.nlanes / 128), dim3(128), 0, s, bases, stride, (const u32*)sorted,
                       (const u32*)starts, (const u32*)ends, NB, prm.nlanes, (const u32*)meta, buckets, pieces, piece_bucket));
  launches++;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[2], s);
  STEP(LAUNCH_NOSYNC(combine_small_kernel, dim3((cap_small + 127) / 128), dim3(128), 0, s, (const u32*)small_list,
                     (const u32*)starts, (const u32*)ends, prm.nlanes, (const u32*)meta, (const G1Xyzz*)pieces,
                     (const u32*)piece_bucket, buckets));
  launches++;
  {
    const u32 g = cap_large < 592 ? cap_large : 592;  // 4 CTAs per SM; the kernel strides over the list
    STEP(LAUNCH(combine_large_kernel, dim3(g), dim3(COMBINE_TPB), COMBINE_TPB * sizeof(G1Xyzz), s, (const u32*)large_list,
                (const u32*)starts, (const u32*)ends, prm.nlanes, (const u32*)meta, (const G1Xyzz*)pieces,
                (const u32*)piece_bucket, buckets));
    launches++;
  }
  // bucket reduction: one launch per level until a single chunk per window is left
  for (size_t l = 0; l < lv.size(); l++) {
    ReduceArgs ra;
    ra.X = (l == 0) ? buckets : WSP(G1Xyzz, o_R[l - 1]);
    ra.x_stride = (l == 0) ? prm.B : lv[l - 1].T;
    ra.x_off = (l == 0) ? 0 : 1;
    ra.m = lv[l].m;
    ra.P = (l == 0) ? nullptr : WSP(G1Xyzz, o_P[l - 1]);
    ra.T_in = (l == 0) ? 0 : lv[l - 1].T;
    ra.level = (u32)l;
    ra.R_out = WSP(G1Xyzz, o_R[l]);
    ra.P_out = WSP(G1Xyzz, o_P[l]);
    ra.T_out = lv[l].T;
    ra.nwin = nwin;
    const u32 threads = nwin * lv[l].T;
    STEP(LAUNCH_NOSYNC(reduce_level_kernel, dim3((threads + 127) / 128), dim3(128), 0, s, ra));
    launches++;
  }
  const G1Xyzz* S = WSP(G1Xyzz, o_P[lv.size() - 1]);
  G1Xyzz* D = WSP(G1Xyzz, o_D);
  // per-window weights 2^(c w): not needed for a resident SRS (they are baked into the expanded bases)
  STEP(LAUNCH_NOSYNC(window_weigh_kernel, dim3((nwin + 31) / 32), dim3(32), 0, s, S, lv.back().T, nwin, prm.c, D));
  launches++;
  STEP(LAUNCH_NOSYNC(final_kernel, dim3(1), dim3(1), 0, s, (const G1Xyzz*)D, nwin, out144));
  launches++;
  if (phase_ev && !dry) cudaEventRecord(phase_ev[3], s);
  if (ws) cudaFreeAsync(ws, s);
#undef STEP
#undef WSP
  if (launches_out) *launches_out = launches;
  return err;
}

}  // namespace msm
