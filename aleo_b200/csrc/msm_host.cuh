// msm_host.cuh -- host-side driver of msm.cuh: window choice, workspace carving, launch sequence.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "msm.cuh"
#include "msm_ba.cuh"

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

// windows of c bits: W(c) = 253 / c + 1 (the +1 absorbs the signed-recoding carry)
static inline u32 windows_for(u32 c) { return SCALAR_BITS / c + 1; }

// Batch-affine levels (msm_ba.cuh) run in front of the XYZZ kernel from this many points on (ALEO_B200_MSM_BA = 0: never).
// Measured on B200 (profiles/r03e_ba_sweep_*.log): 2^20 9.18 ms without / 9.61 ms with, 2^22 26.8 / 24.9, 2^24 88.2 / 79.8;
// with the final level kernels 2^21 14.79 / 14.15 (profiles/r04c_ba_2p21.log).
static inline bool ba_enabled_for(size_t n_points) {
  const char* env = getenv("ALEO_B200_MSM_BA");  // read per call: tests and sweeps switch it
  if (env) return atol(env) > 0;
  return n_points >= ((size_t)1 << 21);
}

// GLV split (msm.cuh glv_split_kernel): plain MSMs below the batch-affine threshold run over 2 n virtual points with
// 127-bit scalars -- half the windows, so half the buckets to reduce and half the doubling chain of the tail, which is
// most of a proof-sized MSM (2^16: 1.65 of 2.6 ms).  Above the threshold the accumulation dominates and the levels
// gather the bases themselves.  ALEO_B200_MSM_GLV = 0 | 1 overrides.
constexpr u32 GLV_BITS = 127;
static inline bool glv_enabled_for(size_t n_points) {
  const char* env = getenv("ALEO_B200_MSM_GLV");  // read per call: tests switch it
  if (env) return atol(env) > 0 && !ba_enabled_for(n_points);
  // measured on B200 (profiles/r03r_glv_sizes.log, with / without): 2^12 1.45 / 1.99 ms, 2^16 2.25 / 2.62, 2^18 3.71 / 4.23,
  // 2^20 8.69 / 8.74, 2^21 15.19 / 14.61 -- the split doubles the points the sort and the gathers see
  return n_points <= ((size_t)1 << 20) && !ba_enabled_for(n_points);
}
static inline u32 windows_for_glv(u32 c) { return GLV_BITS / c + 1; }

// Window bits by cost model, in Fq products: n * W(c) mixed additions (10 each) of bucket accumulation
// against W(c) * 2^(c-1) buckets * 2 full additions (14 each, x1.5: the reduction runs at lower
// occupancy).  2^16 -> 12, 2^20 -> 16.  With the batch-affine levels an entry costs less and a bucket more (every
// bucket is copied through the levels and finished by an XYZZ stage of short runs): fitted on B200 at 2^24,
// 0.262 ns per entry + 1.18 ns per bucket in the accumulation and 1.3 .. 2.4 ns per bucket in the reduction = about 11
// entries per bucket: 2^22 -> 16, 2^24 -> 18 (79.8 ms against 81.2 / 81.0 ms for c = 17 / 20), 2^26 -> 20.
// ALEO_B200_MSM_C overrides (sweeps).
static inline u32 choose_window(size_t n, u32 chunks = 1) {
  const char* env = getenv("ALEO_B200_MSM_C");  // read per call: sweeps switch it
  const long c_env = env ? atol(env) : 0L;
  if (c_env >= 4 && c_env <= 22) return (u32)c_env;
  if (n < 2) return 4;
  const bool ba = ba_enabled_for(n / chunks);
  const bool glv = glv_enabled_for(n / chunks);
  // GLV sizes are latency bound in the reduction: 2.7 ns per bucket against 0.38 ns per entry (r03r: tails of 1.88 / 1.28 ms
  // for 262 k / 41 k buckets)
  const double per_bucket = ba ? 110.0 : (glv ? 80.0 : 2.0 * 14.0 * 1.5);
  u32 best_c = 4;
  double best = 0;
  for (u32 c = 4; c <= 22; c++) {
    const double W = glv ? (double)windows_for_glv(c) : (double)windows_for(c);
    // half-range scalars have 252 bits: only ceil(252 / c) windows receive entries, the one above holds the recoding
    // carry alone (c = 18: 14 windows of entries, c = 17: 15 -- why 18 beats 17 at 2^24); counted in the fitted model only
    const double We = ba ? (double)((SCALAR_BITS - 1 + c - 1) / c) : W;
    // an MSM that arrives in `chunks` point ranges re-opens every bucket once per extra range: one more
    // mixed addition per bucket and range (the first addition into an empty bucket is a copy); weighted 5 rather
    // than 10: measured on B200 at 2^24 in 3 ranges, c = 20 runs 102.3 ms against 105.5 ms for c = 19
    const double cost = (glv ? 2.0 : 1.0) * (double)n * We * 10.0 + W * (double)(1u << (c - 1)) * (per_bucket + 5.0 * (chunks - 1));
    if (c == 4 || cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// a resident SRS expanded by srs_expand_kernel: pre[w][i] = 2^(c w) * P_i, packed 96-byte affine
struct SrsView {
  const unsigned char* pre;
  u32 n_total;  // points per window block
  u32 c, W;
};

// Resident SRS: the doubling tail and the per-window bucket sets are gone (n * W mixed additions against ONE set of
// 2^(c-1) buckets), and the scalars are recoded half-range (252 bits), so W(c) = ceil(253 / c): c = 23 needs 11 windows.
static inline u32 windows_for_srs(u32 c) { return (SCALAR_BITS + c - 1) / c; }

static inline u32 choose_window_srs(size_t n) {
  if (n < 2) return 8;
  u32 best_c = 8;
  double best = 0;
  for (u32 c = 8; c <= 23; c++) {
    const double cost = (double)n * windows_for_srs(c) * 10.0 + (double)(1u << (c - 1)) * 2.0 * 14.0 * 1.5;
    if (c == 8 || cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// Accumulation threads (= equal runs of the sorted entry list) for `n_chunk` points.  At most 16 waves of
// (SM count x 3 CTAs x 128 threads; 148 SMs on B200) -- measured best on B200 at n = 2^24: 4 waves 96.1 ms, 16 waves 93.4 ms.  Below
// that, runs of 64 entries (every run boundary cuts a bucket whose pieces cost two more additions on the latency
// path: 2^20 runs 3 % faster with 64 than with 32), but never fewer than 4 waves while runs of 32 can fill them.
static inline u32 lanes_for_entries(size_t entries) {
  static const long waves_env = []() { const char* e = getenv("ALEO_B200_MSM_WAVES"); return e ? atol(e) : 0L; }();
  const size_t wave = (size_t)dev_props().sms * 384;
  const size_t full = wave * (waves_env > 0 ? (size_t)waves_env : 16);
  size_t lanes = (entries + 63) / 64;
  if (lanes < 4 * wave) {
    const size_t by32 = (entries + 31) / 32;
    lanes = by32 < 4 * wave ? by32 : 4 * wave;
    // ... in WHOLE waves: every lane does the same number of additions, so 1.44 waves of 32-entry runs take as long as
    // two full waves would (2^17: 82 k lanes); k = ceil(entries / (32 wave)) full waves of shorter runs take
    // entries / (k wave) instead.  Not below runs of 12: runs of 8 cut every bucket of 2^14's 64 entries into 8 pieces and the
    // combine doubles the phase (0.55 -> 1.15 ms).  Measured (profiles/r04g_whole_waves.log): 2^16 2.17 -> 2.07 ms, 2^17 2.89 ->
    // 2.81, 13 batched 2^18 commitments 20.5 -> 18.7 ms.  ALEO_B200_MSM_WHOLE_WAVES = 0 switches off.
    static const bool whole = []() { const char* e = getenv("ALEO_B200_MSM_WHOLE_WAVES"); return !(e && e[0] == '0'); }();
    if (whole && by32 > 0 && by32 < 4 * wave) {
      const size_t k = (by32 + wave - 1) / wave;
      const size_t aligned = k * wave;
      if ((entries + aligned - 1) / aligned >= 12) lanes = aligned;
    }
  }
  if (lanes > full) lanes = full;
  if (lanes < 128) lanes = 128;
  return (u32)((lanes + 127) / 128 * 128);
}
static inline u32 lanes_for(size_t n_chunk, u32 W) { return lanes_for_entries((size_t)n_chunk * W); }

// ---- batch-affine pair-tree levels (msm_ba.cuh) in front of the XYZZ accumulation ------------------------------------
// How many levels: after L levels a bucket that received m entries holds ceil(m / 2^L) points, which the XYZZ kernel
// sums (it is the stage that copes with buckets cut by run boundaries).  A level costs 0.26 .. 0.30 ns per addition
// against 0.33 ns in the XYZZ kernel, plus its scans and launch: measured flat within 1 % between log2(load) - 3 and
// log2(load) - 1 levels (load = average entries per bucket; 2^24: c = 20 (load 32) 3 / 4 levels 69.7 / 69.8 ms,
// c = 18 (load 128) 5 / 6 levels 68.3 / 69.3 ms; 2^22, c = 17 (load 64) 3 / 4 / 5 levels 19.6 / 19.8 / 20.0 ms).
// ALEO_B200_MSM_BA = 0 | levels overrides.
static inline u32 ba_levels_for(size_t points, size_t entries, size_t buckets) {
  const char* env = getenv("ALEO_B200_MSM_BA");  // read per call: tests and sweeps switch it
  if (env) {
    const long v = atol(env);
    return v < 0 ? 0u : (v > 12 ? 12u : (u32)v);
  }
  if (!ba_enabled_for(points)) return 0;
  const size_t load = entries / (buckets ? buckets : 1);
  u32 l = 0;
  while (((size_t)8 << l) <= load) l++;  // load 32 -> 3, 64 -> 4, 128 -> 5
  return l;
}
// levels >= 1 at four CTAs per SM (ba::CompactStager: operands re-read from the staging buffer; ALEO_B200_MSM_BA_COMPACT)
static inline bool ba_compact() {
  const char* env = getenv("ALEO_B200_MSM_BA_COMPACT");
  return env ? atoi(env) != 0 : true;  // 2^24: accumulate 67.7 -> 66.2 ms, 2^22: 20.6 -> 20.1 ms (profiles/r03z_ba_compact_sweep.log)
}
// additions one thread shares an inversion over, at most (ALEO_B200_MSM_BA_K)
static inline u32 ba_kmax() {
  const char* env = getenv("ALEO_B200_MSM_BA_K");
  const long v = env ? atol(env) : 0L;
  return (v >= 1 && v <= 4096) ? (u32)v : 256u;
}
// workspace the levels may take, in bytes (ALEO_B200_MSM_BA_MB): about 104 bytes per sorted entry of a group
static inline size_t ba_budget_bytes() {
  const char* env = getenv("ALEO_B200_MSM_BA_MB");
  const long v = env ? atol(env) : 0L;
  return (v > 0 ? (size_t)v : (size_t)40960) << 20;
}
struct BaLevelPlan {
  u32 k, nthreads, run, wave;  // additions per thread at most, threads, positions per thread, threads per wave
};
// level l of a group with at most `entries` sorted entries over `nb` buckets: its input has at most
// (entries >> l) + nb points (sum of ceil(m / 2^l)); a thread takes 2 k positions = at most k additions
static inline BaLevelPlan ba_level_plan(size_t entries, u32 nb, u32 l) {
  const size_t slots = (entries >> l) + (l ? nb : 0);
  // whole waves of 3 CTAs of 128 threads per SM (all runs take the same time: a partial last wave idles the chip), as
  // few as keep a run at <= kmax additions but at least 4 while runs of 16 additions can fill them
  const size_t wave = (size_t)dev_props().sms * ((l > 0 && ba_compact()) ? 4 : 3) * ba::TPB;
  const u32 kmax = ba_kmax();
  size_t waves = (slots / 2 + wave * kmax - 1) / (wave * kmax);
  const size_t kmin = kmax < 32 ? kmax : 32;
  if (waves < 4) {
    const size_t by16 = (slots / 2 + wave * kmin - 1) / (wave * kmin);
    waves = by16 < 4 ? by16 : 4;
  }
  BaLevelPlan pl;
  pl.nthreads = (u32)(waves * wave);
  if (slots / 2 < wave * kmin) {  // not even one wave of 16-addition runs: as many threads as there are such runs
    const size_t th = (slots / 2 + kmin - 1) / kmin;
    pl.nthreads = (u32)((th + ba::TPB - 1) / ba::TPB * ba::TPB);
    if (pl.nthreads == 0) pl.nthreads = ba::TPB;
  }
  const size_t run = (slots + pl.nthreads - 1) / pl.nthreads;  // positions per thread
  pl.k = (u32)((run + 1) / 2 + 1);                              // additions per thread, at most
  pl.run = (u32)(run ? run : 1);
  pl.wave = (u32)wave;
  return pl;
}

// n: points of the whole MSM (decides the window); chunks: how many point ranges it arrives in
static inline Params make_params(size_t n, const SrsView* srs = nullptr, u32 chunks = 1) {
  Params p;
  const bool glv = !srs && n > 0 && glv_enabled_for(n / (chunks ? chunks : 1));
  p.c = srs ? srs->c : choose_window(n, chunks);
  p.W = srs ? srs->W : (glv ? windows_for_glv(p.c) : windows_for(p.c));
  p.half_range = glv ? 0u : 1u;  // split scalars are not residues mod r
  p.glv_n = glv ? 1u : 0u;       // Session::add_chunk sets the range's point count
  p.glv_bx = nullptr;
  p.batch_off = nullptr;
  p.nbatch = 0;
  p.B = 1u << (p.c - 1);
  p.n_stride = srs ? srs->n_total : 0;
  p.first = 0;
  p.nlanes = lanes_for(glv ? 2 * n : n, p.W);
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
                                         cudaStream_t s, int& launches, u32 shift = 0, u32 copy_end = 0) {
  const u32 nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  LAUNCH(scan_block_sums_kernel, dim3(nblocks), dim3(SCAN_TPB), 0, s, in, n, block_sums, shift);
  LAUNCH(scan_sums_kernel, dim3(1), dim3(SCAN_TPB), 0, s, block_sums, nblocks, total);
  LAUNCH(scan_apply_kernel, dim3(nblocks), dim3(SCAN_TPB), 0, s, in, n, (const u32*)block_sums, out, out_copy, shift, copy_end);
  launches += 3;
  return cudaGetLastError();
}

// ALEO_B200_MSM_TRACE=1: finish() times its launches with events and prints them to stderr (tuning aid; synchronises)
struct TailTrace {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  void mark(const char* name, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
    names.push_back(name);
  }
  void report(cudaStream_t s) {
    if (!on) return;
    cudaStreamSynchronize(s);
    for (size_t i = 1; i < ev.size(); i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      fprintf(stderr, "[msm tail] %-14s %.3f ms\n", names[i], ms);
    }
    for (auto e : ev) cudaEventDestroy(e);
  }
};

struct RedLevel {
  u32 m, T, log_kc, scale_log;  // weighted input length per window, chunks out per window, chunk size, weight of WS
  bool cta;                     // computed by a CTA per chunk (reduce_level_cta_kernel) instead of a thread per chunk
};

// A level whose whole input (all bucket sets) has at most this many elements is computed CTA-cooperatively: the extra
// additions (about 4.5x) are cheaper than the serial chain of a chunk level while they still fit a fraction of a
// millisecond of the chip (2^17 elements x 9 additions ~ 0.5 ms).  ALEO_B200_MSM_REDUCE = chunk | cta | ctaK (chunks of at most 2^K: tests) forces one.
constexpr u64 CTA_LEVEL_MAX_ELEMS = (u64)1 << 17;

static inline u32 ceil_log2(u32 v) {
  u32 l = 0;
  while ((1u << l) < v) l++;
  return l;
}

// One MSM = begin() + add_chunk() per point range + finish().  The bucket set lives in the session's
// workspace across chunks: a chunk is sorted on its own and its bucket sums are ADDED to what the earlier
// chunks left, so that the host-pointer entry points can overlap the host->device copy of chunk k + 1
// with the accumulation of chunk k (capi.cu).  Device-resident callers use a single chunk.
// With dry = true nothing is allocated or launched; only the launch count is kept.
struct Session {
  Params prm;
  const SrsView* srs = nullptr;
  u32 nwin = 0, NB = 0, max_lanes = 0, cap_small = 0, cap_large = 0, scan_blocks = 0;
  size_t max_chunk = 0;
  std::vector<RedLevel> lv;
  u32 scan_m = 0, scan_t_in = 0, scan_scale_log = 0, scan_span = 32;  // last reduction stage (one CTA per window)
  // partitioned sort (plain MSM, large ranges): see part_count_kernel
  bool part = false;
  u32 part_fb = 0, part_bps = 0, part_nbin = 0, part_ncta_max = 0, part_wpg = 1;
  size_t o_pcnt = 0, o_poffs = 0, o_part = 0;
  size_t o_counts = 0, o_starts = 0, o_ends = 0, o_piece_bucket = 0, o_bsums = 0, o_meta = 0, o_sorted = 0, o_small = 0,
         o_large = 0, o_buckets = 0, o_pieces = 0, o_D = 0, bytes = 0;
  std::vector<size_t> o_R, o_P;
  // batch-affine levels: L levels per group of `ba_wpg` bucket sets (the workspace bounds a group's entries)
  u32 ba_L = 0, ba_wpg = 0;
  size_t ba_scratch_ops = 0;
  bool glv = false;  // scalars split by u^2, 2 n virtual points (make_params)
  bool ba_denied = false;  // the levels' workspace did not fit the device: this session runs the XYZZ kernel alone
  size_t o_glv_sc = 0, o_glv_bx = 0;
  size_t o_baA = 0, o_baB = 0, o_ba_pre = 0, o_ba_rec = 0, o_ba_s0 = 0, o_ba_s1 = 0, o_ba_e = 0, o_meta2 = 0;
  unsigned char* ws = nullptr;
  bool dry = false;
  int launches = 0;
  u32 chunks_done = 0;

  template <class T>
  T* at(size_t off) const { return reinterpret_cast<T*>(ws + off); }

  // n_total: points of the whole MSM; max_chunk: largest point range add_chunk() will see; chunks: how many
  // batch (resident SRS only): nbatch independent MSMs whose scalar vectors lie back to back, batch_off = nbatch + 1
  // prefix offsets on the device; every member gets its own bucket set and finish_batch() returns nbatch results
  cudaError_t begin(size_t n_total, size_t max_chunk_, u32 chunks, const SrsView* srs_, cudaStream_t s, bool dry_,
                    u32 nbatch = 0, const u32* batch_off = nullptr) {
    srs = srs_;
    dry = dry_;
    max_chunk = max_chunk_;
    prm = make_params(n_total, srs, chunks);
    prm.nbatch = nbatch;
    prm.batch_off = batch_off;
    glv = prm.glv_n != 0;
    nwin = nbatch ? nbatch : (srs ? 1u : prm.W);  // bucket sets (the resident SRS shares one across all windows)
    NB = nwin * prm.B;
    max_lanes = lanes_for(glv ? 2 * max_chunk : max_chunk, prm.W);
    cap_small = max_lanes + 1;  // every run boundary cuts at most one bucket
    cap_large = max_lanes / SMALL_SPLIT_MAX + 1;
    scan_blocks = (NB + SCAN_BLOCK - 1) / SCAN_BLOCK;
    {
      // chunk level l: weighted input of length m (buckets, then R[1..) of the level below), T chunks out; the
      // plain input P of the level below has T_prev entries and needs ceil(T_prev / Kc) chunks too.  Chunk
      // levels run until the scan stage (<= SCAN_MAX elements per window) can take over.
      // Elements per window the last stage (one CTA per window) takes.  64 instead of 256 for small bucket sets -- the
      // CTA-cooperative levels would spread the additions of that stage over more SMs -- measured a wash on B200
      // (profiles/r03v_scan_target.log: 2^16 2.08 against 2.17 ms, 2^15 1.89 against 1.76, resident-SRS commit 2^18 3.04
      // against 2.89): 256 stays.  ALEO_B200_MSM_SCAN_TARGET = 32 | 64 | 128 | 256 overrides.
      u32 scan_target = SCAN_MAX;
      if (const char* te = getenv("ALEO_B200_MSM_SCAN_TARGET")) {
        const long v = atol(te);
        if (v == 32 || v == 64 || v == 128 || v == 256) scan_target = (u32)v;
      }
      u32 m = prm.B, t_prev = 0, scale_log = 0;
      for (;;) {
        const u32 need = t_prev > m ? t_prev : m;
        if (need <= scan_target) break;
        RedLevel l;
        l.m = m;
        const char* renv = getenv("ALEO_B200_MSM_REDUCE");  // read per call: tests force both paths
        l.log_kc = ceil_log2((need + scan_target - 1) / scan_target);
        // ... and only with chunks of >= 32 elements: a CTA of 8 threads still costs whole warp instructions (plain MSM at
        // 2^16, 22 windows x 256 chunks of 8: 0.61 ms against 0.24 ms for the serial chunk levels; the resident SRS's
        // single set of 2^14 .. 2^16 buckets makes chunks of 64 .. 256: 2^16 commit 1.67 -> 1.28 ms, 2^18 3.13 -> 2.86 ms)
        l.cta = renv ? (renv[0] == 'c' && renv[1] == 't') : ((u64)nwin * need <= CTA_LEVEL_MAX_ELEMS && l.log_kc >= 5);
        if (l.cta) {
          if (l.log_kc > 8) l.log_kc = 8;  // one thread per element, at most SCAN_MAX threads
          if (renv && renv[3] >= '1' && renv[3] <= '8' && l.log_kc > (u32)(renv[3] - '0')) l.log_kc = (u32)(renv[3] - '0');  // tests: "cta2" caps the chunk at 4
          if (l.log_kc < 1) l.log_kc = 1;
          l.scale_log = scale_log;
          l.T = (need + (1u << l.log_kc) - 1) >> l.log_kc;
          lv.push_back(l);
          scale_log += l.log_kc;
          t_prev = l.T;
          m = l.T - 1;
          continue;
        }
        // a chunk is serial inside its thread (2 * Kc additions): chunks of 16 while they still fill the chip, chunks of
        // 4 once a level would run on a few warps per SM and its time is the chain, not the work (ncu at 2^22: the second
        // level of 16 ran 3840 threads for 0.82 ms at 26 % of the pipe -- as long as the first with 1/16 of its work)
        const u32 cap = ((u64)nwin * ((need + 15) >> 4) < 32768u) ? 2u : 4u;
        if (l.log_kc > cap) l.log_kc = cap;
        l.scale_log = scale_log;
        l.T = (need + (1u << l.log_kc) - 1) >> l.log_kc;
        lv.push_back(l);
        scale_log += l.log_kc;
        t_prev = l.T;
        m = l.T - 1;
      }
      scan_m = m;
      scan_t_in = t_prev;
      scan_scale_log = scale_log;
      scan_span = 32;
      while (scan_span < (scan_m > scan_t_in ? scan_m : scan_t_in)) scan_span <<= 1;
    }
    Carver cv;
    o_counts = cv.take((size_t)NB * 4);
    o_starts = cv.take((size_t)NB * 4);
    o_ends = cv.take((size_t)NB * 4);
    o_piece_bucket = cv.take((size_t)max_lanes * 2 * 4);
    o_bsums = cv.take((size_t)(scan_blocks + 1) * 4);
    o_meta = cv.take(64);
    o_sorted = cv.take((size_t)max_chunk * (glv ? 2 : 1) * prm.W * 4);
    if (glv) {
      o_glv_sc = cv.take((size_t)max_chunk * 2 * 32);
      o_glv_bx = cv.take((size_t)max_chunk * 48);
    }
    o_small = cv.take((size_t)cap_small * 4);
    o_large = cv.take((size_t)cap_large * 4);
    o_buckets = cv.take((size_t)NB * sizeof(G1Xyzz));
    o_pieces = cv.take((size_t)max_lanes * 2 * sizeof(G1Xyzz));
    o_R.resize(lv.size());
    o_P.resize(lv.size());
    for (size_t l = 0; l < lv.size(); l++) {
      o_R[l] = cv.take((size_t)nwin * lv[l].T * sizeof(G1Xyzz));
      o_P[l] = cv.take((size_t)nwin * lv[l].T * sizeof(G1Xyzz));
    }
    o_D = cv.take((size_t)nwin * sizeof(G1Xyzz));  // S_w: one sum per window
    {
      // ALEO_B200_MSM_SORT = p selects the partitioned sort.  NOT the default: measured on B200 it loses to the atomic
      // sort -- 2^22 (c = 17): 2.6 against 1.6 ms; 2^24 (c = 20): 29.7 against 6.9 ms, because the top window of c = 20
      // holds 13 bits, so 4 of its 512 bins receive all 2^24 of its entries and 4 CTAs of part_finish walk 4 M
      // records each.  Kept (and tested) as the starting point for a version that slices heavy bins.
      const char* sort_env = getenv("ALEO_B200_MSM_SORT");
      const bool forced = sort_env && sort_env[0] == 'p';
      part = !srs && !nbatch && forced;
      if (part) {
        part_fb = prm.c - 1 > 9 ? prm.c - 1 - 9 : 0;
        if (const char* fbe = getenv("ALEO_B200_MSM_PART_FB")) part_fb = (u32)atoi(fbe) < prm.c ? (u32)atoi(fbe) : part_fb;  // tests
        part_bps = prm.B >> part_fb;
        part_nbin = prm.W * part_bps;
        part_wpg = PART_GROUP_BINS / part_bps ? PART_GROUP_BINS / part_bps : 1;
        const size_t ctas = ((glv ? 2 : 1) * max_chunk + PART_TPB - 1) / PART_TPB;  // scalars the sort sees (GLV: 2 n)
        part_ncta_max = (u32)(ctas < (size_t)dev_props().sms * 4 ? ctas : (size_t)dev_props().sms * 4);
        if ((size_t)part_nbin * 4 > (size_t)160 * 1024) part = false;  // histogram of part_count must fit shared memory
      }
      if (part) {
        o_pcnt = cv.take((size_t)part_nbin * part_ncta_max * 4);
        o_poffs = cv.take((size_t)part_nbin * part_ncta_max * 4);
        o_part = cv.take((size_t)max_chunk * (glv ? 2 : 1) * prm.W * sizeof(PartRecord));
        const size_t sb = ((size_t)part_nbin * part_ncta_max + SCAN_BLOCK - 1) / SCAN_BLOCK + 1;
        if (sb > scan_blocks + 1) o_bsums = cv.take(sb * 4);  // the scan of cnt needs more block sums than the bucket scan
      }
    }
    {
      // a group = whole bucket sets whose entries fit the budget; the shared set of a resident SRS and the member sets
      // of a batch (whose sizes the host does not know) form one group
      const size_t total_entries = (size_t)max_chunk * prm.W;
      size_t cap_entries = ba_budget_bytes() / 104;
      if (cap_entries > ((size_t)3 << 29)) cap_entries = (size_t)3 << 29;  // output slots of a level are 30-bit fields of its records
      ba_L = (glv || ba_denied) ? 0u : ba_levels_for(max_chunk, total_entries, NB);  // the level kernel does not know GLV's virtual points
      size_t ge = total_entries;
      if (srs) {
        ba_wpg = nwin;
        if (ge > cap_entries) ba_L = 0;
      } else {
        const size_t fit = cap_entries / (max_chunk ? max_chunk : 1);
        ba_wpg = fit < nwin ? (u32)fit : nwin;
        if (ba_wpg == 0) ba_L = 0;
        // as few groups as fit, of equal size (one group at 2^24; 13 windows at 2^26 that fit 6 by 6 become 5 + 5 + 3):
        // every group pays its levels' launches and partial last waves (2^24 in two groups: accumulate 68.3 -> 72.9 ms)
        else ba_wpg = (nwin + ((nwin + ba_wpg - 1) / ba_wpg) - 1) / ((nwin + ba_wpg - 1) / ba_wpg);
        ge = (size_t)max_chunk * ba_wpg;
      }
      if (ba_L) {
        const u32 gnb = ba_wpg * prm.B;
        size_t scratch_ops = 0;
        for (u32 w0 = 0; w0 < nwin; w0 += ba_wpg) {  // the groups add_chunk() will form (the last one may be smaller)
          const u32 nw = (w0 + ba_wpg <= nwin) ? ba_wpg : nwin - w0;
          for (u32 l = 0; l < ba_L; l++) {
            const BaLevelPlan pl = ba_level_plan(srs ? ge : (size_t)max_chunk * nw, nw * prm.B, l);
            const size_t ops = (size_t)pl.nthreads * pl.k;
            if (ops > scratch_ops) scratch_ops = ops;
          }
        }
        o_baA = cv.take(((ge >> 1) + gnb + 1) * 96);
        o_baB = ba_L > 1 ? cv.take(((ge >> 2) + gnb + 1) * 96) : 0;
        ba_scratch_ops = scratch_ops;
        o_ba_pre = cv.take(scratch_ops * 48);
        o_ba_rec = cv.take(scratch_ops * 16);
        o_ba_s0 = cv.take((size_t)gnb * 4);
        o_ba_s1 = cv.take((size_t)gnb * 4);
        o_ba_e = cv.take((size_t)gnb * 4);
        o_meta2 = cv.take(64);
      }
    }
    bytes = cv.off;
    if (dry) return cudaSuccess;
    {
      // ALEO_B200_TEST_DENY_BA_ALLOC: the tests' stand-in for a device without room for the levels' workspace
      const bool deny = ba_L > 0 && getenv("ALEO_B200_TEST_DENY_BA_ALLOC") != nullptr;
      cudaError_t ea = deny ? cudaErrorMemoryAllocation : aleo::pool_malloc_async((void**)&ws, bytes, s);
      if (ea == cudaErrorMemoryAllocation && ba_L > 0) {
        // the batch-affine levels want ~104 bytes per sorted entry; on a device that cannot spare them the MSM still runs,
        // on the XYZZ kernel alone (same result, the round-1 speed), instead of failing
        (void)cudaGetLastError();
        ws = nullptr;
        ba_denied = true;
        lv.clear();
        launches = 0;
        return begin(n_total, max_chunk_, chunks, srs_, s, dry_, nbatch, batch_off);
      }
      MSM_CK(ea);
    }
    MSM_CK(cudaMemsetAsync(at<G1Xyzz>(o_buckets), 0, (size_t)NB * sizeof(G1Xyzz), s));
    return cudaSuccess;
  }

  // Sorts one point range by (window, bucket) and adds its bucket sums into the session's buckets.
  // bases: the range's own base array (entry indices are range-local), or ignored for a resident SRS,
  // where `first` is the range's offset into the SRS.  phase_ev (optional, 3 events): before the sort,
  // before and after the accumulation kernel.
  cudaError_t add_chunk(const unsigned char* bases, u32 stride, const u32* scalars, size_t n_chunk, size_t first,
                        cudaStream_t s, cudaEvent_t* phase_ev = nullptr) {
    if (n_chunk == 0) return cudaSuccess;
    if (n_chunk > max_chunk) return cudaErrorInvalidValue;
    u32 n = (u32)n_chunk;
    Params p = prm;
    p.first = (u32)first;
    p.nlanes = lanes_for(glv ? 2 * n_chunk : n_chunk, p.W);
    if (p.nlanes > max_lanes) p.nlanes = max_lanes;  // whole-wave rounding is not monotone in the range size; the piece arrays are sized for max_lanes
    if (srs) {
      bases = srs->pre;
      stride = 96;
    }
    const u32 into = chunks_done > 0 ? 1u : 0u;  // later chunks start every bucket from its stored sum
    chunks_done++;
    {
      // sort (count, 3 scan kernels, scatter; the partitioned sort has one more) + per accumulation: plan, XYZZ kernel,
      // two combine kernels; with batch-affine levels every group of bucket sets adds (3 scan kernels + 1 level) per level
      const u32 groups = ba_L ? (nwin + ba_wpg - 1) / ba_wpg : 1;
      launches += (part ? 6 : 5) + (int)groups * (4 + 4 * (int)ba_L) + (glv ? 1 : 0);
    }
    if (dry) return cudaSuccess;
    if (phase_ev) cudaEventRecord(phase_ev[0], s);
    if (glv) {  // k = k1 + k2 u^2: from here on the MSM has 2 n virtual points and 127-bit scalars
      u32* sc2 = at<u32>(o_glv_sc);
      unsigned char* bx = at<unsigned char>(o_glv_bx);
      LAUNCH_NOSYNC(glv_split_kernel, dim3((n + 127) / 128), dim3(128), 0, s, scalars, n, bases, stride, sc2, bx);
      p.glv_n = n;
      p.glv_bx = bx;
      scalars = sc2;
      n *= 2;
    }
    u32* counts = at<u32>(o_counts);
    u32* starts = at<u32>(o_starts);
    u32* ends = at<u32>(o_ends);
    u32* piece_bucket = at<u32>(o_piece_bucket);
    u32* meta = at<u32>(o_meta);
    u32* sorted = at<u32>(o_sorted);
    u32* small_list = at<u32>(o_small);
    u32* large_list = at<u32>(o_large);
    G1Xyzz* buckets = at<G1Xyzz>(o_buckets);
    G1Xyzz* pieces = at<G1Xyzz>(o_pieces);
    const u32 NB = this->NB, cap_small = this->cap_small, cap_large = this->cap_large;  // plain values for the launch macros
    MSM_CK(cudaMemsetAsync(counts, 0, (size_t)NB * 4, s));
    MSM_CK(cudaMemsetAsync(meta, 0, 64, s));
    const u32 g_all = (n + 255) / 256, g_n = g_all < dev_props().sms * 8 ? g_all : dev_props().sms * 8;  // <= 8 CTAs of 256 per SM and window
    TailTrace st;
    st.on = getenv("ALEO_B200_MSM_TRACE") != nullptr;
    st.mark("start", s);
    if (part) {
      PartArgs pa;
      pa.scalars = scalars;
      pa.n = n;
      pa.prm = p;
      pa.fb = part_fb;
      pa.bps = part_bps;
      pa.nbin = part_nbin;
      pa.ncta = g_all < dev_props().sms * 4 ? g_all : dev_props().sms * 4;
      pa.wpg = part_wpg;
      u32* pcnt = at<u32>(o_pcnt);
      u32* poffs = at<u32>(o_poffs);
      PartRecord* precs = at<PartRecord>(o_part);
      const u32 ngroups = (p.W + pa.wpg - 1) / pa.wpg;
      const u32 nfine = 1u << pa.fb;
#ifndef ALEO_EMU
      static thread_local int attr_dev = -1;
      int devnow = 0;
      MSM_CK(cudaGetDevice(&devnow));
      if (attr_dev != devnow) {
        MSM_CK(cudaFuncSetAttribute(part_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_dev = devnow;
      }
#endif
      LAUNCH(part_count_kernel, dim3(pa.ncta), dim3(PART_TPB), (size_t)pa.nbin * 4, s, pa, pcnt);
      st.mark("part count", s);
      int scan_launches = 0;
      MSM_CK(exclusive_scan(pcnt, pa.nbin * pa.ncta, at<u32>(o_bsums), poffs, nullptr, meta + 0, s, scan_launches));
      st.mark("part scan", s);
      LAUNCH(part_scatter_kernel, dim3(pa.ncta, ngroups), dim3(PART_TPB), (size_t)pa.wpg * pa.bps * 4, s, pa, (const u32*)poffs, precs);
      st.mark("part scatter", s);
      LAUNCH(part_finish_kernel, dim3(pa.nbin), dim3(PART_TPB), (size_t)(nfine + PART_TPB) * 4, s, pa, (const u32*)poffs,
             (const u32*)(meta + 0), (const PartRecord*)precs, starts, ends, sorted);
    } else {
    // count: scalar-major (every warp iteration spreads its 32 atomics over one window's counters and moves on);
    // scatter: window-major once the bucket heads of all windows (W * 2^(c-1) sectors of 32 B) outgrow L2 --
    // measured on B200 at n = 2^24, c = 20: count 3.3 (scalar) / 4.5 (window) ms, scatter 9.2 / 5.3 ms;
    // at c = 16: count 4.0 / 5.6, scatter 4.2 / 5.8.  ALEO_B200_MSM_SORT = s | w forces one order for both.
    const char* sort_env = getenv("ALEO_B200_MSM_SORT");
    const bool heads_fit_l2 = (size_t)NB * 32 <= l2_resident_budget();
    const bool count_wm = sort_env && sort_env[0] == 'w';
    const bool scatter_wm = sort_env ? sort_env[0] == 'w' : (!heads_fit_l2 && !srs);  // a resident SRS shares one bucket set: nothing to gain
    if (count_wm)
      LAUNCH_NOSYNC(count_kernel, dim3(g_n, p.W), dim3(256), 0, s, scalars, n, p, counts);
    else
      LAUNCH_NOSYNC(count_kernel_sm, dim3(g_all), dim3(256), 0, s, scalars, n, p, counts);
    st.mark("count", s);
    int scan_launches = 0;
    MSM_CK(exclusive_scan(counts, NB, at<u32>(o_bsums), starts, ends, meta + 0, s, scan_launches));
    st.mark("scan", s);
    if (scatter_wm)
      LAUNCH_NOSYNC(scatter_kernel, dim3(g_n, p.W), dim3(256), 0, s, scalars, n, p, ends, sorted);
    else {
      // single bucket set (resident SRS) whose heads outgrow L2: bucket-range passes (power of two, <= 8)
      u32 passes = 1;
      if (srs && !prm.nbatch && !sort_env)
        while (passes < 8 && ((size_t)p.B * 32) / passes > l2_resident_budget() && (p.B / passes) > 1) passes <<= 1;
      if (const char* pe = getenv("ALEO_B200_MSM_SRS_PASSES"))  // tests: force the number of bucket-range passes
        if (srs && !prm.nbatch && (atoi(pe) == 2 || atoi(pe) == 4 || atoi(pe) == 8) && p.B >= 8) passes = (u32)atoi(pe);
      LAUNCH_NOSYNC(scatter_kernel_sm, dim3(g_all, passes), dim3(256), 0, s, scalars, n, p, ends, sorted);
    }
    }
    st.mark("scatter", s);
    st.report(s);
    if (phase_ev) cudaEventRecord(phase_ev[1], s);
    // ALEO_B200_MSM_ACC=i: inlined field products instead of the two out-of-line functions (A/B switch behind the
    // I-cache finding of DESIGN.md: 85.0 against 72.9 ms at 2^24); n: no prefetch of the next base (73.00 -> 72.71 ms)
    static const bool acc_inline = []() { const char* e = getenv("ALEO_B200_MSM_ACC"); return e && e[0] == 'i'; }();
    static const bool acc_prefetch = []() { const char* e = getenv("ALEO_B200_MSM_ACC"); return !(e && e[0] == 'n'); }();
    TailTrace tr;
    tr.on = st.on;
    tr.mark("start", s);
    // plan + XYZZ accumulation + combine over one bucket range: `pts` / `ent` are the bases and the sorted entry list,
    // or (direct) the dense level array the batch-affine levels left and no entry list
    auto flat = [&](const unsigned char* pts, u32 pstride, const u32* ent, const u32* st_, const u32* en_, u32 nb, u32 lanes,
                    u32* mt, G1Xyzz* bk, bool direct) -> cudaError_t {
      LAUNCH_NOSYNC(plan_pieces_kernel, dim3((nb + 255) / 256), dim3(256), 0, s, st_, en_, nb, lanes, small_list, large_list, cap_small,
                    cap_large, mt);
#define ACC_LAUNCH(...)                                                                                                   \
  LAUNCH_NOSYNC((accumulate_kernel<__VA_ARGS__>), dim3(lanes / 128), dim3(128), 0, s, pts, pstride, ent, st_, en_, nb, lanes, \
                (const u32*)mt, bk, pieces, piece_bucket, into, direct ? 0u : p.glv_n, p.glv_bx)
      if (direct)
        ACC_LAUNCH(true, false, true);
      else if (acc_inline)
        ACC_LAUNCH(false);
      else if (acc_prefetch)
        ACC_LAUNCH(true, true);
      else
        ACC_LAUNCH(true);
#undef ACC_LAUNCH
      tr.mark(direct ? "xyzz (direct)" : "accumulate", s);
      LAUNCH_NOSYNC(combine_small_kernel, dim3((lanes + 1 + 127) / 128), dim3(128), 0, s, (const u32*)small_list, st_, en_, lanes,
                    (const u32*)mt, (const G1Xyzz*)pieces, (const u32*)piece_bucket, bk);
      tr.mark("combine small", s);
      const u32 cl = lanes / SMALL_SPLIT_MAX + 1;
      const u32 g = cl < dev_props().sms * 4 ? cl : dev_props().sms * 4;  // 4 CTAs per SM; the kernel strides over the list
      LAUNCH(combine_large_kernel, dim3(g), dim3(COMBINE_TPB), 0, s, (const u32*)large_list, st_, en_, lanes, (const u32*)mt,
             (const G1Xyzz*)pieces, (const u32*)piece_bucket, bk);
      tr.mark("combine large", s);
      return cudaGetLastError();
    };
    if (!ba_L) {
      MSM_CK(flat(bases, stride, sorted, starts, ends, NB, p.nlanes, meta, buckets, false));
      if (phase_ev) cudaEventRecord(phase_ev[2], s);
    } else {
#ifndef ALEO_EMU
      {  // the staging buffers of the level kernel need the large shared-memory carve-out (once per device and thread)
        static thread_local int attr_dev = -1;
        int devnow = 0;
        MSM_CK(cudaGetDevice(&devnow));
        if (attr_dev != devnow) {
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<true, ba::AsyncStagerShifted>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba::STAGE_BYTES));
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<false, ba::AsyncStager>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba::STAGE_BYTES));
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<false, ba::CompactStager, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba::COMPACT_BYTES));
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<false, ba::CompactStager, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<true, ba::AsyncStagerShifted>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          MSM_CK(cudaFuncSetAttribute(ba::level_kernel<false, ba::AsyncStager>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          attr_dev = devnow;
        }
      }
#endif
      u32* meta2 = at<u32>(o_meta2);
      u32* sa[2] = {at<u32>(o_ba_s0), at<u32>(o_ba_s1)};
      u32* e_out = at<u32>(o_ba_e);
      unsigned char* lvbuf[2] = {at<unsigned char>(o_baA), o_baB ? at<unsigned char>(o_baB) : nullptr};
      for (u32 w0 = 0; w0 < nwin; w0 += ba_wpg) {
        const u32 nw = (w0 + ba_wpg <= nwin) ? ba_wpg : nwin - w0;
        const u32 g0 = w0 * p.B, nb = nw * p.B;
        const size_t ge = srs ? (size_t)n * p.W : (size_t)n * nw;  // entries of this group, at most
        MSM_CK(cudaMemsetAsync(meta2, 0, 64, s));
        ba::LevelArgs la;
        la.bases = bases;
        la.stride = stride;
        la.sorted = sorted;
        la.in = nullptr;
        la.start_in = starts + g0;
        la.cnt0 = counts + g0;
        la.nb = nb;
        la.pre = at<uint4>(o_ba_pre);
        la.rec = at<uint4>(o_ba_rec);
        for (u32 l = 0; l < ba_L; l++) {
          const bool last = (l + 1 == ba_L);
          int sl = 0;
          MSM_CK(exclusive_scan(counts + g0, nb, at<u32>(o_bsums), sa[l & 1], last ? e_out : nullptr, meta2 + 0, s, sl, l + 1,
                                last ? 1u : 0u));
          const BaLevelPlan pl = ba_level_plan(ge, nb, l);
          if ((size_t)pl.nthreads * pl.k > ba_scratch_ops) return cudaErrorInvalidValue;  // cannot happen: begin() sized it over these groups
          la.level = l;
          la.start_out = sa[l & 1];
          la.out = lvbuf[l & 1];
          la.nthreads = pl.nthreads;
          la.run_target = pl.run;
          {
            const char* ce = getenv("ALEO_B200_MSM_BA_CA");  // bit 0: level 0, bit 1: levels >= 1 (A/B switch)
            const int cm = ce ? atoi(ce) : 0;
            la.copy_via_l1 = (l == 0) ? (cm & 1) : ((cm >> 1) & 1);
          }
          la.wave = pl.wave;
          // levels >= 1: operands staged through shared memory (cp.async.cg); level 0 gathers the caller's bases with plain
          // loads -- staged it is slower whichever way (2^24, accumulate phase: 67.9 ms plain; 84.4 ms with 16-byte .cg
          // copies of the enclosing words, 77.9 ms with .ca, 42.2 ms for the level alone with 8-byte copies against 35.1:
          // a gathered 96-byte point is two L1 lines that plain loads fetch once, an L2-only copy asks for every 16 bytes).
          // ALEO_B200_MSM_BA_STAGE = 0: plain loads everywhere, 2: staged everywhere (bases 16-byte aligned) -- A/B switch;
          // ALEO_B200_MSM_BA_CA bit 0 / 1: .ca copies at level 0 / above (levels >= 1: 70.7 ms against 67.9 with .cg)
          const char* stage_env = getenv("ALEO_B200_MSM_BA_STAGE");
          const int stage_mode = stage_env ? atoi(stage_env) : 1;
          const bool bases_ok = ((size_t)bases & 15u) == 0 && (stride & 7u) == 0;
          const bool staged = l > 0 ? stage_mode >= 1 : (stage_mode >= 2 && bases_ok);
          const dim3 grid(pl.nthreads / ba::TPB), block(ba::TPB);
          if (l == 0 && staged)
            LAUNCH_NOSYNC((ba::level_kernel<true, ba::AsyncStagerShifted>), grid, block, ba::STAGE_BYTES, s, la);
          else if (l == 0)
            LAUNCH_NOSYNC((ba::level_kernel<true, ba::DirectStager>), grid, block, 0, s, la);
          else if (staged && ba_compact())
            LAUNCH_NOSYNC((ba::level_kernel<false, ba::CompactStager, 4>), grid, block, ba::COMPACT_BYTES, s, la);
          else if (staged)
            LAUNCH_NOSYNC((ba::level_kernel<false, ba::AsyncStager>), grid, block, ba::STAGE_BYTES, s, la);
          else
            LAUNCH_NOSYNC((ba::level_kernel<false, ba::DirectStager>), grid, block, 0, s, la);
          tr.mark("affine level", s);
          la.in = la.out;
          la.start_in = la.start_out;
        }
        u32 lanes = lanes_for_entries((ge >> ba_L) + nb);
        if (lanes > max_lanes) lanes = max_lanes;
        MSM_CK(flat(lvbuf[(ba_L - 1) & 1], 96, nullptr, sa[(ba_L - 1) & 1], e_out, nb, lanes, meta2, buckets + g0, true));
      }
      if (phase_ev) cudaEventRecord(phase_ev[2], s);
    }
    tr.report(s);
    return cudaGetLastError();
  }

  // bucket reduction + window combination + normalisation -> 144-byte Jacobian; frees the workspace.
  // Batch sessions: `out144` receives nbatch results, 48-byte compressed points when batch_compressed, else 144 bytes each.
  cudaError_t finish(unsigned char* out144, cudaStream_t s, cudaEvent_t* done_ev = nullptr, bool batch_compressed = true) {
    launches += (int)lv.size() + 2;
    if (dry) return cudaSuccess;
    G1Xyzz* buckets = at<G1Xyzz>(o_buckets);
    const u32 nwin = this->nwin, c = prm.c;  // plain values for the launch macros
    TailTrace tr;
    tr.on = getenv("ALEO_B200_MSM_TRACE") != nullptr;
    tr.mark("start", s);
    for (size_t l = 0; l < lv.size(); l++) {
      ReduceArgs ra;
      ra.X = (l == 0) ? buckets : at<G1Xyzz>(o_R[l - 1]);
      ra.x_stride = (l == 0) ? prm.B : lv[l - 1].T;
      ra.x_off = (l == 0) ? 0 : 1;
      ra.m = lv[l].m;
      ra.P = (l == 0) ? nullptr : at<G1Xyzz>(o_P[l - 1]);
      ra.T_in = (l == 0) ? 0 : lv[l - 1].T;
      ra.log_kc = lv[l].log_kc;
      ra.scale_log = lv[l].scale_log;
      ra.R_out = at<G1Xyzz>(o_R[l]);
      ra.P_out = at<G1Xyzz>(o_P[l]);
      ra.T_out = lv[l].T;
      ra.nwin = nwin;
      const u32 threads = nwin * lv[l].T;
      if (lv[l].cta) {
        const u32 kc = 1u << lv[l].log_kc;
        LAUNCH(reduce_level_cta_kernel, dim3(threads), dim3(kc), kc * sizeof(G1Xyzz), s, ra);
        tr.mark("cta level", s);
      } else {
        LAUNCH_NOSYNC(reduce_level_kernel, dim3((threads + 127) / 128), dim3(128), 0, s, ra);
        tr.mark("chunk level", s);
      }
    }
    G1Xyzz* S = at<G1Xyzz>(o_D);
    {
      const size_t L = lv.size();
      ScanArgs sa;
      sa.X = L ? at<G1Xyzz>(o_R[L - 1]) : buckets;
      sa.x_stride = L ? lv[L - 1].T : prm.B;
      sa.x_off = L ? 1 : 0;
      sa.m = scan_m;
      sa.P = L ? at<G1Xyzz>(o_P[L - 1]) : nullptr;
      sa.T_in = scan_t_in;
      sa.scale_log = scan_scale_log;
      sa.S = S;
      const u32 span = scan_span;
      LAUNCH(reduce_scan_kernel, dim3(nwin), dim3(span), span * sizeof(G1Xyzz), s, sa);
      tr.mark("scan", s);
    }
    if (prm.nbatch) {
      LAUNCH(batch_finalize_kernel, dim3(nwin), dim3(32), 0, s, (const G1Xyzz*)S, nwin, out144, (u32)(batch_compressed ? 1 : 0));
      tr.mark("batch finalize", s);
      tr.report(s);
      if (done_ev) cudaEventRecord(*done_ev, s);
      cudaError_t eb = cudaGetLastError();
      release(s);
      return eb;
    }
    // per-window weights 2^(c w) (none for a resident SRS: they are baked into the expanded bases), sum, normalise
    unsigned long long* dbg = nullptr;
    if (tr.on && cudaMalloc((void**)&dbg, 32) != cudaSuccess) dbg = nullptr;
    LAUNCH(weigh_sum_kernel, dim3(1), dim3(TAIL_TPB), 0, s, (const G1Xyzz*)S, nwin, c, out144, dbg);
    tr.mark("weigh+sum+inv", s);
    tr.report(s);
    if (dbg) {
      unsigned long long h[4] = {0, 0, 0, 0};
      cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[msm tail] clocks: doubling chain %llu (%u doublings), window tree %llu, normalise %llu\n", h[1] - h[0],
              c * (nwin - 1), h[2] - h[1], h[3] - h[2]);
      cudaFree(dbg);
    }
    if (done_ev) cudaEventRecord(*done_ev, s);
    cudaError_t e = cudaGetLastError();
    release(s);
    return e;
  }

  void release(cudaStream_t s) {
    if (ws) cudaFreeAsync(ws, s);
    ws = nullptr;
  }
};

// Runs (or, with dry = true, only counts the launches of) one MSM over device-resident operands.
// Everything is asynchronous on `s` except the workspace allocation call itself.
// phase_ev (optional, 4 events): recorded before the sort phase, before / after the bucket
// accumulation kernel, and after the final kernel -- bench.py's per-kernel timing.
static inline cudaError_t run(const unsigned char* bases, u32 stride, const u32* scalars, size_t n_sz,
                              unsigned char* out144, cudaStream_t s, bool dry, int* launches_out,
                              cudaEvent_t* phase_ev = nullptr, const SrsView* srs = nullptr) {
  if (n_sz == 0) {
    if (!dry) LAUNCH_NOSYNC(write_identity_kernel, dim3(1), dim3(1), 0, s, out144);
    if (launches_out) *launches_out = 1;
    return dry ? cudaSuccess : cudaGetLastError();
  }
  Session ss;
  cudaError_t e = ss.begin(n_sz, n_sz, 1, srs, s, dry);
  if (e == cudaSuccess) e = ss.add_chunk(bases, stride, scalars, n_sz, 0, s, phase_ev);
  if (e == cudaSuccess) e = ss.finish(out144, s, phase_ev ? &phase_ev[3] : nullptr);
  else ss.release(s);
  if (launches_out) *launches_out = ss.launches;
  return e;
}

}  // namespace msm
