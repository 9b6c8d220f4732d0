// ntt_host.cuh -- host-side plan (pass split, twiddle tables) and launcher for ntt.cuh.
#pragma once
#include <cstdlib>
#include <mutex>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "ntt.cuh"

namespace ntt {

#define NTT_CK(expr)                       \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return _e;      \
  } while (0)

// split L bits into P = ceil(L/8) passes, larger radices first (so that the last pass's tile
// width G = 2048/R never exceeds the first pass's radix, and M >= G holds for strided passes)
static inline int split_passes(int L, int* K) {
  const int P = (L + MAX_PASS_BITS - 1) / MAX_PASS_BITS;
  const int base = L / P, rem = L % P;
  for (int i = 0; i < P; i++) K[i] = base + (i < rem ? 1 : 0);
  return P;
}

constexpr u32 DIRECT_TW_MAX_LOG = 18;
constexpr u32 DIRECT0_MIN_LOG = 12, DIRECT0_MAX_LOG = 24;

static inline void set_single_gpu(PassArgs& a) {
  a.d_k2l = 31;
  a.d_klast = 0;
  a.d_rank_bits = 0;
  a.d_log_chunk = 0;
  a.d_exchange = 0;
  a.d_rank = 0;
  a.d_lg = 0;
  for (int i = 0; i < 8; i++) a.peer[i] = nullptr;
  for (int i = 0; i < 8; i++) a.d_flag_peer[i] = nullptr;
  a.d_flag_local = nullptr;
  a.d_epoch = 0;
  a.d_tile_rot = 0;
  a.pf_ahead = 0;
}

// cross-rank flags of one distributed transform (PassArgs d_flag_*): where this rank signals, where it waits
struct DistFlags {
  u32* signal[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // slot [rank] of every rank's array
  const u32* local = nullptr;                                                                  // this rank's slots
  u32 epoch = 0;
};

// Pass split of a transform spread over 2^lg GPUs (ntt_dist_*): every radix 2^5..2^8, and
//   K_p + (bits below digit p) - lg >= 11 for every strided pass (its tile is 2^(11 - K_p) local columns wide),
//   K_0 + K_last - lg >= 11 for the last pass (its tile takes 2^(11 - K_last) local values of the first digit),
//   K_last >= lg (every rank holds at least one value of the last digit).
// Returns the number of passes (2..4) or 0 when the size cannot be spread that far.
static inline bool split_ok(const int* k, int P, int lg) {
  bool ok = k[P - 1] >= lg && k[0] + k[P - 1] - lg >= TILE_LOG;
  int below = 0;
  for (int i = P - 1; i >= 1 && ok; i--) {
    below += k[i];
    ok = k[i - 1] + below - lg >= TILE_LOG;
  }
  return ok;
}

static inline int split_passes_dist(int L, int lg, int* K) {
  for (int P = (L + MAX_PASS_BITS - 1) / MAX_PASS_BITS; P <= 4; P++) {
    if (P < 2) continue;
    int best[4] = {0, 0, 0, 0}, best_spread = 99, k[4] = {0, 0, 0, 0};
    const int nk = MAX_PASS_BITS - 5 + 1;
    int lim = 1;
    for (int i = 0; i < P; i++) lim *= nk;
    for (int code = 0; code < lim; code++) {
      int sum = 0, c = code;
      for (int i = 0; i < P; i++) {
        k[i] = 5 + c % nk;
        c /= nk;
        sum += k[i];
      }
      if (sum != L || !split_ok(k, P, lg)) continue;
      int mx = 0, mn = 99;
      for (int i = 0; i < P; i++) {
        mx = k[i] > mx ? k[i] : mx;
        mn = k[i] < mn ? k[i] : mn;
      }
      // most balanced split; among equals the one with the larger first radix (code order visits small k[0] first)
      if (mx - mn <= best_spread) {
        best_spread = mx - mn;
        for (int i = 0; i < P; i++) best[i] = k[i];
      }
    }
    if (best_spread != 99) {
      for (int i = 0; i < P; i++) K[i] = best[i];
      return P;
    }
  }
  return 0;
}

// ALEO_B200_NTT_SPLIT="7,9,5": forces the pass split of sizes with that digit sum (tests: a 9-bit middle pass at 2^21)
static inline int forced_split(int L, int* K) {
  const char* e = getenv("ALEO_B200_NTT_SPLIT");
  if (!e) return 0;
  int k[4] = {0, 0, 0, 0}, n = 0, sum = 0;
  for (const char* p = e; *p && n < 4;) {
    k[n] = atoi(p);
    sum += k[n];
    n++;
    while (*p && *p != ',') p++;
    if (*p == ',') p++;
  }
  if (n < 2 || sum != L) return 0;
  for (int i = 0; i < n; i++) {
    if (k[i] < 5 || k[i] > MAX_PASS_BITS) return 0;
    K[i] = k[i];
  }
  return split_ok(k, n, 0) ? n : 0;
}

struct Plan {
  int device = -1;
  u32 log_n = 0;
  bool inverse = false, coset = false;
  int npass = 0;
  int K[4] = {0, 0, 0, 0};
  Fr* consts = nullptr;          // [w, n^-1, g (or g^-1), 1]
  Fr* inner[4] = {nullptr, nullptr, nullptr, nullptr};
  Fr* small_inner = nullptr;     // n <= 2^11: w^j, j < n
  Fr *tw_lo = nullptr, *tw_hi = nullptr, *tw_hi_scaled = nullptr;
  Fr* tw_direct[4] = {nullptr, nullptr, nullptr, nullptr};  // per pass i >= 1 (not the last): w^(x << (L - log_cur_i)), x < 2^log_cur_i
  Fr *cs_lo = nullptr, *cs_hi = nullptr;  // coset powers (inverse: hi carries n^-1)
  u32 lo_bits = 0;
  Fr scale_host;                 // n^-1 (Montgomery) for the small kernel
  std::vector<void*> owned;
};

static inline cudaError_t plan_alloc(Plan& p, Fr** out, size_t count) {
  void* d = nullptr;
  NTT_CK(cudaMalloc(&d, count * sizeof(Fr)));
  p.owned.push_back(d);
  *out = (Fr*)d;
  return cudaSuccess;
}

static inline cudaError_t fill_pow_table(Fr* out, const Fr* base, const Fr* scale, u32 count, u32 shift, cudaStream_t s) {
  LAUNCH_NOSYNC(pow_table_kernel, dim3((count + 127) / 128), dim3(128), 0, s, out, base, scale, count, shift);
  return cudaGetLastError();
}

// dist_lg > 0: the plan of a transform spread over 2^dist_lg GPUs (same tables, the pass split of split_passes_dist)
static inline cudaError_t build_plan(Plan& p, int device, u32 log_n, bool inverse, bool coset, cudaStream_t s, int dist_lg = 0) {
  p.device = device;
  p.log_n = log_n;
  p.inverse = inverse;
  p.coset = coset;
  NTT_CK(plan_alloc(p, &p.consts, 4));
  LAUNCH_NOSYNC(domain_consts_kernel, dim3(1), dim3(1), 0, s, p.consts, log_n, (u32)(inverse ? 1 : 0));
  NTT_CK(cudaGetLastError());
  const Fr* w = p.consts + 0;
  const Fr* ninv = p.consts + 1;
  const Fr* cg = p.consts + 2;
  const Fr* one = p.consts + 3;
  p.lo_bits = (log_n + 1) / 2;
  const u32 hi_bits = log_n - p.lo_bits;
  if (coset) {
    NTT_CK(plan_alloc(p, &p.cs_lo, (size_t)1 << p.lo_bits));
    NTT_CK(plan_alloc(p, &p.cs_hi, (size_t)1 << hi_bits));
    NTT_CK(fill_pow_table(p.cs_lo, cg, one, 1u << p.lo_bits, 0, s));
    NTT_CK(fill_pow_table(p.cs_hi, cg, inverse ? ninv : one, 1u << hi_bits, p.lo_bits, s));
  }
  if (log_n <= (u32)SMALL_MAX_LOG && dist_lg == 0) {
    NTT_CK(plan_alloc(p, &p.small_inner, (size_t)1 << log_n));
    NTT_CK(fill_pow_table(p.small_inner, w, one, 1u << log_n, 0, s));
    NTT_CK(cudaMemcpyAsync(&p.scale_host, ninv, sizeof(Fr), cudaMemcpyDeviceToHost, s));
  } else {
    p.npass = dist_lg ? split_passes_dist((int)log_n, dist_lg, p.K) : forced_split((int)log_n, p.K);
    if (p.npass == 0 && !dist_lg) p.npass = split_passes((int)log_n, p.K);
    if (p.npass == 0) return cudaErrorInvalidValue;
    NTT_CK(plan_alloc(p, &p.tw_lo, (size_t)1 << p.lo_bits));
    NTT_CK(plan_alloc(p, &p.tw_hi, (size_t)1 << hi_bits));
    NTT_CK(fill_pow_table(p.tw_lo, w, one, 1u << p.lo_bits, 0, s));
    NTT_CK(fill_pow_table(p.tw_hi, w, one, 1u << hi_bits, p.lo_bits, s));
    if (inverse && !coset) {
      NTT_CK(plan_alloc(p, &p.tw_hi_scaled, (size_t)1 << hi_bits));
      NTT_CK(fill_pow_table(p.tw_hi_scaled, w, ninv, 1u << hi_bits, p.lo_bits, s));
    }
    u32 log_cur = log_n;
    for (int i = 0; i < p.npass; i++) {
      NTT_CK(plan_alloc(p, &p.inner[i], (size_t)1 << p.K[i]));
      NTT_CK(fill_pow_table(p.inner[i], w, one, 1u << p.K[i], log_n - p.K[i], s));
      // direct inter-pass twiddles for the middle passes while the table stays L2 sized (<= 2^18 entries = 8 MB)
      if (i >= 1 && i + 1 < p.npass && log_cur <= DIRECT_TW_MAX_LOG) {
        NTT_CK(plan_alloc(p, &p.tw_direct[i], (size_t)1 << log_cur));
        NTT_CK(fill_pow_table(p.tw_direct[i], w, one, 1u << log_cur, log_n - log_cur, s));
      }
      // The FIRST boundary has n distinct twiddles.  Up to 2^24 they come from a direct table too (L2 resident up to
      // 2^20: 2^16 0.054 -> 0.051 ms, 2^20 0.213 -> 0.202 ms).  For 2^21 <= n <= 2^24 they are streamed
      // (n x 32 B in HBM per plan, read once per transform: +32 B per element on a pass that uses 13 % of the HBM
      // roof) instead of one table product per element: 2^24 3.80 -> 3.59 ms, 2^22 0.956 -> 0.902 ms on B200; no gain
      // at 2^20; above 2^24 (9-bit first pass, 4-element runs) the scattered table reads LOSE: 2^25 3.26 -> 3.97 ms.  ALEO_B200_NTT_DIRECT0=0 switches it off (memory).
      u32 direct0_min = DIRECT0_MIN_LOG;
      if (const char* em = getenv("ALEO_B200_NTT_DIRECT0_MIN")) direct0_min = (u32)atoi(em);  // sweeps
      if (i == 0 && p.npass >= 2 && log_n >= direct0_min && log_n <= DIRECT0_MAX_LOG) {
        const char* e0 = getenv("ALEO_B200_NTT_DIRECT0");
        if (!(e0 && e0[0] == '0')) {
          NTT_CK(plan_alloc(p, &p.tw_direct[0], (size_t)1 << log_n));
          NTT_CK(fill_pow_table(p.tw_direct[0], w, (inverse && !coset) ? ninv : one, 1u << log_n, 0, s));
        }
      }
      log_cur -= (u32)p.K[i];
    }
  }
  return cudaStreamSynchronize(s);
}

static inline void destroy_plan(Plan& p) {
  for (void* d : p.owned) cudaFree(d);
  p.owned.clear();
}

template <int K, bool LAST, int TL>
static inline cudaError_t launch_pass(const PassArgs& a, u32 grid, u32 batch, cudaStream_t s) {
  const size_t smem = 2 * sizeof(uint4) * tile_plane_elems<K, LAST, TL>();
#ifndef ALEO_EMU
  static thread_local int attr_done_dev = -1;  // per instantiation, per thread: cheap and race free
  int dev = 0;
  NTT_CK(cudaGetDevice(&dev));
  if (attr_done_dev != dev) {
    NTT_CK(cudaFuncSetAttribute(pass_kernel<K, LAST, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done_dev = dev;
  }
#endif
  LAUNCH((pass_kernel<K, LAST, TL>), dim3(grid, batch), dim3((1u << TL) / 8), smem, s, a);
  return cudaGetLastError();
}

// tile_log: TILE_LOG (2048-element tiles) or SMALL_TILE_LOG; grid = local elements >> tile_log
template <bool LAST>
static inline cudaError_t launch_pass_k(int K, const PassArgs& a, u32 grid, u32 batch, cudaStream_t s, int tile_log = TILE_LOG) {
  if (tile_log == SMALL_TILE_LOG) {
    switch (K) {
      case 5: return launch_pass<5, LAST, SMALL_TILE_LOG>(a, grid, batch, s);
      case 6: return launch_pass<6, LAST, SMALL_TILE_LOG>(a, grid, batch, s);
      case 7: return launch_pass<7, LAST, SMALL_TILE_LOG>(a, grid, batch, s);
      case 8: return launch_pass<8, LAST, SMALL_TILE_LOG>(a, grid, batch, s);
      case 9: return launch_pass<9, LAST, SMALL_TILE_LOG>(a, grid, batch, s);
    }
    return cudaErrorInvalidValue;
  }
  switch (K) {
    case 5: return launch_pass<5, LAST, TILE_LOG>(a, grid, batch, s);
    case 6: return launch_pass<6, LAST, TILE_LOG>(a, grid, batch, s);
    case 7: return launch_pass<7, LAST, TILE_LOG>(a, grid, batch, s);
    case 8: return launch_pass<8, LAST, TILE_LOG>(a, grid, batch, s);
    case 9: return launch_pass<9, LAST, TILE_LOG>(a, grid, batch, s);
  }
  return cudaErrorInvalidValue;
}

// tile size of a single-GPU transform (ALEO_B200_NTT_TILE=10|11 forces one: A/B sweeps and tests)
static inline int tile_log_for(u32 log_n) {
  if (const char* e = getenv("ALEO_B200_NTT_TILE")) {
    const int v = atoi(e);
    if (v == SMALL_TILE_LOG || v == TILE_LOG) return v;
  }
  return log_n <= (u32)SMALL_TILE_MAX_LOG ? SMALL_TILE_LOG : TILE_LOG;
}

// number of kernel launches one transform of this plan issues (bench.py's gpu_launches)
static inline int launches_per_transform(const Plan& p) { return p.log_n <= (u32)SMALL_MAX_LOG ? 1 : p.npass; }

// scratch elements run() needs for `batch` transforms of 2^log_n (0 for the single-CTA sizes)
static inline size_t scratch_batch(u32 log_n, size_t batch) {
  if (log_n <= (u32)SMALL_MAX_LOG) return 0;
  size_t cap = ((size_t)1 << 25) >> log_n;  // <= 2^25 elements (1 GiB) of scratch per launch group
  if (cap < 1) cap = 1;
  if (cap > 65535) cap = 65535;             // gridDim.y limit
  return batch < cap ? batch : cap;
}

// data: batch x n elements, in place.  scratch: scratch_batch(log_n, batch) x n elements.
// pass_ev (optional, npass + 1 events, batch must be 1): recorded around every pass.
static inline cudaError_t run(const Plan& p, Fr* data, size_t batch, Fr* scratch, cudaStream_t s,
                              cudaEvent_t* pass_ev = nullptr) {
  if (p.log_n == 0) return cudaSuccess;  // size-1 transform is the identity (also for coset: g^0 = 1)
  const size_t n = (size_t)1 << p.log_n;
  if (p.log_n <= (u32)SMALL_MAX_LOG) {
    SmallArgs a;
    a.data = data;
    a.log_n = p.log_n;
    a.inner = p.small_inner;
    a.pre = PowTable{p.cs_lo, p.cs_hi, p.lo_bits};
    a.post = a.pre;
    a.scale = p.scale_host;
    a.use_pre = (p.coset && !p.inverse) ? 1 : 0;
    a.use_post = (p.coset && p.inverse) ? 1 : 0;
    a.use_scale = (!p.coset && p.inverse) ? 1 : 0;
    const size_t smem = 2 * sizeof(uint4) * n;
#ifndef ALEO_EMU
    static thread_local int attr_done_dev = -1;
    int dev = 0;
    NTT_CK(cudaGetDevice(&dev));
    if (attr_done_dev != dev) {
      NTT_CK(cudaFuncSetAttribute(small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(uint4) << SMALL_MAX_LOG)));
      attr_done_dev = dev;
    }
#endif
    const u32 threads = (u32)(n / 2 < 32 ? 32 : (n / 2 > (size_t)TPB ? (size_t)TPB : n / 2));
    if (pass_ev) cudaEventRecord(pass_ev[0], s);
    LAUNCH(small_kernel, dim3((u32)batch), dim3(threads), smem, s, a);
    if (pass_ev) cudaEventRecord(pass_ev[1], s);
    return cudaGetLastError();
  }
  const size_t group = scratch_batch(p.log_n, batch);
  for (size_t b = 0; b < batch; b += group) {
    Fr* x = data + b * n;
    const u32 nb = (u32)(batch - b < group ? batch - b : group);
    u32 log_cur = p.log_n;
    for (int i = 0; i < p.npass; i++) {
      const bool last = (i == p.npass - 1);
      PassArgs a;
      set_single_gpu(a);
      a.src = (i == 0) ? x : scratch;
      a.dst = last ? x : scratch;
      a.log_n = p.log_n;
      a.log_cur = log_cur;
      a.log_r1 = (u32)p.K[0];
      a.log_r2 = p.npass >= 3 ? (u32)p.K[1] : 0;
      a.log_r3 = p.npass >= 4 ? (u32)p.K[2] : 0;
      a.inner = p.inner[i];
      a.tw = PowTable{p.tw_lo, (i == 0 && p.tw_hi_scaled) ? p.tw_hi_scaled : p.tw_hi, p.lo_bits};
      a.tw_direct = p.tw_direct[i];
      a.pre = PowTable{p.cs_lo, p.cs_hi, p.lo_bits};
      a.post = a.pre;
      const int tl = tile_log_for(p.log_n);
      const u32 grid = (u32)(n >> tl);
      a.use_pre = ((i == 0) && p.coset && !p.inverse) ? 1 : 0;
      a.use_post = (last && p.coset && p.inverse) ? 1 : 0;
      {
        // ALEO_B200_NTT_PREFETCH = tiles ahead (experiment); read per call: sweeps flip it
        const char* pe = getenv("ALEO_B200_NTT_PREFETCH");
        a.pf_ahead = pe ? (u32)atoi(pe) : 0u;
      }
      if (pass_ev && i == 0) cudaEventRecord(pass_ev[0], s);
      NTT_CK(last ? launch_pass_k<true>(p.K[i], a, grid, nb, s, tl) : launch_pass_k<false>(p.K[i], a, grid, nb, s, tl));
      if (pass_ev) cudaEventRecord(pass_ev[i + 1], s);
      log_cur -= (u32)p.K[i];
    }
  }
  return cudaSuccess;
}

// ---- one transform over 2^lg GPUs ---------------------------------------------------------------------------------
// Rank j holds the column block of the (N / R_last) x R_last row-major matrix (local[a * R_last / g + b] =
// x[a * R_last + j * R_last / g + b]).  Stage 1 = passes 0 .. P-2 on the local N / g elements (twiddles by global
// index); the last of them stores chunk t of its result straight into rank t's receive buffer.  After a cross-rank
// barrier stage 2 = the last pass from the receive buffer; rank t ends with X[k], k mod R_0 in its range, as the column
// block of the (N / R_0) x R_0 matrix.
static inline cudaError_t run_dist_stage1(const Plan& p, int lg, int rank, const Fr* in, Fr* scratch, Fr* const* peers,
                                          cudaStream_t s, const DistFlags* fl = nullptr, cudaEvent_t* ev2 = nullptr) {
  const u32 L = p.log_n, Ll = L - (u32)lg;
  const u32 klast = (u32)p.K[p.npass - 1];
  u32 log_cur = L;
  for (int i = 0; i + 1 < p.npass; i++) {
    const bool exchange = (i + 2 == p.npass);
    PassArgs a;
    a.src = (i == 0) ? in : scratch;
    a.dst = scratch;
    a.log_n = Ll;
    a.log_cur = log_cur - (u32)lg;
    a.log_r1 = a.log_r2 = a.log_r3 = 0;
    a.inner = p.inner[i];
    a.tw = PowTable{p.tw_lo, (i == 0 && p.tw_hi_scaled) ? p.tw_hi_scaled : p.tw_hi, p.lo_bits};
    a.tw_direct = p.tw_direct[i];
    a.pre = PowTable{p.cs_lo, p.cs_hi, p.lo_bits};
    a.post = a.pre;
    a.use_pre = (i == 0 && p.coset && !p.inverse) ? 1 : 0;
    a.use_post = 0;
    a.d_lg = (u32)lg;
    a.d_k2l = klast - (u32)lg;
    a.d_klast = klast;
    a.d_rank_bits = (u32)rank << a.d_k2l;
    a.d_log_chunk = L - 2 * (u32)lg;
    a.d_exchange = exchange ? 1 : 0;
    a.d_rank = (u32)rank;
    for (int r = 0; r < 8; r++) a.peer[r] = (r < (1 << lg)) ? peers[r] : nullptr;
    for (int r = 0; r < 8; r++) a.d_flag_peer[r] = (exchange && fl && r < (1 << lg)) ? fl->signal[r] : nullptr;
    a.d_flag_local = nullptr;
    a.d_epoch = fl ? fl->epoch : 0;
    const u32 tiles = (u32)(((size_t)1 << Ll) >> TILE_LOG);
    static const bool no_rot = getenv("ALEO_B200_NTT_DIST_NOROT") != nullptr;  // A/B switch: every rank starts with rank 0's tiles
    a.pf_ahead = 0;
    a.d_tile_rot = (exchange && !no_rot) ? (u32)(((u64)((rank + 1) & ((1 << lg) - 1)) * tiles) >> lg) : 0u;
    if (ev2 && i == 0) cudaEventRecord(ev2[0], s);        // profiling: [0] start, [1] before the exchange pass
    if (ev2 && exchange) cudaEventRecord(ev2[1], s);
    NTT_CK(launch_pass_k<false>(p.K[i], a, tiles, 1, s));
    if (exchange && fl) {
      LAUNCH_NOSYNC(dist_signal_kernel, dim3(1), dim3(32), 0, s, a);
      NTT_CK(cudaGetLastError());
    }
    log_cur -= (u32)p.K[i];
  }
  return cudaSuccess;
}

static inline cudaError_t run_dist_stage2(const Plan& p, int lg, int rank, const Fr* recv, Fr* out, cudaStream_t s,
                                          const DistFlags* fl = nullptr) {
  const u32 L = p.log_n, Ll = L - (u32)lg;
  const int last = p.npass - 1;
  PassArgs a;
  a.src = recv;
  a.dst = out;
  a.log_n = Ll;
  a.log_cur = (u32)p.K[last];
  a.log_r1 = (u32)p.K[0] - (u32)lg;
  a.log_r2 = p.npass >= 3 ? (u32)p.K[1] : 0;
  a.log_r3 = p.npass >= 4 ? (u32)p.K[2] : 0;
  a.inner = p.inner[last];
  a.tw = PowTable{p.tw_lo, p.tw_hi, p.lo_bits};
  a.tw_direct = nullptr;
  a.pre = PowTable{p.cs_lo, p.cs_hi, p.lo_bits};
  a.post = a.pre;
  a.use_pre = 0;
  a.use_post = (p.coset && p.inverse) ? 1 : 0;
  a.d_lg = (u32)lg;
  a.d_k2l = (u32)p.K[last] - (u32)lg;
  a.d_klast = (u32)p.K[last];
  a.d_rank_bits = (u32)rank << a.d_k2l;
  a.d_log_chunk = L - 2 * (u32)lg;
  a.d_exchange = 0;
  a.d_rank = (u32)rank;
  for (int r = 0; r < 8; r++) a.peer[r] = nullptr;
  for (int r = 0; r < 8; r++) a.d_flag_peer[r] = nullptr;
  a.d_flag_local = fl ? fl->local : nullptr;
  a.d_epoch = fl ? fl->epoch : 0;
  a.d_tile_rot = 0;
  a.pf_ahead = 0;
  return launch_pass_k<true>(p.K[last], a, (u32)(((size_t)1 << Ll) >> TILE_LOG), 1, s);
}

// ---- plan cache (immutable after construction; guarded by one mutex) -----------------------------
class PlanCache {
 public:
  cudaError_t get(int device, u32 log_n, bool inverse, bool coset, cudaStream_t s, const Plan** out, int dist_lg = 0) {
    std::lock_guard<std::mutex> lk(mu_);
    // build_plan also reads tuning switches from the environment (sweeps and tests flip them inside one process): they
    // are part of the key, so a plan built under one setting is never handed out under another
    std::string env_sig;
    for (const char* name : {"ALEO_B200_NTT_SPLIT", "ALEO_B200_NTT_DIRECT0", "ALEO_B200_NTT_DIRECT0_MIN"}) {
      const char* v = getenv(name);
      env_sig += v ? v : "";
      env_sig += '|';
    }
    auto key = std::make_tuple(device, log_n, inverse, coset, dist_lg, env_sig);
    auto it = plans_.find(key);
    if (it == plans_.end()) {
      Plan p;
      cudaError_t e = build_plan(p, device, log_n, inverse, coset, s, dist_lg);
      if (e != cudaSuccess) {
        destroy_plan(p);
        return e;
      }
      it = plans_.emplace(key, std::move(p)).first;
    }
    *out = &it->second;
    return cudaSuccess;
  }
  void clear() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : plans_) destroy_plan(kv.second);
    plans_.clear();
  }

 private:
  std::mutex mu_;
  std::map<std::tuple<int, u32, bool, bool, int, std::string>, Plan> plans_;
};

}  // namespace ntt
