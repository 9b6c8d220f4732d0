// ntt_host.cuh -- host-side plan (pass split, twiddle tables) and launcher for ntt.cuh.
#pragma once
#include <mutex>
#include <map>
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

static inline cudaError_t build_plan(Plan& p, int device, u32 log_n, bool inverse, bool coset, cudaStream_t s) {
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
  if (log_n <= (u32)SMALL_MAX_LOG) {
    NTT_CK(plan_alloc(p, &p.small_inner, (size_t)1 << log_n));
    NTT_CK(fill_pow_table(p.small_inner, w, one, 1u << log_n, 0, s));
    NTT_CK(cudaMemcpyAsync(&p.scale_host, ninv, sizeof(Fr), cudaMemcpyDeviceToHost, s));
  } else {
    p.npass = split_passes((int)log_n, p.K);
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
      log_cur -= (u32)p.K[i];
    }
  }
  return cudaStreamSynchronize(s);
}

static inline void destroy_plan(Plan& p) {
  for (void* d : p.owned) cudaFree(d);
  p.owned.clear();
}

template <int K, bool LAST>
static inline cudaError_t launch_pass(const PassArgs& a, u32 grid, u32 batch, cudaStream_t s) {
  const size_t smem = 2 * sizeof(uint4) * tile_plane_elems<K, LAST>();
#ifndef ALEO_EMU
  static thread_local int attr_done_dev = -1;  // per instantiation, per thread: cheap and race free
  int dev = 0;
  NTT_CK(cudaGetDevice(&dev));
  if (attr_done_dev != dev) {
    NTT_CK(cudaFuncSetAttribute(pass_kernel<K, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done_dev = dev;
  }
#endif
  LAUNCH((pass_kernel<K, LAST>), dim3(grid, batch), dim3(TPB), smem, s, a);
  return cudaGetLastError();
}

template <bool LAST>
static inline cudaError_t launch_pass_k(int K, const PassArgs& a, u32 grid, u32 batch, cudaStream_t s) {
  switch (K) {
    case 5: return launch_pass<5, LAST>(a, grid, batch, s);
    case 6: return launch_pass<6, LAST>(a, grid, batch, s);
    case 7: return launch_pass<7, LAST>(a, grid, batch, s);
    case 8: return launch_pass<8, LAST>(a, grid, batch, s);
  }
  return cudaErrorInvalidValue;
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
      const u32 grid = (u32)(n >> TILE_LOG);
      a.use_pre = ((i == 0) && p.coset && !p.inverse) ? 1 : 0;
      a.use_post = (last && p.coset && p.inverse) ? 1 : 0;
      if (pass_ev && i == 0) cudaEventRecord(pass_ev[0], s);
      NTT_CK(last ? launch_pass_k<true>(p.K[i], a, grid, nb, s) : launch_pass_k<false>(p.K[i], a, grid, nb, s));
      if (pass_ev) cudaEventRecord(pass_ev[i + 1], s);
      log_cur -= (u32)p.K[i];
    }
  }
  return cudaSuccess;
}

// ---- plan cache (immutable after construction; guarded by one mutex) -----------------------------
class PlanCache {
 public:
  cudaError_t get(int device, u32 log_n, bool inverse, bool coset, cudaStream_t s, const Plan** out) {
    std::lock_guard<std::mutex> lk(mu_);
    auto key = std::make_tuple(device, log_n, inverse, coset);
    auto it = plans_.find(key);
    if (it == plans_.end()) {
      Plan p;
      cudaError_t e = build_plan(p, device, log_n, inverse, coset, s);
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
  std::map<std::tuple<int, u32, bool, bool>, Plan> plans_;
};

}  // namespace ntt
