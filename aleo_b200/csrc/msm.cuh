// msm.cuh -- BLS12-377 G1 variable-base multi-scalar multiplication (Pippenger) for sm_100a.
//
// Stands in for snarkvm_algorithms::msm::VariableBase::msm::<G1Affine> (snarkvm-algorithms 0.14.5
// src/msm/variable_base/{mod,batched,standard}.rs), called by KZG10::commit /
// commit_lagrange / open and reached from the reference at rust/src/program/execute.rs:74,177.
//
// Pipeline (all on device, nothing returns to the host until the 144-byte result):
//   1 count     signed-window recoding of every scalar (digits in [-2^(c-1), 2^(c-1)]), histogram
//               of (window, |digit|-1) with L2 atomics                                  [HBM/atomic]
//   2 scan      exclusive scan of the histogram -> bucket start offsets                 [tiny]
//   3 scatter   recode again, claim a slot per (window, bucket) with an atomic cursor, write
//               point index | sign<<31   (order inside a bucket is irrelevant: + commutes) [HBM]
//   4 tasks     buckets longer than T entries are cut into tasks of <= T entries (skewed
//               scalar distributions: KZG witnesses are full of 0 / 1) + scan of task counts
//   5 accumulate one thread per task: XYZZ mixed additions of gathered affine bases      [IMAD]
//               -- this is >95 % of the work: n * W additions of 8M + 2S in Fq
//   6 combine   partial sums of split buckets are added (thread per bucket, or CTA per bucket)
//   7 reduce    per window sum_b (b+1) * bucket[b] by chunked running sums, recursively
//   8 final     per window 2^(c w) * S_w (windows in parallel), sum, one inversion -> normalised Jacobian
#pragma once
#include "g1.cuh"

namespace msm {

constexpr u32 SCALAR_BITS = 253;
constexpr u32 RED_LOG_KC = 5;            // bucket-reduction chunk: 32 buckets per thread
constexpr u32 RED_KC = 1u << RED_LOG_KC;
constexpr u32 SMALL_SPLIT_MAX = 64;      // split buckets with <= 64 tasks are combined by one thread
constexpr u32 COMBINE_TPB = 128;

struct Params {
  u32 c;       // window bits
  u32 W;       // number of windows = 253 / c + 1 (always absorbs the recoding carry)
  u32 B;       // buckets per window = 2^(c-1)
  u32 T;       // longest run of entries one accumulation task handles
};

// signed digit of window w given the carry from window w-1; returns |digit| and updates carry/neg
DEV u32 recode_digit(const u32* s /* 8 limbs in global memory */, u32 w, u32 c, u32& carry, u32& neg) {
  const u32 bit = w * c;
  const u32 limb = bit >> 5, sh = bit & 31u;
  u32 raw = 0;
  if (limb < 8) {
    raw = s[limb] >> sh;
    if (sh + c > 32 && limb + 1 < 8) raw |= s[limb + 1] << (32 - sh);
  }
  raw &= (1u << c) - 1u;
  const u32 d = raw + carry;
  if (d > (1u << (c - 1))) {
    neg = 1;
    carry = 1;
    return (1u << c) - d;
  }
  neg = 0;
  carry = 0;
  return d;
}

KERNEL void count_kernel(const u32* scalars, u32 n, Params prm, u32* counts) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32* s = scalars + (size_t)i * 8;
  u32 carry = 0, neg = 0;
  for (u32 w = 0; w < prm.W; w++) {
    const u32 mag = recode_digit(s, w, prm.c, carry, neg);
    if (mag) atomic_add_u32(&counts[w * prm.B + (mag - 1)], 1u);
  }
}

KERNEL void scatter_kernel(const u32* scalars, u32 n, Params prm, u32* cursor, u32* sorted) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32* s = scalars + (size_t)i * 8;
  u32 carry = 0, neg = 0;
  for (u32 w = 0; w < prm.W; w++) {
    const u32 mag = recode_digit(s, w, prm.c, carry, neg);
    if (mag) {
      const u32 pos = atomic_add_u32(&cursor[w * prm.B + (mag - 1)], 1u);
      sorted[pos] = i | (neg << 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of u32 (three launches; 2048 items per CTA)
// ---------------------------------------------------------------------------------------------
constexpr u32 SCAN_TPB = 256;
constexpr u32 SCAN_ITEMS = 8;
constexpr u32 SCAN_BLOCK = SCAN_TPB * SCAN_ITEMS;

// block-wide inclusive scan of one value per thread through shared memory
DEV u32 block_inclusive_scan(u32 v, u32* sh /* >= blockDim.x */) {
  const u32 t = threadIdx.x;
  sh[t] = v;
  SYNC_THREADS();
  for (u32 off = 1; off < blockDim.x; off <<= 1) {
    u32 add = (t >= off) ? sh[t - off] : 0u;
    SYNC_THREADS();
    sh[t] += add;
    SYNC_THREADS();
  }
  return sh[t];
}

KERNEL void scan_block_sums_kernel(const u32* in, u32 n, u32* block_sums) {
  SHARED u32 sh[SCAN_TPB];
  const u32 base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  u32 sum = 0;
  for (u32 k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) sum += in[base + k];
  const u32 incl = block_inclusive_scan(sum, sh);
  if (threadIdx.x == blockDim.x - 1) block_sums[blockIdx.x] = incl;
}

// single CTA: exclusive scan of block_sums in place, grand total to *total
KERNEL void scan_sums_kernel(u32* block_sums, u32 nblocks, u32* total) {
  SHARED u32 sh[SCAN_TPB];
  SHARED u32 running;
  if (threadIdx.x == 0) running = 0;
  SYNC_THREADS();
  for (u32 base = 0; base < nblocks; base += blockDim.x) {
    const u32 i = base + threadIdx.x;
    const u32 v = (i < nblocks) ? block_sums[i] : 0u;
    const u32 incl = block_inclusive_scan(v, sh);
    const u32 run = running;
    if (i < nblocks) block_sums[i] = run + incl - v;
    SYNC_THREADS();
    if (threadIdx.x == blockDim.x - 1) running = run + incl;
    SYNC_THREADS();
  }
  if (threadIdx.x == 0) *total = running;
}

KERNEL void scan_apply_kernel(const u32* in, u32 n, const u32* block_sums, u32* out, u32* out_copy) {
  SHARED u32 sh[SCAN_TPB];
  const u32 base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u32 sum = 0;
  for (u32 k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  const u32 incl = block_inclusive_scan(sum, sh);
  u32 run = block_sums[blockIdx.x] + incl - sum;
  for (u32 k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
      out[base + k] = run;
      if (out_copy) out_copy[base + k] = run;
    }
    run += v[k];
  }
}

// ---------------------------------------------------------------------------------------------
// task planning
// ---------------------------------------------------------------------------------------------
// meta[0] = total entries, meta[1] = total tasks, meta[2] = #small split buckets, meta[3] = #large
KERNEL void plan_tasks_kernel(const u32* starts, const u32* ends, u32 nb, u32 T, u32* ntasks, u32* small_list,
                              u32* large_list, u32 list_cap_small, u32 list_cap_large, u32* meta) {
  const u32 wb = blockIdx.x * blockDim.x + threadIdx.x;
  if (wb >= nb) return;
  const u32 cnt = ends[wb] - starts[wb];
  const u32 nt = (cnt + T - 1) / T;
  ntasks[wb] = nt;
  if (nt > 1) {
    if (nt <= SMALL_SPLIT_MAX) {
      const u32 k = atomic_add_u32(&meta[2], 1u);
      if (k < list_cap_small) small_list[k] = wb;
    } else {
      const u32 k = atomic_add_u32(&meta[3], 1u);
      if (k < list_cap_large) large_list[k] = wb;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation: one thread per task
// ---------------------------------------------------------------------------------------------
KERNEL void __launch_bounds__(128) accumulate_kernel(const unsigned char* bases, u32 stride, const u32* sorted,
                                                      const u32* starts, const u32* ends, const u32* ntasks,
                                                      const u32* task_off, u32 nb, u32 T, const u32* meta,
                                                      G1Xyzz* buckets, G1Xyzz* partials) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= meta[1]) return;
  u32 lo = 0, hi = nb - 1;  // last bucket whose first task id is <= t
  while (lo < hi) {
    const u32 mid = (lo + hi + 1) >> 1;
    if (task_off[mid] <= t)
      lo = mid;
    else
      hi = mid - 1;
  }
  const u32 wb = lo;
  const u32 k = t - task_off[wb];
  const u32 begin = starts[wb] + k * T;
  u32 end = begin + T;
  const u32 bend = ends[wb];
  if (end > bend) end = bend;
  G1Xyzz acc = xyzz_identity();
  for (u32 pos = begin; pos < end; pos++) {
    const u32 e = sorted[pos];
    G1Affine p = affine_load(bases, stride, e & 0x7fffffffu);
    if (p.inf) continue;
    if (e >> 31) p.y = fp_neg(p.y);
    xyzz_add_affine(acc, p.x, p.y);
  }
  if (ntasks[wb] == 1)
    buckets[wb] = acc;
  else
    partials[t] = acc;
}

KERNEL void __launch_bounds__(128) combine_small_kernel(const u32* small_list, const u32* ntasks, const u32* task_off,
                                                         const u32* meta, const G1Xyzz* partials, G1Xyzz* buckets) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= meta[2]) return;
  const u32 wb = small_list[j];
  const u32 first = task_off[wb], nt = ntasks[wb];
  G1Xyzz acc = partials[first];
  for (u32 k = 1; k < nt; k++) xyzz_add_ni(acc, partials[first + k]);
  buckets[wb] = acc;
}

// one CTA per heavily split bucket: strided partial sums, then a shared-memory tree
KERNEL void __launch_bounds__(COMBINE_TPB) combine_large_kernel(const u32* large_list, const u32* ntasks,
                                                                 const u32* task_off, const u32* meta,
                                                                 const G1Xyzz* partials, G1Xyzz* buckets) {
  DYN_SMEM(G1Xyzz, sh);
  for (u32 j = blockIdx.x; j < meta[3]; j += gridDim.x) {
    const u32 wb = large_list[j];
    const u32 first = task_off[wb], nt = ntasks[wb];
    G1Xyzz acc = xyzz_identity();
    for (u32 k = threadIdx.x; k < nt; k += blockDim.x) xyzz_add_ni(acc, partials[first + k]);
    sh[threadIdx.x] = acc;
    SYNC_THREADS();
    for (u32 off = blockDim.x >> 1; off > 0; off >>= 1) {
      if (threadIdx.x < off) {
        G1Xyzz a = sh[threadIdx.x];
        xyzz_add_ni(a, sh[threadIdx.x + off]);
        sh[threadIdx.x] = a;
      }
      SYNC_THREADS();
    }
    if (threadIdx.x == 0) buckets[wb] = sh[0];
    SYNC_THREADS();
  }
}

// ---------------------------------------------------------------------------------------------
// bucket reduction:  f(X[0..m)) = sum_i (i+1) * X[i]  =  sum_t WS_t + Kc * f(R[1..T))
//   with chunk sums R_t = sum X[tKc .. tKc+Kc) and WS_t = sum (local index + 1) * X[...]
// ---------------------------------------------------------------------------------------------
KERNEL void __launch_bounds__(128) wsum_kernel(const G1Xyzz* X, u32 x_stride, u32 m, u32 nwin, u32 T, G1Xyzz* R,
                                                G1Xyzz* WS) {
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nwin * T) return;
  const u32 w = id / T, t = id % T;
  const G1Xyzz* x = X + (size_t)w * x_stride;
  const u32 lo = t * RED_KC;
  u32 hi = lo + RED_KC;
  if (hi > m) hi = m;
  G1Xyzz run = xyzz_identity(), acc = xyzz_identity();
  for (u32 i = hi; i > lo; i--) {
    xyzz_add_ni(run, x[i - 1]);
    xyzz_add_ni(acc, run);
  }
  R[(size_t)w * T + t] = run;
  WS[(size_t)w * T + t] = acc;
}

KERNEL void __launch_bounds__(128) psum_kernel(const G1Xyzz* X, u32 x_stride, u32 m, u32 nwin, u32 T, G1Xyzz* PS) {
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nwin * T) return;
  const u32 w = id / T, t = id % T;
  const G1Xyzz* x = X + (size_t)w * x_stride;
  const u32 lo = t * RED_KC;
  u32 hi = lo + RED_KC;
  if (hi > m) hi = m;
  G1Xyzz acc = xyzz_identity();
  for (u32 i = lo; i < hi; i++) xyzz_add_ni(acc, x[i]);
  PS[(size_t)w * T + t] = acc;
}

constexpr int MAX_RED_LEVELS = 6;
struct FinalArgs {
  const G1Xyzz* f[MAX_RED_LEVELS];  // f[l][w * stride[l]] = per-window plain sum of level l's WS
  u32 stride[MAX_RED_LEVELS];
  int levels;
  u32 W, c;
};

// Tail, step 1 (one thread per window): S_w = f0 + Kc (f1 + Kc (f2 + ...)), then D_w = 2^(c w) S_w.
// The c*w doublings of the different windows run in parallel; the critical path is the top window's.
KERNEL void window_weigh_kernel(FinalArgs a, G1Xyzz* D) {
  const u32 w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.W) return;
  G1Xyzz s = xyzz_identity();
  for (int l = a.levels - 1; l >= 0; l--) {
    for (u32 i = 0; i < RED_LOG_KC; i++) s = xyzz_double(s);
    xyzz_add_ni(s, a.f[l][(size_t)w * a.stride[l]]);
  }
  for (u32 i = 0; i < a.c * w; i++) s = xyzz_double(s);
  D[w] = s;
}

// Tail, step 2 (one thread): result = sum_w D_w, normalised (one inversion).
KERNEL void final_kernel(const G1Xyzz* D, u32 W, unsigned char* out144) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  G1Xyzz total = xyzz_identity();
  for (u32 w = 0; w < W; w++) xyzz_add_ni(total, D[w]);
  jacobian_store_normalised(out144, total);
}

// sum of `count` Jacobian points -> normalised Jacobian (multi-GPU combine; count is tiny)
KERNEL void g1_sum_kernel(const unsigned char* pts144, u32 count, unsigned char* out144) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  G1Xyzz total = xyzz_identity();
  for (u32 i = 0; i < count; i++) {
    const unsigned char* p = pts144 + (size_t)i * 144;
    xyzz_add_ni(total, xyzz_from_jacobian(fq_load8(p), fq_load8(p + 48), fq_load8(p + 96)));
  }
  jacobian_store_normalised(out144, total);
}

KERNEL void write_identity_kernel(unsigned char* out144) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  jacobian_store_normalised(out144, xyzz_identity());
}

}  // namespace msm
