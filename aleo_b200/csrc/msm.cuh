// msm.cuh -- BLS12-377 G1 variable-base multi-scalar multiplication (Pippenger) for sm_100a.
//
// Stands in for snarkvm_algorithms::msm::VariableBase::msm::<G1Affine> (snarkvm-algorithms 0.14.5
// src/msm/variable_base/{mod,batched,standard}.rs), called by KZG10::commit /
// commit_lagrange / open and reached from the reference at rust/src/program/execute.rs:74,177.
//
// Pipeline (all on device, nothing returns to the host until the 144-byte result):
//   1 count     half-range scalars (s > (r-1)/2 -> r - s on the negated point), signed-window recoding (digits in
//               [-2^(c-1), 2^(c-1)]), histogram of (window, |digit|-1) with L2 atomics, one atomic per distinct
//               slot of a warp when the warp has duplicates                               [atomic]
//   2 scan      exclusive scan of the histogram -> bucket start offsets                   [tiny]
//   3 scatter   recode again, claim a slot per (window, bucket) with an atomic cursor, write
//               point index | sign<<31 (order inside a bucket is irrelevant: + commutes); window-major, or in
//               bucket-range passes for a resident SRS, once the bucket heads outgrow L2   [atomic / HBM]
//   4 plan      the sorted entry list is cut into equal runs, one per thread; buckets cut by a run
//               boundary are listed for the combine step (skew-proof: KZG witnesses are full of 0 / 1)
//   5 accumulate one thread per run: XYZZ mixed additions (8 products + 2 dedicated squares, two out-of-line field
//               functions) of gathered affine bases                                        [IMAD.WIDE]
//               -- ~80 % of the time: n * W additions
//   6 combine   the pieces of cut buckets are added (thread per bucket, or CTA per bucket)
//   7 reduce    per window sum_b (b+1) * bucket[b]: chunk levels (running sums) down to <= 256 elements per
//               window, then one CTA per window (log-depth suffix scan + tree)
//   8 tail      one CTA: a quad of lanes per window doubles 2^(c w) S_w cooperatively, tree over the windows,
//               one binary-GCD inversion -> normalised Jacobian
// A resident SRS (bases expanded to 2^(c w) P_i) shares one bucket set across windows and has no step-8 doublings;
// a batch of MSMs against one SRS gives every member its own bucket set.  Host-pointer calls run steps 1-6 once
// per point range while the next range is copied, the buckets accumulating across ranges.
#pragma once
#include "g1.cuh"

namespace msm {

constexpr u32 SCALAR_BITS = 253;
constexpr u32 SMALL_SPLIT_MAX = 8;       // buckets cut into <= 8 pieces are combined by one thread, the others by a CTA each (a narrow top
                                         // window -- GLV at c = 12: 70 buckets of 1900 entries -- makes buckets of ~60 pieces: 60 serial
                                         // additions in combine_small doubled the 2^16 accumulation phase, 0.9 -> 1.8 ms)
constexpr u32 COMBINE_TPB = 128;

struct Params {
  u32 c;       // window bits
  u32 W;       // number of windows = 253 / c + 1 (always absorbs the recoding carry)
  u32 B;       // buckets per window = 2^(c-1)
  u32 nlanes;  // threads of the accumulation kernel = equal runs the sorted entry list is cut into
  // Resident-SRS mode (n_stride != 0): bases were expanded once to P[w][i] = 2^(c w) * P_i, so every
  // window shares ONE set of 2^(c-1) buckets and the entry for (point i, window w) is w * n_stride + i.
  u32 n_stride;
  u32 first;   // resident-SRS mode: offset of this point range into the SRS (scalar i belongs to P_(first + i))
  // Half-range recoding: a scalar s > (r-1)/2 is replaced by r - s and all its digits change sign
  // (s P = (r - s)(-P)).  The recoded scalars have 252 bits, which is what lets c = 23 cover them with 11
  // windows instead of 12 (253 = 11 * 23 leaves no room for the recoding carry) for the resident SRS.  Always on:
  // it also turns the "negative small" scalars of a witness (r - k) into k, whose digits are zero above window 0,
  // instead of one crowded bucket in every window.
  u32 half_range;
  // Batch of independent MSMs against ONE resident SRS (KZG10 commits the ~13 polynomials of a proof against the
  // same powers): the scalar vectors lie back to back, batch_off[m] .. batch_off[m + 1] are the scalars of MSM m,
  // and MSM m owns the bucket set m (a "window" of the reduction kernels).  nbatch = 0: a single MSM.
  const u32* batch_off;
  u32 nbatch;
  // GLV (plain MSM below the batch-affine threshold): the scalars were split k = k1 + k2 u^2 (glv_split_kernel) and the
  // MSM runs over 2 n "virtual" points with 127-bit scalars -- virtual point i < glv_n is P_i, virtual point
  // glv_n + i is [u^2] P_i = (beta x_i, -y_i), whose x coordinates lie in `glv_bx` (48 bytes each).  Half as many
  // windows: half as many buckets to reduce and half the doubling chain of the tail, which is what a proof-sized MSM
  // spends most of its time in.  glv_n = 0: off.
  u32 glv_n;
  const unsigned char* glv_bx;
};

// which MSM of a batch scalar i belongs to (nbatch <= 64: a short walk), and its index inside that MSM
DEV u32 batch_member(const Params& prm, u32 i, u32& local) {
  u32 m = 0;
  while (m + 1 < prm.nbatch && prm.batch_off[m + 1] <= i) m++;
  local = i - prm.batch_off[m];
  return m;
}

// set = bucket set: the window (plain MSM), 0 (resident SRS) or the batch member
DEV u32 bucket_slot(const Params& prm, u32 w, u32 mag, u32 member = 0) {
  if (prm.nbatch) return member * prm.B + (mag - 1);
  return prm.n_stride ? (mag - 1) : (w * prm.B + (mag - 1));
}
DEV u32 entry_index(const Params& prm, u32 w, u32 i) { return prm.n_stride ? (w * prm.n_stride + prm.first + i) : i; }

// scalar i -> t (local copy, indexed by window position); returns 1 when half-range recoding replaced it by r - s
DEV u32 load_scalar(const u32* s, u32 (&t)[8], u32 half_range) {
#pragma unroll
  for (int k = 0; k < 8; k++) t[k] = s[k];
  if (!half_range) return 0;
  // s > (r-1)/2  <=>  2s >= r  <=>  2s - r does not borrow   (2s < 2^254 fits 8 limbs)
  u32 d[8];
  d[0] = ptx::sub_cc(t[0] << 1, FrParams::MOD(0));
#pragma unroll
  for (int k = 1; k < 8; k++) d[k] = ptx::subc_cc((t[k] << 1) | (t[k - 1] >> 31), FrParams::MOD(k));
  const u32 borrow = ptx::subc(0, 0);
  if (borrow) return 0;
  t[0] = ptx::sub_cc(FrParams::MOD(0), t[0]);
#pragma unroll
  for (int k = 1; k < 7; k++) t[k] = ptx::subc_cc(FrParams::MOD(k), t[k]);
  t[7] = ptx::subc(FrParams::MOD(7), t[7]);
  return 1;
}

// signed digit of window w given the carry from window w-1; returns |digit| and updates carry/neg
DEV u32 recode_digit(const u32* s /* 8 limbs */, u32 w, u32 c, u32& carry, u32& neg) {
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

// signed digit of window w of scalar s: walks the carry up from window 0 (W <= 64 cheap iterations)
DEV u32 digit_of_window(const u32* s, u32 w, u32 c, u32& neg) {
  u32 carry = 0, mag = 0;
  for (u32 j = 0; j <= w; j++) mag = recode_digit(s, j, c, carry, neg);
  return mag;
}

// slot += 1 for every active lane; returns the lane's position.  When lanes of the warp share a slot the warp
// issues ONE atomic per distinct slot: KZG scalars are full of 0 / 1 / small values and the top window of some
// (c, 253) pairs has one or two bits, and without aggregation millions of atomics serialise on a handful of
// counters (measured on B200 at n = 2^24: c = 21 -> 41 ms of "sort" against 10 ms for c = 20).  MATCH is slow
// (one per ~64 clocks per SM: unconditional use cost 1.5 ms per sort kernel at 2^24), so it only runs when a
// cheap test -- some lane holds the same slot as its neighbour -- says the warp has duplicates; hot slots put
// duplicates next to each other with near certainty, uniform scalars almost never do.
DEV u32 warp_aggregated_inc(u32* counters, u32 slot, bool active) {
#ifndef ALEO_EMU
  const unsigned act = __ballot_sync(0xffffffffu, active);
  const u32 key = active ? slot : (0xffffff00u | (threadIdx.x & 31u));  // inactive lanes never match anybody
  const u32 next = __shfl_down_sync(0xffffffffu, key, 1);
  const bool dup = __any_sync(0xffffffffu, (threadIdx.x & 31u) != 31u && next == key);
  u32 pos = 0;
  if (!dup) {
    if (active) pos = atomicAdd(&counters[slot], 1u);
    return pos;
  }
  if (active) {
    const unsigned peers = __match_any_sync(act, slot);
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    u32 base = 0;
    if ((int)lane == leader) base = atomicAdd(&counters[slot], (u32)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    pos = base + (u32)__popc(peers & ((1u << lane) - 1u));
  }
  return pos;
#else
  return active ? atomic_add_u32(&counters[slot], 1u) : 0u;
#endif
}

// Both sort kernels run window-major (blockIdx.y = window): the CTAs in flight then write to the 2^(c-1)
// bucket heads of ONE window (16 MB of 32-byte sectors at c = 20, L2 resident) instead of W * 2^(c-1)
// heads at once, which at c >= 18 overflowed L2 and turned every 4-byte entry into a DRAM sector write.
// gridDim.x CTAs per window stride over the scalars (a CTA per 256 scalars and window was launch bound:
// 2^20 short CTAs at n = 2^24).
KERNEL void count_kernel(const u32* scalars, u32 n, Params prm, u32* counts) {
  const u32 w = blockIdx.y;
  for (u32 base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {  // uniform trip count per CTA
    const u32 i = base + threadIdx.x;
    u32 neg = 0, mag = 0, t[8];
    if (i < n) {
      load_scalar(scalars + (size_t)i * 8, t, prm.half_range);
      mag = digit_of_window(t, w, prm.c, neg);
    }
    warp_aggregated_inc(counts, mag ? bucket_slot(prm, w, mag) : 0u, mag != 0);
  }
}

KERNEL void scatter_kernel(const u32* scalars, u32 n, Params prm, u32* cursor, u32* sorted) {
  const u32 w = blockIdx.y;
  for (u32 base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const u32 i = base + threadIdx.x;
    u32 neg = 0, mag = 0, flip = 0, t[8];
    if (i < n) {
      flip = load_scalar(scalars + (size_t)i * 8, t, prm.half_range);
      mag = digit_of_window(t, w, prm.c, neg);
    }
    const u32 pos = warp_aggregated_inc(cursor, mag ? bucket_slot(prm, w, mag) : 0u, mag != 0);
    if (mag) sorted[pos] = entry_index(prm, w, i) | ((neg ^ flip) << 31);
  }
}

// scalar-major variants: one thread per scalar walks all windows (the carry is free), the warp's 32 atomics of an
// iteration go to 32 different counters of ONE window, and consecutive iterations move on to the next window's
// counter region
KERNEL void count_kernel_sm(const u32* scalars, u32 n, Params prm, u32* counts) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 t[8];
  load_scalar(scalars + (size_t)(i < n ? i : 0) * 8, t, prm.half_range);
  u32 li = i, member = 0;
  if (prm.nbatch && i < n) member = batch_member(prm, i, li);
  u32 carry = 0, neg = 0;
  for (u32 w = 0; w < prm.W; w++) {
    const u32 mag = (i < n) ? recode_digit(t, w, prm.c, carry, neg) : 0u;
    warp_aggregated_inc(counts, mag ? bucket_slot(prm, w, mag, member) : 0u, mag != 0);
  }
}

// gridDim.y > 1: bucket-range passes.  A resident SRS shares one bucket set, so window-major order cannot shrink
// the set of live bucket heads; instead pass blockIdx.y only emits the entries whose bucket lies in its
// 1 / gridDim.y slice of the set (2^22 buckets x 32-byte sectors = 134 MB of heads at c = 23: four passes of 32 MB).
KERNEL void scatter_kernel_sm(const u32* scalars, u32 n, Params prm, u32* cursor, u32* sorted) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 slice_shift = (gridDim.y > 1) ? (prm.c - 1 - (31 - __clz((int)gridDim.y))) : 31;
  u32 t[8];
  const u32 flip = load_scalar(scalars + (size_t)(i < n ? i : 0) * 8, t, prm.half_range);
  u32 li = i, member = 0;
  if (prm.nbatch && i < n) member = batch_member(prm, i, li);
  u32 carry = 0, neg = 0;
  for (u32 w = 0; w < prm.W; w++) {
    u32 mag = (i < n) ? recode_digit(t, w, prm.c, carry, neg) : 0u;
    if (mag && gridDim.y > 1 && ((mag - 1) >> slice_shift) != blockIdx.y) mag = 0;  // another pass owns this bucket
    const u32 pos = warp_aggregated_inc(cursor, mag ? bucket_slot(prm, w, mag, member) : 0u, mag != 0);
    if (mag) sorted[pos] = entry_index(prm, w, li) | ((neg ^ flip) << 31);
  }
}

// ---------------------------------------------------------------------------------------------
// Partitioned sort (plain MSM, large sizes).  The count + scatter pair above pays one global atomic WITH return per
// entry (ATOM: ~45 G/s on B200 against ~140 G/s for the fire-and-forget RED of the count kernel) and 4-byte stores
// at random heads.  Here the sort is two levels and every atomic is a shared-memory one:
//   part_count   CTA x walks ITS scalars (all windows), histogram over coarse BINS (bin = bucket id >> fb, 2^fb fine
//                buckets per bin) in shared memory, one plain store per (bin, CTA) of the count -- layout
//                cnt[bin][cta], so that ONE flat exclusive scan yields the exact output offset of every (bin, CTA)
//   part_scatter grid (cta, window group): the same walk, shared-memory cursors initialised from the scan, 8-byte
//                records {entry, fine bucket} appended to the bin's region.  A group holds few enough bins
//                (<= PART_GROUP_BINS) that the partially written sectors of all CTAs stay L2 resident.
//   part_finish  one CTA per bin: shared-memory histogram of the fine buckets, block scan -> the bin's bucket
//                starts / ends, then the final 4-byte entries in bucket order (the bin's output region is a few
//                hundred KB: L2 resident, written back as whole sectors).
// ---------------------------------------------------------------------------------------------
constexpr u32 PART_TPB = 256;
constexpr u32 PART_GROUP_BINS = 2048;

struct PartArgs {
  const u32* scalars;
  u32 n;
  Params prm;
  u32 fb;          // fine bits: 2^fb buckets per bin
  u32 bps;         // bins per bucket set (window) = B >> fb
  u32 nbin;        // W * bps
  u32 ncta;        // CTAs that split the scalars (gridDim.x of part_count / part_scatter)
  u32 wpg;         // windows per scatter group
};

struct PartRecord {
  u32 entry;  // point index | sign << 31
  u32 fine;   // bucket id inside the bin
};

// cnt[bin * ncta + cta] = entries of bin produced by CTA cta
KERNEL void __launch_bounds__(PART_TPB) part_count_kernel(PartArgs a, u32* cnt) {
  DYN_SMEM(u32, hist);
  for (u32 b = threadIdx.x; b < a.nbin; b += PART_TPB) hist[b] = 0;
  SYNC_THREADS();
  for (u32 base = blockIdx.x * PART_TPB; base < a.n; base += a.ncta * PART_TPB) {
    const u32 i = base + threadIdx.x;
    if (i >= a.n) continue;
    u32 t[8];
    load_scalar(a.scalars + (size_t)i * 8, t, a.prm.half_range);
    u32 carry = 0, neg = 0;
    for (u32 w = 0; w < a.prm.W; w++) {
      const u32 mag = recode_digit(t, w, a.prm.c, carry, neg);
      if (mag) atomic_add_shared_u32(&hist[w * a.bps + ((mag - 1) >> a.fb)], 1u);
    }
  }
  SYNC_THREADS();
  for (u32 b = threadIdx.x; b < a.nbin; b += PART_TPB) cnt[(size_t)b * a.ncta + blockIdx.x] = hist[b];
}

// offs = exclusive scan of cnt (same layout); part[pos] = record
KERNEL void __launch_bounds__(PART_TPB) part_scatter_kernel(PartArgs a, const u32* offs, PartRecord* part) {
  DYN_SMEM(u32, cur);
  const u32 w0 = blockIdx.y * a.wpg;
  const u32 w1 = (w0 + a.wpg < a.prm.W) ? w0 + a.wpg : a.prm.W;
  const u32 bin0 = w0 * a.bps, nb = (w1 - w0) * a.bps;
  for (u32 b = threadIdx.x; b < nb; b += PART_TPB) cur[b] = offs[(size_t)(bin0 + b) * a.ncta + blockIdx.x];
  SYNC_THREADS();
  for (u32 base = blockIdx.x * PART_TPB; base < a.n; base += a.ncta * PART_TPB) {
    const u32 i = base + threadIdx.x;
    if (i >= a.n) continue;
    u32 t[8];
    const u32 flip = load_scalar(a.scalars + (size_t)i * 8, t, a.prm.half_range);
    u32 carry = 0, neg = 0;
    for (u32 w = 0; w < w1; w++) {  // the carry walks up from window 0; only the group's windows are emitted
      const u32 mag = recode_digit(t, w, a.prm.c, carry, neg);
      if (w < w0 || !mag) continue;
      const u32 pos = atomic_add_shared_u32(&cur[(w - w0) * a.bps + ((mag - 1) >> a.fb)], 1u);
      PartRecord r;
      r.entry = i | ((neg ^ flip) << 31);
      r.fine = (mag - 1) & ((1u << a.fb) - 1u);
      part[pos] = r;
    }
  }
}

// one CTA per bin: starts / ends of its 2^fb buckets, and the bin's entries in bucket order
KERNEL void __launch_bounds__(PART_TPB) part_finish_kernel(PartArgs a, const u32* offs, const u32* total, const PartRecord* part,
                                                          u32* starts, u32* ends, u32* sorted) {
  DYN_SMEM(u32, sm);  // [0, nf): histogram, then cursors; [nf, nf + PART_TPB): scan scratch
  const u32 nf = 1u << a.fb;
  u32* hist = sm;
  u32* scratch = sm + nf;
  const u32 bin = blockIdx.x;
  const u32 lo = offs[(size_t)bin * a.ncta];
  const u32 hi = (bin + 1 < a.nbin) ? offs[(size_t)(bin + 1) * a.ncta] : *total;
  for (u32 f = threadIdx.x; f < nf; f += PART_TPB) hist[f] = 0;
  SYNC_THREADS();
  for (u32 p = lo + threadIdx.x; p < hi; p += PART_TPB) atomic_add_shared_u32(&hist[part[p].fine], 1u);
  SYNC_THREADS();
  // exclusive scan of hist (nf values, nf / PART_TPB consecutive values per thread)
  const u32 per = (nf + PART_TPB - 1) / PART_TPB;
  const u32 f0 = threadIdx.x * per;
  u32 sum = 0;
  for (u32 k = 0; k < per; k++)
    if (f0 + k < nf) sum += hist[f0 + k];
  scratch[threadIdx.x] = sum;
  SYNC_THREADS();
  for (u32 off = 1; off < PART_TPB; off <<= 1) {
    const u32 add = (threadIdx.x >= off) ? scratch[threadIdx.x - off] : 0u;
    SYNC_THREADS();
    scratch[threadIdx.x] += add;
    SYNC_THREADS();
  }
  u32 run = lo + scratch[threadIdx.x] - sum;
  const u32 g0 = bin << a.fb;  // first global bucket of the bin
  for (u32 k = 0; k < per; k++) {
    if (f0 + k < nf) {
      const u32 c = hist[f0 + k];
      starts[g0 + f0 + k] = run;
      ends[g0 + f0 + k] = run + c;
      hist[f0 + k] = run;  // becomes the bucket's cursor
      run += c;
    }
  }
  SYNC_THREADS();
  for (u32 p = lo + threadIdx.x; p < hi; p += PART_TPB) {
    const PartRecord r = part[p];
    sorted[atomic_add_shared_u32(&hist[r.fine], 1u)] = r.entry;
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

// shift != 0: the scanned value is ceil(in / 2^shift) -- the bucket sizes after `shift` pair-tree levels (msm_ba.cuh)
DEV u32 scan_value(u32 v, u32 shift) { return (v + ((1u << shift) - 1u)) >> shift; }

KERNEL void scan_block_sums_kernel(const u32* in, u32 n, u32* block_sums, u32 shift) {
  SHARED u32 sh[SCAN_TPB];
  const u32 base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  u32 sum = 0;
  for (u32 k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) sum += scan_value(in[base + k], shift);
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

// out = exclusive offsets; out_copy (optional) = the same (a cursor for the scatter) or, with copy_end, offset + value
KERNEL void scan_apply_kernel(const u32* in, u32 n, const u32* block_sums, u32* out, u32* out_copy, u32 shift, u32 copy_end) {
  SHARED u32 sh[SCAN_TPB];
  const u32 base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  u32 v[SCAN_ITEMS];
  u32 sum = 0;
  for (u32 k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? scan_value(in[base + k], shift) : 0u;
    sum += v[k];
  }
  const u32 incl = block_inclusive_scan(sum, sh);
  u32 run = block_sums[blockIdx.x] + incl - sum;
  for (u32 k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
      out[base + k] = run;
      if (out_copy) out_copy[base + k] = copy_end ? run + v[k] : run;
    }
    run += v[k];
  }
}

// ---------------------------------------------------------------------------------------------
// Bucket accumulation, "flat" scheme: the sorted entry list (all windows, all buckets, back to back)
// is cut into `nlanes` equal runs of L = ceil(total / nlanes) entries and every thread (lane) sums
// exactly one run -- all lanes of a warp execute the same number of mixed additions no matter how
// the scalars are distributed (uniform: bucket sizes fluctuate; KZG witnesses: a few huge buckets of
// 0 / 1 digits).  A bucket that lies inside one run is written straight to buckets[]; a bucket cut
// by run boundaries leaves one PIECE per run it touches (a run has at most two cut buckets: the one
// it starts in and the one it ends in) and the pieces are added by the combine kernels.
// ---------------------------------------------------------------------------------------------
constexpr u32 NO_BUCKET = 0xffffffffu;

DEV u32 run_length(u32 total, u32 nlanes) { return (total + nlanes - 1) / nlanes; }

// last bucket whose start offset is <= pos: the non-empty bucket that contains entry `pos`
DEV u32 bucket_of_entry(const u32* starts, u32 nb, u32 pos) {
  u32 lo = 0, hi = nb - 1;
  while (lo < hi) {
    const u32 mid = (lo + hi + 1) >> 1;
    if (starts[mid] <= pos)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// meta[0] = total entries, meta[2] = #buckets cut into 2..64 pieces, meta[3] = #buckets cut into more
KERNEL void plan_pieces_kernel(const u32* starts, const u32* ends, u32 nb, u32 nlanes, u32* small_list, u32* large_list,
                               u32 cap_small, u32 cap_large, u32* meta) {
  const u32 wb = blockIdx.x * blockDim.x + threadIdx.x;
  if (wb >= nb) return;
  const u32 s = starts[wb], e = ends[wb];
  if (e == s) return;
  const u32 L = run_length(meta[0], nlanes);
  const u32 span = (e - 1) / L - s / L + 1;
  if (span > 1) {
    if (span <= SMALL_SPLIT_MAX) {
      const u32 k = atomic_add_u32(&meta[2], 1u);
      if (k < cap_small) small_list[k] = wb;
    } else {
      const u32 k = atomic_add_u32(&meta[3], 1u);
      if (k < cap_large) large_list[k] = wb;
    }
  }
}

// into != 0: the buckets already hold the sums of earlier point ranges of the same MSM (Session::add_chunk);
// a bucket that lies inside the run then starts from its stored sum instead of the identity.
// The gather / bucket-boundary bookkeeping of a lane (cheap, divergent) is separated from the warp's common addition
// (expensive, convergent).  Variants that were measured and dropped: letting a lane with an empty accumulator copy its
// point and move straight on to its next entry (no gain: 72.9 ms either way at 2^24, 1 % slower at c = 16), and 4 CTAs
// per SM at 128 registers (75.3 ms).
// pulls the two cache lines of base `idx` towards L1 (no registers held, unlike loading it early)
DEV void prefetch_base(const unsigned char* bases, size_t stride, u32 idx) {
#ifndef ALEO_EMU
  const unsigned char* p = bases + (size_t)idx * stride;
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 88));
#else
  (void)bases; (void)stride; (void)idx;
#endif
}

// DIRECT: the entries ARE the points -- position pos of the (dense, bucket-ordered) packed affine array `bases` that the
// batch-affine levels left (msm_ba.cuh); `sorted` is not read and there is no sign.
// virtual point of a GLV entry -> base index; returns true for the endomorphism image ([u^2] P = (beta x, -y))
DEV bool glv_map(u32& idx, u32 glv_n) {
  if (glv_n == 0 || idx < glv_n) return false;
  idx -= glv_n;
  return true;
}

template <bool CALL, bool PREFETCH = false, bool DIRECT = false>
KERNEL void __launch_bounds__(128, 3) accumulate_kernel(const unsigned char* bases, u32 stride, const u32* sorted,
                                                      const u32* starts, const u32* ends, u32 nb, u32 nlanes,
                                                      const u32* meta, G1Xyzz* buckets, G1Xyzz* pieces,
                                                      u32* piece_bucket, u32 into, u32 glv_n, const unsigned char* glv_bx) {
  const u32 lane = blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= nlanes) return;
  piece_bucket[2 * lane] = NO_BUCKET;
  piece_bucket[2 * lane + 1] = NO_BUCKET;
  const u32 total = meta[0];
  if (total == 0) return;
  const u32 L = run_length(total, nlanes);
  const u32 begin = lane * L;
  if (begin >= total) return;
  const u32 end = (begin + L < total) ? begin + L : total;
  u32 wb = bucket_of_entry(starts, nb, begin);
  u32 cur_end = ends[wb];
  bool cut_at_start = starts[wb] < begin;  // the bucket began in an earlier run: what we sum is a piece
  G1Xyzz acc = xyzz_identity();
  if (into && !cut_at_start && cur_end <= end) acc = buckets[wb];
  u32 pos = begin;
  for (;;) {
    bool have = false;
    Fq px, py;
    while (pos < end) {
      if (pos == cur_end) {  // bucket wb ends inside this run (short, divergent)
        if (cut_at_start) {
          pieces[2 * lane] = acc;
          piece_bucket[2 * lane] = wb;
        } else {
          buckets[wb] = acc;
        }
        cut_at_start = false;
        do {  // next non-empty bucket: almost always wb + 1 (pos < total = ends[nb - 1] bounds the walk)
          wb++;
          cur_end = ends[wb];
        } while (cur_end <= pos);
        if (into && cur_end <= end)
          acc = buckets[wb];
        else
          acc = xyzz_identity();
      }
      const u32 e = DIRECT ? pos : sorted[pos];
      pos++;
      u32 idx = e & 0x7fffffffu;
      const bool phi = !DIRECT && glv_map(idx, glv_n);
      G1Affine p;
      if (phi) {  // x from the beta x array, y and the flag from the base: one round of loads, not two
        const unsigned char* pb = bases + (size_t)idx * stride;
        p.x = fq_load8(glv_bx + (size_t)idx * 48);
        p.y = fq_load8(pb + 48);
        p.inf = (stride >= 97) ? (pb[96] != 0) : (fp_is_zero(p.x) && fp_is_zero(p.y));  // beta x of a packed identity is 0
      } else {
        p = affine_load(bases, stride, idx);
      }
      if (PREFETCH && !DIRECT && pos < end) {  // the next gather, while this addition runs
        u32 nidx = sorted[pos] & 0x7fffffffu;
        const bool nphi = glv_map(nidx, glv_n);
        prefetch_base(bases, stride, nidx);
#ifndef ALEO_EMU
        if (nphi) asm volatile("prefetch.global.L1 [%0];" ::"l"(glv_bx + (size_t)nidx * 48));
#else
        (void)nphi;
#endif
      }
      if (p.inf) continue;
      if (!DIRECT && ((e >> 31) != (phi ? 1u : 0u))) p.y = fp_neg(p.y);
      px = p.x;
      py = p.y;
      have = true;
      break;
    }
    if (!have) break;
    xyzz_add_affine_t<CALL>(acc, px, py);
  }
  if (cur_end == end && !cut_at_start) {
    buckets[wb] = acc;  // the bucket ends exactly with the run and began inside it
  } else {
    const u32 slot = cut_at_start ? 0u : 1u;
    pieces[2 * lane + slot] = acc;
    piece_bucket[2 * lane + slot] = wb;
  }
}

// adds the pieces of bucket wb held by runs first_run .. last_run (strided by `step` from `offset`)
DEV void gather_pieces(G1Xyzz& acc, u32 wb, u32 first_run, u32 last_run, u32 offset, u32 step, const G1Xyzz* pieces,
                       const u32* piece_bucket) {
  for (u32 r = first_run + offset; r <= last_run; r += step) {
    if (piece_bucket[2 * r] == wb) xyzz_add_ni(acc, pieces[2 * r]);
    if (piece_bucket[2 * r + 1] == wb) xyzz_add_ni(acc, pieces[2 * r + 1]);
  }
}

KERNEL void __launch_bounds__(128) combine_small_kernel(const u32* small_list, const u32* starts, const u32* ends,
                                                         u32 nlanes, const u32* meta, const G1Xyzz* pieces,
                                                         const u32* piece_bucket, G1Xyzz* buckets) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= meta[2]) return;
  const u32 wb = small_list[j];
  const u32 L = run_length(meta[0], nlanes);
  G1Xyzz acc = buckets[wb];  // identity, or the sum the earlier point ranges of this MSM left (accumulate never writes a cut bucket)
  gather_pieces(acc, wb, starts[wb] / L, (ends[wb] - 1) / L, 0, 1, pieces, piece_bucket);
  buckets[wb] = acc;
}

// one CTA per bucket that was cut into many pieces: strided partial sums, then a shared-memory tree
KERNEL void __launch_bounds__(COMBINE_TPB) combine_large_kernel(const u32* large_list, const u32* starts, const u32* ends,
                                                                 u32 nlanes, const u32* meta, const G1Xyzz* pieces,
                                                                 const u32* piece_bucket, G1Xyzz* buckets) {
  SHARED G1Xyzz sh[COMBINE_TPB];
  const u32 L = run_length(meta[0], nlanes);
  for (u32 j = blockIdx.x; j < meta[3]; j += gridDim.x) {
    const u32 wb = large_list[j];
    G1Xyzz acc = xyzz_identity();
    if (threadIdx.x == 0) acc = buckets[wb];
    gather_pieces(acc, wb, starts[wb] / L, (ends[wb] - 1) / L, threadIdx.x, blockDim.x, pieces, piece_bucket);
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
// bucket reduction, per window:  f(X[0..m)) = sum_i (i+1) * X[i].
//   With chunks of Kc:  f(X) = sum_t WS_t + Kc * f(R[1..T)),  R_t = sum of chunk t,
//   WS_t = sum (local index + 1) * X[..] (running sums).  A CHUNK LEVEL (reduce_level_kernel) computes, for chunk t,
//     R'_t = sum of its chunk of the weighted input (X at level 0, R[1..) above),
//     P'_t = scale * WS_t + sum of its chunk of the previous level's P        (plain partial sums)
//   with scale = product of the earlier levels' chunk sizes.  Chunk levels are work efficient (2 additions
//   per element) but serial inside a thread, so they only run until <= SCAN_MAX elements per window are
//   left; the rest is one CTA per window (reduce_scan_kernel): f = sum of all suffix sums, a log-depth
//   suffix scan plus a log-depth tree -- 17 dependent additions instead of ~50 per level.
// ---------------------------------------------------------------------------------------------
constexpr u32 SCAN_MAX = 256;  // elements per window the scan stage takes (256 x 192 B = 48 KB of shared memory)

struct ReduceArgs {
  const G1Xyzz* X;   // weighted input, per window `x_stride` apart, first element at X[x_off]
  u32 x_stride, x_off, m;
  const G1Xyzz* P;   // plain input of the previous level (T_in per window) or nullptr at level 0
  u32 T_in;
  u32 log_kc;        // chunk size of this level = 2^log_kc
  u32 scale_log;     // log2 of the product of the earlier levels' chunk sizes: weight of this level's WS
  G1Xyzz* R_out;     // T_out per window
  G1Xyzz* P_out;
  u32 T_out, nwin;
};

KERNEL void __launch_bounds__(128) reduce_level_kernel(ReduceArgs a) {
  const u32 id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= a.nwin * a.T_out) return;
  const u32 w = id / a.T_out, t = id % a.T_out;
  const G1Xyzz* x = a.X + (size_t)w * a.x_stride + a.x_off;
  const u32 kc = 1u << a.log_kc;
  const u32 lo = t * kc;
  u32 hi = lo + kc;
  if (hi > a.m) hi = a.m;
  G1Xyzz run = xyzz_identity(), acc = xyzz_identity();
  for (u32 i = hi; i > lo; i--) {
    xyzz_add_ni(run, x[i - 1]);
    xyzz_add_ni(acc, run);
  }
  a.R_out[(size_t)w * a.T_out + t] = run;
  for (u32 i = 0; i < a.scale_log; i++) acc = xyzz_double(acc);
  if (a.P) {
    const G1Xyzz* p = a.P + (size_t)w * a.T_in;
    u32 phi = lo + kc;
    if (phi > a.T_in) phi = a.T_in;
    for (u32 i = lo; i < phi; i++) xyzz_add_ni(acc, p[i]);
  }
  a.P_out[(size_t)w * a.T_out + t] = acc;
}

// The same level computed by a whole CTA per chunk (blockDim.x = chunk size Kc <= 256, one thread per element): a
// log-depth suffix scan gives the running sums, a tree adds them up, thread 0 scales and joins the previous level's plain
// sums.  Depth 2 log2(Kc) + scale_log additions instead of 2 Kc + scale_log, for (log2(Kc) + 1) / 2 times the work: used
// while the whole level is small (proof-sized commitments: a 2^16-point KZG commit spent 1.0 of its 1.66 ms in three
// serial chunk levels and the scan stage, profiles/r02_small_msm_trace.log).
KERNEL void __launch_bounds__(SCAN_MAX) reduce_level_cta_kernel(ReduceArgs a) {
  DYN_SMEM(G1Xyzz, sh);
  const u32 w = blockIdx.x / a.T_out, t = blockIdx.x % a.T_out;
  const u32 j = threadIdx.x, kc = blockDim.x;
  const G1Xyzz* x = a.X + (size_t)w * a.x_stride + a.x_off;
  const u32 idx = t * kc + j;
  G1Xyzz mine = xyzz_identity();
  if (idx < a.m) mine = x[idx];
  sh[j] = mine;
  SYNC_THREADS();
  // suffix scan over the chunk (Hillis-Steele): mine = sum of the chunk's elements at positions >= j
  for (u32 off = 1; off < kc; off <<= 1) {
    const bool take = (j + off < kc);
    G1Xyzz other;
    if (take) other = sh[j + off];
    SYNC_THREADS();
    if (take) {
      xyzz_add_ni(mine, other);
      sh[j] = mine;
    }
    SYNC_THREADS();
  }
  if (j == 0) a.R_out[(size_t)w * a.T_out + t] = mine;  // the chunk's plain sum
  // WS = sum of the suffix sums = sum (local index + 1) * X
  for (u32 off = kc >> 1; off > 0; off >>= 1) {
    if (j < off) {
      xyzz_add_ni(mine, sh[j + off]);
      sh[j] = mine;
    }
    SYNC_THREADS();
  }
  G1Xyzz acc = mine;  // thread 0: WS
  if (j == 0)
    for (u32 i = 0; i < a.scale_log; i++) acc = xyzz_double(acc);
  if (a.P) {  // uniform branch: the previous level's plain sums of this chunk join through a second tree
    G1Xyzz pj = xyzz_identity();
    if (idx < a.T_in) pj = a.P[(size_t)w * a.T_in + idx];
    SYNC_THREADS();
    sh[j] = pj;
    SYNC_THREADS();
    for (u32 off = kc >> 1; off > 0; off >>= 1) {
      if (j < off) {
        xyzz_add_ni(pj, sh[j + off]);
        sh[j] = pj;
      }
      SYNC_THREADS();
    }
    if (j == 0) xyzz_add_ni(acc, pj);
  }
  if (j == 0) a.P_out[(size_t)w * a.T_out + t] = acc;
}

// Last reduction stage, one CTA per window: S_w = 2^scale_log * f(X[0..m)) + sum P[0..T_in), m, T_in <= blockDim.x.
struct ScanArgs {
  const G1Xyzz* X;
  u32 x_stride, x_off, m;
  const G1Xyzz* P;
  u32 T_in, scale_log;
  G1Xyzz* S;  // one per window
};

KERNEL void __launch_bounds__(SCAN_MAX) reduce_scan_kernel(ScanArgs a) {
  DYN_SMEM(G1Xyzz, sh);
  const u32 w = blockIdx.x, i = threadIdx.x, span = blockDim.x;
  G1Xyzz mine = xyzz_identity();
  if (i < a.m) {
    mine = a.X[(size_t)w * a.x_stride + a.x_off + i];
    for (u32 k = 0; k < a.scale_log; k++) mine = xyzz_double(mine);
  }
  sh[i] = mine;
  SYNC_THREADS();
  // suffix scan (Hillis-Steele): after the loop mine = sum_{j >= i} X[j]
  for (u32 off = 1; off < span; off <<= 1) {
    const bool take = (i + off < a.m);
    G1Xyzz other;
    if (take) other = sh[i + off];
    SYNC_THREADS();
    if (take) {
      xyzz_add_ni(mine, other);
      sh[i] = mine;
    }
    SYNC_THREADS();
  }
  // f(X) = sum of all suffix sums; the plain partial sums of the chunk levels join the same tree
  if (a.P && i < a.T_in) {
    xyzz_add_ni(mine, a.P[(size_t)w * a.T_in + i]);
    sh[i] = mine;
  }
  SYNC_THREADS();
  for (u32 off = span >> 1; off > 0; off >>= 1) {
    if (i < off) {
      xyzz_add_ni(mine, sh[i + off]);
      sh[i] = mine;
    }
    SYNC_THREADS();
  }
  if (i == 0) a.S[w] = mine;
}

// ---------------------------------------------------------------------------------------------
// Tail: result = sum_w 2^(c w) * S_w, normalised.  The c * (W - 1) doublings of the top window are a
// dependency chain no schedule removes, so the kernel shortens each link instead: a QUAD of lanes owns one
// window and computes the independent Fq products of a doubling side by side (XYZZ dbl-2008-s-1 has
// 9 products in 3 dependent levels of 2 / 4 / 3), exchanging them with warp shuffles.  Every lane of a
// warp runs the same multiplier code on its own operands, so there is no divergence.  One CTA, 4 lanes
// per window (W <= 64).  Then a tree over the windows and ONE inversion (binary GCD).
// ---------------------------------------------------------------------------------------------
constexpr u32 TAIL_TPB = 256;

// operand of the calling lane: x_r for role r, chosen with bit masks (a ternary per limb compiles to divergent
// branches here: measured 4400 clocks per product level instead of ~1700)
DEV Fq role_select(u32 role, const Fq& x0, const Fq& x1, const Fq& x2, const Fq& x3) {
  const u32 m0 = 0u - (u32)(role == 0), m1 = 0u - (u32)(role == 1), m2 = 0u - (u32)(role == 2), m3 = 0u - (u32)(role == 3);
  Fq r;
#pragma unroll
  for (int k = 0; k < FqParams::N; k++) r.l[k] = (x0.l[k] & m0) | (x1.l[k] & m1) | (x2.l[k] & m2) | (x3.l[k] & m3);
  return r;
}

// out[r] = a_r * b_r for the four roles of a quad, delivered to every lane of the quad
DEV void quad_products(u32 role, const Fq& a0, const Fq& b0, const Fq& a1, const Fq& b1, const Fq& a2, const Fq& b2,
                       const Fq& a3, const Fq& b3, Fq& o0, Fq& o1, Fq& o2, Fq& o3) {
#ifndef ALEO_EMU
  const Fq mine = fp_mul(role_select(role, a0, a1, a2, a3), role_select(role, b0, b1, b2, b3));  // inlined: 3 per loop body
  const unsigned base = (threadIdx.x & 31u) & ~3u;
#pragma unroll
  for (int k = 0; k < FqParams::N; k++) {
    o0.l[k] = __shfl_sync(0xffffffffu, mine.l[k], base + 0);
    o1.l[k] = __shfl_sync(0xffffffffu, mine.l[k], base + 1);
    o2.l[k] = __shfl_sync(0xffffffffu, mine.l[k], base + 2);
    o3.l[k] = __shfl_sync(0xffffffffu, mine.l[k], base + 3);
  }
#else
  (void)role;  // the emulator runs CUDA threads one after the other: role 0 computes all four products
  o0 = fq_mul_v(a0, b0);
  o1 = fq_mul_v(a1, b1);
  o2 = fq_mul_v(a2, b2);
  o3 = fq_mul_v(a3, b3);
#endif
}

// p <- 2 p, all four lanes of the quad hold (and update) the whole point
DEV void quad_double(u32 role, G1Xyzz& p) {
  const Fq u = fp_dbl(p.y);
  Fq v, xx, t2, t3;
  quad_products(role, u, u, p.x, p.x, u, u, p.x, p.x, v, xx, t2, t3);                     // level 1: V = U^2, X^2
  const Fq m = fp_add(fp_dbl(xx), xx);
  Fq w, mm, s, zz3;
  quad_products(role, u, v, m, m, p.x, v, p.zz, v, w, mm, s, zz3);                        // level 2: W, M^2, S, ZZ3
  const Fq x3 = fp_sub(mm, fp_dbl(s));
  const Fq d = fp_sub(s, x3);
  Fq t, wy, zzz3, t4;
  quad_products(role, m, d, w, p.y, w, p.zzz, m, d, t, wy, zzz3, t4);                     // level 3: M (S - X3), W Y, ZZZ3
  p.x = x3;
  p.y = fp_sub(t, wy);
  p.zz = zz3;    // identity (zz = 0) and y = 0 (u = 0 => v = 0) both give zz3 = 0: the identity, as xyzz_double does
  p.zzz = zzz3;
}

#ifndef ALEO_EMU
#define TAIL_CLOCK(slot) \
  if (dbg && threadIdx.x == ((nwin - 1) << 2)) dbg[slot] = (unsigned long long)clock64()
#else
#define TAIL_CLOCK(slot) ((void)0)
#endif

// dbg (optional, 4 x u64): clock64 of the top window's lane before / after the doubling chain, after the tree,
// after the normalisation (tuning aid, ALEO_B200_MSM_TRACE)
KERNEL void __launch_bounds__(TAIL_TPB, 1) weigh_sum_kernel(const G1Xyzz* S, u32 nwin, u32 c, unsigned char* out144,
                                                            unsigned long long* dbg) {
  SHARED G1Xyzz D[TAIL_TPB / 4];
  const u32 tid = threadIdx.x;
  TAIL_CLOCK(0);
  const u32 w = tid >> 2, role = tid & 3u;
  G1Xyzz p = xyzz_identity();
  if (w < nwin) p = S[w];
  // every lane of a warp iterates to the warp's largest count (shuffles need the whole warp); a quad that is
  // done keeps its point
  const u32 my_count = (w < nwin) ? c * w : 0u;
#ifndef ALEO_EMU
  const u32 w_first = (tid & ~31u) >> 2, w_last = (tid | 31u) >> 2;  // windows of this warp's quads
  const u32 warp_count = (w_first < nwin) ? c * (w_last < nwin ? w_last : nwin - 1) : 0u;
  for (u32 it = 0; it < warp_count; it++) {
    G1Xyzz q = p;
    quad_double(role, q);
    const u32 keep = 0u - (u32)(it >= my_count);  // all ones: this quad is done, its point stays
#pragma unroll
    for (int k = 0; k < FqParams::N; k++) {
      p.x.l[k] = (p.x.l[k] & keep) | (q.x.l[k] & ~keep);
      p.y.l[k] = (p.y.l[k] & keep) | (q.y.l[k] & ~keep);
      p.zz.l[k] = (p.zz.l[k] & keep) | (q.zz.l[k] & ~keep);
      p.zzz.l[k] = (p.zzz.l[k] & keep) | (q.zzz.l[k] & ~keep);
    }
  }
#else
  if (role == 0)
    for (u32 it = 0; it < my_count; it++) quad_double(role, p);
#endif
  TAIL_CLOCK(1);
  if (role == 0) D[w] = p;
  SYNC_THREADS();
  for (u32 off = TAIL_TPB / 8; off > 0; off >>= 1) {
    if (tid < off && tid + off < nwin) {
      G1Xyzz a = D[tid];
      xyzz_add_ni(a, D[tid + off]);
      D[tid] = a;
    }
    SYNC_THREADS();
  }
  TAIL_CLOCK(2);
  if (tid == 0) jacobian_store_normalised(out144, D[0]);
  SYNC_THREADS();
  TAIL_CLOCK(3);
}

// Batch tail: MSM m's result is S[m] (no window weights against a resident SRS); one CTA per member normalises it
// (one inversion each, in parallel) and writes the 48-byte compressed point or the 144-byte Jacobian image.
KERNEL void batch_finalize_kernel(const G1Xyzz* S, u32 count, unsigned char* out, u32 compressed);

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

// ---------------------------------------------------------------------------------------------
// Resident SRS: one thread per point expands P_i into W packed affine copies 2^(c w) * P_i
// (window-major: pre[(w * n + i) * 96]).  c doublings per window in XYZZ, then ONE inversion per
// point normalises all windows (Montgomery's trick).  Runs once per SRS; every later commitment
// then needs neither per-window buckets nor the 253 final doublings.
// ---------------------------------------------------------------------------------------------
constexpr u32 SRS_MAX_WINDOWS = 32;

KERNEL void __launch_bounds__(128) srs_expand_kernel(const unsigned char* bases, u32 stride, u32 n, u32 c, u32 W,
                                                      unsigned char* pre) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = affine_load(bases, stride, i);
  affine_store(pre, 96, i, p);
  G1Xyzz pts[SRS_MAX_WINDOWS - 1];
  Fq prefix[SRS_MAX_WINDOWS - 1];
  G1Xyzz cur = xyzz_from_affine(p);
  Fq run = fp_one<FqParams>();
  for (u32 w = 1; w < W; w++) {
    for (u32 k = 0; k < c; k++) cur = xyzz_double(cur);
    pts[w - 1] = cur;
    prefix[w - 1] = run;
    if (!xyzz_is_identity(cur)) run = fq_mul_ni(run, fq_mul_ni(cur.zz, cur.zzz));
  }
  Fq inv = fq_inv_ni(run);
  for (u32 w = W - 1; w >= 1; w--) {
    const G1Xyzz q = pts[w - 1];
    G1Affine a;
    if (xyzz_is_identity(q)) {
      a.inf = true;
      a.x = fp_zero<FqParams>();
      a.y = fp_zero<FqParams>();
    } else {
      const Fq zinv = fq_mul_ni(inv, prefix[w - 1]);
      inv = fq_mul_ni(inv, fq_mul_ni(q.zz, q.zzz));
      a.inf = false;
      a.x = fq_mul_ni(q.x, fq_mul_ni(zinv, q.zzz));
      a.y = fq_mul_ni(q.y, fq_mul_ni(zinv, q.zz));
    }
    affine_store(pre, 96, (size_t)w * n + i, a);
  }
}

// ---------------------------------------------------------------------------------------------
// GLV split.  On G1 the endomorphism phi(x, y) = (beta x, y) is multiplication by -u^2 (poly.cuh uses the same fact
// for its subgroup test), so for k < r:  k P = k1 P + k2 [u^2] P = k1 P + k2 (beta x, -y)  with k2 = k div u^2,
// k1 = k mod u^2, both below 2^127 -- no lattice rounding: r = u^4 - u^2 + 1 makes plain division by u^2 a balanced
// split.  Holds for points of the prime-order subgroup, which is what a snarkVM G1Affine is (deserialisation checks
// it and the group law preserves it).  out: 2 n scalars of 8 limbs (k1 of point i at i, k2 at n + i); bx: beta x_i.
// ---------------------------------------------------------------------------------------------
KERNEL void glv_split_kernel(const u32* scalars, u32 n, const unsigned char* bases, u32 stride, u32* out, unsigned char* bx) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 k[8];
#pragma unroll
  for (int j = 0; j < 8; j++) k[j] = scalars[(size_t)i * 8 + j];
  // q = floor(k * BARRETT / 2^384): limbs 12 .. 15 of the 17-limb product (column sums; q <= floor(k / u^2) <= q + 2)
  u32 q[4] = {0, 0, 0, 0};
  {
    u64 carry = 0;
    for (int col = 0; col < 16; col++) {
      u64 lo = carry & 0xffffffffull, hi = carry >> 32;  // running column sum in two halves so that nothing overflows
      for (int a = 0; a < 8; a++) {
        const int b = col - a;
        if (b < 0 || b > 8) continue;
        const u64 prod = (u64)k[a] * GlvParams::BARRETT(b);
        lo += prod & 0xffffffffull;
        hi += prod >> 32;
      }
      hi += lo >> 32;
      if (col >= 12) q[col - 12] = (u32)lo;
      carry = hi;
    }
  }
  // rem = k - q * u^2 (fits 5 limbs: < 3 u^2), then at most two corrections
  u32 rem[5];
  {
    u32 qd[5] = {0, 0, 0, 0, 0};  // low 5 limbs of q * u^2
    u64 carry = 0;
    for (int col = 0; col < 5; col++) {
      u64 lo = carry & 0xffffffffull, hi = carry >> 32;
      for (int a = 0; a < 4; a++) {
        const int b = col - a;
        if (b < 0 || b > 3) continue;
        const u64 prod = (u64)q[a] * GlvParams::U2(b);
        lo += prod & 0xffffffffull;
        hi += prod >> 32;
      }
      hi += lo >> 32;
      qd[col] = (u32)lo;
      carry = hi;
    }
    u64 borrow = 0;
    for (int j = 0; j < 5; j++) {
      const u64 d = (u64)k[j] - qd[j] - borrow;
      rem[j] = (u32)d;
      borrow = (d >> 32) & 1ull;
    }
  }
  for (int round = 0; round < 3; round++) {
    bool ge = rem[4] != 0;
    if (!ge) {
      ge = true;
      for (int j = 3; j >= 0; j--) {
        if (rem[j] != GlvParams::U2(j)) {
          ge = rem[j] > GlvParams::U2(j);
          break;
        }
      }
    }
    if (!ge) break;
    u64 borrow = 0;
    for (int j = 0; j < 5; j++) {
      const u64 d = (u64)rem[j] - (j < 4 ? GlvParams::U2(j) : 0u) - borrow;
      rem[j] = (u32)d;
      borrow = (d >> 32) & 1ull;
    }
    for (int j = 0; j < 4; j++) {
      q[j] += 1u;
      if (q[j] != 0) break;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    out[(size_t)i * 8 + j] = j < 4 ? rem[j] : 0u;
    out[(size_t)(n + i) * 8 + j] = j < 4 ? q[j] : 0u;
  }
  const G1Affine p = affine_load(bases, stride, i);
  fq_store8(bx + (size_t)i * 48, p.inf ? fp_zero<FqParams>() : fq_mul_ni(fp_const<FqParams, FqParams::BETA_M>(), p.x));
}

// Montgomery Fr -> canonical BigInteger256 (PrimeField::to_bigint), the first step of KZG10::commit
KERNEL void fr_to_bigint_kernel(const Fr* in, Fr* out, u32 n) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fp_from_mont(in[i]);
}

// normalised Jacobian (144 B, Montgomery) -> 48-byte compressed G1 (snarkVM wire format, pinned by the
// reference's proof fixture: x little-endian canonical, bit 383 = y is the larger root, bit 382 = infinity)
KERNEL void g1_compress_kernel(const unsigned char* jac144, unsigned char* out48);

DEV void g1_compress_to(const unsigned char* jac144, unsigned char* out48) {
  const Fq z = fq_load8(jac144 + 96);
  Fq x = fp_zero<FqParams>();
  u32 flags = 0;
  if (fp_is_zero(z)) {
    flags = 1u << 30;
  } else {
    x = fp_from_mont(fq_load8(jac144));
    const Fq y = fp_from_mont(fq_load8(jac144 + 48));
    const Fq ny = fp_neg(y);
    bool larger = false;  // y > p - y  <=>  y > (p - 1) / 2
    for (int k = FqParams::N - 1; k >= 0; k--) {
      if (y.l[k] != ny.l[k]) {
        larger = y.l[k] > ny.l[k];
        break;
      }
    }
    if (larger) flags = 1u << 31;
  }
  x.l[FqParams::N - 1] |= flags;
  fq_store8(out48, x);
}

KERNEL void batch_finalize_kernel(const G1Xyzz* S, u32 count, unsigned char* out, u32 compressed) {
  SHARED unsigned char jac[144];
  const u32 m = blockIdx.x;
  if (m >= count || threadIdx.x != 0) return;
  if (compressed) {
    jacobian_store_normalised(jac, S[m]);
    g1_compress_to(jac, out + (size_t)m * 48);
  } else {
    jacobian_store_normalised(out + (size_t)m * 144, S[m]);
  }
}

KERNEL void g1_compress_kernel(const unsigned char* jac144, unsigned char* out48) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  g1_compress_to(jac144, out48);
}

KERNEL void write_identity_kernel(unsigned char* out144) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  jacobian_store_normalised(out144, xyzz_identity());
}

}  // namespace msm
