// ntt.cuh -- radix-2 NTT over the BLS12-377 scalar field Fr for sm_100a.
//
// Replaces (drop-in, same numeric contract) snarkvm-algorithms 0.14.5
//   EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place   (src/fft/domain.rs)
// reached from the reference at rust/src/program/execute.rs:74,177 (prove_execution / vm.execute).
// Natural order in, natural order out; omega = TWO_ADIC_ROOT^(2^(47-k)); inverse scales by n^-1;
// coset variants shift by g = 22 (forward: c_i *= g^i before; inverse: c_i *= g^-i after).
//
// Design (B200 first, not a translation of any CPU loop nest):
//   * N = 2^L is split into P = ceil(L/9) passes of K_p <= 9 bits (a 9-bit pass = three radix-8 steps).
//     One pass = one HBM round trip (64 B per element), so the algorithmic traffic is 64*P bytes per element.
//   * A pass handles a 2048-element tile per 256-thread CTA: every thread owns 8 elements in
//     registers, does radix-8 / radix-4 / radix-2 butterflies on them, and trades elements with
//     the other threads of the CTA through a 64 KB shared-memory tile (two uint4 planes, so all
//     shared accesses are 128-bit and bank-conflict free) -- at most 3 __syncthreads per pass.
//   * Global accesses are 256-bit (LDG.E.256 / STG.E.256), each warp touching runs of >= 256
//     contiguous bytes.  The last pass writes the digit-reversed (natural-order) positions
//     directly, so no separate permutation pass exists.
//   * Twiddles: the first inter-pass factor w^(i2*k1) is rebuilt from two L2-resident power tables
//     (w^lo * w^(hi<<lo_bits)), the later ones are read from a direct table; the in-tile factors come from
//     an 8-16 KB table of the tile's own root.  The inverse transform's n^-1 and the coset powers ride on
//     the two-table scheme.
//   * The same pass kernels run one transform over 2 / 4 / 8 GPUs: local geometry, twiddles at the global
//     index, and the last-but-one pass stores straight into the peers' receive buffers (PassArgs d_*).
//   The transform is integer-pipe bound (12.75 Fr products per element at L = 24, 120
//   IMAD.WIDE each), not HBM bound -- see DESIGN.md for the two rooflines.
#pragma once
#include "mont.cuh"

namespace ntt {

constexpr int TILE_LOG = 11;            // elements per CTA tile
constexpr int TILE = 1 << TILE_LOG;     // 2048 elements = 64 KB
constexpr int TPB = TILE / 8;           // 256 threads, 8 elements each
// Transforms of <= 2^SMALL_TILE_MAX_LOG elements use 1024-element tiles (128 threads, 4 CTAs per SM): twice the CTAs
// for the same data, which is what a grid of 32 .. 512 CTAs on 148 SMs needs (B200: 2^16 0.061 -> 0.054 ms, 2^20
// 0.238 -> 0.214 ms; above 2^24 the 9-bit passes would fall to 64-byte global runs and lose 2 %).
constexpr int SMALL_TILE_LOG = 10;
constexpr int SMALL_TILE_MAX_LOG = 20;
constexpr int SMALL_MAX_LOG = 11;       // single-CTA kernel handles n <= 2^11
constexpr int MAX_PASS_BITS = 9;
constexpr int MAX_LOG_N = 32;           // memory bound, far below the field's two-adicity (47)

// Fr product used by the transform kernels.  Inlined by default: the out-of-line variant (by-value
// operands in registers, ~20 KB of SASS per pass kernel instead of > 200 KB) measured 5-7 % SLOWER on
// B200 -- a pass kernel is straight-line code executed once per tile, so unlike the MSM accumulation
// loop it gains nothing from instruction-cache residency and pays ~24 MOVs per 120-IMAD product.
#ifndef ALEO_NTT_CALL_MUL
#define ALEO_NTT_CALL_MUL 0
#endif
#if ALEO_NTT_CALL_MUL
DEV_NOINLINE Fr fr_mul_v(Fr a, Fr b) { return fp_mul(a, b); }
#else
DEV Fr fr_mul_v(const Fr& a, const Fr& b) { return fp_mul(a, b); }
#endif

// w^e from a two-level table: lo[e & mask] * hi[e >> lo_bits]
struct PowTable {
  const Fr* lo;
  const Fr* hi;
  u32 lo_bits;
};

DEV Fr pow_lookup(const PowTable& t, u32 e) {
  Fr a = t.lo[e & ((1u << t.lo_bits) - 1u)];
  Fr b = t.hi[e >> t.lo_bits];
  return fr_mul_v(a, b);
}

// ---------------------------------------------------------------------------------------------
// 2 / 4 / 8 point transforms on registers.  Natural order in, natural order out.
// roots: w8[1], w8[2], w8[3] = powers of the primitive 8th root used by this direction
// (w8[2] is the primitive 4th root).
// ---------------------------------------------------------------------------------------------
DEV void bfly(Fr& a, Fr& b) {
  Fr s = fp_add(a, b);
  b = fp_sub(a, b);
  a = s;
}

DEV void ntt2(Fr* x) { bfly(x[0], x[1]); }

DEV void ntt4(Fr* x, const Fr& w4) {
  bfly(x[0], x[2]);
  bfly(x[1], x[3]);
  x[3] = fr_mul_v(x[3], w4);
  bfly(x[0], x[1]);
  bfly(x[2], x[3]);
  Fr t = x[1];  // bit-reversed -> natural
  x[1] = x[2];
  x[2] = t;
}

DEV void ntt8(Fr* x, const Fr* w8 /* w8[1..3] valid */) {
  bfly(x[0], x[4]);
  bfly(x[1], x[5]);
  bfly(x[2], x[6]);
  bfly(x[3], x[7]);
  x[5] = fr_mul_v(x[5], w8[1]);
  x[6] = fr_mul_v(x[6], w8[2]);
  x[7] = fr_mul_v(x[7], w8[3]);
  bfly(x[0], x[2]);
  bfly(x[1], x[3]);
  bfly(x[4], x[6]);
  bfly(x[5], x[7]);
  x[3] = fr_mul_v(x[3], w8[2]);
  x[7] = fr_mul_v(x[7], w8[2]);
  bfly(x[0], x[1]);
  bfly(x[2], x[3]);
  bfly(x[4], x[5]);
  bfly(x[6], x[7]);
  // positions hold X[bitrev3(pos)]: swap 1<->4, 3<->6
  Fr t = x[1];
  x[1] = x[4];
  x[4] = t;
  t = x[3];
  x[3] = x[6];
  x[6] = t;
}

// ---------------------------------------------------------------------------------------------
// The same transforms on semi-reduced values (mont.cuh "Lazy arithmetic"): inputs and outputs in [0, 2r), products
// without a final subtraction, and a difference (or, at the last level, a sum) that is about to be multiplied is
// left in [0, 4r) with no test at all.  POSTMUL: the caller multiplies every output except x[0] by a twiddle next,
// so the last level leaves those wide.  The pass kernels use only these; the canonical ones above serve the
// single-CTA kernel.  ALU-pipe instructions per element drop by about a fifth (the passes are co-limited by the
// ALU and the heavy FMA pipe).
// ---------------------------------------------------------------------------------------------
#ifdef ALEO_NTT_LZM_CALL
DEV_NOINLINE Fr lzm(Fr a, Fr b) { return lz_mul(a, b); }
#else
DEV Fr lzm(const Fr& a, const Fr& b) { return lz_mul(a, b); }
#endif

DEV Fr pow_lookup_lz(const PowTable& t, u32 e) {  // < 2r
  Fr a = t.lo[e & ((1u << t.lo_bits) - 1u)];
  Fr b = t.hi[e >> t.lo_bits];
  return lzm(a, b);
}

DEV void bfly_lz(Fr& a, Fr& b) {  // (a + b, a - b), both < 2r
  Fr s = lz_add(a, b);
  b = lz_sub(a, b);
  a = s;
}
DEV void bfly_lz_mul(Fr& a, Fr& b, const Fr& w) {  // (a + b, (a - b) * w)
  Fr s = lz_add(a, b);
  b = lzm(lz_sub_wide(a, b), w);
  a = s;
}
template <bool WIDE>
DEV void bfly_lz_out(Fr& a, Fr& b) {  // last level: both outputs wide when the caller multiplies them next
  Fr s = WIDE ? lz_add_wide(a, b) : lz_add(a, b);
  b = WIDE ? lz_sub_wide(a, b) : lz_sub(a, b);
  a = s;
}
template <bool WIDE>
DEV void bfly_lz_out0(Fr& a, Fr& b) {  // ... except x[0], which is never multiplied
  Fr s = lz_add(a, b);
  b = WIDE ? lz_sub_wide(a, b) : lz_sub(a, b);
  a = s;
}

template <bool POSTMUL>
DEV void ntt2_lz(Fr* x) { bfly_lz_out0<POSTMUL>(x[0], x[1]); }

template <bool POSTMUL>
DEV void ntt4_lz(Fr* x, const Fr& w4) {
  bfly_lz(x[0], x[2]);
  bfly_lz_mul(x[1], x[3], w4);
  bfly_lz_out0<POSTMUL>(x[0], x[1]);
  bfly_lz_out<POSTMUL>(x[2], x[3]);
  Fr t = x[1];  // bit-reversed -> natural
  x[1] = x[2];
  x[2] = t;
}

template <bool POSTMUL>
DEV void ntt8_lz(Fr* x, const Fr* w8 /* w8[1..3] valid */) {
  bfly_lz(x[0], x[4]);
  bfly_lz_mul(x[1], x[5], w8[1]);
  bfly_lz_mul(x[2], x[6], w8[2]);
  bfly_lz_mul(x[3], x[7], w8[3]);
  bfly_lz(x[0], x[2]);
  bfly_lz_mul(x[1], x[3], w8[2]);
  bfly_lz(x[4], x[6]);
  bfly_lz_mul(x[5], x[7], w8[2]);
  bfly_lz_out0<POSTMUL>(x[0], x[1]);
  bfly_lz_out<POSTMUL>(x[2], x[3]);
  bfly_lz_out<POSTMUL>(x[4], x[5]);
  bfly_lz_out<POSTMUL>(x[6], x[7]);
  // positions hold X[bitrev3(pos)]: swap 1<->4, 3<->6
  Fr t = x[1];
  x[1] = x[4];
  x[4] = t;
  t = x[3];
  x[3] = x[6];
  x[6] = t;
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tile: two planes of uint4 (low / high half of an element).
// ---------------------------------------------------------------------------------------------
template <int K, bool LAST, int TL>
DEV u32 tile_phys(u32 p, u32 g) {
  constexpr u32 R = 1u << K;
  constexpr u32 G = (1u << TL) >> K;
  if (LAST) return g * (R + 1) + (p ^ ((p >> 3) & 7u));  // lanes run along p (loads) or g (stores)
  return p * G + g;                                      // lanes always run along g
}
template <int K, bool LAST, int TL>
CONSTFN u32 tile_plane_elems() {
  return LAST ? ((1u << TL) >> K) * ((1u << K) + 1u) : (1u << TL);
}

DEV void smem_put(uint4* plane0, uint4* plane1, u32 idx, const Fr& v) {
  plane0[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  plane1[idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
DEV Fr smem_get(const uint4* plane0, const uint4* plane1, u32 idx) {
  uint4 a = plane0[idx], b = plane1[idx];
  Fr v;
  v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
  v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
  return v;
}

struct PassArgs {
  const Fr* src;
  Fr* dst;
  u32 log_n;        // L
  u32 log_cur;      // log2 of the sub-problem size entering this pass
  u32 log_r1;       // LAST only: log2 of the first pass's radix
  u32 log_r2;       // LAST only: log2 of the second pass's radix when P >= 3, else 0
  u32 log_r3;       // LAST only: log2 of the third pass's radix when P == 4, else 0
  const Fr* inner;  // w_R^j, j < R, R = this pass's radix
  PowTable tw;      // powers of w_N (inverse: hi table carries n^-1 when fold_scale)
  const Fr* tw_direct;  // passes after the first: w_N^(x << (log_n - log_cur)), x < 2^log_cur, one lookup and no product
                        // (2^16 entries = 2 MB at L = 24, L2 resident); nullptr = use the two-level table
  PowTable pre;     // PRE : coset powers c^i applied to inputs of the first pass
  PowTable post;    // POST: powers applied to outputs of the last pass (coset inverse, carries n^-1)
  u32 use_pre, use_post;
  // ---- one transform spread over 2^lg GPUs (ntt_dist_*; all zero / neutral on a single GPU) -------------------------
  // The local array holds N / 2^lg elements whose last digit is shortened to `d_k2l` = K_last - lg bits: geometry
  // (log_n, log_cur, log_r1) is LOCAL, twiddles use the GLOBAL index (the rank's bits spliced in above d_k2l).
  u32 d_k2l;         // bits of the last digit held locally (>= 31: single GPU, every mask below is the identity)
  u32 d_klast;       // K_last
  u32 d_rank_bits;   // rank << d_k2l
  u32 d_log_chunk;   // log2(N / 4^lg): elements one rank sends to one peer
  u32 d_exchange;    // this (last-but-one) pass stores straight into the peers' receive buffers (NVLink P2P)
  u32 d_rank;
  u32 d_lg;          // log2 of the number of GPUs
  Fr* peer[8];       // receive buffer of every rank (own included), device pointers valid on this GPU
  // Cross-rank ordering without a collective: a one-warp kernel after the exchange pass stores `d_epoch` into slot
  // [d_rank] of every rank's flag array (release, system scope, after a system fence); the last pass spins on its own
  // `world` slots (acquire) before it reads the receive buffer.  nullptr / 0: no signalling / no waiting.
  u32* d_flag_peer[8];   // &flags_of_rank_r[d_rank]
  const u32* d_flag_local;  // this rank's slots, one per source rank
  u32 d_epoch;
  // Exchange pass only: tiles are taken in rotated order, starting with the tiles whose results go to rank d_rank + 1.
  // All results of a tile go to ONE rank (tile t of T belongs to chunk t / (T / g)), so with the rotation every rank
  // receives from exactly one sender at a time instead of all g - 1 senders converging on rank 0, then on rank 1, ...
  u32 d_tile_rot;
  // Single-GPU passes: while a tile's results are being stored, its CTA prefetches into L2 the inputs of the tile
  // `pf_ahead` blocks further on (the tile some SM starts about when this one ends).  0 = off.
  u32 pf_ahead;
};

// global index of the local index i2l of a strided pass: the rank's bits go in above the shortened last digit
DEV u32 dist_expand(const PassArgs& a, u32 i2l) {
  if (a.d_k2l >= 31u) return i2l;
  return ((i2l >> a.d_k2l) << a.d_klast) | a.d_rank_bits | (i2l & ((1u << a.d_k2l) - 1u));
}

// One pass.  K = bits handled, LAST = writes the natural-order result.  Coset scaling (use_pre on
// the first pass, use_post on the last) is a CTA-uniform runtime branch.
template <int K, bool LAST, int TL>
KERNEL void __launch_bounds__((1 << TL) / 8, (TL >= 11) ? 2 : 4) pass_kernel(PassArgs a) {
  constexpr u32 R = 1u << K;
  constexpr u32 LG = TL - K;
  constexpr u32 G = 1u << LG;
  constexpr u32 RT = R / 8;  // threads along r
  constexpr int S1 = 3;
  constexpr int S2 = (K - 3 >= 3) ? 3 : (K - 3);
  constexpr int S3 = K - S1 - S2;
  static_assert(K >= 5 && K <= 9, "pass radix");
  DYN_SMEM(uint4, plane0);
  uint4* plane1 = plane0 + tile_plane_elems<K, LAST, TL>();

  const u32 tid = threadIdx.x;
  // gridDim.y = batch of independent transforms laid out back to back
  const Fr* src = a.src + ((u64)blockIdx.y << a.log_n);
  Fr* dst = a.dst + ((u64)blockIdx.y << a.log_n);
  // loads: lanes along g (strided passes) or along r (last pass, rows are contiguous)
  const u32 tr = LAST ? (tid & (RT - 1)) : (tid >> LG);
  const u32 tg = LAST ? (tid / RT) : (tid & (G - 1));

  // ---- tile geometry -------------------------------------------------------------------------
  const u32 log_m = a.log_cur - K;  // log2(M), M = cur / R
  u64 in_base, out_base;
  u64 stride_r, stride_g, ostride_r;
  u32 i2_base = 0;
  if (!LAST) {
    u32 bx = blockIdx.x;
    if (a.d_exchange) {
      bx += a.d_tile_rot;
      if (bx >= gridDim.x) bx -= gridDim.x;
    }
    const u32 tiles_per_blk = 1u << (log_m - LG);
    const u32 blk = bx / tiles_per_blk;
    const u32 cg = bx % tiles_per_blk;
    i2_base = cg << LG;
    in_base = ((u64)blk << a.log_cur) + i2_base;
    out_base = in_base;
    stride_r = (u64)1 << log_m;
    stride_g = 1;
    ostride_r = stride_r;
  } else {
    // blocks b = k1 * S + rest ; the tile takes G consecutive k1 for one `rest`
    const u32 log_s = a.log_r2 + a.log_r3;  // S = product of the middle passes' radices
    const u32 S = 1u << log_s;
    const u32 rest = blockIdx.x & (S - 1);
    const u32 k1_0 = (blockIdx.x >> log_s) << LG;
    in_base = (((u64)k1_0 << log_s) + rest) << K;
    stride_r = 1;
    stride_g = (u64)S << K;
    // digit reversal of the middle digits: rest = k2 * R3 + k3  ->  k2 + R2 * k3
    const u32 k2 = rest >> a.log_r3, k3 = rest & ((1u << a.log_r3) - 1u);
    const u64 rev_rest = k2 + ((u64)k3 << a.log_r2);
    out_base = k1_0 + (rev_rest << a.log_r1);
    ostride_r = (u64)1 << (a.log_n - K);
  }

  if (LAST && a.d_flag_local) {  // every source rank has finished storing into this rank's receive buffer
    if (tid < (1u << a.d_lg)) {
      while ((int)(load_acquire_sys_u32(a.d_flag_local + tid) - a.d_epoch) < 0) spin_pause();
    }
    SYNC_THREADS();
  }
  Fr x[8];
  // ---- load + step 1 (radix 8 over the top three bits of r) ------------------------------------
  {
    constexpr u32 C1 = R >> 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const u32 p = j * C1 + tr;
      u64 gi = in_base + p * stride_r + tg * stride_g;
      if (LAST && a.d_k2l < 31u)  // a row's R elements arrived as 2^lg pieces of 2^d_k2l, one per source rank
        gi = ((u64)(p >> a.d_k2l) << a.d_log_chunk) + (((in_base >> K) + ((u64)tg * (stride_g >> K))) << a.d_k2l) +
             (p & ((1u << a.d_k2l) - 1u));
      x[j] = src[gi];
      if (!LAST && a.use_pre) x[j] = lzm(x[j], pow_lookup_lz(a.pre, dist_expand(a, (u32)gi)));  // coset: g^(global index)
    }
    Fr w8[4];
    w8[1] = a.inner[R / 8];
    w8[2] = a.inner[R / 4];
    w8[3] = a.inner[3 * (R / 8)];
    ntt8_lz<true>(x, w8);
#pragma unroll
    for (int k = 1; k < 8; k++) x[k] = lzm(x[k], a.inner[tr * k]);
#pragma unroll
    for (int k = 0; k < 8; k++) smem_put(plane0, plane1, tile_phys<K, LAST, TL>(k * C1 + tr, tg), x[k]);
  }
  SYNC_THREADS();
  // ---- step 2 ------------------------------------------------------------------------------------
  {
    constexpr u32 CP = R >> 3;        // sub-problem size entering the step
    constexpr u32 C = CP >> S2;       // ... leaving it
    constexpr int GROUPS = 8 >> S2;   // independent small transforms per thread
    constexpr int PTS = 1 << S2;
    Fr w8[4];
    if (S2 == 3) {
      w8[1] = a.inner[R / 8];
      w8[3] = a.inner[3 * (R / 8)];
    }
    if (S2 >= 2) w8[2] = a.inner[R / 4];
#pragma unroll
    for (int h = 0; h < GROUPS; h++) {
      const u32 gam = tr + h * RT;
      const u32 blk = gam / C, t = gam % C;
      const u32 bp = blk * CP + t;
#pragma unroll
      for (int j = 0; j < PTS; j++) x[h * PTS + j] = smem_get(plane0, plane1, tile_phys<K, LAST, TL>(bp + j * C, tg));
      if (S2 == 3) ntt8_lz<(C > 1)>(x + h * PTS, w8);
      if (S2 == 2) ntt4_lz<(C > 1)>(x + h * PTS, w8[2]);
      if (S2 == 1) ntt2_lz<(C > 1)>(x + h * PTS);
      if (C > 1) {
#pragma unroll
        for (int k = 1; k < PTS; k++) x[h * PTS + k] = lzm(x[h * PTS + k], a.inner[(R / CP) * t * k]);
      }
#pragma unroll
      for (int k = 0; k < PTS; k++) smem_put(plane0, plane1, tile_phys<K, LAST, TL>(bp + k * C, tg), x[h * PTS + k]);
    }
  }
  SYNC_THREADS();
  // ---- step 3 (only when K > 6): sub-problems of size 2^S3, no twiddles afterwards ---------------
  if (S3 > 0) {
    constexpr u32 CP = 1u << S3;
    constexpr int GROUPS = 8 >> S3;
    constexpr int PTS = 1 << S3;
    Fr w4;
    if (S3 == 2) w4 = a.inner[R / 4];
    Fr w8[4];
    if (S3 == 3) {  // 9-bit pass: a third radix-8 step
      w8[1] = a.inner[R / 8];
      w8[2] = a.inner[R / 4];
      w8[3] = a.inner[3 * (R / 8)];
    }
#pragma unroll
    for (int h = 0; h < GROUPS; h++) {
      const u32 gam = tr + h * RT;
      const u32 bp = gam * CP;
#pragma unroll
      for (int j = 0; j < PTS; j++) x[h * PTS + j] = smem_get(plane0, plane1, tile_phys<K, LAST, TL>(bp + j, tg));
      if (S3 == 3) ntt8_lz<false>(x + h * PTS, w8);
      if (S3 == 2) ntt4_lz<false>(x + h * PTS, w4);
      if (S3 == 1) ntt2_lz<false>(x + h * PTS);
#pragma unroll
      for (int k = 0; k < PTS; k++) smem_put(plane0, plane1, tile_phys<K, LAST, TL>(bp + k, tg), x[h * PTS + k]);
    }
    SYNC_THREADS();
  }
  // ---- L2 prefetch of a later tile's inputs (same thread -> element mapping as the load phase above) ------------
  if (a.pf_ahead && a.d_k2l >= 31u && blockIdx.x + a.pf_ahead < gridDim.x) {
    const u32 bx = blockIdx.x + a.pf_ahead;
    u64 pbase, pstride_r, pstride_g;
    if (!LAST) {
      const u32 tiles_per_blk = 1u << (log_m - LG);
      pbase = ((u64)(bx / tiles_per_blk) << a.log_cur) + ((bx % tiles_per_blk) << LG);
      pstride_r = (u64)1 << log_m;
      pstride_g = 1;
    } else {
      const u32 log_s = a.log_r2 + a.log_r3;
      pbase = (((u64)((bx >> log_s) << LG) << log_s) + (bx & ((1u << log_s) - 1u))) << K;
      pstride_r = 1;
      pstride_g = ((u64)1 << log_s) << K;
    }
    constexpr u32 C1 = R >> 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const Fr* ptr = src + pbase + (j * C1 + tr) * pstride_r + tg * pstride_g;
#ifndef ALEO_EMU
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
      (void)ptr;
#endif
    }
  }
  // ---- store: thread takes outputs kappa = j*RT + sr for its column sg (lanes along g) ------------
  // Rolled on purpose (2 iterations per trip): nothing here indexes the register array, and 8 inlined copies of
  // the two products below were a third of the kernel's code (ncu: 1.0-1.2 issue slots per instruction lost to
  // instruction fetch).
  {
    const u32 sr = tid >> LG, sg = tid & (G - 1);
    constexpr u32 C1 = R >> S1, C2 = C1 >> S2;
#pragma unroll 2
    for (int j = 0; j < 8; j++) {
      const u32 kap = j * RT + sr;
      // position of output kappa after the in-place decimation-in-frequency steps
      const u32 d1 = kap & 7u, d2 = (kap >> S1) & ((1u << S2) - 1u), d3 = kap >> (S1 + S2);
      const u32 pos = d1 * C1 + d2 * C2 + d3;
      Fr v = smem_get(plane0, plane1, tile_phys<K, LAST, TL>(pos, sg));
      const u64 go = out_base + kap * ostride_r + sg;
      if (!LAST) {
        const u32 x = dist_expand(a, i2_base + sg) * kap;
        if (a.tw_direct)  // the passes in between exchange semi-reduced values (< 2r)
          v = lzm(v, a.tw_direct[x]);
        else
          v = lzm(v, pow_lookup_lz(a.tw, x << (a.log_n - a.log_cur)));
      }
      if (LAST && a.use_post) {
        u32 gk = (u32)go;  // global output index: the rank's bits sit above the local part of the first digit
        if (a.d_k2l < 31u) gk = (((u32)go >> a.log_r1) << (a.log_r1 + a.d_lg)) | (a.d_rank << a.log_r1) | ((u32)go & ((1u << a.log_r1) - 1u));
        v = fp_mul(v, pow_lookup_lz(a.post, gk));  // full reduction: this is the result
      } else if (LAST) {
        fp_reduce_once(v);  // [0, 2r) -> canonical
      }
      if (!LAST && a.d_exchange)  // flat all-to-all: chunk t of the local result goes to rank t, slot = my rank
        a.peer[go >> a.d_log_chunk][((u64)a.d_rank << a.d_log_chunk) + (go & (((u64)1 << a.d_log_chunk) - 1u))] = v;
      else
        dst[go] = v;
    }
  }
}

// After the exchange pass (same stream: its stores into the peers are complete when this kernel starts): raise this
// rank's slot in every rank's flag array.  One warp, lane r signals rank r.
KERNEL void dist_signal_kernel(PassArgs a) {
  fence_system();
  if (threadIdx.x < (1u << a.d_lg)) store_release_sys_u32(a.d_flag_peer[threadIdx.x], a.d_epoch);
}

// standalone form of the last pass's prologue (profiling: times the cross-rank wait on its own)
KERNEL void dist_wait_kernel(const u32* flags, u32 world, u32 epoch) {
  if (threadIdx.x < world) {
    while ((int)(load_acquire_sys_u32(flags + threadIdx.x) - epoch) < 0) spin_pause();
  }
}

// ---------------------------------------------------------------------------------------------
// n <= 2^11: one CTA per transform, whole vector in shared memory (latency bound by nature).
// ---------------------------------------------------------------------------------------------
struct SmallArgs {
  Fr* data;         // batch x n, in place
  u32 log_n;
  const Fr* inner;  // w_n^j, j < n
  PowTable pre;     // coset forward: g^i on the inputs
  PowTable post;    // coset inverse: n^-1 * g^-i on the outputs
  Fr scale;         // plain inverse: n^-1
  u32 use_pre, use_post, use_scale;
};

KERNEL void __launch_bounds__(TPB) small_kernel(SmallArgs a) {
  DYN_SMEM(uint4, plane0);
  const u32 n = 1u << a.log_n;
  uint4* plane1 = plane0 + n;
  Fr* x = a.data + ((u64)blockIdx.x << a.log_n);
  for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
    Fr v = x[i];
    if (a.use_pre) v = fr_mul_v(v, pow_lookup(a.pre, i));
    u32 r = 0;
    for (u32 b = 0; b < a.log_n; b++) r |= ((i >> b) & 1u) << (a.log_n - 1 - b);
    smem_put(plane0, plane1, r, v);
  }
  SYNC_THREADS();
  for (u32 lm = 0; lm < a.log_n; lm++) {
    const u32 m = 1u << lm;
    for (u32 idx = threadIdx.x; idx < n / 2; idx += blockDim.x) {
      const u32 k = idx & (m - 1);
      const u32 lo = ((idx >> lm) << (lm + 1)) + k;
      Fr u = smem_get(plane0, plane1, lo);
      Fr t = smem_get(plane0, plane1, lo + m);
      if (k) t = fr_mul_v(t, a.inner[k << (a.log_n - lm - 1)]);
      smem_put(plane0, plane1, lo, fp_add(u, t));
      smem_put(plane0, plane1, lo + m, fp_sub(u, t));
    }
    SYNC_THREADS();
  }
  for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
    Fr v = smem_get(plane0, plane1, i);
    if (a.use_post) v = fr_mul_v(v, pow_lookup(a.post, i));
    if (a.use_scale) v = fr_mul_v(v, a.scale);
    x[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// data[r][c] *= w^((r + row0) * (c + col0)) for a rows x cols matrix: the twiddle step between the two
// local transform stages of the multi-GPU four-step NTT (w = root of the GLOBAL domain).
// ---------------------------------------------------------------------------------------------
KERNEL void twiddle_matrix_kernel(Fr* data, u32 rows, u32 cols, u32 row0, u32 col0, PowTable tw, u32 log_n_global) {
  const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (u64)rows * cols) return;
  const u32 r = (u32)(idx / cols), c = (u32)(idx % cols);
  const u64 e = ((u64)(r + row0) * (u64)(c + col0)) & (((u64)1 << log_n_global) - 1);
  data[idx] = fr_mul_v(data[idx], pow_lookup(tw, (u32)e));
}

// ---------------------------------------------------------------------------------------------
// In-place bit-reversal permutation of 2^log_n elements (blockIdx.y = transform of a batch): the out-of-order
// variants of upstream's transforms (FFTOrder::IO / OI in snarkvm-algorithms 0.14.5 src/fft/domain.rs, used by
// the prover with its FFTPrecomputation) are the in-order transform plus this pass.
// ---------------------------------------------------------------------------------------------
DEV u32 bit_reverse(u32 v, u32 bits) {
  u32 r = 0;
  for (u32 b = 0; b < bits; b++) r |= ((v >> b) & 1u) << (bits - 1 - b);
  return r;
}

KERNEL void bitrev_permute_kernel(Fr* data, u32 log_n) {
  Fr* x = data + ((u64)blockIdx.y << log_n);
  const u64 n = (u64)1 << log_n;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u32 j = bit_reverse((u32)i, log_n);
    if (i < j) {
      const Fr a = x[i], b = x[j];
      x[i] = b;
      x[j] = a;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Table builders (run once per (log_n, direction, kind) and cached by the host plan).
// ---------------------------------------------------------------------------------------------
// out[j] = scale * base^(j << shift)      (base, scale: device scalars)
KERNEL void pow_table_kernel(Fr* out, const Fr* base, const Fr* scale, u32 count, u32 shift) {
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  out[j] = fp_mul(*scale, fp_pow_u64(*base, (u64)j << shift));
}

// roots[0] = w_N (or its inverse), roots[1] = n^-1 (Montgomery), roots[2] = g or g^-1, roots[3] = 1
KERNEL void domain_consts_kernel(Fr* roots, u32 log_n, u32 inverse) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fr w = inverse ? fp_const<FrParams, FrParams::TWO_ADIC_ROOT_INV_M>() : fp_const<FrParams, FrParams::TWO_ADIC_ROOT_M>();
  for (u32 i = log_n; i < (u32)FR_TWO_ADICITY; i++) w = fp_sqr(w);
  Fr ninv = fp_one<FrParams>();
  Fr half = fp_const<FrParams, FrParams::TWO_INV_M>();
  for (u32 i = 0; i < log_n; i++) ninv = fp_mul(ninv, half);
  roots[0] = w;
  roots[1] = ninv;
  roots[2] = inverse ? fp_const<FrParams, FrParams::GENERATOR_INV_M>() : fp_const<FrParams, FrParams::GENERATOR_M>();
  roots[3] = fp_one<FrParams>();
}

}  // namespace ntt
