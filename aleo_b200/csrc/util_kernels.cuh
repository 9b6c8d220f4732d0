// util_kernels.cuh -- synthetic workload generation, on-device checks and the integer-pipe
// microbenchmark that defines the MSM roofline denominator (bench / tests support; no product path
// depends on these).
#pragma once
#include "g1.cuh"
#include "mont_fp64.cuh"

namespace util {

constexpr u32 GEN_CHUNK = 32;

DEV u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// scalars[i] = hash(seed, first + i) masked to 252 bits (< r); optionally stored in Montgomery form
KERNEL void gen_scalars_kernel(Fr* out, u32 n, u64 seed, u64 first, u32 montgomery) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 ctr = (first + i) * 4;
  Fr v;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const u64 z = splitmix64(seed * 0xD1342543DE82EF95ull + ctr + k);
    v.l[2 * k] = (u32)z;
    v.l[2 * k + 1] = (u32)(z >> 32);
  }
  v.l[7] &= 0x0fffffffu;  // 252 bits
  if (montgomery) v = fp_to_mont(v);
  out[i] = v;
}

DEV Fr fr_from_u64_mont(u64 v) {
  Fr a = fp_zero<FrParams>();
  a.l[0] = (u32)v;
  a.l[1] = (u32)(v >> 32);
  return fp_to_mont(a);
}

// canonical k = s0 + idx * d mod r
DEV Fr dlog_of_index(const Fr& s0_m, const Fr& d_m, u64 idx) {
  return fp_from_mont(fp_add(s0_m, fp_mul(fr_from_u64_mont(idx), d_m)));
}

// k * G by double-and-add, k canonical 253-bit
DEV G1Xyzz g1_mul_generator(const Fr& k) {
  const Fq gx = fp_const<FqParams, FqParams::GX_M>(), gy = fp_const<FqParams, FqParams::GY_M>();
  G1Xyzz acc = xyzz_identity();
  for (int i = 7; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      acc = xyzz_double(acc);
      if ((k.l[i] >> b) & 1u) xyzz_add_affine_ni(acc, gx, gy);
    }
  }
  return acc;
}

// qaff[0], qaff[1] = affine coordinates of d * G
KERNEL void gen_step_point_kernel(Fq* qaff, Fr d_canon) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  G1Xyzz q = g1_mul_generator(d_canon);
  Fq x, y;
  if (!xyzz_to_affine(q, x, y)) {
    x = fp_zero<FqParams>();
    y = fp_zero<FqParams>();
  }
  qaff[0] = x;
  qaff[1] = y;
}

// thread t: XYZZ points for indices first + t*32 + [0, 32)
KERNEL void __launch_bounds__(128) gen_bases_xyzz_kernel(G1Xyzz* scratch, u32 n, Fr s0_canon, Fr d_canon, u64 first,
                                                          const Fq* qaff) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 base = t * GEN_CHUNK;
  if (base >= n) return;
  const Fr s0_m = fp_to_mont(s0_canon), d_m = fp_to_mont(d_canon);
  G1Xyzz p = g1_mul_generator(dlog_of_index(s0_m, d_m, first + base));
  const Fq qx = qaff[0], qy = qaff[1];
  const bool q_inf = fp_is_zero(qx) && fp_is_zero(qy);
  for (u32 j = 0; j < GEN_CHUNK && base + j < n; j++) {
    scratch[base + j] = p;
    if (!q_inf) xyzz_add_affine_ni(p, qx, qy);
  }
}

// thread t: batch-normalise 32 XYZZ points (one inversion) and store them as affine bases
KERNEL void __launch_bounds__(128) gen_bases_normalise_kernel(unsigned char* bases, u32 stride, const G1Xyzz* scratch,
                                                               u32 n, u64 out_first) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 base = t * GEN_CHUNK;
  if (base >= n) return;
  const u32 cnt = (n - base < GEN_CHUNK) ? (n - base) : GEN_CHUNK;
  Fq prefix[GEN_CHUNK];
  Fq run = fp_one<FqParams>();
  for (u32 j = 0; j < cnt; j++) {
    prefix[j] = run;
    const G1Xyzz& p = scratch[base + j];
    if (!xyzz_is_identity(p)) run = fq_mul_ni(run, fq_mul_ni(p.zz, p.zzz));
  }
  Fq inv = fq_inv_ni(run);
  for (u32 j = cnt; j > 0; j--) {
    const G1Xyzz p = scratch[base + j - 1];
    G1Affine a;
    if (xyzz_is_identity(p)) {
      a.inf = true;
      a.x = fp_zero<FqParams>();
      a.y = fp_zero<FqParams>();
    } else {
      const Fq zinv = fq_mul_ni(inv, prefix[j - 1]);  // 1 / (zz * zzz)
      inv = fq_mul_ni(inv, fq_mul_ni(p.zz, p.zzz));
      a.inf = false;
      a.x = fq_mul_ni(p.x, fq_mul_ni(zinv, p.zzz));
      a.y = fq_mul_ni(p.y, fq_mul_ni(zinv, p.zz));
    }
    affine_store(bases, stride, out_first + base + j - 1, a);
  }
}

// block_partials[b] (Montgomery) = sum over this block's grid-stride share of s_i * (s0 + (first+i) d)
KERNEL void dlog_dot_kernel(const Fr* scalars, u32 n, Fr s0_canon, Fr d_canon, u64 first, Fr* block_partials) {
  DYN_SMEM(Fr, sh);
  const Fr s0_m = fp_to_mont(s0_canon), d_m = fp_to_mont(d_canon);
  Fr acc = fp_zero<FrParams>();
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Fr k = fp_add(s0_m, fp_mul(fr_from_u64_mont(first + i), d_m));
    acc = fp_add(acc, fp_mul(fp_to_mont(scalars[i]), k));
  }
  sh[threadIdx.x] = acc;
  SYNC_THREADS();
  for (u32 off = blockDim.x >> 1; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = fp_add(sh[threadIdx.x], sh[threadIdx.x + off]);
    SYNC_THREADS();
  }
  if (threadIdx.x == 0) block_partials[blockIdx.x] = sh[0];
}

KERNEL void dlog_dot_final_kernel(const Fr* block_partials, u32 nblocks, Fr* out_canon) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fr acc = fp_zero<FrParams>();
  for (u32 i = 0; i < nblocks; i++) acc = fp_add(acc, block_partials[i]);
  *out_canon = fp_from_mont(acc);
}

KERNEL void check_on_curve_kernel(const unsigned char* bases, u32 stride, u32 n, u32* bad) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine a = affine_load(bases, stride, i);
  if (a.inf) return;
  const Fq lhs = fp_sqr(a.y);
  const Fq rhs = fp_add(fp_mul(fp_sqr(a.x), a.x), fp_one<FqParams>());
  if (!fp_eq(lhs, rhs)) atomic_add_u32(bad, 1u);
}

// ---- integer-pipe microbenchmark -------------------------------------------------------------------
// kind 0: independent mad.lo.u32 (IMAD);  kind 1: independent 32x32+64 (IMAD.WIDE.U32);
// kind 2: carry-chained lo/hi pairs (IMAD.WIDE.U32.X), 4 pairs per chain, as in the multiplier rows;
// kind 3: dependent Fq Montgomery products, two independent chains per thread.
template <int KIND>
KERNEL void __launch_bounds__(256) imad_bench_kernel(u32* sink, u32 iters, u32 seed) {
  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (KIND == 0) {
    u32 a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = seed + tid * 16 + k;
    const u32 x = seed | 1u, y = tid;
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = a[k] * x + y;
      }
    }
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= a[k];
    if (s == 0x12345678u) sink[0] = s;
  } else if (KIND == 1) {
    u64 a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = seed + tid * 16 + k;
    const u32 x = seed | 1u;
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = (u64)(u32)a[k] * x + a[k];
      }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= a[k];
    if (s == 0x12345678u) sink[0] = (u32)s;
  } else if (KIND == 2) {
    u32 acc[4][9];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 9; k++) acc[r][k] = seed + tid + r * 9 + k;
    u32 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = seed * (k + 3) + tid;
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
      for (int rep = 0; rep < 4; rep++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const u32 y = acc[(r + 1) & 3][8];
          acc[r][0] = ptx::mad_lo_cc(x[0], y, acc[r][0]);
          acc[r][1] = ptx::madc_hi_cc(x[0], y, acc[r][1]);
#pragma unroll
          for (int j = 2; j < 8; j += 2) {
            acc[r][j] = ptx::madc_lo_cc(x[j], y, acc[r][j]);
            acc[r][j + 1] = ptx::madc_hi_cc(x[j], y, acc[r][j + 1]);
          }
          acc[r][8] = ptx::addc(acc[r][8], 0);
        }
      }
    }
    u32 s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 9; k++) s ^= acc[r][k];
    if (s == 0x12345678u) sink[0] = s;
  } else if (KIND == 4 || KIND == 5) {
    // KIND 4: independent DFMA chains (the FP64 pipe: candidate carrier of 52-bit-limb products);
    // KIND 5: the same DFMAs interleaved one to one with 32 x 32 + 64 IMAD.WIDE chains -- do the two pipes overlap?
    double d[16];
    u64 a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      d[k] = (double)(seed + tid * 16 + k) * 1e-9;
      a[k] = seed + tid * 16 + k;
    }
    const double x = 1.0 + (double)(seed & 7u) * 1e-12, y = (double)tid * 1e-15;
    const u32 xi = seed | 1u;
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          d[k] = d[k] * x + y;  // contracted into one DFMA
          if (KIND == 5) a[k] = (u64)(u32)a[k] * xi + a[k];
        }
      }
    }
    double sd = 0.0;
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      sd += d[k];
      s ^= a[k];
    }
    if (sd == 0.12345678 || s == 0x12345678u) sink[0] = (u32)s;
  } else {
    // KIND 3: dependent Fq products on IMAD.WIDE; KIND 6: the same chains through the FP64-pipe multiplier
    // (mont_fp64.cuh); KIND 7: FP64-pipe squares
    Fq a = fp_zero<FqParams>(), b = fp_zero<FqParams>();
#pragma unroll
    for (int k = 0; k < 11; k++) {
      a.l[k] = seed + tid + k;
      b.l[k] = seed * 7 + tid + k;
    }
    const Fq c = fp_const<FqParams, FqParams::GX_M>();
    for (u32 it = 0; it < iters; it++) {
      if (KIND == 6) {
        a = fq_mul_fp64(a, c);
        b = fq_mul_fp64(b, c);
      } else if (KIND == 7) {
        a = fq_sqr_fp64(a);
        b = fq_sqr_fp64(b);
      } else if (KIND == 8) {
        a = fp_sqr(a);
        b = fp_sqr(b);
      } else {
        a = fp_mul(a, c);
        b = fp_mul(b, c);
      }
    }
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) s ^= a.l[k] ^ b.l[k];
    if (s == 0x12345678u) sink[0] = s;
  }
}

// ---- SASS-shape probes (tests/test_sass_shape.py) ----------------------------------------------------------------------
// One field product each, nothing else on the heavy FMA pipe: `cuobjdump -sass` of these three kernels must show exactly
// the IMAD.WIDE counts the roofline arithmetic of DESIGN.md rests on (Fq product 276, dedicated Fq square 210, Fr raw
// product 120).  The carry chains are separate asm statements that ptxas fuses pairwise into IMAD.WIDE.U32.X; a compiler
// or source change that breaks the fusion doubles the count and shows up here, on the CPU, before any GPU time is spent.
// out[i] = a[i] * b[i] (or a[i]^2) over Fq through the FP64-pipe multiplier: parity of the experiment with fp_mul
KERNEL void __launch_bounds__(128) fq_mul_fp64_kernel(Fq* out, const Fq* a, const Fq* b, u32 n, u32 square) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = square ? fq_sqr_fp64(a[i]) : fq_mul_fp64(a[i], b[i]);
}

#ifndef ALEO_EMU
extern "C" KERNEL void aleo_probe_fq_mul_fp64(Fq* out, const Fq* a, const Fq* b) { out[threadIdx.x] = fq_mul_fp64(a[threadIdx.x], b[threadIdx.x]); }
extern "C" KERNEL void aleo_probe_fq_sqr_fp64(Fq* out, const Fq* a) { out[threadIdx.x] = fq_sqr_fp64(a[threadIdx.x]); }
extern "C" KERNEL void aleo_probe_fq_mul(Fq* out, const Fq* a, const Fq* b) { out[threadIdx.x] = fp_mul(a[threadIdx.x], b[threadIdx.x]); }
extern "C" KERNEL void aleo_probe_fq_sqr(Fq* out, const Fq* a) { out[threadIdx.x] = fp_sqr(a[threadIdx.x]); }
extern "C" KERNEL void aleo_probe_fr_mul_raw(Fr* out, const Fr* a, const Fr* b) { out[threadIdx.x] = lz_mul(a[threadIdx.x], b[threadIdx.x]); }
#endif

}  // namespace util
