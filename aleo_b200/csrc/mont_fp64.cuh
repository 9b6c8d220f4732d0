// mont_fp64.cuh -- the Fq Montgomery product on the FP64 pipe (DFMA), 48-bit limbs.  EXPERIMENT:
// same contract as fp_mul<FqParams> / fp_sqr<FqParams> of mont.cuh -- operands and result are the 12 x 32-bit Montgomery
// images (R = 2^384), result fully reduced -- so the two multipliers are interchangeable bit for bit.
//
// Why it could pay: IMAD.WIDE and DFMA share an issue port on B200 and one IMAD.WIDE costs two DFMA slots
// (tools/pipe_mix.cu), but a DFMA multiplies 48 x 48 bits where an IMAD.WIDE multiplies 32 x 32.  An exact 48 x 48-bit
// product takes three FP64-pipe instructions:
//     h = fma_rz(a, b, 2^100)              = 2^100 + floor(a b / 2^48) 2^48     (a double in [2^100, 2^101) has ulp 2^48)
//     s = (2^100 + 2^52) - h               = 2^52 - floor(a b / 2^48) 2^48      (exact)
//     l = fma_rz(a, b, s)                  = 2^52 + (a b mod 2^48)              (exact)
// and the mantissa fields of h and l ARE the two 48-bit halves: the bit patterns are added as 64-bit integers into
// column accumulators (ALU pipe) and the exponent fields they drag along are subtracted as compile-time constants.
// 8 limbs x 48 bits = 384 bits = R, so the reduction is the usual limb-by-limb Montgomery reduction with 2^48 limbs; the
// modulus is 1 + 3 * 2^46 mod 2^48, so the Montgomery factor needs no multiplication.
// FP64-pipe instructions per product: 64 * 3 (a b) + 64 * 3 (q p) + 8 (q -> double) + 16 (operands -> double) = 408,
// against 276 IMAD.WIDE = 552 slots of the same port.
#pragma once
#include "mont.cuh"

namespace fp64mul {

constexpr u64 MASK48 = (1ull << 48) - 1ull;
constexpr u64 LO_BITS = 0x4330000000000000ull;  // bit pattern of 2^52: l = 2^52 + L
constexpr u64 HI_BITS = 0x4630000000000000ull;  // bit pattern of 2^100: h = 2^100 + H 2^48

// 48-bit limb j of the modulus
template <class P>
CONSTFN u64 mod48(int j) {
  // bits [48 j, 48 j + 48) of the 32-bit limb array
  const int w = (48 * j) / 32, sh = (48 * j) % 32;  // sh is 0 or 16
  const u64 a = P::MOD(w), b = (w + 1 < P::N) ? P::MOD(w + 1 < P::N ? w + 1 : 0) : 0ull, c = (w + 2 < P::N) ? P::MOD(w + 2 < P::N ? w + 2 : 0) : 0ull;
  return sh == 0 ? ((a | (b << 32)) & MASK48) : (((a >> 16) | (b << 16) | (c << 48)) & MASK48);
}

// number of partial products a_i b_j with i + j == k, 0 <= i, j < 8
CONSTFN u64 npairs(int k) { return (k < 0 || k > 14) ? 0ull : (u64)(k <= 7 ? k + 1 : 15 - k); }
// exponent-field bias column k has collected after the product AND the reduction (both add npairs(k) low halves and
// npairs(k - 1) high halves)
CONSTFN u64 bias_of(int k) { return 2ull * (LO_BITS * npairs(k) + HI_BITS * npairs(k - 1)); }

DEV double limb_to_double(u64 v) {
#ifndef ALEO_EMU
  return __longlong_as_double((long long)(LO_BITS | v)) - 4503599627370496.0;  // (2^52 + v) - 2^52, exact
#else
  return (double)v;
#endif
}

// bit patterns of the two halves of the exact product of two 48-bit integers held in doubles
DEV void split_mul(double a, double b, u64& hi, u64& lo) {
#ifndef ALEO_EMU
  const double h = __fma_rz(a, b, 0x1p100);
  const double l = __fma_rz(a, b, (0x1p100 + 0x1p52) - h);
  hi = (u64)__double_as_longlong(h);
  lo = (u64)__double_as_longlong(l);
#else
  const unsigned __int128 p = (unsigned __int128)(u64)a * (u64)b;
  hi = HI_BITS | (u64)(p >> 48);
  lo = LO_BITS | ((u64)p & MASK48);
#endif
}

// 12 x 32-bit limbs -> 8 doubles holding 48-bit limbs
DEV void load48(const Fq& a, double (&A)[8]) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const u32 w0 = a.l[3 * p], w1 = a.l[3 * p + 1], w2 = a.l[3 * p + 2];
#ifndef ALEO_EMU
    A[2 * p] = __hiloint2double((int)(0x43300000u | (w1 & 0xffffu)), (int)w0) - 4503599627370496.0;
    A[2 * p + 1] = __hiloint2double((int)__byte_perm(w2, 0x43300000u, 0x7632), (int)__funnelshift_r(w1, w2, 16)) - 4503599627370496.0;
#else
    A[2 * p] = (double)((u64)w0 | ((u64)(w1 & 0xffffu) << 32));
    A[2 * p + 1] = (double)((u64)(w1 >> 16) | ((u64)w2 << 16));
#endif
  }
}

// SQUARE: the off-diagonal products are computed once and added twice
template <bool SQUARE>
DEV Fq mul(const Fq& a, const Fq& b) {
  typedef FqParams P;
  static_assert(P::N == 12 && mod48<P>(0) == 0xc00000000001ull, "48-bit limbs: Fq only; Montgomery factor assumes p = 1 + 3 * 2^46 mod 2^48");
  double A[8], B[8];
  load48(a, A);
  if (!SQUARE) load48(b, B);
  u64 c[16];
#pragma unroll
  for (int k = 0; k < 16; k++) c[k] = 0;
  // ---- a * b into the columns ---------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (SQUARE && j < i) continue;
      u64 hi, lo;
      split_mul(A[i], SQUARE ? A[j] : B[j], hi, lo);
      if (SQUARE && j > i) {
        c[i + j] += lo + lo;
        c[i + j + 1] += hi + hi;
      } else {
        c[i + j] += lo;
        c[i + j + 1] += hi;
      }
    }
  }
  // ---- Montgomery reduction, one 48-bit limb per round ------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // exponent-field biases are multiples of 2^52: the low 48 bits of a column are always its true low bits
    const u64 t = c[i] & MASK48;
    const u64 q = ((((3ull * t) & 3ull) << 46) - t) & MASK48;  // t * (-p^-1) mod 2^48, -p^-1 = 3 * 2^46 - 1
    const double Q = limb_to_double(q);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      u64 hi, lo;
      split_mul(Q, (double)mod48<P>(j), hi, lo);
      c[i + j] += lo;
      c[i + j + 1] += hi;
    }
    // column i is complete and == 0 mod 2^48: its upper part moves into the next column
    c[i + 1] += (c[i] - bias_of(i)) >> 48;
  }
  // ---- columns 8 .. 15 hold the result (< 2p): normalise, repack, reduce ----------------------------------------
  u64 limb[8], carry = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const u64 v = c[8 + k] - bias_of(8 + k) + carry;
    limb[k] = v & MASK48;
    carry = v >> 48;
  }
  Fq r;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const u64 l0 = limb[2 * p], l1 = limb[2 * p + 1];
    r.l[3 * p] = (u32)l0;
    r.l[3 * p + 1] = (u32)(l0 >> 32) | ((u32)l1 << 16);
    r.l[3 * p + 2] = (u32)(l1 >> 16);
  }
  fp_reduce_once(r);
  return r;
}

}  // namespace fp64mul

DEV Fq fq_mul_fp64(const Fq& a, const Fq& b) { return fp64mul::mul<false>(a, b); }
DEV Fq fq_sqr_fp64(const Fq& a) { return fp64mul::mul<true>(a, a); }
