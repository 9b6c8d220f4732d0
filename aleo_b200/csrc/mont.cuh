// mont.cuh -- Montgomery-form prime-field arithmetic on 32-bit limbs for the sm_100a integer pipes.
//
// Number representation contract (snarkvm-fields 0.14.5 Fp256 / Fp384, SURVEY.md section 8a row 7/8,
// App. C): a field element is the Montgomery residue a*R mod m, R = 2^(32*N), stored as
// little-endian limbs and always fully reduced (< m).  An Fp<P> is bit-identical to the Rust
// struct's memory image (4 or 6 little-endian u64 limbs == 8 or 12 little-endian u32 limbs).
//
// Multiplier: operand-scanning Montgomery product with the row products split by the parity of
// their absolute limb position into two accumulators (E: pairs starting at even positions,
// O: pairs starting at odd positions).  Every row is then two independent carry chains made of
// (mad.lo.cc, madc.hi.cc) pairs, each pair = one IMAD.WIDE.U32.X in SASS.  The limb that would
// straddle the two accumulators after the implicit one-limb shift is folded with a single add.cc
// whose carry-out is consumed as carry-in by the next trailing chain.  fma-pipe instruction
// count: N*(2N) IMAD.WIDE for N limbs (both moduli are == 1 mod 2^32, so the Montgomery factor
// is a negation, not a multiplication).
#pragma once
#include "bls12_377_constants.cuh"
#include "ptx_arith.cuh"

// 32-byte alignment for Fr makes every element move one LDG.E.256 / STG.E.256 on sm_100a.
template <class P>
struct alignas((P::N % 8 == 0) ? 32 : 16) Fp {
  u32 l[P::N];
};
typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

template <class P>
DEV Fp<P> fp_zero() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.l[i] = 0;
  return r;
}

// constant from a generated accessor, e.g. fp_const<FrParams, FrParams::GENERATOR_M>()
template <class P, u32 (*F)(int)>
DEV Fp<P> fp_const() {
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < P::N; i++) r.l[i] = F(i);
  return r;
}

template <class P>
DEV Fp<P> fp_one() {
  return fp_const<P, P::ONE>();
}

template <class P>
DEV bool fp_is_zero(const Fp<P>& a) {
  u32 acc = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) acc |= a.l[i];
  return acc == 0;
}

template <class P>
DEV bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
  u32 acc = 0;
#pragma unroll
  for (int i = 0; i < P::N; i++) acc |= a.l[i] ^ b.l[i];
  return acc == 0;
}

// r = a - m if a >= m else a      (input < 2m)
template <class P>
DEV void fp_reduce_once(Fp<P>& a) {
  constexpr int N = P::N;
  u32 s[N];
  s[0] = ptx::sub_cc(a.l[0], P::MOD(0));
#pragma unroll
  for (int i = 1; i < N; i++) s[i] = ptx::subc_cc(a.l[i], P::MOD(i));
  u32 borrow = ptx::subc(0, 0);  // 0xffffffff when a < m
#pragma unroll
  for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : s[i];
}

template <class P>
DEV Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
  r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);  // m < 2^(32N-1): no carry out
  fp_reduce_once(r);
  return r;
}

template <class P>
DEV Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
  u32 mask = ptx::subc(0, 0);  // all ones when a < b
  r.l[0] = ptx::add_cc(r.l[0], P::MOD(0) & mask);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], P::MOD(i) & mask);
  r.l[N - 1] = ptx::addc(r.l[N - 1], P::MOD(N - 1) & mask);
  return r;
}

template <class P>
DEV Fp<P> fp_neg(const Fp<P>& a) {
  constexpr int N = P::N;
  Fp<P> r;
  u32 nz = 0;
#pragma unroll
  for (int i = 0; i < N; i++) nz |= a.l[i];
  r.l[0] = ptx::sub_cc(P::MOD(0), a.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(P::MOD(i), a.l[i]);
  r.l[N - 1] = ptx::subc(P::MOD(N - 1), a.l[N - 1]);
  if (nz == 0) r = a;  // -0 = 0 (keep the canonical representative)
  return r;
}

template <class P>
DEV Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add(a, a);
}

// ---------------------------------------------------------------------------------------------
// Montgomery product.  Absolute-position accumulators E / O, see the header comment.
// ---------------------------------------------------------------------------------------------
namespace montdetail {
// acc pairs (pos+j, pos+j+1) += x[j] * y for j = J0, J0+2, ... < N      (one carry chain)
// CARRY_IN: first instruction consumes CF.  TOP: add the chain's carry-out into acc[pos+N'].
// X_IS_MOD: the row operand is the modulus, read from the constant bank (uniform registers; an
// immediate would stop ptxas from fusing the pair into IMAD.WIDE) instead of x[].
// UNIT0 (modulus rows of the leading accumulator): limb 0 of both moduli is 1 and y = -acc[POS], so the first
// pair adds nothing but the carry of acc[POS] + y = (acc[POS] != 0), which the caller left in CF (neg_cc): one
// addc instead of a product (ptxas otherwise keeps an IMAD.HI.U32 on the heavy pipe for it).
template <class P, int POS, int J0, bool CARRY_IN, bool X_IS_MOD, int LEN, bool UNIT0 = false>
DEV void row_chain(u32 (&acc)[LEN], const u32* x, u32 y) {
  constexpr int N = P::N;
  static_assert(!UNIT0 || (X_IS_MOD && J0 == 0 && CARRY_IN && P::MOD(0) == 1u), "UNIT0: leading modulus row only");
  if (UNIT0) acc[POS + 1] = ptx::addc_cc(acc[POS + 1], 0);
#pragma unroll
  for (int j = UNIT0 ? 2 : J0; j < N; j += 2) {
    const u32 xj = X_IS_MOD ? P::MODC()[j] : x[j];
    if (j == J0 && !CARRY_IN)
      acc[POS + j] = ptx::mad_lo_cc(xj, y, acc[POS + j]);
    else
      acc[POS + j] = ptx::madc_lo_cc(xj, y, acc[POS + j]);
    if (J0 == 1 && j + 2 >= N)
      acc[POS + j + 1] = ptx::madc_hi(xj, y, acc[POS + j + 1]);  // trailing chain: bounded, no carry out
    else
      acc[POS + j + 1] = ptx::madc_hi_cc(xj, y, acc[POS + j + 1]);
  }
  if (J0 == 0) acc[POS + N] = ptx::addc(acc[POS + N], 0);
}

template <class P, int I, int LEN>
DEV void iter(u32 (&E)[LEN], u32 (&O)[LEN], const u32* a, u32 bi) {
  u32(&L)[LEN] = (I & 1) ? O : E;  // leading accumulator: owns the pair that starts at position I
  u32(&T)[LEN] = (I & 1) ? E : O;  // trailing accumulator: its limb at position I is a pair's high half
  if (I > 0) {
    L[I] = ptx::add_cc(L[I], T[I]);  // fold; carry-out belongs to position I+1 = first trailing pair
    row_chain<P, I, 1, true, false>(T, a, bi);
  } else {
    row_chain<P, I, 1, false, false>(T, a, bi);
  }
  row_chain<P, I, 0, false, false>(L, a, bi);
  u32 m = ptx::neg_cc(L[I]);  // Montgomery factor: INV == -1 mod 2^32 for both BLS12-377 moduli; CF = (L[I] != 0)
  row_chain<P, I, 0, true, true, LEN, true>(L, a, m);
  row_chain<P, I, 1, false, true>(T, a, m);
}

template <class P, int I, int LEN>
struct Unroll {
  static DEV void run(u32 (&E)[LEN], u32 (&O)[LEN], const u32* a, const u32* b) {
    Unroll<P, I - 1, LEN>::run(E, O, a, b);
    iter<P, I, LEN>(E, O, a, b[I]);
  }
};
template <class P, int LEN>
struct Unroll<P, -1, LEN> {
  static DEV void run(u32 (&)[LEN], u32 (&)[LEN], const u32*, const u32*) {}
};
}  // namespace montdetail

// (a * b + q * m) / R without the final conditional subtraction: < a*b/R + m.
template <class P>
DEV Fp<P> fp_mul_raw(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  static_assert(P::INV == 0xffffffffu, "multiplier assumes modulus == 1 mod 2^32");
  constexpr int LEN = 2 * N + 1;
  u32 E[LEN], O[LEN];
#pragma unroll
  for (int k = 0; k < LEN; k++) E[k] = O[k] = 0;
  montdetail::Unroll<P, N - 1, LEN>::run(E, O, a.l, b.l);
  Fp<P> r;
  r.l[0] = ptx::add_cc(E[N], O[N]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(E[N + k], O[N + k]);
  r.l[N - 1] = ptx::addc(E[2 * N - 1], O[2 * N - 1]);
  return r;
}

template <class P>
DEV Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
  Fp<P> r = fp_mul_raw(a, b);
  fp_reduce_once(r);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Lazy ("semi-reduced") arithmetic for kernels that chain butterflies and products (the NTT passes): values live
// in [0, 2m) and only the final result is brought back to [0, m).  The modulus must leave spare bits: with
// rho = m / R, a raw Montgomery product of a < A*m and b < B*m is < (A*B*rho + 1) * m.  For Fr rho = 0.0729
// (R / m = 13.7): A*B <= 8 gives < 1.59 m, so
//   lz_mul  (a < 4m, b < 2m)  -> < 2m   no conditional subtraction at all
//   lz_add  (a, b < 2m)       -> < 2m   a + b, minus 2m when that is not negative
//   lz_sub  (a, b < 2m)       -> < 2m   a - b, plus 2m when that borrowed
//   lz_add_wide / lz_sub_wide -> < 4m   a + b / a - b + 2m with no test: only as the operand of the next lz_mul
// ---------------------------------------------------------------------------------------------
template <class P>
CONSTFN u32 fp_mod2_limb(int i) {  // limb i of 2m
  return (P::MOD(i) << 1) | (i > 0 ? (P::MOD(i > 0 ? i - 1 : 0) >> 31) : 0u);
}

template <class P>
DEV Fp<P> lz_mul(const Fp<P>& a, const Fp<P>& b) {
  static_assert(32 * P::N - P::BITS >= 3, "lazy products need three spare bits");
  return fp_mul_raw(a, b);
}

template <class P>
DEV Fp<P> lz_add_wide(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
  r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);
  return r;
}

template <class P>
DEV Fp<P> lz_sub_wide(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  // (a + 2m) - b, in this order: the intermediate never goes negative
  r.l[0] = ptx::add_cc(a.l[0], fp_mod2_limb<P>(0));
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], fp_mod2_limb<P>(i));
  r.l[N - 1] = ptx::addc(a.l[N - 1], fp_mod2_limb<P>(N - 1));
  r.l[0] = ptx::sub_cc(r.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(r.l[i], b.l[i]);
  r.l[N - 1] = ptx::subc(r.l[N - 1], b.l[N - 1]);
  return r;
}

template <class P>
DEV Fp<P> lz_add(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r = lz_add_wide(a, b);
  u32 s[N];
  s[0] = ptx::sub_cc(r.l[0], fp_mod2_limb<P>(0));
#pragma unroll
  for (int i = 1; i < N; i++) s[i] = ptx::subc_cc(r.l[i], fp_mod2_limb<P>(i));
  u32 borrow = ptx::subc(0, 0);  // 0xffffffff when a + b < 2m
#pragma unroll
  for (int i = 0; i < N; i++) r.l[i] = borrow ? r.l[i] : s[i];
  return r;
}

template <class P>
DEV Fp<P> lz_sub(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
  u32 mask = ptx::subc(0, 0);  // all ones when a < b
  r.l[0] = ptx::add_cc(r.l[0], fp_mod2_limb<P>(0) & mask);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], fp_mod2_limb<P>(i) & mask);
  r.l[N - 1] = ptx::addc(r.l[N - 1], fp_mod2_limb<P>(N - 1) & mask);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Montgomery square.  The 2N x N multiplier rows of fp_mul become N(N-1)/2 off-diagonal products
// a_i * (2a)_j (i < j), N diagonal squares and N reduction rows: N(N+1)/2 + N^2 wide MACs instead of 2N^2
// (Fq: 222 instead of 288).  Same two-accumulator layout (E: pairs starting at even limb positions,
// O: odd), every row again two carry chains of IMAD.WIDE pairs.
//   phase 1  T = a^2 into E + O (2N limbs).  A row chain ends with one addc into the next limb: that limb has
//            so far only received carry bits of earlier rows (rows i' < i end below position i + N), so it
//            cannot overflow; the diagonal chain runs last and propagates through all 2N limbs.
//   phase 2  Montgomery-reduce the LOW halves only (upper limbs start at zero, exactly the situation the
//            fp_mul rows are bounded for) and add the high halves at the end:
//            (T + M q) / R = T_hi + (T_lo + M q) / R < 2q.
// ---------------------------------------------------------------------------------------------
namespace montdetail {
// acc pairs (POS+j, POS+j+1) += x[j] * y for j = JS, JS+2, ... < N (x[JS] replaced by `first`), carry-out into
// the limb above
template <class P, int POS, int JS, int LEN>
DEV void sqr_row_chain(u32 (&acc)[LEN], const u32* x, u32 first, u32 y) {
  constexpr int N = P::N;
  if (JS < N) {
#pragma unroll
    for (int j = JS; j < N; j += 2) {
      const u32 xj = (j == JS) ? first : x[j];
      if (j == JS)
        acc[POS + j] = ptx::mad_lo_cc(xj, y, acc[POS + j]);
      else
        acc[POS + j] = ptx::madc_lo_cc(xj, y, acc[POS + j]);
      acc[POS + j + 1] = ptx::madc_hi_cc(xj, y, acc[POS + j + 1]);
    }
    constexpr int LAST = JS + 2 * ((N - 1 - JS) / 2);
    acc[POS + LAST + 2] = ptx::addc(acc[POS + LAST + 2], 0);
  }
}

template <class P, int I, int LEN>
struct SqrRows {
  static DEV void run(u32 (&E)[LEN], u32 (&O)[LEN], const u32* a, const u32* b2) {
    SqrRows<P, I - 1, LEN>::run(E, O, a, b2);
    // row I multiplies a_I by the limbs of 2 * (a >> 32 (I+1)): limb I+1 of 2a without the bit that a_I shifts in
    sqr_row_chain<P, I, I + 1, LEN>(O, b2, a[I + 1] << 1, a[I]);                            // positions 2I+1, 2I+3, ...: odd
    sqr_row_chain<P, I, I + 2, LEN>(E, b2, (I + 2 < P::N) ? b2[(I + 2 < P::N) ? I + 2 : 0] : 0u, a[I]);  // 2I+2, 2I+4, ...: even
  }
};
template <class P, int LEN>
struct SqrRows<P, -1, LEN> {
  static DEV void run(u32 (&)[LEN], u32 (&)[LEN], const u32*, const u32*) {}
};

// reduction step I on the low halves: fold position I, m = -L[I], add m * modulus at positions I .. I+N
template <class P, int I, int LEN>
DEV void red_iter(u32 (&E)[LEN], u32 (&O)[LEN]) {
  u32(&L)[LEN] = (I & 1) ? O : E;
  u32(&T)[LEN] = (I & 1) ? E : O;
  if (I > 0) L[I] = ptx::add_cc(L[I], T[I]);  // carry-out belongs to position I+1 = first trailing pair
  // the fold's carry-out (position I+1) must be consumed by the trailing chain before neg_cc overwrites CF
  if (I > 0)
    row_chain<P, I, 1, true, true>(T, (const u32*)nullptr, ptx::neg_opaque(L[I]));
  else
    row_chain<P, I, 1, false, true>(T, (const u32*)nullptr, ptx::neg_opaque(L[I]));
  const u32 m = ptx::neg_cc(L[I]);
  row_chain<P, I, 0, true, true, LEN, true>(L, (const u32*)nullptr, m);
}
template <class P, int I, int LEN>
struct RedRows {
  static DEV void run(u32 (&E)[LEN], u32 (&O)[LEN]) {
    RedRows<P, I - 1, LEN>::run(E, O);
    red_iter<P, I, LEN>(E, O);
  }
};
template <class P, int LEN>
struct RedRows<P, -1, LEN> {
  static DEV void run(u32 (&)[LEN], u32 (&)[LEN]) {}
};
}  // namespace montdetail

template <class P>
DEV Fp<P> fp_sqr(const Fp<P>& a) {
  constexpr int N = P::N;
  static_assert(P::INV == 0xffffffffu, "multiplier assumes modulus == 1 mod 2^32");
  static_assert(P::BITS + 1 <= 32 * N, "2a must fit N limbs");
  constexpr int LEN = 2 * N + 1;
  u32 E[LEN], O[LEN], b2[N];
#pragma unroll
  for (int k = 0; k < LEN; k++) E[k] = O[k] = 0;
  b2[0] = a.l[0] << 1;
#pragma unroll
  for (int k = 1; k < N; k++) b2[k] = (a.l[k] << 1) | (a.l[k - 1] >> 31);
  // phase 1: off-diagonal rows, then the diagonal chain through all 2N limbs of E
  montdetail::SqrRows<P, N - 2, LEN>::run(E, O, a.l, b2);
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (i == 0)
      E[0] = ptx::mad_lo_cc(a.l[0], a.l[0], E[0]);
    else
      E[2 * i] = ptx::madc_lo_cc(a.l[i], a.l[i], E[2 * i]);
    if (i == N - 1)
      E[2 * i + 1] = ptx::madc_hi(a.l[i], a.l[i], E[2 * i + 1]);  // a^2 < 2^(64N): no carry out
    else
      E[2 * i + 1] = ptx::madc_hi_cc(a.l[i], a.l[i], E[2 * i + 1]);
  }
  // high halves aside (H = E_hi + O_hi < 2^(32N)), low halves reduced with fresh upper limbs
  u32 H[N];
  H[0] = ptx::add_cc(E[N], O[N]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) H[k] = ptx::addc_cc(E[N + k], O[N + k]);
  H[N - 1] = ptx::addc(E[2 * N - 1], O[2 * N - 1]);
#pragma unroll
  for (int k = N; k < LEN; k++) E[k] = O[k] = 0;
  montdetail::RedRows<P, N - 1, LEN>::run(E, O);
  Fp<P> r;
  r.l[0] = ptx::add_cc(E[N], O[N]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(E[N + k], O[N + k]);
  r.l[N - 1] = ptx::addc(E[2 * N - 1], O[2 * N - 1]);
  r.l[0] = ptx::add_cc(r.l[0], H[0]);
#pragma unroll
  for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(r.l[k], H[k]);
  r.l[N - 1] = ptx::addc(r.l[N - 1], H[N - 1]);
  fp_reduce_once(r);
  return r;
}

// Montgomery -> canonical (PrimeField::to_bigint): multiply by 1.
template <class P>
DEV Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> one = fp_zero<P>();
  one.l[0] = 1;
  return fp_mul(a, one);
}

template <class P>
DEV Fp<P> fp_to_mont(const Fp<P>& a) {
  return fp_mul(a, fp_const<P, P::R2>());
}

// a^e for a little-endian 32-bit-limb exponent (runtime loop: used off the hot path only).
template <class P>
DEV Fp<P> fp_pow(const Fp<P>& a, const u32* e, int nlimbs) {
  Fp<P> r = fp_one<P>();
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      if (started) r = fp_sqr(r);
      if ((e[i] >> b) & 1) {
        r = fp_mul(r, a);
        started = true;
      }
    }
  }
  return r;
}

template <class P>
DEV Fp<P> fp_pow_u64(const Fp<P>& a, u64 e) {
  u32 limbs[2] = {(u32)e, (u32)(e >> 32)};
  return fp_pow(a, limbs, 2);
}

// Fermat inversion a^(m-2); 0 -> 0.
template <class P>
DEV Fp<P> fp_inv(const Fp<P>& a) {
  u32 e[P::N];
#pragma unroll
  for (int i = 0; i < P::N; i++) e[i] = P::MOD_MINUS_2(i);
  return fp_pow(a, e, P::N);
}

// ---------------------------------------------------------------------------------------------
// Inversion by the binary extended Euclid algorithm (shifts and subtractions only).  One inversion ends
// every MSM (normalisation of the result) and sits on its latency path: Fermat costs ~570 dependent
// products (~0.45 ms on one B200 thread), this ~750 short carry chains (< 0.1 ms).
// Invariants: x1 * a == u, x2 * a == v (mod m); u, v odd after the halving loops; ends with u == 1 or v == 1.
// a is a Montgomery residue A*R; the loop yields (A*R)^-1 as a plain residue, and a Montgomery product with R^3
// (done by the caller, so that it can use its out-of-line multiplier) turns that into A^-1 * R.  0 -> 0.
// ---------------------------------------------------------------------------------------------
namespace bingcd {
template <int N>
DEV bool is_one(const u32* a) {
  u32 acc = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < N; i++) acc |= a[i];
  return acc == 0;
}
template <int N>
DEV void shr1(u32* a, u32 top) {
#pragma unroll
  for (int i = 0; i < N - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[N - 1] = (a[N - 1] >> 1) | (top << 31);
}
// x <- x / 2 mod m
template <class P>
DEV void halve(u32* x) {
  constexpr int N = P::N;
  u32 carry = 0;
  if (x[0] & 1u) {
    x[0] = ptx::add_cc(x[0], P::MOD(0));
#pragma unroll
    for (int i = 1; i < N; i++) x[i] = ptx::addc_cc(x[i], P::MOD(i));
    carry = ptx::addc(0, 0);
  }
  shr1<N>(x, carry);
}
template <int N>
DEV bool geq(const u32* a, const u32* b) {
  ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N; i++) ptx::subc_cc(a[i], b[i]);
  return ptx::subc(0, 0) == 0;  // no borrow
}
template <int N>
DEV void sub(u32* a, const u32* b) {
  a[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) a[i] = ptx::subc_cc(a[i], b[i]);
  a[N - 1] = ptx::subc(a[N - 1], b[N - 1]);
}
}  // namespace bingcd

template <class P>
DEV Fp<P> fp_inv_bingcd_raw(const Fp<P>& a) {
  constexpr int N = P::N;
  if (fp_is_zero(a)) return a;
  Fp<P> u = a, v, x1 = fp_zero<P>(), x2 = fp_zero<P>();
#pragma unroll
  for (int i = 0; i < N; i++) v.l[i] = P::MOD(i);
  x1.l[0] = 1;
  while (!bingcd::is_one<N>(u.l) && !bingcd::is_one<N>(v.l)) {
    while (!(u.l[0] & 1u)) {
      bingcd::shr1<N>(u.l, 0);
      bingcd::halve<P>(x1.l);
    }
    while (!(v.l[0] & 1u)) {
      bingcd::shr1<N>(v.l, 0);
      bingcd::halve<P>(x2.l);
    }
    if (bingcd::geq<N>(u.l, v.l)) {
      bingcd::sub<N>(u.l, v.l);
      x1 = fp_sub(x1, x2);
    } else {
      bingcd::sub<N>(v.l, u.l);
      x2 = fp_sub(x2, x1);
    }
  }
  return bingcd::is_one<N>(u.l) ? x1 : x2;  // (A*R)^-1 as a plain residue; the caller multiplies by R^3
}
