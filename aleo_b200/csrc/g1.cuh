// g1.cuh -- BLS12-377 G1 (y^2 = x^3 + 1 over Fq) point arithmetic for the MSM kernels.
//
// Data contracts (snarkvm-curves 0.14.5 bls12_377::{G1Affine, G1Projective}):
//   affine in memory : x | y Montgomery Fq, stride 104 (Rust struct, bool `infinity` at byte 96)
//                      or stride 96 (packed; identity = (0, 0)).  Only 8-byte alignment is assumed.
//   result           : Jacobian (X, Y, Z), x = X/Z^2, y = Y/Z^3, identity z = 0 -- returned normalised.
// Buckets are kept in XYZZ coordinates (X, Y, ZZ, ZZZ), x = X/ZZ, y = Y/ZZZ, identity ZZ = 0:
// a mixed addition costs 8M + 2S, a full addition 12M + 2S, no inversions anywhere in the hot loop.
#pragma once
#include "mont.cuh"

struct G1Affine {
  Fq x, y;
  bool inf;
};

struct G1Xyzz {
  Fq x, y, zz, zzz;
};

DEV G1Xyzz xyzz_identity() {
  G1Xyzz r;
  r.x = fp_zero<FqParams>();
  r.y = fp_zero<FqParams>();
  r.zz = fp_zero<FqParams>();
  r.zzz = fp_zero<FqParams>();
  return r;
}

DEV bool xyzz_is_identity(const G1Xyzz& p) { return fp_is_zero(p.zz); }

DEV G1Xyzz xyzz_from_affine(const G1Affine& a) {
  if (a.inf) return xyzz_identity();
  G1Xyzz r;
  r.x = a.x;
  r.y = a.y;
  r.zz = fp_one<FqParams>();
  r.zzz = r.zz;
  return r;
}

// 48-byte Fq from memory that is only guaranteed 8-byte aligned
DEV Fq fq_load8(const void* p) {
  const uint2* q = reinterpret_cast<const uint2*>(p);
  Fq r;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint2 v = q[i];
    r.l[2 * i] = v.x;
    r.l[2 * i + 1] = v.y;
  }
  return r;
}
DEV void fq_store8(void* p, const Fq& v) {
  uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
  for (int i = 0; i < 6; i++) q[i] = make_uint2(v.l[2 * i], v.l[2 * i + 1]);
}

DEV G1Affine affine_load(const unsigned char* bases, size_t stride, size_t idx) {
  const unsigned char* p = bases + idx * stride;
  G1Affine a;
  a.x = fq_load8(p);
  a.y = fq_load8(p + 48);
  if (stride >= 97)
    a.inf = p[96] != 0;
  else
    a.inf = fp_is_zero(a.x) && fp_is_zero(a.y);
  return a;
}

DEV void affine_store(unsigned char* bases, size_t stride, size_t idx, const G1Affine& a) {
  unsigned char* p = bases + idx * stride;
  if (a.inf) {
    // Rust: Affine::zero() = (0, 1, infinity = true); packed: (0, 0)
    fq_store8(p, fp_zero<FqParams>());
    fq_store8(p + 48, stride >= 97 ? fp_one<FqParams>() : fp_zero<FqParams>());
  } else {
    fq_store8(p, a.x);
    fq_store8(p + 48, a.y);
  }
  if (stride >= 97) {
    uint2* flag = reinterpret_cast<uint2*>(p + 96);  // flag byte + 7 bytes of struct padding
    *flag = make_uint2(a.inf ? 1u : 0u, 0u);
  }
}

// Out-of-line Fq product with by-value operands: ptxas passes them in registers (no stack), so a call
// costs ~36 MOVs on the ALU pipe while the loop body that uses it shrinks ~10x and stays resident in
// the instruction cache (the fully inlined accumulation loop is ~115 KB of SASS and was stalled on
// instruction fetch for 10-40 % of its issue slots, profiles/r01_ncu_hot_kernels.md).
DEV_NOINLINE Fq fq_mul_v(Fq a, Fq b) { return fp_mul(a, b); }
// dedicated square (mont.cuh fp_sqr: 222 wide MACs instead of 288), same calling convention
DEV_NOINLINE Fq fq_sqr_v(Fq a) { return fp_sqr(a); }

// 2 * (x, y) for an affine point that is not the identity (mdbl-2008-s-1, a = 0).
// Out of line: only reached when a bucket receives the same point twice.
// (Inlined products on purpose: routing them through fq_mul_v changes the register assignment around the
// call sites of the accumulation loop and cost 4 % of its throughput on B200.)
DEV_NOINLINE G1Xyzz xyzz_double_affine(const Fq& x, const Fq& y) {
  G1Xyzz r;
  Fq u = fp_dbl(y);
  Fq v = fp_sqr(u);
  Fq w = fp_mul(u, v);
  Fq s = fp_mul(x, v);
  Fq xx = fp_sqr(x);
  Fq m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fp_sqr(m), fp_dbl(s));
  r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * p (dbl-2008-s-1, a = 0); identity and y = 0 map to the identity.  Out of line (tail kernels only).
DEV_NOINLINE G1Xyzz xyzz_double(const G1Xyzz& p) {
  if (xyzz_is_identity(p) || fp_is_zero(p.y)) return xyzz_identity();
  G1Xyzz r;
  Fq u = fp_dbl(p.y);
  Fq v = fq_sqr_v(u);
  Fq w = fq_mul_v(u, v);
  Fq s = fq_mul_v(p.x, v);
  Fq xx = fq_sqr_v(p.x);
  Fq m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fq_sqr_v(m), fp_dbl(s));
  r.y = fp_sub(fq_mul_v(m, fp_sub(s, r.x)), fq_mul_v(w, p.y));
  r.zz = fq_mul_v(v, p.zz);
  r.zzz = fq_mul_v(w, p.zzz);
  return r;
}

template <bool CALL>
DEV Fq fq_mul_sel(const Fq& a, const Fq& b) {
  if (CALL) return fq_mul_v(a, b);
  return fp_mul(a, b);
}
template <bool CALL>
DEV Fq fq_sqr_sel(const Fq& a) {
  if (CALL) return fq_sqr_v(a);
  return fp_sqr(a);
}

// acc += (x2, y2)   (madd-2008-s; the affine operand is not the identity)
template <bool CALL>
DEV void xyzz_add_affine_t(G1Xyzz& acc, const Fq& x2, const Fq& y2);

DEV void xyzz_add_affine(G1Xyzz& acc, const Fq& x2, const Fq& y2) { xyzz_add_affine_t<false>(acc, x2, y2); }

template <bool CALL>
DEV void xyzz_add_affine_t(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
  if (xyzz_is_identity(acc)) {
    acc.x = x2;
    acc.y = y2;
    acc.zz = fp_one<FqParams>();
    acc.zzz = acc.zz;
    return;
  }
  Fq u2 = fq_mul_sel<CALL>(x2, acc.zz);
  Fq s2 = fq_mul_sel<CALL>(y2, acc.zzz);
  Fq p = fp_sub(u2, acc.x);
  Fq r = fp_sub(s2, acc.y);
  if (fp_is_zero(p)) {  // same x: doubling or cancellation (rare; the reference's edge-case tests hit it)
    if (fp_is_zero(r))
      acc = xyzz_double_affine(x2, y2);
    else
      acc = xyzz_identity();
    return;
  }
  Fq pp = fq_sqr_sel<CALL>(p);
  Fq ppp = fq_mul_sel<CALL>(p, pp);
  Fq q = fq_mul_sel<CALL>(acc.x, pp);
  Fq x3 = fp_sub(fp_sub(fq_sqr_sel<CALL>(r), ppp), fp_dbl(q));
  Fq y3 = fp_sub(fq_mul_sel<CALL>(r, fp_sub(q, x3)), fq_mul_sel<CALL>(acc.y, ppp));
  acc.zz = fq_mul_sel<CALL>(acc.zz, pp);
  acc.zzz = fq_mul_sel<CALL>(acc.zzz, ppp);
  acc.x = x3;
  acc.y = y3;
}

// acc += b   (add-2008-s).  All products go through the out-of-line multiplier: the tail kernels run a handful of
// warps, and 14 inlined products (~80 KB of SASS per addition) made them instruction-fetch bound.
DEV void xyzz_add(G1Xyzz& acc, const G1Xyzz& b) {
  if (xyzz_is_identity(b)) return;
  if (xyzz_is_identity(acc)) {
    acc = b;
    return;
  }
  Fq u1 = fq_mul_v(acc.x, b.zz);
  Fq u2 = fq_mul_v(b.x, acc.zz);
  Fq s1 = fq_mul_v(acc.y, b.zzz);
  Fq s2 = fq_mul_v(b.y, acc.zzz);
  Fq p = fp_sub(u2, u1);
  Fq r = fp_sub(s2, s1);
  if (fp_is_zero(p)) {
    if (fp_is_zero(r))
      acc = xyzz_double(acc);
    else
      acc = xyzz_identity();
    return;
  }
  Fq pp = fq_sqr_v(p);
  Fq ppp = fq_mul_v(p, pp);
  Fq q = fq_mul_v(u1, pp);
  Fq x3 = fp_sub(fp_sub(fq_sqr_v(r), ppp), fp_dbl(q));
  Fq y3 = fp_sub(fq_mul_v(r, fp_sub(q, x3)), fq_mul_v(s1, ppp));
  acc.zz = fq_mul_v(fq_mul_v(acc.zz, b.zz), pp);
  acc.zzz = fq_mul_v(fq_mul_v(acc.zzz, b.zzz), ppp);
  acc.x = x3;
  acc.y = y3;
}

DEV_NOINLINE Fq fq_mul_ni(const Fq& a, const Fq& b);
DEV_NOINLINE Fq fq_inv_ni(const Fq& a);

// Jacobian (X, Y, Z) -> XYZZ
DEV G1Xyzz xyzz_from_jacobian(const Fq& X, const Fq& Y, const Fq& Z) {
  if (fp_is_zero(Z)) return xyzz_identity();
  G1Xyzz r;
  r.x = X;
  r.y = Y;
  r.zz = fq_sqr_v(Z);
  r.zzz = fq_mul_v(r.zz, Z);
  return r;
}

// XYZZ -> normalised affine coordinates (one Fermat inversion).  Returns false for the identity.
DEV bool xyzz_to_affine(const G1Xyzz& p, Fq& x, Fq& y) {
  if (xyzz_is_identity(p)) return false;
  Fq inv = fq_inv_ni(fq_mul_ni(p.zz, p.zzz));
  x = fq_mul_ni(p.x, fq_mul_ni(inv, p.zzz));
  y = fq_mul_ni(p.y, fq_mul_ni(inv, p.zz));
  return true;
}

// 144-byte normalised Jacobian image: (x, y, 1) or (0, 1, 0)
DEV void jacobian_store_normalised(unsigned char* out, const G1Xyzz& p) {
  Fq x, y;
  if (xyzz_to_affine(p, x, y)) {
    fq_store8(out, x);
    fq_store8(out + 48, y);
    fq_store8(out + 96, fp_one<FqParams>());
  } else {
    fq_store8(out, fp_zero<FqParams>());
    fq_store8(out + 48, fp_one<FqParams>());
    fq_store8(out + 96, fp_zero<FqParams>());
  }
}

// out-of-line variants for the tail kernels (reduction, combine, generators): keeps their code
// size and ptxas time small; the bucket-accumulation loop uses the inlined xyzz_add_affine above.
DEV_NOINLINE void xyzz_add_ni(G1Xyzz& acc, const G1Xyzz& b) { xyzz_add(acc, b); }
DEV_NOINLINE void xyzz_add_affine_ni(G1Xyzz& acc, const Fq& x2, const Fq& y2) { xyzz_add_affine(acc, x2, y2); }
DEV_NOINLINE Fq fq_mul_ni(const Fq& a, const Fq& b) { return fq_mul_v(a, b); }
DEV_NOINLINE Fq fq_inv_ni(const Fq& a) { return fq_mul_v(fp_inv_bingcd_raw(a), fp_const<FqParams, FqParams::R3>()); }
