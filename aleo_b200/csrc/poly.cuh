// poly.cuh -- elementwise field arithmetic on device vectors: the pointwise work snarkVM's Varuna prover does on
// evaluation vectors between its FFTs (snarkvm-algorithms 0.14.5 src/fft/evaluations.rs `Evaluations` Mul / Sub /
// Add, src/fft/polynomial/dense.rs, snarkvm-fields `batch_inversion`), reached from the
// reference through prove_execution (rust/src/program/execute.rs:74,219).  HBM bound: 64-96 bytes per element
// against 120 IMAD.WIDE for a product, so every kernel is one coalesced pass with 256-bit accesses (Fr).
// Also the entry point the parity tests use to check the field core (mul / sqr / inv) against the oracle.
#pragma once
#include "mont.cuh"

namespace poly {

enum Op : int { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SQR = 3, OP_INV = 4, OP_NEG = 5 };

template <class P>
DEV Fp<P> load_elem(const void* base, size_t i) {
  Fp<P> v;
  if (P::N % 8 == 0) {
    v = reinterpret_cast<const Fp<P>*>(base)[i];  // 32-byte aligned: one 256-bit load
  } else {
    const uint2* q = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(base) + i * (P::N * 4));
#pragma unroll
    for (int k = 0; k < P::N / 2; k++) {
      const uint2 t = q[k];
      v.l[2 * k] = t.x;
      v.l[2 * k + 1] = t.y;
    }
  }
  return v;
}

template <class P>
DEV void store_elem(void* base, size_t i, const Fp<P>& v) {
  if (P::N % 8 == 0) {
    reinterpret_cast<Fp<P>*>(base)[i] = v;
  } else {
    uint2* q = reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(base) + i * (P::N * 4));
#pragma unroll
    for (int k = 0; k < P::N / 2; k++) q[k] = make_uint2(v.l[2 * k], v.l[2 * k + 1]);
  }
}

// Montgomery inverse by binary GCD (mont.cuh); 0 -> 0, as snarkVM's batch_inversion leaves zeros untouched
template <class P>
DEV Fp<P> inv_elem(const Fp<P>& a) {
  return fp_mul(fp_inv_bingcd_raw(a), fp_const<P, P::R3>());
}

template <class P, int OP>
KERNEL void __launch_bounds__(256) field_op_kernel(void* out, const void* a, const void* b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const Fp<P> x = load_elem<P>(a, i);
    Fp<P> r;
    if (OP == OP_ADD) r = fp_add(x, load_elem<P>(b, i));
    if (OP == OP_SUB) r = fp_sub(x, load_elem<P>(b, i));
    if (OP == OP_MUL) r = fp_mul(x, load_elem<P>(b, i));
    if (OP == OP_SQR) r = fp_sqr(x);
    if (OP == OP_INV) r = inv_elem<P>(x);
    if (OP == OP_NEG) r = fp_neg(x);
    store_elem<P>(out, i, r);
  }
}

}  // namespace poly
