// poly.cuh -- elementwise field arithmetic on device vectors: the pointwise work snarkVM's Varuna prover does on
// evaluation vectors between its FFTs (snarkvm-algorithms 0.14.5 src/fft/evaluations.rs `Evaluations` Mul / Sub /
// Add, src/fft/polynomial/dense.rs, snarkvm-fields `batch_inversion`), reached from the
// reference through prove_execution (rust/src/program/execute.rs:74,219).  HBM bound: 64-96 bytes per element
// against 120 IMAD.WIDE for a product, so every kernel is one coalesced pass with 256-bit accesses (Fr).
// Also the entry point the parity tests use to check the field core (mul / sqr / inv) against the oracle.
#pragma once
#include "g1.cuh"

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

// =============================================================================================
// Polynomial helpers of the KZG / Varuna call sites (Fr only; all vectors in Montgomery form).
// =============================================================================================
namespace poly {

constexpr u32 PW_TPB = 256;   // threads per CTA
constexpr u32 PW_RUN = 32;    // elements per thread, strided by PW_TPB: a CTA covers 8192 contiguous elements

// scalars every kernel below needs, computed once per call by setup_kernel:
//   s[0] = k (or 1), s[1] = g, s[2] = g^PW_TPB, s[3] = g^-1 (0 if g = 0), s[4] = g^-PW_TPB
KERNEL void setup_kernel(Fr* s, Fr g, Fr k, u32 have_k) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  s[0] = have_k ? k : fp_one<FrParams>();
  s[1] = g;
  Fr p = g;
  for (u32 i = 1; i < PW_TPB; i <<= 1) p = fp_sqr(p);
  s[2] = p;
  const Fr gi = inv_elem<FrParams>(g);
  s[3] = gi;
  p = gi;
  for (u32 i = 1; i < PW_TPB; i <<= 1) p = fp_sqr(p);
  s[4] = p;
}

// first index of this thread and the power base^(first + shift), times `scale`
DEV Fr thread_power(const Fr& base, const Fr& scale, u64 first, u64 shift) {
  return fp_mul(scale, fp_pow_u64(base, first + shift));
}

// a[i] <- a[i] * k * g^i      (EvaluationDomain::distribute_powers / distribute_powers_and_mul_by_const)
KERNEL void __launch_bounds__(PW_TPB) distribute_powers_kernel(Fr* a, u64 n, const Fr* s) {
  const u64 cta_base = (u64)blockIdx.x * PW_TPB * PW_RUN;
  const u64 first = cta_base + threadIdx.x;
  if (first >= n) return;
  Fr p = thread_power(s[1], s[0], first, 0);
  const Fr step = s[2];
#pragma unroll 4
  for (u32 j = 0; j < PW_RUN; j++) {
    const u64 i = first + (u64)j * PW_TPB;
    if (i >= n) break;
    a[i] = fp_mul(a[i], p);
    p = fp_mul(p, step);
  }
}

// y[i] <- y[i] + a * x[i]: the linear-combination step of SonicKZG10::open_combinations / batch_open (sum_i xi^i p_i)
KERNEL void __launch_bounds__(256) axpy_kernel(Fr* y, const Fr* x, Fr a, u64 n) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) y[i] = fp_add(y[i], fp_mul(a, x[i]));
}

// block-wide sum of one Fr per thread (blockDim.x = PW_TPB), result valid in thread 0
DEV Fr block_sum(Fr v, Fr* sh) {
  sh[threadIdx.x] = v;
  SYNC_THREADS();
  for (u32 off = PW_TPB / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] = fp_add(sh[threadIdx.x], sh[threadIdx.x + off]);
    SYNC_THREADS();
  }
  return sh[0];
}

// partial[cta] = sum over the CTA's 8192 elements of c[i] * z^i      (DensePolynomial::evaluate, first stage)
KERNEL void __launch_bounds__(PW_TPB) eval_partial_kernel(const Fr* c, u64 n, const Fr* s, Fr* partial) {
  SHARED Fr sh[PW_TPB];
  const u64 first = (u64)blockIdx.x * PW_TPB * PW_RUN + threadIdx.x;
  Fr acc = fp_zero<FrParams>();
  if (first < n) {
    Fr p = thread_power(s[1], fp_one<FrParams>(), first, 0);
    const Fr step = s[2];
    for (u32 j = 0; j < PW_RUN; j++) {
      const u64 i = first + (u64)j * PW_TPB;
      if (i >= n) break;
      acc = fp_add(acc, fp_mul(c[i], p));
      p = fp_mul(p, step);
    }
  }
  const Fr tot = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// out = sum of `count` partials (single CTA)
KERNEL void __launch_bounds__(PW_TPB) sum_partials_kernel(const Fr* partial, u32 count, Fr* out) {
  SHARED Fr sh[PW_TPB];
  Fr acc = fp_zero<FrParams>();
  for (u32 i = threadIdx.x; i < count; i += PW_TPB) acc = fp_add(acc, partial[i]);
  const Fr tot = block_sum(acc, sh);
  if (threadIdx.x == 0) *out = tot;
}

// ---- witness polynomial of KZG10::open: q(x) = (p(x) - p(z)) / (x - z), q_j = sum_{i > j} p_i z^(i-j-1) ----------
// With t_i = p_i z^i and the suffix sums S_j = sum_{i >= j} t_i:  q_j = S_(j+1) * z^-(j+1).  Three stages:
//   suffix_local : t_i and the suffix sums inside a CTA's 8192-element slab (thread-strided layout turned into
//                  a contiguous one through the scan), slab totals out
//   suffix_carry : exclusive suffix scan of the slab totals (single CTA; <= 2^19 slabs for 2^32 coefficients)
//   witness_fix  : q_j = (local S_(j+1) + carry of the slab) * z^-(j+1)
// z = 0 needs no arithmetic (q_j = p_(j+1)) and is handled by the host.
constexpr u32 SLAB = PW_TPB * 8;  // 2048 contiguous elements per CTA for the scan kernels, 8 consecutive per thread

KERNEL void __launch_bounds__(PW_TPB) suffix_local_kernel(const Fr* p, u64 n, const Fr* s, Fr* S, Fr* slab_total) {
  SHARED Fr sh[PW_TPB];
  const u64 base = (u64)blockIdx.x * SLAB + (u64)threadIdx.x * 8;
  // t_i for this thread's 8 consecutive elements, then their local suffix sums
  Fr t[8];
  Fr pw = (base < n) ? thread_power(s[1], fp_one<FrParams>(), base, 0) : fp_zero<FrParams>();
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const u64 i = base + k;
    t[k] = (i < n) ? fp_mul(p[i], pw) : fp_zero<FrParams>();
    pw = fp_mul(pw, s[1]);
  }
#pragma unroll
  for (int k = 6; k >= 0; k--) t[k] = fp_add(t[k], t[k + 1]);
  // exclusive suffix scan of the per-thread totals across the CTA (Hillis-Steele on 256 values)
  sh[threadIdx.x] = t[0];
  SYNC_THREADS();
  for (u32 off = 1; off < PW_TPB; off <<= 1) {
    Fr add = fp_zero<FrParams>();
    const bool take = threadIdx.x + off < PW_TPB;
    if (take) add = sh[threadIdx.x + off];
    SYNC_THREADS();
    if (take) sh[threadIdx.x] = fp_add(sh[threadIdx.x], add);
    SYNC_THREADS();
  }
  const Fr after = (threadIdx.x + 1 < PW_TPB) ? sh[threadIdx.x + 1] : fp_zero<FrParams>();
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const u64 i = base + k;
    if (i < n) S[i] = fp_add(t[k], after);
  }
  if (threadIdx.x == 0) slab_total[blockIdx.x] = sh[0];
}

// carry[b] = sum of slab totals of the slabs after b (single CTA, walks the slabs from the top down)
KERNEL void __launch_bounds__(PW_TPB) suffix_carry_kernel(const Fr* slab_total, u32 nslabs, Fr* carry) {
  SHARED Fr sh[PW_TPB];
  SHARED Fr running;
  if (threadIdx.x == 0) running = fp_zero<FrParams>();
  SYNC_THREADS();
  const u32 rounds = (nslabs + PW_TPB - 1) / PW_TPB;
  for (u32 r = 0; r < rounds; r++) {
    // this round covers slabs hi-1 .. hi-256 (descending); thread t takes slab hi-1-t
    const long long idx = (long long)nslabs - 1 - (long long)r * PW_TPB - threadIdx.x;
    const Fr v = (idx >= 0) ? slab_total[idx] : fp_zero<FrParams>();
    // inclusive prefix scan over t (= suffix over slabs)
    sh[threadIdx.x] = v;
    SYNC_THREADS();
    for (u32 off = 1; off < PW_TPB; off <<= 1) {
      Fr add = fp_zero<FrParams>();
      const bool take = threadIdx.x >= off;
      if (take) add = sh[threadIdx.x - off];
      SYNC_THREADS();
      if (take) sh[threadIdx.x] = fp_add(sh[threadIdx.x], add);
      SYNC_THREADS();
    }
    const Fr incl = sh[threadIdx.x];
    const Fr run = running;
    if (idx >= 0) carry[idx] = fp_add(run, fp_sub(incl, v));  // slabs strictly after idx
    SYNC_THREADS();
    if (threadIdx.x == PW_TPB - 1) running = fp_add(run, incl);
    SYNC_THREADS();
  }
}

// q[j] = (S[j+1] + carry of slab(j+1)) * z^-(j+1), j < n-1
KERNEL void __launch_bounds__(PW_TPB) witness_fix_kernel(const Fr* S, const Fr* carry, u64 n, const Fr* s, Fr* q) {
  const u64 first = (u64)blockIdx.x * PW_TPB * PW_RUN + threadIdx.x;  // j
  if (first + 1 >= n) return;
  Fr pw = thread_power(s[3], fp_one<FrParams>(), first, 1);  // z^-(j+1)
  const Fr step = s[4];
  for (u32 k = 0; k < PW_RUN; k++) {
    const u64 j = first + (u64)k * PW_TPB;
    if (j + 1 >= n) break;
    const Fr full = fp_add(S[j + 1], carry[(j + 1) / SLAB]);
    q[j] = fp_mul(full, pw);
    pw = fp_mul(pw, step);
  }
}

// EvaluationDomain::evaluate_all_lagrange_coefficients(tau): L_i(tau) = (tau^n - 1) / n * w^i / (tau - w^i), i < n.
// s[0] = (tau^n - 1) / n, s[1] = w, s[2] = w^256 (setup as for distribute_powers with g = w, k = that constant).
// When tau lies in the domain (some tau - w^i = 0) upstream returns the indicator vector: *hit records that index + 1.
KERNEL void __launch_bounds__(PW_TPB) lagrange_kernel(Fr* out, u64 n, const Fr* s, Fr tau, u32* hit) {
  const u64 first = (u64)blockIdx.x * PW_TPB * PW_RUN + threadIdx.x;
  if (first >= n) return;
  Fr w = thread_power(s[1], fp_one<FrParams>(), first, 0);
  const Fr step = s[2], c = s[0];
  for (u32 j = 0; j < PW_RUN; j++) {
    const u64 i = first + (u64)j * PW_TPB;
    if (i >= n) break;
    const Fr d = fp_sub(tau, w);
    if (fp_is_zero(d)) *hit = (u32)i + 1u;
    out[i] = fp_mul(fp_mul(c, w), inv_elem<FrParams>(d));
    w = fp_mul(w, step);
  }
}

// tau = w^(idx): out = indicator of idx
KERNEL void lagrange_indicator_kernel(Fr* out, u64 n, const u32* hit) {
  const u32 h = *hit;
  if (h == 0) return;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    out[i] = (i + 1 == h) ? fp_one<FrParams>() : fp_zero<FrParams>();
}

// s[0] <- (tau^n - 1) * n^-1 for n = 2^log_n (single thread; n^-1 = (1/2)^log_n)
KERNEL void lagrange_const_kernel(Fr* s, Fr tau, u32 log_n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fr t = tau;
  for (u32 i = 0; i < log_n; i++) t = fp_sqr(t);
  Fr ninv = fp_one<FrParams>();
  const Fr half = fp_const<FrParams, FrParams::TWO_INV_M>();
  for (u32 i = 0; i < log_n; i++) ninv = fp_mul(ninv, half);
  s[0] = fp_mul(fp_sub(t, fp_one<FrParams>()), ninv);
}

// domain generator w = TWO_ADIC_ROOT^(2^(47 - log_n)) (Montgomery), for the host wrapper
KERNEL void domain_root_kernel(Fr* out, u32 log_n) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Fr w = fp_const<FrParams, FrParams::TWO_ADIC_ROOT_M>();
  for (u32 i = log_n; i < (u32)FR_TWO_ADICITY; i++) w = fp_sqr(w);
  *out = w;
}

// Division by the vanishing polynomial Z_n(x) = x^n - 1 of the domain of size n = 2^log_n on the coset g H_m of a
// domain of size m = 2^log_m >= n (EvaluationDomain::divide_by_vanishing_poly_on_coset_in_place, src/fft/domain.rs, and
// the mul-domain quotients of the prover rounds): x_i^n = g^n * w_k^i with k = m / n and w_k the generator of the
// domain of size k, so the divisor has period k.  table[j] = (g^n * w_k^j - 1)^-1 (a zero divisor stays 0, as
// batch_inversion leaves zeros untouched).
KERNEL void __launch_bounds__(256) vanishing_table_kernel(Fr* table, u32 log_k, u32 log_n, Fr g) {
  const u32 k = 1u << log_k;
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    Fr gn = g;
    for (u32 i = 0; i < log_n; i++) gn = fp_sqr(gn);
    Fr w = fp_const<FrParams, FrParams::TWO_ADIC_ROOT_M>();
    for (u32 i = log_k; i < (u32)FR_TWO_ADICITY; i++) w = fp_sqr(w);
    Fr acc = gn;                                  // g^n * w^j by square-and-multiply over the bits of j
    for (u32 b = 0; b < log_k; b++) {
      if ((j >> b) & 1) acc = fp_mul(acc, w);
      w = fp_sqr(w);
    }
    table[j] = inv_elem<FrParams>(fp_sub(acc, fp_one<FrParams>()));
  }
}

KERNEL void __launch_bounds__(256) mul_periodic_kernel(Fr* e, u64 n, const Fr* table, u32 mask) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
    e[i] = fp_mul(e[i], table[(u32)i & mask]);
}

// z = 0: q_j = p_(j+1)
KERNEL void shift_down_kernel(const Fr* p, u64 n, Fr* q) {
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j + 1 < n; j += (u64)gridDim.x * blockDim.x) q[j] = p[j + 1];
}

}  // namespace poly

// =============================================================================================
// Compressed G1 wire format (snarkVM CanonicalSerialize / CanonicalDeserialize of G1Affine, the format of every
// commitment inside a proof and of the points in key files; pinned by the reference's own proof string
// wasm/src/programs/transaction.rs:100): 48 bytes, x canonical little-endian, bit 383 = y is the
// larger root, bit 382 = infinity.  SURVEY.md 8f rank 3.
// =============================================================================================
namespace wire {

// square root in Fq by Tonelli-Shanks (p - 1 = 2^46 t; z = 15^t); returns false when a is not a square
DEV bool fq_sqrt(const Fq& a, Fq& root) {
  if (fp_is_zero(a)) {
    root = a;
    return true;
  }
  u32 e[FqParams::N];
#pragma unroll
  for (int i = 0; i < FqParams::N; i++) e[i] = FqParams::TS_T_MINUS_1_DIV_2(i);
  const Fq w0 = fp_pow(a, e, FqParams::N);  // a^((t-1)/2)
  Fq x = fp_mul(a, w0);                     // a^((t+1)/2)
  Fq b = fp_mul(x, w0);                     // a^t
  Fq z = fp_const<FqParams, FqParams::TS_ROOT_M>();
  const Fq one = fp_one<FqParams>();
  u32 v = FQ_TWO_ADICITY;
  while (!fp_eq(b, one)) {
    u32 k = 0;
    Fq b2 = b;
    while (!fp_eq(b2, one)) {  // least k with b^(2^k) = 1
      b2 = fp_sqr(b2);
      k++;
      if (k == v) return false;  // a is not a square
    }
    Fq w = z;
    for (u32 i = 0; i + k + 1 < v; i++) w = fp_sqr(w);  // z^(2^(v-k-1))
    z = fp_sqr(w);
    b = fp_mul(b, z);
    x = fp_mul(x, w);
    v = k;
  }
  root = x;
  return true;
}

// y > (p - 1) / 2 for a canonical (non-Montgomery) y
DEV bool is_larger_root(const Fq& y_canon) {
  for (int k = FqParams::N - 1; k >= 0; k--) {
    const u32 h = FqParams::HALF_P_MINUS_1(k);
    if (y_canon.l[k] != h) return y_canon.l[k] > h;
  }
  return false;
}

DEV void store_affine(unsigned char* out, u32 stride, size_t i, const Fq& x, const Fq& y, bool inf) {
  uint2* q = reinterpret_cast<uint2*>(out + i * stride);
  const Fq zero = fp_zero<FqParams>();
  const Fq ox = inf ? zero : x;
  const Fq oy = inf ? (stride >= 97 ? fp_one<FqParams>() : zero) : y;  // Rust: Affine::zero() = (0, 1, infinity)
#pragma unroll
  for (int k = 0; k < 6; k++) q[k] = make_uint2(ox.l[2 * k], ox.l[2 * k + 1]);
#pragma unroll
  for (int k = 0; k < 6; k++) q[6 + k] = make_uint2(oy.l[2 * k], oy.l[2 * k + 1]);
  if (stride >= 97) q[12] = make_uint2(inf ? 1u : 0u, 0u);
}

// Is the curve point (x, y) in the prime-order subgroup G1?  BLS12-377's G1 cofactor is (u - 1)^2 / 3, so a point
// that merely satisfies the curve equation can have a component outside the r-torsion; snarkVM's
// deserialize_compressed (Validate::Yes) calls is_in_correct_subgroup_assuming_on_curve for exactly that.
// Test: phi(P) == -[u^2] P with phi(x, y) = (beta x, y) (tools/gen_constants.py subgroup_beta: phi^2 + phi + 1 = 0 on
// the whole curve, so the equation forces [u^4 - u^2 + 1] P = [r] P = O; on G1 phi is multiplication by -u^2).
// Cost: two multiplications by the 64-bit u = 0x8508c00000000001 (7 bits set): 126 doublings + 12 additions.
DEV G1Xyzz mul_by_u(const G1Xyzz& p) {
  constexpr unsigned long long U = 0x8508c00000000001ull;
  G1Xyzz acc = p;
  for (int b = 62; b >= 0; b--) {
    acc = xyzz_double(acc);
    if ((U >> b) & 1ull) xyzz_add_ni(acc, p);
  }
  return acc;
}

DEV bool g1_in_subgroup(const Fq& x, const Fq& y) {
  G1Xyzz p;
  p.x = x;
  p.y = y;
  p.zz = fp_one<FqParams>();
  p.zzz = p.zz;
  const G1Xyzz q = mul_by_u(mul_by_u(p));  // [u^2] P
  if (xyzz_is_identity(q)) return false;    // phi(P) is a finite point
  // -phi(P) = (beta x, -y) == (q.x / q.zz, q.y / q.zzz)
  const Fq bx = fq_mul_ni(fp_const<FqParams, FqParams::BETA_M>(), x);
  return fp_eq(fq_mul_ni(bx, q.zz), q.x) && fp_eq(fq_mul_ni(fp_neg(y), q.zzz), q.y);
}

// out[i] = affine point of in48[i]; invalid encodings (x >= p, x^3 + 1 not a square, and -- unless UNCHECKED -- a curve
// point outside the prime-order subgroup; stray bits with the infinity flag are ignored as upstream does) are stored as
// the identity and counted in *bad
template <bool UNCHECKED>
KERNEL void __launch_bounds__(128) g1_decompress_kernel(const unsigned char* in48, u32 n, unsigned char* out, u32 stride, u32* bad) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2* q = reinterpret_cast<const uint2*>(in48 + (size_t)i * 48);
  Fq xc;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const uint2 t = q[k];
    xc.l[2 * k] = t.x;
    xc.l[2 * k + 1] = t.y;
  }
  const u32 top = xc.l[FqParams::N - 1];
  const bool y_flag = (top >> 31) & 1u, inf_flag = (top >> 30) & 1u;
  xc.l[FqParams::N - 1] = top & 0x3fffffffu;
  if (inf_flag) {
    store_affine(out, stride, i, xc, xc, true);
    return;
  }
  // x < p ?
  bool lt = false;
  for (int k = FqParams::N - 1; k >= 0; k--) {
    if (xc.l[k] != FqParams::MOD(k)) {
      lt = xc.l[k] < FqParams::MOD(k);
      break;
    }
  }
  Fq x = fp_to_mont(xc), y;
  const Fq rhs = fp_add(fp_mul(fp_sqr(x), x), fp_one<FqParams>());
  if (!lt || !fq_sqrt(rhs, y)) {
    atomic_add_u32(bad, 1u);
    store_affine(out, stride, i, x, x, true);
    return;
  }
  if (is_larger_root(fp_from_mont(y)) != y_flag) y = fp_neg(y);
  if (!UNCHECKED && !g1_in_subgroup(x, y)) {
    atomic_add_u32(bad, 1u);
    store_affine(out, stride, i, x, x, true);
    return;
  }
  store_affine(out, stride, i, x, y, false);
}

KERNEL void __launch_bounds__(128) g1_compress_affine_kernel(const unsigned char* in, u32 stride, u32 n, unsigned char* out48) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = in + (size_t)i * stride;
  const uint2* q = reinterpret_cast<const uint2*>(p);
  Fq x, y;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    const uint2 a = q[k], b = q[6 + k];
    x.l[2 * k] = a.x;
    x.l[2 * k + 1] = a.y;
    y.l[2 * k] = b.x;
    y.l[2 * k + 1] = b.y;
  }
  const bool inf = (stride >= 97) ? (p[96] != 0) : (fp_is_zero(x) && fp_is_zero(y));
  Fq xc = fp_zero<FqParams>();
  u32 flags = 0;
  if (inf) {
    flags = 1u << 30;
  } else {
    xc = fp_from_mont(x);
    if (is_larger_root(fp_from_mont(y))) flags = 1u << 31;
  }
  xc.l[FqParams::N - 1] |= flags;
  uint2* o = reinterpret_cast<uint2*>(out48 + (size_t)i * 48);
#pragma unroll
  for (int k = 0; k < 6; k++) o[k] = make_uint2(xc.l[2 * k], xc.l[2 * k + 1]);
}

}  // namespace wire
