// poly_lib.cu -- translation unit of the elementwise field / polynomial kernels (poly.cuh).
#include "internal.h"
#include "poly.cuh"

namespace aleo {
cudaError_t poly_upload_constants() { return aleo_upload_field_constants(); }

template <class P>
static cudaError_t field_op_t(int op, void* out, const void* a, const void* b, size_t n, cudaStream_t s) {
  size_t blocks = (n + 255) / 256;
  if (blocks > dev_props().sms * 16) blocks = dev_props().sms * 16;  // grid-stride above 16 CTAs per SM
  const dim3 g((u32)blocks), t(256);
  switch (op) {
    case poly::OP_ADD: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_ADD>), g, t, 0, s, out, a, b, n); break;
    case poly::OP_SUB: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_SUB>), g, t, 0, s, out, a, b, n); break;
    case poly::OP_MUL: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_MUL>), g, t, 0, s, out, a, b, n); break;
    case poly::OP_SQR: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_SQR>), g, t, 0, s, out, a, b, n); break;
    case poly::OP_INV: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_INV>), g, t, 0, s, out, a, b, n); break;
    case poly::OP_NEG: LAUNCH_NOSYNC((poly::field_op_kernel<P, poly::OP_NEG>), g, t, 0, s, out, a, b, n); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

#define PL_CK(expr)                       \
  do {                                    \
    cudaError_t _e = (expr);              \
    if (_e != cudaSuccess) return _e;     \
  } while (0)

static Fr host_fr(const void* p) {
  Fr v;
  std::memcpy(v.l, p, 32);
  return v;
}

// the per-call scalars (k, g, g^256, g^-1, g^-256) on the device
static cudaError_t make_setup(Fr** s_out, const void* g32, const void* k32, cudaStream_t s) {
  Fr* d = nullptr;
  PL_CK(aleo::pool_malloc_async((void**)&d, 8 * sizeof(Fr), s));
  Fr k = host_fr(g32);
  if (k32) k = host_fr(k32);
  LAUNCH_NOSYNC(poly::setup_kernel, dim3(1), dim3(1), 0, s, d, host_fr(g32), k, (u32)(k32 ? 1 : 0));
  *s_out = d;
  return cudaGetLastError();
}

static inline u32 pw_grid(size_t n) { return (u32)((n + (size_t)poly::PW_TPB * poly::PW_RUN - 1) / ((size_t)poly::PW_TPB * poly::PW_RUN)); }

// out[i] = L_i(tau) over the domain of size 2^log_n (EvaluationDomain::evaluate_all_lagrange_coefficients)
cudaError_t fr_lagrange_coeffs(void* out_dev, u32 log_n, const void* tau32, cudaStream_t s) {
  const size_t n = (size_t)1 << log_n;
  Fr* sc = nullptr;
  PL_CK(aleo::pool_malloc_async((void**)&sc, 10 * sizeof(Fr), s));
  u32* hit = reinterpret_cast<u32*>(sc + 8);
  cudaError_t e = cudaMemsetAsync(hit, 0, 4, s);
  const Fr tau = host_fr(tau32);
  // s[1] = w, s[2] = w^256 through the common setup (g = w read back from the device scalar), s[0] = (tau^n - 1) / n
  LAUNCH_NOSYNC(poly::domain_root_kernel, dim3(1), dim3(1), 0, s, sc + 7, log_n);
  Fr w_host;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&w_host, sc + 7, sizeof(Fr), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) {
    LAUNCH_NOSYNC(poly::setup_kernel, dim3(1), dim3(1), 0, s, sc, w_host, w_host, (u32)0);
    LAUNCH_NOSYNC(poly::lagrange_const_kernel, dim3(1), dim3(1), 0, s, sc, tau, log_n);
    LAUNCH_NOSYNC(poly::lagrange_kernel, dim3(pw_grid(n)), dim3(poly::PW_TPB), 0, s, (Fr*)out_dev, (u64)n, (const Fr*)sc, tau, hit);
    u32 blocks = (u32)((n + 255) / 256 < dev_props().sms * 8 ? (n + 255) / 256 : dev_props().sms * 8);
    LAUNCH_NOSYNC(poly::lagrange_indicator_kernel, dim3(blocks), dim3(256), 0, s, (Fr*)out_dev, (u64)n, (const u32*)hit);
    e = cudaGetLastError();
  }
  cudaFreeAsync(sc, s);
  return e;
}

cudaError_t fr_axpy(void* y_dev, const void* x_dev, const void* a32, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  size_t blocks = (n + 255) / 256;
  if (blocks > dev_props().sms * 16) blocks = dev_props().sms * 16;
  LAUNCH_NOSYNC(poly::axpy_kernel, dim3((u32)blocks), dim3(256), 0, s, (Fr*)y_dev, (const Fr*)x_dev, host_fr(a32), (u64)n);
  return cudaGetLastError();
}

// evals (on the coset g H_m, m = 2^log_m) /= Z_n(x) = x^n - 1, n = 2^log_n <= m
cudaError_t fr_divide_by_vanishing_on_coset(void* inout_dev, u32 log_m, u32 log_n, const void* g32, cudaStream_t s) {
  const u32 log_k = log_m - log_n;
  const size_t m = (size_t)1 << log_m, k = (size_t)1 << log_k;
  Fr* table = nullptr;
  PL_CK(aleo::pool_malloc_async((void**)&table, k * sizeof(Fr), s));
  size_t tb = (k + 255) / 256;
  if (tb > dev_props().sms * 8) tb = dev_props().sms * 8;
  LAUNCH_NOSYNC(poly::vanishing_table_kernel, dim3((u32)tb), dim3(256), 0, s, table, log_k, log_n, host_fr(g32));
  size_t blocks = (m + 255) / 256;
  if (blocks > dev_props().sms * 16) blocks = dev_props().sms * 16;
  LAUNCH_NOSYNC(poly::mul_periodic_kernel, dim3((u32)blocks), dim3(256), 0, s, (Fr*)inout_dev, (u64)m, (const Fr*)table, (u32)(k - 1));
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(table, s);
  return e;
}

cudaError_t fr_distribute_powers(void* inout_dev, size_t n, const void* g32, const void* k32, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  Fr* sc = nullptr;
  PL_CK(make_setup(&sc, g32, k32, s));
  LAUNCH_NOSYNC(poly::distribute_powers_kernel, dim3(pw_grid(n)), dim3(poly::PW_TPB), 0, s, (Fr*)inout_dev, (u64)n, (const Fr*)sc);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(sc, s);
  return e;
}

cudaError_t fr_poly_eval(void* out_dev, const void* coeffs_dev, size_t n, const void* z32, cudaStream_t s) {
  if (n == 0) return cudaMemsetAsync(out_dev, 0, sizeof(Fr), s);
  Fr* sc = nullptr;
  Fr* partial = nullptr;
  PL_CK(make_setup(&sc, z32, nullptr, s));
  const u32 g = pw_grid(n);
  cudaError_t e = aleo::pool_malloc_async((void**)&partial, (size_t)g * sizeof(Fr), s);
  if (e == cudaSuccess) {
    LAUNCH(poly::eval_partial_kernel, dim3(g), dim3(poly::PW_TPB), 0, s, (const Fr*)coeffs_dev, (u64)n, (const Fr*)sc, partial);
    LAUNCH(poly::sum_partials_kernel, dim3(1), dim3(poly::PW_TPB), 0, s, (const Fr*)partial, g, (Fr*)out_dev);
    e = cudaGetLastError();
    cudaFreeAsync(partial, s);
  }
  cudaFreeAsync(sc, s);
  return e;
}

// quotient_dev: n elements; q[n-1] = 0 (the witness polynomial has one coefficient less)
cudaError_t fr_divide_by_linear(void* quotient_dev, const void* coeffs_dev, size_t n, const void* z32, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  Fr* q = (Fr*)quotient_dev;
  PL_CK(cudaMemsetAsync(q + (n - 1), 0, sizeof(Fr), s));
  if (n == 1) return cudaSuccess;
  const Fr z = host_fr(z32);
  bool z_is_zero = true;
  for (int i = 0; i < 8; i++) z_is_zero = z_is_zero && z.l[i] == 0;
  if (z_is_zero) {
    u32 blocks = (u32)((n + 255) / 256 < dev_props().sms * 8 ? (n + 255) / 256 : dev_props().sms * 8);
    LAUNCH_NOSYNC(poly::shift_down_kernel, dim3(blocks), dim3(256), 0, s, (const Fr*)coeffs_dev, (u64)n, q);
    return cudaGetLastError();
  }
  const u32 nslabs = (u32)((n + poly::SLAB - 1) / poly::SLAB);
  Fr* sc = nullptr;
  Fr* S = nullptr;
  Fr* totals = nullptr;
  PL_CK(make_setup(&sc, z32, nullptr, s));
  cudaError_t e = aleo::pool_malloc_async((void**)&S, n * sizeof(Fr), s);
  if (e == cudaSuccess) e = aleo::pool_malloc_async((void**)&totals, (size_t)2 * nslabs * sizeof(Fr), s);
  if (e == cudaSuccess) {
    Fr* carry = totals + nslabs;
    LAUNCH(poly::suffix_local_kernel, dim3(nslabs), dim3(poly::PW_TPB), 0, s, (const Fr*)coeffs_dev, (u64)n, (const Fr*)sc, S, totals);
    LAUNCH(poly::suffix_carry_kernel, dim3(1), dim3(poly::PW_TPB), 0, s, (const Fr*)totals, nslabs, carry);
    LAUNCH_NOSYNC(poly::witness_fix_kernel, dim3(pw_grid(n)), dim3(poly::PW_TPB), 0, s, (const Fr*)S, (const Fr*)carry, (u64)n,
                  (const Fr*)sc, q);
    e = cudaGetLastError();
  }
  if (totals) cudaFreeAsync(totals, s);
  if (S) cudaFreeAsync(S, s);
  cudaFreeAsync(sc, s);
  return e;
}

// 48-byte compressed G1 -> affine (stride 104 / 96); *bad_out = number of invalid encodings.  Synchronises s.
cudaError_t g1_decompress(const void* in48_dev, size_t n, void* out_affine_dev, u32 stride, cudaStream_t s, u32* bad_out, bool unchecked) {
  *bad_out = 0;
  if (n == 0) return cudaSuccess;
  u32* bad = nullptr;
  PL_CK(aleo::pool_malloc_async((void**)&bad, 4, s));
  cudaMemsetAsync(bad, 0, 4, s);
  if (unchecked)
    LAUNCH_NOSYNC(wire::g1_decompress_kernel<true>, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)in48_dev, (u32)n,
                  (unsigned char*)out_affine_dev, stride, bad);
  else
    LAUNCH_NOSYNC(wire::g1_decompress_kernel<false>, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)in48_dev, (u32)n,
                  (unsigned char*)out_affine_dev, stride, bad);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(bad_out, bad, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(bad, s);
  return e;
}

cudaError_t g1_compress_affine(const void* affine_dev, u32 stride, size_t n, void* out48_dev, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  LAUNCH_NOSYNC(wire::g1_compress_affine_kernel, dim3((u32)((n + 127) / 128)), dim3(128), 0, s, (const unsigned char*)affine_dev,
                stride, (u32)n, (unsigned char*)out48_dev);
  return cudaGetLastError();
}

cudaError_t field_op(int field, int op, void* out, const void* a, const void* b, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  return field == 0 ? field_op_t<FrParams>(op, out, a, b, n, s) : field_op_t<FqParams>(op, out, a, b, n, s);
}
}  // namespace aleo
