// poly_lib.cu -- translation unit of the elementwise field / polynomial kernels (poly.cuh).
#include "internal.h"
#include "poly.cuh"

namespace aleo {
cudaError_t poly_upload_constants() { return aleo_upload_field_constants(); }

template <class P>
static cudaError_t field_op_t(int op, void* out, const void* a, const void* b, size_t n, cudaStream_t s) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride above 16 CTAs per SM
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

cudaError_t field_op(int field, int op, void* out, const void* a, const void* b, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  return field == 0 ? field_op_t<FrParams>(op, out, a, b, n, s) : field_op_t<FqParams>(op, out, a, b, n, s);
}
}  // namespace aleo
