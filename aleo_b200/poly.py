"""Elementwise field arithmetic on device vectors (include/aleo_b200.h ``aleo_b200_field_op_dev``): the pointwise
work snarkVM's prover does on evaluation vectors between its FFTs (snarkvm-algorithms 0.14.5 src/fft/evaluations.rs
``Evaluations`` Mul / Sub / Add; snarkvm-fields ``batch_inversion``).  Elements are
Montgomery-form memory images (Fr: 32 bytes, Fq: 48 bytes) in torch CUDA tensors; no CPU fallback exists."""
from __future__ import annotations

from . import _lib

FR, FQ = 0, 1
ADD, SUB, MUL, SQR, INV, NEG = 0, 1, 2, 3, 4, 5
_ELEM_BYTES = {FR: 32, FQ: 48}


def field_op_dev(field: int, op: int, a, b=None, out=None):
    """out[i] = a[i] op b[i]; a, b, out: CUDA tensors holding n Montgomery-form elements (any dtype; sized in
    bytes).  In-place (out is a) is allowed.  Asynchronous on torch's current stream."""
    import torch

    lib = _lib.get_lib()
    nbytes = a.numel() * a.element_size()
    n = nbytes // _ELEM_BYTES[field]
    if out is None:
        out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        stream = torch.cuda.current_stream().cuda_stream
        lib.check(lib.field_op_dev(field, op, out.data_ptr(), a.data_ptr(), b.data_ptr() if b is not None else None, n,
                                   stream), "aleo_b200_field_op_dev")
    return out


class Evaluations:
    """mirror of the operator set of snarkVM's ``Evaluations<F>`` over Fr vectors resident on the device"""

    @staticmethod
    def mul(a, b, out=None):
        return field_op_dev(FR, MUL, a, b, out)

    @staticmethod
    def add(a, b, out=None):
        return field_op_dev(FR, ADD, a, b, out)

    @staticmethod
    def sub(a, b, out=None):
        return field_op_dev(FR, SUB, a, b, out)

    @staticmethod
    def square(a, out=None):
        return field_op_dev(FR, SQR, a, None, out)

    @staticmethod
    def batch_inversion(a, out=None):
        """elementwise inverse, zeros stay zero (snarkvm_fields::batch_inversion semantics)"""
        return field_op_dev(FR, INV, a, None, out)

    @staticmethod
    def neg(a, out=None):
        return field_op_dev(FR, NEG, a, None, out)
