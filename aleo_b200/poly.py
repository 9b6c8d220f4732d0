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


def _fr_host(v) -> bytes:
    """32-byte Montgomery image of a canonical integer (or pass 32 bytes through)"""
    if isinstance(v, (bytes, bytearray)):
        assert len(v) == 32
        return bytes(v)
    r = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
    return ((int(v) % r) * (1 << 256) % r).to_bytes(32, "little")


def distribute_powers_dev(t, g, k=None):
    """t[i] *= k * g^i in place (EvaluationDomain::distribute_powers[_and_mul_by_const]); g, k canonical ints"""
    import torch

    lib = _lib.get_lib()
    n = t.numel() * t.element_size() // 32
    with torch.cuda.device(t.device):
        lib.check(lib.fr_distribute_powers_dev(t.data_ptr(), n, _fr_host(g), _fr_host(k) if k is not None else None,
                                               torch.cuda.current_stream().cuda_stream), "aleo_b200_fr_distribute_powers_dev")
    return t


def evaluate_dev(coeffs_t, z) -> int:
    """DensePolynomial::evaluate: sum_i c_i z^i as a canonical integer (synchronises)"""
    import torch

    lib = _lib.get_lib()
    n = coeffs_t.numel() * coeffs_t.element_size() // 32
    out = torch.empty(4, dtype=torch.int64, device=coeffs_t.device)
    with torch.cuda.device(coeffs_t.device):
        lib.check(lib.fr_poly_eval_dev(out.data_ptr(), coeffs_t.data_ptr(), n, _fr_host(z), torch.cuda.current_stream().cuda_stream),
                  "aleo_b200_fr_poly_eval_dev")
    r = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
    return int.from_bytes(out.cpu().numpy().tobytes(), "little") * pow(1 << 256, -1, r) % r


def divide_by_linear_dev(coeffs_t, z, out=None):
    """KZG10::compute_witness_polynomial: (p(x) - p(z)) / (x - z), same length as p with a trailing zero"""
    import torch

    lib = _lib.get_lib()
    n = coeffs_t.numel() * coeffs_t.element_size() // 32
    if out is None:
        out = torch.empty_like(coeffs_t)
    with torch.cuda.device(coeffs_t.device):
        lib.check(lib.fr_divide_by_linear_dev(out.data_ptr(), coeffs_t.data_ptr(), n, _fr_host(z),
                                              torch.cuda.current_stream().cuda_stream), "aleo_b200_fr_divide_by_linear_dev")
    return out


def divide_by_vanishing_on_coset_dev(evals_t, log_n: int, g=22):
    """evals of a polynomial on the coset g * H_m (m = len(evals)) divided in place by Z_n(x) = x^n - 1, n = 2^log_n
    (EvaluationDomain::divide_by_vanishing_poly_on_coset_in_place when m == n); g canonical int, 22 = the coset_fft shift"""
    import torch

    lib = _lib.get_lib()
    m = evals_t.numel() * evals_t.element_size() // 32
    if m == 0 or m & (m - 1):
        raise ValueError("the evaluation vector must have a power-of-two length")
    with torch.cuda.device(evals_t.device):
        lib.check(lib.fr_divide_by_vanishing_on_coset_dev(evals_t.data_ptr(), m.bit_length() - 1, log_n, _fr_host(g),
                                                          torch.cuda.current_stream().cuda_stream),
                  "aleo_b200_fr_divide_by_vanishing_on_coset_dev")
    return evals_t


def axpy_dev(y_t, x_t, a):
    """y += a * x in place (the linear-combination step of SonicKZG10::open_combinations); a canonical int"""
    import torch

    lib = _lib.get_lib()
    n = y_t.numel() * y_t.element_size() // 32
    with torch.cuda.device(y_t.device):
        lib.check(lib.fr_axpy_dev(y_t.data_ptr(), x_t.data_ptr(), _fr_host(a), n, torch.cuda.current_stream().cuda_stream),
                  "aleo_b200_fr_axpy_dev")
    return y_t


def lagrange_coeffs_dev(log_n: int, tau, device=None):
    """EvaluationDomain::evaluate_all_lagrange_coefficients(tau) over the domain of size 2^log_n: (n, 4) int64 CUDA tensor"""
    import torch

    lib = _lib.get_lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((1 << log_n, 4), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        lib.check(lib.fr_lagrange_coeffs_dev(out.data_ptr(), log_n, _fr_host(tau), torch.cuda.current_stream().cuda_stream),
                  "aleo_b200_fr_lagrange_coeffs_dev")
    return out


class PolyMultiplier:
    """mirror of snarkVM's ``PolyMultiplier`` (src/fft/polynomial/multiplier.rs [U]; its CUDA arm calls
    ``snarkvm_polymul``): collect polynomials (coefficient form) and evaluation vectors, then
    ``multiply`` returns the coefficients of their product over the domain of size 2^log_n -- every operand crosses
    PCIe once (host form) or not at all (device form)."""

    def __init__(self):
        self.polynomials, self.evaluations = [], []

    def add_polynomial(self, p):
        self.polynomials.append(p)
        return self

    def add_evaluation(self, e):
        self.evaluations.append(e)
        return self

    def multiply(self, log_n: int) -> bytes:
        """host buffers (bytes / bytearray / numpy / CPU tensors of Montgomery Fr) -> n * 32 bytes"""
        import ctypes as C

        from .msm import _host_ptr
        lib = _lib.get_lib()
        ps = [_host_ptr(p) for p in self.polynomials]
        es = [_host_ptr(e) for e in self.evaluations]
        out = C.create_string_buffer(32 << log_n)
        pp = (C.c_void_p * max(len(ps), 1))(*[p[0] for p in ps])
        pl = (C.c_size_t * max(len(ps), 1))(*[p[1] // 32 for p in ps])
        ep = (C.c_void_p * max(len(es), 1))(*[e[0] for e in es])
        el = (C.c_size_t * max(len(es), 1))(*[e[1] // 32 for e in es])
        lib.check(lib.polymul(C.cast(out, C.c_void_p), len(ps), pp, pl, len(es), ep, el, log_n), "aleo_b200_polymul")
        return out.raw

    def multiply_dev(self, log_n: int, out=None):
        """CUDA tensors of Montgomery Fr -> (n, 4) int64 CUDA tensor; asynchronous on torch's current stream"""
        import ctypes as C

        import torch
        lib = _lib.get_lib()
        ts = self.polynomials + self.evaluations
        dev = ts[0].device
        if out is None:
            out = torch.empty((1 << log_n, 4), dtype=torch.int64, device=dev)
        np_, ne = len(self.polynomials), len(self.evaluations)
        pp = (C.c_void_p * max(np_, 1))(*[t.data_ptr() for t in self.polynomials])
        pl = (C.c_size_t * max(np_, 1))(*[t.numel() * t.element_size() // 32 for t in self.polynomials])
        ep = (C.c_void_p * max(ne, 1))(*[t.data_ptr() for t in self.evaluations])
        el = (C.c_size_t * max(ne, 1))(*[t.numel() * t.element_size() // 32 for t in self.evaluations])
        with torch.cuda.device(dev):
            lib.check(lib.polymul_dev(out.data_ptr(), np_, pp, pl, ne, ep, el, log_n, torch.cuda.current_stream().cuda_stream),
                      "aleo_b200_polymul_dev")
        return out
