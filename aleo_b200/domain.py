"""Host-side mirror of snarkVM's ``EvaluationDomain`` for Fr (snarkvm-algorithms 0.14.5
src/fft/domain.rs), backed by the sm_100a NTT kernels through the
C ABI (include/aleo_b200.h).  Same names, argument meaning and error behaviour as upstream:

* ``EvaluationDomain.new(num_coeffs)`` -> domain with ``size = next_pow2(num_coeffs)`` or ``None``
  when ``log2(size)`` exceeds the field's two-adicity (47);
* ``fft_in_place(v)`` first resizes ``v`` to the domain size with zeros (a longer input is truncated,
  as ``Vec::resize`` does upstream);
* all vectors are images of ``Vec<Fp256<FrParameters>>``: n x 32 bytes, Montgomery form.

Host vectors are ``bytearray`` / ``bytes`` / numpy uint64 ``(n, 4)`` arrays; device vectors are torch
uint64/int64 ``(n, 4)`` CUDA tensors (``*_dev`` methods, asynchronous on torch's current stream).
No CPU fallback exists: without the CUDA library every transform raises ``AleoB200Error``.
"""
from __future__ import annotations

import ctypes as C

from . import _lib

FR_TWO_ADICITY = 47
MAX_DEVICE_LOG_N = 32   # aleo_b200_ntt_fr returns ETOOLARGE above this (memory bound)

# BLS12-377 scalar field (derived in tools/gen_constants.py / oracle; repeated here for the
# domain's public fields, which upstream exposes)
_R_MOD = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
_GENERATOR = 22
_TWO_ADIC_ROOT = pow(_GENERATOR, (_R_MOD - 1) >> FR_TWO_ADICITY, _R_MOD)


def _as_bytearray(v) -> bytearray:
    if isinstance(v, bytearray):
        return v
    if isinstance(v, (bytes, memoryview)):
        return bytearray(v)
    return bytearray(memoryview(v).cast("B"))      # numpy arrays and friends


class EvaluationDomain:
    """``EvaluationDomain<Fr>``."""

    def __init__(self, size: int, log_size: int):
        self.size = size
        self.log_size_of_group = log_size
        self.size_as_field_element = size % _R_MOD
        self.size_inv = pow(size, -1, _R_MOD)
        self.group_gen = pow(_TWO_ADIC_ROOT, 1 << (FR_TWO_ADICITY - log_size), _R_MOD)
        self.group_gen_inv = pow(self.group_gen, -1, _R_MOD)
        self.generator_inv = pow(_GENERATOR, -1, _R_MOD)

    # -- construction -----------------------------------------------------------------------------
    @staticmethod
    def compute_size_of_domain(num_coeffs: int):
        size = 1
        while size < num_coeffs:
            size <<= 1
        return size if size.bit_length() - 1 <= FR_TWO_ADICITY else None

    @classmethod
    def new(cls, num_coeffs: int):
        size = cls.compute_size_of_domain(num_coeffs)
        if size is None:
            return None
        return cls(size, size.bit_length() - 1)

    # -- host vectors -------------------------------------------------------------------------------
    def _resize(self, v) -> bytearray:
        buf = _as_bytearray(v)
        want = self.size * 32
        if len(buf) % 32:
            raise ValueError("vector length is not a multiple of 32 bytes")
        if len(buf) < want:
            buf.extend(bytes(want - len(buf)))
        elif len(buf) > want:
            del buf[want:]
        return buf

    def _run_host(self, v, direction: int, kind: int) -> bytearray:
        lib = _lib.get_lib()
        buf = self._resize(v)
        cbuf = (C.c_char * len(buf)).from_buffer(buf)
        lib.check(lib.ntt_fr(C.cast(cbuf, C.c_void_p), self.log_size_of_group, direction, kind), "aleo_b200_ntt_fr")
        return buf

    def fft_in_place(self, v: bytearray) -> bytearray:
        return self._run_host(v, _lib.NTT_FORWARD, _lib.NTT_STANDARD)

    def ifft_in_place(self, v: bytearray) -> bytearray:
        return self._run_host(v, _lib.NTT_INVERSE, _lib.NTT_STANDARD)

    def coset_fft_in_place(self, v: bytearray) -> bytearray:
        return self._run_host(v, _lib.NTT_FORWARD, _lib.NTT_COSET)

    def coset_ifft_in_place(self, v: bytearray) -> bytearray:
        return self._run_host(v, _lib.NTT_INVERSE, _lib.NTT_COSET)

    def fft(self, coeffs) -> bytes:
        return bytes(self.fft_in_place(bytearray(_as_bytearray(coeffs))))

    def ifft(self, evals) -> bytes:
        return bytes(self.ifft_in_place(bytearray(_as_bytearray(evals))))

    def coset_fft(self, coeffs) -> bytes:
        return bytes(self.coset_fft_in_place(bytearray(_as_bytearray(coeffs))))

    def coset_ifft(self, evals) -> bytes:
        return bytes(self.coset_ifft_in_place(bytearray(_as_bytearray(evals))))

    # -- device vectors (torch) ---------------------------------------------------------------------
    def _run_dev(self, t, direction: int, kind: int, batch: int = 1):
        import torch

        lib = _lib.get_lib()
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("expected a contiguous CUDA tensor")
        if t.numel() * t.element_size() != batch * self.size * 32:
            raise ValueError("tensor holds %d bytes, domain needs %d" % (t.numel() * t.element_size(), batch * self.size * 32))
        with torch.cuda.device(t.device):
            stream = torch.cuda.current_stream().cuda_stream
            lib.check(lib.ntt_fr_dev(t.data_ptr(), self.log_size_of_group, batch, direction, kind, stream),
                      "aleo_b200_ntt_fr_dev")
        return t

    def fft_in_place_dev(self, t, batch: int = 1):
        return self._run_dev(t, _lib.NTT_FORWARD, _lib.NTT_STANDARD, batch)

    def ifft_in_place_dev(self, t, batch: int = 1):
        return self._run_dev(t, _lib.NTT_INVERSE, _lib.NTT_STANDARD, batch)

    def coset_fft_in_place_dev(self, t, batch: int = 1):
        return self._run_dev(t, _lib.NTT_FORWARD, _lib.NTT_COSET, batch)

    def coset_ifft_in_place_dev(self, t, batch: int = 1):
        return self._run_dev(t, _lib.NTT_INVERSE, _lib.NTT_COSET, batch)

    # -- upstream's FFTOrder variants (fft_helper_in_place_with_pc & co.; the precomputation argument is accepted
    #    and ignored: twiddle tables are cached per size inside the library) ---------------------------------------
    def _run_dev_ordered(self, t, direction: int, kind: int, order: int, batch: int = 1):
        import torch

        lib = _lib.get_lib()
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("expected a contiguous CUDA tensor")
        if t.numel() * t.element_size() != batch * self.size * 32:
            raise ValueError("tensor holds %d bytes, domain needs %d" % (t.numel() * t.element_size(), batch * self.size * 32))
        with torch.cuda.device(t.device):
            stream = torch.cuda.current_stream().cuda_stream
            lib.check(lib.ntt_fr_ordered_dev(t.data_ptr(), self.log_size_of_group, batch, direction, kind, order, stream),
                      "aleo_b200_ntt_fr_ordered_dev")
        return t

    def fft_helper_in_place_with_pc_dev(self, t, order: int, pc=None, batch: int = 1):
        return self._run_dev_ordered(t, _lib.NTT_FORWARD, _lib.NTT_STANDARD, order, batch)

    def ifft_helper_in_place_with_pc_dev(self, t, order: int, pc=None, batch: int = 1):
        return self._run_dev_ordered(t, _lib.NTT_INVERSE, _lib.NTT_STANDARD, order, batch)

    def out_order_fft_in_place_with_pc_dev(self, t, pc=None, batch: int = 1):
        return self._run_dev_ordered(t, _lib.NTT_FORWARD, _lib.NTT_STANDARD, _lib.NTT_ORDER_IO, batch)

    def in_order_ifft_in_place_with_pc_dev(self, t, pc=None, batch: int = 1):
        return self._run_dev_ordered(t, _lib.NTT_INVERSE, _lib.NTT_STANDARD, _lib.NTT_ORDER_II, batch)

    def in_order_coset_ifft_in_place_with_pc_dev(self, t, pc=None, batch: int = 1):
        return self._run_dev_ordered(t, _lib.NTT_INVERSE, _lib.NTT_COSET, _lib.NTT_ORDER_II, batch)

    def ntt_host_buffer(self, buf, direction: int, kind: int, order: int = _lib.NTT_ORDER_II):
        """in-place transform of a host buffer that already holds `size` elements (numpy array or
        torch CPU tensor, pinned or pageable) -- the zero-copy form of aleo_b200_ntt_fr / aleo_b200_ntt_fr_ordered"""
        from .msm import _host_ptr

        lib = _lib.get_lib()
        ptr, nbytes, _keep = _host_ptr(buf)
        if nbytes != self.size * 32:
            raise ValueError("buffer holds %d bytes, domain needs %d" % (nbytes, self.size * 32))
        if order == _lib.NTT_ORDER_II:
            lib.check(lib.ntt_fr(ptr, self.log_size_of_group, direction, kind), "aleo_b200_ntt_fr")
        else:
            lib.check(lib.ntt_fr_ordered(ptr, self.log_size_of_group, direction, kind, order), "aleo_b200_ntt_fr_ordered")
        return buf

    # host-vector forms of upstream's FFTOrder entry points (the precomputation argument is accepted and ignored)
    def _run_host_ordered(self, v, direction: int, kind: int, order: int) -> bytearray:
        lib = _lib.get_lib()
        buf = self._resize(v)
        cbuf = (C.c_char * len(buf)).from_buffer(buf)
        lib.check(lib.ntt_fr_ordered(C.cast(cbuf, C.c_void_p), self.log_size_of_group, direction, kind, order),
                  "aleo_b200_ntt_fr_ordered")
        return buf

    def fft_helper_in_place_with_pc(self, v: bytearray, order: int, pc=None) -> bytearray:
        return self._run_host_ordered(v, _lib.NTT_FORWARD, _lib.NTT_STANDARD, order)

    def ifft_helper_in_place_with_pc(self, v: bytearray, order: int, pc=None) -> bytearray:
        return self._run_host_ordered(v, _lib.NTT_INVERSE, _lib.NTT_STANDARD, order)

    def out_order_fft_in_place_with_pc(self, v: bytearray, pc=None) -> bytearray:
        return self._run_host_ordered(v, _lib.NTT_FORWARD, _lib.NTT_STANDARD, _lib.NTT_ORDER_IO)

    def launches(self) -> int:
        """kernel launches one transform of this domain issues"""
        return _lib.get_lib().ntt_launches(self.log_size_of_group)
