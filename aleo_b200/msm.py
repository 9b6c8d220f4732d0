"""Host-side mirror of snarkVM's ``VariableBase::msm`` for BLS12-377 G1 (snarkvm-algorithms 0.14.5
src/msm/variable_base/mod.rs), backed by the sm_100a Pippenger kernels.

``VariableBase.msm(bases, scalars)`` takes the memory images upstream passes to its own FFI:
``&[G1Affine]`` (104-byte stride; 96-byte packed also accepted) and ``&[BigInteger256]`` (canonical
little-endian).  Like upstream the two slices are zipped (the shorter length wins) and the result is
a ``G1Projective`` (144-byte Jacobian image; returned normalised, z = 1 or (0, 1, 0)).
No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C

from . import _lib

AFFINE_STRIDE_RUST = 104
AFFINE_STRIDE_PACKED = 96
PROJECTIVE_BYTES = 144


def _host_ptr(buf):
    """(void*, nbytes, keepalive) of a host buffer without copying: bytes, bytearray, numpy array
    or torch CPU tensor (pinned or pageable)."""
    if hasattr(buf, "data_ptr") and hasattr(buf, "element_size"):      # torch CPU tensor
        if buf.is_cuda:
            raise ValueError("host entry point got a CUDA tensor; use msm_dev")
        return C.c_void_p(buf.data_ptr()), buf.numel() * buf.element_size(), buf
    if hasattr(buf, "__array_interface__"):                              # numpy
        return C.c_void_p(buf.ctypes.data), buf.nbytes, buf
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p), len(buf), buf
    if isinstance(buf, bytearray):
        arr = (C.c_char * len(buf)).from_buffer(buf)
        return C.cast(arr, C.c_void_p), len(buf), arr
    raise TypeError("unsupported host buffer type %r" % type(buf))


class VariableBase:
    @staticmethod
    def msm(bases, scalars, affine_stride: int = AFFINE_STRIDE_RUST) -> bytes:
        lib = _lib.get_lib()
        bp, bbytes, _keep_b = _host_ptr(bases)
        sp, sbytes, _keep_s = _host_ptr(scalars)
        n = min(bbytes // affine_stride, sbytes // 32)
        out = C.create_string_buffer(PROJECTIVE_BYTES)
        lib.check(lib.msm_g1(C.cast(out, C.c_void_p), bp, n, sp, affine_stride), "aleo_b200_msm_g1")
        return out.raw

    @staticmethod
    def msm_multi(bases, scalars, n_devices: int, affine_stride: int = AFFINE_STRIDE_RUST) -> bytes:
        """the same MSM spread over the first n_devices GPUs from this one process (aleo_b200_msm_g1_multi)"""
        lib = _lib.get_lib()
        bp, bbytes, _keep_b = _host_ptr(bases)
        sp, sbytes, _keep_s = _host_ptr(scalars)
        n = min(bbytes // affine_stride, sbytes // 32)
        out = C.create_string_buffer(PROJECTIVE_BYTES)
        lib.check(lib.msm_g1_multi(C.cast(out, C.c_void_p), bp, n, sp, affine_stride, n_devices), "aleo_b200_msm_g1_multi")
        return out.raw

    @staticmethod
    def msm_dev(bases_t, scalars_t, n: int, affine_stride: int = AFFINE_STRIDE_RUST, out=None):
        """device-resident operands (torch CUDA tensors); asynchronous on torch's current stream.
        Returns a uint8 CUDA tensor of 144 bytes."""
        import torch

        lib = _lib.get_lib()
        if out is None:
            out = torch.empty(PROJECTIVE_BYTES, dtype=torch.uint8, device=bases_t.device)
        with torch.cuda.device(bases_t.device):
            stream = torch.cuda.current_stream().cuda_stream
            lib.check(lib.msm_g1_dev(out.data_ptr(), bases_t.data_ptr(), n, scalars_t.data_ptr(), affine_stride, stream),
                      "aleo_b200_msm_g1_dev")
        return out

    @staticmethod
    def sum_partials_dev(partials_t, count: int, out=None):
        """sum of `count` Jacobian partial results (count x 144 B on device) -> normalised Jacobian"""
        import torch

        lib = _lib.get_lib()
        if out is None:
            out = torch.empty(PROJECTIVE_BYTES, dtype=torch.uint8, device=partials_t.device)
        with torch.cuda.device(partials_t.device):
            stream = torch.cuda.current_stream().cuda_stream
            lib.check(lib.g1_sum_dev(out.data_ptr(), partials_t.data_ptr(), count, stream), "aleo_b200_g1_sum_dev")
        return out

    @staticmethod
    def window_bits(n: int) -> int:
        return _lib.get_lib().msm_window_bits(n)

    @staticmethod
    def batch_affine_levels(n: int) -> int:
        """pair-tree levels of batch-affine additions in front of the XYZZ accumulation (0 below 2^22 points)"""
        return _lib.get_lib().msm_ba_levels(n)

    @staticmethod
    def launches(n: int) -> int:
        return _lib.get_lib().msm_launches(n)

    @staticmethod
    def host_plan(n: int) -> dict:
        """how the host-pointer entry point cuts n points into overlapped point ranges, and its window size"""
        lib = _lib.get_lib()
        k, c = C.c_int(), C.c_int()
        lib.check(lib.msm_host_plan(n, C.byref(k), C.byref(c)), "aleo_b200_msm_host_plan")
        return {"ranges": k.value, "window_bits": c.value}


# ---- synthetic workloads and on-device checks ---------------------------------
def gen_bases_dev(n: int, s0: int, d: int, first_index: int = 0, affine_stride: int = AFFINE_STRIDE_RUST, device=None):
    """bases[i] = (s0 + (first_index + i) d) G as a uint8 CUDA tensor of n * stride bytes"""
    import torch

    lib = _lib.get_lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = torch.empty(max(n, 1) * affine_stride, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        lib.check(lib.gen_bases_dev(t.data_ptr(), n, affine_stride, int(s0).to_bytes(32, "little"),
                                    int(d).to_bytes(32, "little"), first_index, stream), "aleo_b200_gen_bases_dev")
    return t


def gen_scalars_dev(n: int, seed: int, first_index: int = 0, montgomery: bool = False, device=None):
    import torch

    lib = _lib.get_lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = torch.empty((max(n, 1), 4), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        lib.check(lib.gen_scalars_dev(t.data_ptr(), n, seed, first_index, 1 if montgomery else 0, stream),
                  "aleo_b200_gen_scalars_dev")
    return t


def dlog_dot_dev(scalars_t, n: int, s0: int, d: int, first_index: int = 0) -> int:
    """sum_i scalars[i] * (s0 + (first_index + i) d) mod r, computed on the device, returned as int"""
    import torch

    lib = _lib.get_lib()
    out = torch.empty(4, dtype=torch.int64, device=scalars_t.device)
    with torch.cuda.device(scalars_t.device):
        stream = torch.cuda.current_stream().cuda_stream
        lib.check(lib.dlog_dot_dev(out.data_ptr(), scalars_t.data_ptr(), n, int(s0).to_bytes(32, "little"),
                                   int(d).to_bytes(32, "little"), first_index, stream), "aleo_b200_dlog_dot_dev")
    return int.from_bytes(out.cpu().numpy().tobytes(), "little")


def check_on_curve_dev(bases_t, n: int, affine_stride: int = AFFINE_STRIDE_RUST) -> bool:
    import torch

    lib = _lib.get_lib()
    with torch.cuda.device(bases_t.device):
        stream = torch.cuda.current_stream().cuda_stream
        return lib.check(lib.check_on_curve_dev(bases_t.data_ptr(), n, affine_stride, stream), "check_on_curve") == 1
