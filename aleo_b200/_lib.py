"""ctypes binding of libaleo_b200.so (include/aleo_b200.h).

The product library is the in-tree ``aleo_b200/libaleo_b200.so`` built by ``make`` (or
``__graft_entry__.build()``).  There is NO fallback: if the shared library is missing, or no
sm_100 device is present, every compute call raises ``AleoB200Error``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "libaleo_b200.so")

OK, EINVAL, ETOOLARGE, ENODEVICE, ECUDA, ENOMEM = 0, -1, -2, -3, -4, -5
NTT_FORWARD, NTT_INVERSE = 0, 1
NTT_STANDARD, NTT_COSET = 0, 1
NTT_ORDER_II, NTT_ORDER_IO, NTT_ORDER_OI = 0, 1, 2

# every symbol include/aleo_b200.h declares: (name, restype, argtypes)
_vp, _sz, _u32, _u64, _int = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
SYMBOLS = [
    ("aleo_b200_version", C.c_char_p, []),
    ("aleo_b200_strerror", C.c_char_p, [_int]),
    ("aleo_b200_last_cuda_error", C.c_char_p, []),
    ("aleo_b200_init", _int, [_int]),
    ("aleo_b200_device_count", _int, []),
    ("aleo_b200_shutdown", _int, []),
    ("aleo_b200_ntt_fr", _int, [_vp, _u32, _int, _int]),
    ("aleo_b200_ntt_fr_dev", _int, [_vp, _u32, _sz, _int, _int, _vp]),
    ("aleo_b200_ntt_fr_ordered_dev", _int, [_vp, _u32, _sz, _int, _int, _int, _vp]),
    ("aleo_b200_ntt_fr_dev_profile", _int, [_vp, _u32, _int, _int, _vp, C.POINTER(C.c_float)]),
    ("aleo_b200_ntt_twiddle_dev", _int, [_vp, _u32, _int, _u32, _u32, _u32, _u32, _vp]),
    ("aleo_b200_ntt_launches", _int, [_u32]),
    ("aleo_b200_ntt_dist_layout", _int, [_u32, _int, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_int)]),
    ("aleo_b200_ntt_dist_create", _int, [C.POINTER(_vp), _u32, _int, _int]),
    ("aleo_b200_ntt_dist_handles", _int, [_vp, _vp]),
    ("aleo_b200_ntt_dist_open", _int, [_vp, _vp]),
    ("aleo_b200_ntt_dist_stage1", _int, [_vp, _vp, _int, _int, _vp]),
    ("aleo_b200_ntt_fr_ordered", _int, [_vp, _u32, _int, _int, _int]),
    ("aleo_b200_polymul", _int, [_vp, _sz, C.POINTER(_vp), C.POINTER(_sz), _sz, C.POINTER(_vp), C.POINTER(_sz), _u32]),
    ("aleo_b200_polymul_dev", _int, [_vp, _sz, C.POINTER(_vp), C.POINTER(_sz), _sz, C.POINTER(_vp), C.POINTER(_sz), _u32, _vp]),
    ("aleo_b200_kzg_open_combinations_dev", _int, [_vp, _vp, C.POINTER(_vp), C.POINTER(_sz), _sz, _vp, _vp, _sz, _vp]),
    ("aleo_b200_ntt_dist_stage2", _int, [_vp, _vp, _int, _int, _vp]),
    ("aleo_b200_ntt_dist_transform", _int, [_vp, _vp, _vp, _int, _int, _vp]),
    ("aleo_b200_ntt_dist_profile", _int, [_vp, _vp, _vp, _int, _int, _vp, C.POINTER(C.c_float)]),
    ("aleo_b200_ntt_dist_destroy", _int, [_vp]),
    ("aleo_b200_msm_g1", _int, [_vp, _vp, _sz, _vp, _sz]),
    ("aleo_b200_msm_g1_multi", _int, [_vp, _vp, _sz, _vp, _sz, _int]),
    ("aleo_b200_msm_g1_dev", _int, [_vp, _vp, _sz, _vp, _sz, _vp]),
    ("aleo_b200_msm_g1_dev_profile", _int, [_vp, _vp, _sz, _vp, _sz, _vp, C.POINTER(C.c_float)]),
    ("aleo_b200_g1_sum_dev", _int, [_vp, _vp, _sz, _vp]),
    ("aleo_b200_srs_create", _int, [C.POINTER(_vp), _vp, _sz, _sz]),
    ("aleo_b200_srs_create_dev", _int, [C.POINTER(_vp), _vp, _sz, _sz, _vp]),
    ("aleo_b200_srs_destroy", _int, [_vp]),
    ("aleo_b200_srs_info", _int, [_vp, C.POINTER(_sz), C.POINTER(_int), C.POINTER(_int), C.POINTER(_sz)]),
    ("aleo_b200_srs_msm", _int, [_vp, _vp, _vp, _sz]),
    ("aleo_b200_srs_msm_dev", _int, [_vp, _vp, _vp, _sz, _vp]),
    ("aleo_b200_srs_msm_dev_profile", _int, [_vp, _vp, _vp, _sz, _vp, C.POINTER(C.c_float)]),
    ("aleo_b200_srs_msm_launches", _int, [_vp, _sz]),
    ("aleo_b200_kzg_commit", _int, [_vp, _vp, _vp, _sz]),
    ("aleo_b200_kzg_commit_dev", _int, [_vp, _vp, _vp, _sz, _vp]),
    ("aleo_b200_kzg_commit_hiding_dev", _int, [_vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    ("aleo_b200_kzg_commit_batch_dev", _int, [_vp, _vp, C.POINTER(_vp), C.POINTER(_sz), _sz, _vp]),
    ("aleo_b200_field_op_dev", _int, [_int, _int, _vp, _vp, _vp, _sz, _vp]),
    ("aleo_b200_fr_lagrange_coeffs_dev", _int, [_vp, _u32, _vp, _vp]),
    ("aleo_b200_fr_axpy_dev", _int, [_vp, _vp, _vp, _sz, _vp]),
    ("aleo_b200_fr_distribute_powers_dev", _int, [_vp, _sz, _vp, _vp, _vp]),
    ("aleo_b200_fr_divide_by_vanishing_on_coset_dev", _int, [_vp, _u32, _u32, _vp, _vp]),
    ("aleo_b200_fr_poly_eval_dev", _int, [_vp, _vp, _sz, _vp, _vp]),
    ("aleo_b200_fr_divide_by_linear_dev", _int, [_vp, _vp, _sz, _vp, _vp]),
    ("aleo_b200_kzg_open_dev", _int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    ("aleo_b200_g1_decompress_dev", _int, [_vp, _sz, _vp, _sz, _vp]),
    ("aleo_b200_g1_decompress_unchecked_dev", _int, [_vp, _sz, _vp, _sz, _vp]),
    ("aleo_b200_g1_compress_dev", _int, [_vp, _vp, _sz, _sz, _vp]),
    ("aleo_b200_msm_window_bits", _int, [_sz]),
    ("aleo_b200_msm_ba_levels", _int, [_sz]),
    ("aleo_b200_msm_launches", _int, [_sz]),
    ("aleo_b200_msm_host_plan", _int, [_sz, C.POINTER(_int), C.POINTER(_int)]),
    ("aleo_b200_gen_bases_dev", _int, [_vp, _sz, _sz, _vp, _vp, _u64, _vp]),
    ("aleo_b200_gen_scalars_dev", _int, [_vp, _sz, _u64, _u64, _int, _vp]),
    ("aleo_b200_dlog_dot_dev", _int, [_vp, _vp, _sz, _vp, _vp, _u64, _vp]),
    ("aleo_b200_check_on_curve_dev", _int, [_vp, _sz, _sz, _vp]),
    ("aleo_b200_fq_mul_fp64_dev", _int, [_vp, _vp, _vp, _sz, _int, _vp]),
    ("aleo_b200_bench_imad", _int, [_int, _int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]


class AleoB200Error(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__("%s failed: %s%s" % (what, code, (" (" + detail + ")") if detail else ""))


class Lib:
    """A loaded libaleo_b200 (or, in tests/emu only, the development emulator build)."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise AleoB200Error(ENODEVICE, "load",
                                "%s is missing: run `make` (there is no CPU fallback)" % path)
        self.path = path
        self.dll = C.CDLL(path)
        for name, res, args in SYMBOLS:
            fn = getattr(self.dll, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("aleo_b200_"):], fn)

    def check(self, rc: int, what: str) -> int:
        if rc < 0:
            detail = self.strerror(rc).decode()
            cuda = self.last_cuda_error().decode()
            raise AleoB200Error(rc, what, detail + ((": " + cuda) if cuda else ""))
        return rc


_product = None


def get_lib() -> Lib:
    """The product library.  Raises (loudly) when it has not been built."""
    global _product
    if _product is None:
        _product = Lib(PRODUCT_LIB)
    return _product
