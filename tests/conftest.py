import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _make(target, out):
    if not os.path.exists(out):
        subprocess.run(["make", "-j8", target], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    return out


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def c_oracle():
    """the C restatement (oracle/oracle.c) as a ctypes library -- checker only"""
    import ctypes as C

    path = _make("oracle", os.path.join(ROOT, "oracle", "_build", "liboracle.so"))
    lib = C.CDLL(path)
    lib.oracle_ntt_fr.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    lib.oracle_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    lib.oracle_gen_bases.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.oracle_msm_window_bits.argtypes = [C.c_size_t]
    return lib


@pytest.fixture(scope="session")
def emu_lib():
    """development emulator build of the kernels (tests/emu): kernel-logic checks without a GPU"""
    from aleo_b200._lib import Lib

    return Lib(_make("emu", os.path.join(ROOT, "build", "libaleo_b200_emu.so")))


@pytest.fixture(scope="session")
def product_lib_path():
    return _make("all", os.path.join(ROOT, "aleo_b200", "libaleo_b200.so"))
