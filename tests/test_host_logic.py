"""Host-side mirror of the reference interface (EvaluationDomain / VariableBase): the parts that run
without a device -- domain construction, resize semantics, buffer handling."""
import pytest

import aleo_b200
from aleo_b200.domain import EvaluationDomain
from oracle import bls12_377 as o


def test_domain_new_matches_upstream_rules():
    assert EvaluationDomain.new(1).size == 1
    assert EvaluationDomain.new(2).size == 2
    assert EvaluationDomain.new(3).size == 4
    assert EvaluationDomain.new(1 << 20).log_size_of_group == 20
    assert EvaluationDomain.new((1 << 20) + 1).log_size_of_group == 21
    assert EvaluationDomain.new(1 << 47).log_size_of_group == 47
    assert EvaluationDomain.new((1 << 47) + 1) is None          # above the field's two-adicity
    assert EvaluationDomain.compute_size_of_domain(0) == 1


def test_domain_fields_match_oracle():
    for k in (0, 1, 8, 16, 26):
        d = EvaluationDomain.new(1 << k)
        assert d.group_gen == o.fr_root_of_unity(k)
        assert d.group_gen * d.group_gen_inv % o.R_MOD == 1
        assert d.size_inv * d.size % o.R_MOD == 1
        assert d.generator_inv * o.FR_GENERATOR % o.R_MOD == 1
        assert pow(d.group_gen, d.size, o.R_MOD) == 1


def test_resize_pads_with_zero_and_truncates():
    d = EvaluationDomain.new(8)
    v = bytearray(o.fr_vec_to_bytes(o.random_fr_vec(5, 1)))
    out = d._resize(v)
    assert len(out) == 8 * 32 and bytes(out[5 * 32:]) == bytes(3 * 32)
    v = bytearray(o.fr_vec_to_bytes(o.random_fr_vec(11, 1)))
    assert len(d._resize(v)) == 8 * 32
    with pytest.raises(ValueError):
        d._resize(bytearray(33))


def test_host_pointer_helper_accepts_numpy_and_torch():
    import numpy as np
    import torch

    from aleo_b200.msm import _host_ptr

    a = np.zeros((4, 4), dtype=np.uint64)
    p, n, _ = _host_ptr(a)
    assert n == 128 and p.value == a.ctypes.data
    t = torch.zeros(13, dtype=torch.int64)
    p, n, _ = _host_ptr(t)
    assert n == 104 and p.value == t.data_ptr()
    assert _host_ptr(b"abcd")[1] == 4 and _host_ptr(bytearray(7))[1] == 7
    with pytest.raises(TypeError):
        _host_ptr([1, 2, 3])


def test_package_surface():
    for name in ("EvaluationDomain", "VariableBase", "AleoB200Error", "gen_bases_dev", "gen_scalars_dev"):
        assert hasattr(aleo_b200, name)


def test_glv_constants_in_the_generated_header():
    """u^2 and floor(2^384 / u^2) of the GLV split (csrc/msm.cuh glv_split_kernel) against big integers; r = u^4 - u^2 + 1"""
    import os
    import re
    from oracle import bls12_377 as o
    src = open(os.path.join(os.path.dirname(__file__), "..", "aleo_b200", "csrc", "bls12_377_constants.cuh")).read()
    body = src[src.index("struct GlvParams"):]

    def limbs(name):
        m = re.search(name + r"\(int i\) \{\s*constexpr u32 t\[\d+\] = \{([^}]*)\}", body)
        return sum(int(x.strip().rstrip("u"), 16) << (32 * i) for i, x in enumerate(m.group(1).split(",")))

    u = 0x8508c00000000001
    assert u ** 4 - u ** 2 + 1 == o.R_MOD
    assert limbs("U2") == u * u and limbs("BARRETT") == (1 << 384) // (u * u)
    # the endomorphism's eigenvalue on G1 is -u^2: lambda^2 + lambda + 1 = 0 mod r
    lam = (-u * u) % o.R_MOD
    assert (lam * lam + lam + 1) % o.R_MOD == 0
