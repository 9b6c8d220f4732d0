"""The C restatement (oracle/oracle.c: the large-size checker and the timed CPU baseline) against the
Python big-integer oracle and the golden vectors."""
import ctypes as C
import json
import os
import random

import pytest

from oracle import bls12_377 as o


def _limbs(v, n):
    return (C.c_uint64 * n)(*[(v >> (64 * i)) & (2**64 - 1) for i in range(n)])


def test_field_core(c_oracle):
    rng = random.Random(5)
    for mod, n, fn, R in ((o.R_MOD, 4, c_oracle.oracle_fr_mul, 1 << 256), (o.P_MOD, 6, c_oracle.oracle_fq_mul, 1 << 384)):
        for a, b in [(0, 0), (1, mod - 1), (mod - 1, mod - 1)] + [(rng.randrange(mod), rng.randrange(mod)) for _ in range(300)]:
            out = (C.c_uint64 * n)()
            fn(out, _limbs(a, n), _limbs(b, n))
            assert sum(out[i] << (64 * i) for i in range(n)) == a * b * pow(R, -1, mod) % mod
    a = rng.randrange(1, o.P_MOD)
    out = (C.c_uint64 * 6)()
    c_oracle.oracle_fq_inv(out, _limbs(o.fq_to_mont(a), 6))
    assert o.fq_from_mont(sum(out[i] << (64 * i) for i in range(6))) == pow(a, -1, o.P_MOD)


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 13])
def test_ntt_variants(c_oracle, log_n):
    n = 1 << log_n
    v = o.random_fr_vec(n, 40 + log_n)
    for inv, coset, ref in ((0, 0, o.fft), (1, 0, o.ifft), (0, 1, o.coset_fft), (1, 1, o.coset_ifft)):
        for threads in (1, 3):
            buf = C.create_string_buffer(o.fr_vec_to_bytes(v), n * 32)
            assert c_oracle.oracle_ntt_fr(buf, log_n, inv, coset, threads) == 0
            assert o.fr_vec_from_bytes(buf.raw) == ref(v)


def test_gen_bases_and_msm(c_oracle):
    for n, threads in ((1, 1), (7, 2), (200, 3), (4096, 8)):
        for stride in (104, 96):
            s0, d = o.base_dlogs(n, n + 3)
            bb = C.create_string_buffer(n * stride)
            c_oracle.oracle_gen_bases(bb, n, stride, o.int_to_le_bytes(s0, 32), o.int_to_le_bytes(d, 32), 2, threads)
            if n <= 200:
                pts = [o.g1_mul(o.G1_GEN, (s0 + (2 + i) * d) % o.R_MOD) for i in range(n)]
                assert bb.raw == o.g1_affine_vec_to_bytes(pts, stride)
            s = o.random_fr_vec(n, n + 9)
            out = C.create_string_buffer(144)
            c_oracle.oracle_msm_g1(out, bb, n, o.fr_vec_to_bytes(s, mont=False), stride, threads)
            k = sum(s[i] * (s0 + (2 + i) * d) for i in range(n)) % o.R_MOD
            assert out.raw == o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, k))


def test_golden_msm(c_oracle, golden_dir):
    msm = json.load(open(os.path.join(golden_dir, "msm_golden.json")))
    for name, g in msm.items():
        out = C.create_string_buffer(144)
        c_oracle.oracle_msm_g1(out, bytes.fromhex(g["bases104"]), g["n"], bytes.fromhex(g["scalars"]), 104, 2)
        assert out.raw.hex() == g["result"], name
    out = C.create_string_buffer(144)
    c_oracle.oracle_msm_g1(out, None, 0, None, 104, 1)
    assert out.raw == o.g1_projective_to_bytes(None)
