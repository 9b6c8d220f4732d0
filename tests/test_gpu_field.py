"""Parity of the sm_100a field core (Montgomery add / sub / mul, the dedicated square, the binary-GCD inverse)
with big-integer arithmetic, through aleo_b200_field_op_dev -- the entry point that also serves the prover's
pointwise evaluation-vector arithmetic.  Bit-exact."""
import random

import numpy as np
import pytest
import torch

import aleo_b200 as ab
from aleo_b200 import poly
from oracle import bls12_377 as o

pytestmark = pytest.mark.gpu


def _enc(field, vals):
    if field == poly.FR:
        raw = o.fr_vec_to_bytes(vals)
    else:
        raw = b"".join(o.int_to_le_bytes(o.fq_to_mont(x), 48) for x in vals)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()


def _dec(field, t):
    raw = t.cpu().numpy().tobytes()
    if field == poly.FR:
        return o.fr_vec_from_bytes(raw)
    return [o.fq_from_mont(o.le_bytes_to_int(raw[i:i + 48])) for i in range(0, len(raw), 48)]


def _edge_values(mod, nlimbs):
    R = 1 << (32 * nlimbs)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, R % mod, (R * R) % mod, pow(R, -1, mod)]
    vals += [((1 << k) - 1) % mod for k in (31, 32, 33, 63, 64, 65, 32 * nlimbs - 8)]
    vals += [(1 << k) % mod for k in range(0, mod.bit_length(), 29)]
    vals += [(x * pow(R, -1, mod)) % mod for x in list(vals)]      # Montgomery images with the same limb shapes
    return vals


@pytest.mark.parametrize("field", [poly.FR, poly.FQ])
def test_field_ops_match_big_integers(field):
    mod, nl = (o.R_MOD, 8) if field == poly.FR else (o.P_MOD, 12)
    rng = random.Random(77 + field)
    a = _edge_values(mod, nl) + [rng.randrange(mod) for _ in range(3000)]
    b = list(reversed(_edge_values(mod, nl))) + [rng.randrange(mod) for _ in range(3000)]
    # limbs made of all-ones / all-zero / single-bit words: every carry chain saturates somewhere
    for _ in range(2000):
        limbs = [rng.choice([0, 0xFFFFFFFF, 0xFFFFFFFE, 1, 0x80000000, 0x7FFFFFFF, rng.getrandbits(32)]) for _ in range(nl)]
        a.append(sum(v << (32 * i) for i, v in enumerate(limbs)) % mod)
        b.append(rng.randrange(mod))
    A, B = _enc(field, a), _enc(field, b)
    assert _dec(field, poly.field_op_dev(field, poly.ADD, A, B)) == [(x + y) % mod for x, y in zip(a, b)]
    assert _dec(field, poly.field_op_dev(field, poly.SUB, A, B)) == [(x - y) % mod for x, y in zip(a, b)]
    assert _dec(field, poly.field_op_dev(field, poly.MUL, A, B)) == [(x * y) % mod for x, y in zip(a, b)]
    assert _dec(field, poly.field_op_dev(field, poly.SQR, A)) == [(x * x) % mod for x in a]
    assert _dec(field, poly.field_op_dev(field, poly.INV, A)) == [pow(x, -1, mod) if x else 0 for x in a]
    assert _dec(field, poly.field_op_dev(field, poly.NEG, A)) == [(-x) % mod for x in a]


def test_evaluations_identities_at_scale():
    """2^22 random Fr: (a*b)*a^-1 == b where a != 0, a^2 == a*a, (a+b)-b == a, in place as the prover uses them"""
    n = 1 << 22
    a = ab.gen_scalars_dev(n, 11, 0, True)
    b = ab.gen_scalars_dev(n, 12, 0, True)
    E = ab.Evaluations
    prod = E.mul(a, b)
    back = E.mul(prod, E.batch_inversion(a))
    assert torch.equal(back, b)
    assert torch.equal(E.square(a), E.mul(a, a))
    t = E.add(a, b)
    E.sub(t, b, out=t)
    assert torch.equal(t, a)
    assert torch.equal(E.add(a, E.neg(a)), torch.zeros_like(a))


def test_empty_and_bad_arguments():
    lib = ab.get_lib()
    assert lib.field_op_dev(0, 2, None, None, None, 0, None) == 0
    assert lib.field_op_dev(2, 2, None, None, None, 1, None) == -1
    assert lib.field_op_dev(0, 9, None, None, None, 1, None) == -1
    x = torch.zeros(32, dtype=torch.uint8, device="cuda")
    assert lib.field_op_dev(0, 2, x.data_ptr(), x.data_ptr(), None, 1, None) == -1     # MUL needs b


def test_fp64_pipe_multiplier_is_bit_identical():
    """EXPERIMENT (csrc/mont_fp64.cuh): the Fq product / square on the FP64 pipe (DFMA, 48-bit limbs,
    exact hi / lo halves through fma_rz) against big integers and against the IMAD.WIDE multiplier: edge values whose
    48-bit limbs are all ones / zero / single bits (every hi / lo split and every column carry saturates somewhere) and
    2 * 10^5 random pairs"""
    mod = o.P_MOD
    rng = random.Random(4848)
    a = _edge_values(mod, 12) + [rng.randrange(mod) for _ in range(2000)]
    b = list(reversed(_edge_values(mod, 12))) + [rng.randrange(mod) for _ in range(2000)]
    for _ in range(3000):
        limbs = [rng.choice([0, (1 << 48) - 1, (1 << 48) - 2, 1, 1 << 47, (1 << 47) - 1, rng.getrandbits(48)]) for _ in range(8)]
        # as Montgomery IMAGES: the multiplier sees exactly these limb patterns
        a.append(sum(v << (48 * i) for i, v in enumerate(limbs)) % mod * pow(1 << 384, -1, mod) % mod)
        b.append(rng.randrange(mod))
    A, B = _enc(poly.FQ, a), _enc(poly.FQ, b)
    lib = ab.get_lib()
    out = torch.empty_like(A)
    lib.check(lib.fq_mul_fp64_dev(out.data_ptr(), A.data_ptr(), B.data_ptr(), len(a), 0, None), "fq_mul_fp64")
    assert _dec(poly.FQ, out) == [(x * y) % mod for x, y in zip(a, b)]
    lib.check(lib.fq_mul_fp64_dev(out.data_ptr(), A.data_ptr(), None, len(a), 1, None), "fq_sqr_fp64")
    assert _dec(poly.FQ, out) == [(x * x) % mod for x in a]
    # at scale against the IMAD.WIDE multiplier (raw 48-byte images, any bit pattern below p)
    n = 200000
    raw = np.random.default_rng(5).integers(0, 2**63, size=(n, 6), dtype=np.int64)
    raw[:, 5] &= (1 << 55) - 1                       # < 2^375 < p
    X = torch.from_numpy(raw).cuda()
    Y = torch.from_numpy(np.roll(raw, 1, axis=0).copy()).cuda()
    got = torch.empty_like(X)
    lib.check(lib.fq_mul_fp64_dev(got.data_ptr(), X.data_ptr(), Y.data_ptr(), n, 0, None), "fq_mul_fp64")
    assert torch.equal(got, poly.field_op_dev(poly.FQ, poly.MUL, X, Y))
    lib.check(lib.fq_mul_fp64_dev(got.data_ptr(), X.data_ptr(), None, n, 1, None), "fq_sqr_fp64")
    assert torch.equal(got, poly.field_op_dev(poly.FQ, poly.SQR, X))
