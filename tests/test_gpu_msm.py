"""Parity of the sm_100a MSM (through the C ABI / the VariableBase mirror) with the oracle.
Bit-exact: the 144-byte normalised Jacobian image must equal the oracle's."""
import ctypes as C
import json
import os
import threading

import numpy as np
import pytest

import aleo_b200 as ab
from oracle import bls12_377 as o

pytestmark = pytest.mark.gpu


def _host(bases, scalars, stride=104):
    return ab.VariableBase.msm(o.g1_affine_vec_to_bytes(bases, stride), o.fr_vec_to_bytes(scalars, mont=False), stride)


def test_golden_vectors(golden_dir):
    msm = json.load(open(os.path.join(golden_dir, "msm_golden.json")))
    for name, g in msm.items():
        got = ab.VariableBase.msm(bytes.fromhex(g["bases104"]), bytes.fromhex(g["scalars"]), 104)
        assert got.hex() == g["result"], name


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 255, 1000, 4097])
def test_parity_small_sizes_both_strides(n):
    B = o.synthetic_bases(n, 40 + n)
    s = o.random_fr_vec(n, 41 + n)
    want = o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 40 + n, s) if n else None)
    assert _host(B, s, 104) == want
    assert _host(B, s, 96) == want


def test_edge_cases():
    n = 300
    B = o.synthetic_bases(n, 5)
    s = o.random_fr_vec(n, 6)
    R = o.R_MOD

    def check(bases, scalars, stride=104):
        assert _host(bases, scalars, stride) == o.g1_projective_to_bytes(o.msm_pippenger(bases, scalars))

    check(B, [0] * n)
    check(B, [1] * n)
    check(B, [R - 1] * n)
    check(B, [0 if i % 2 == 0 else (1 if i % 4 == 1 else s[i]) for i in range(n)])      # witness-like
    check([B[0]] * n, s)                                                               # all points equal
    check([B[0]] * n, [s[0]] * n)                                                      # ... and equal scalars
    check([B[i // 2] if i % 2 == 0 else o.g1_neg(B[i // 2]) for i in range(n)], [s[i // 2] for i in range(n)])
    Binf = [None if i % 7 == 0 else B[i] for i in range(n)]
    check(Binf, s, 104)
    check(Binf, s, 96)
    check([None] * n, s)
    # ragged inputs: upstream zips the slices (shorter length wins)
    assert ab.VariableBase.msm(o.g1_affine_vec_to_bytes(B, 104), o.fr_vec_to_bytes(s[:77], mont=False)) == \
        o.g1_projective_to_bytes(o.msm_pippenger(B[:77], s[:77]))


@pytest.mark.parametrize("log_n", [14, 16, 18])
def test_parity_with_c_oracle(c_oracle, log_n):
    """GPU-generated bases copied to the host, C oracle as the independent checker"""
    n = 1 << log_n
    s0, d = o.base_dlogs(n, 5000 + log_n)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, 999 + log_n)
    got = ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes()
    hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
    # the generator itself is checked against the C oracle's own generator, byte for byte
    ref_b = np.zeros(n * 104, dtype=np.uint8)
    c_oracle.oracle_gen_bases(ref_b.ctypes.data, n, 104, o.int_to_le_bytes(s0, 32), o.int_to_le_bytes(d, 32), 0, os.cpu_count() or 1)
    assert np.array_equal(hb, ref_b)
    out = C.create_string_buffer(144)
    c_oracle.oracle_msm_g1(out, hb.ctypes.data, n, hs.ctypes.data, 104, os.cpu_count() or 1)
    assert got == out.raw
    # host-pointer entry point on the same data (numpy buffers, zero copy)
    assert ab.VariableBase.msm(hb, hs, 104) == out.raw


@pytest.mark.parametrize("log_n,dist", [(20, "uniform"), (20, "witness"), (22, "witness-signed"), (22, "uniform"), (24, "uniform"),
                                          (26, "uniform"), (26, "witness-signed")])   # 2^26: the top of BASELINE.json's metric range
def test_known_discrete_logs_at_full_size(log_n, dist):
    """bases (s0 + i d) G: the result must be (sum_i s_i (s0 + i d)) G -- checkable at any size"""
    import torch

    n = 1 << log_n
    s0, d = o.base_dlogs(n, 7000 + log_n)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    assert ab.check_on_curve_dev(bases, n, 104)
    raw = bases.view(-1, 104)[[0, 1, n // 3, n - 1]].cpu().numpy().tobytes()
    for k, i in enumerate([0, 1, n // 3, n - 1]):
        assert o.g1_affine_from_bytes(raw[k * 104:(k + 1) * 104]) == o.g1_mul(o.G1_GEN, (s0 + i * d) % o.R_MOD)
    sc = ab.gen_scalars_dev(n, 31337 + log_n)
    if dist.startswith("witness"):     # 50 % zero, 25 % one, 25 % uniform -- KZG witness polynomials look like this
        idx = torch.arange(n, device="cuda")
        sc[idx % 2 == 0] = 0
        one = torch.tensor([1, 0, 0, 0], dtype=torch.int64, device="cuda")
        sc[idx % 4 == 1] = one
    if dist == "witness-signed":       # ... plus "negative" small values r - 1, r - 2, r - 3 (half-range recoding)
        rm = [o.R_MOD - k for k in (1, 2, 3)]
        for j, v in enumerate(rm):
            limbs = np.frombuffer(o.int_to_le_bytes(v, 32), dtype=np.int64).copy()
            sc[idx % 16 == 2 * j + 2] = torch.from_numpy(limbs).cuda()
    k = ab.dlog_dot_dev(sc, n, s0, d)
    # the on-device dot product is itself checked on a prefix against Python big integers
    m = 2048
    pref = o.fr_vec_from_bytes(sc[:m].cpu().numpy().tobytes(), mont=False)
    assert ab.dlog_dot_dev(sc, m, s0, d) == sum(pref[i] * (s0 + i * d) for i in range(m)) % o.R_MOD
    got = ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes()
    assert got == o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, k))


def test_linearity_and_point_range_sharding():
    """msm(P, s) = sum of msm over disjoint point ranges (the multi-GPU decomposition), combined by
    aleo_b200_g1_sum_dev"""
    import torch

    n = 1 << 16
    s0, d = o.base_dlogs(n, 1234)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, 55)
    full = ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes()
    parts = torch.empty(4 * 144, dtype=torch.uint8, device="cuda")
    q = n // 4
    for r in range(4):
        ab.VariableBase.msm_dev(bases[r * q * 104:], sc[r * q:], q, 104, out=parts[r * 144:(r + 1) * 144])
    assert ab.VariableBase.sum_partials_dev(parts, 4).cpu().numpy().tobytes() == full


def test_concurrent_host_calls_are_reentrant():
    """snarkVM commits polynomials concurrently from a rayon pool: the ABI must be re-entrant"""
    cases = []
    for t in range(6):
        n = 500 + 37 * t
        B = o.synthetic_bases(n, 70 + t)
        s = o.random_fr_vec(n, 80 + t)
        cases.append((o.g1_affine_vec_to_bytes(B, 104), o.fr_vec_to_bytes(s, mont=False),
                      o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 70 + t, s))))
    results = [None] * len(cases)

    def work(i):
        for _ in range(3):
            results[i] = ab.VariableBase.msm(cases[i][0], cases[i][1], 104)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(results[i] == cases[i][2] for i in range(len(cases)))


def test_error_codes():
    lib = ab.get_lib()
    buf = C.create_string_buffer(256)
    p = C.cast(buf, C.c_void_p)
    assert lib.msm_g1(p, p, 1, p, 100) == ab._lib.EINVAL
    assert lib.msm_g1(None, p, 1, p, 104) == ab._lib.EINVAL
    assert lib.msm_g1(p, None, 1, p, 104) == ab._lib.EINVAL
    assert lib.msm_g1(p, None, 0, None, 104) == ab._lib.OK          # n = 0 -> identity
    assert buf.raw[:144] == o.g1_projective_to_bytes(None)
    assert lib.msm_g1_dev(p, p, 1 << 40, p, 104, None) == ab._lib.ETOOLARGE


@pytest.mark.parametrize("chunks", [1, 2, 3, 4])
def test_host_call_streams_point_ranges(chunks, monkeypatch):
    """aleo_b200_msm_g1 copies and accumulates point range by point range (H2D overlapped); the result must
    not depend on the number of ranges.  2^20 points, known discrete logs, witness-like scalars mixed in."""
    monkeypatch.setenv("ALEO_B200_MSM_CHUNKS", str(chunks))
    n = (1 << 20) + 12345
    seed = 4242
    s0, d = o.base_dlogs(n, seed)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, seed)
    sc[::3] = 0                                  # a third zero, a sixth one: the hot buckets of a witness
    sc[1::6, 0] = 1
    sc[1::6, 1:] = 0
    k = ab.dlog_dot_dev(sc, n, s0, d, 0)
    want = o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, k))
    assert ab.VariableBase.msm(bases.cpu().numpy(), sc.cpu().numpy(), 104) == want
    srs = ab.ResidentSRS.from_device(bases, n, 104)
    try:
        assert srs.msm(sc.cpu().numpy()) == want
    finally:
        srs.close()


def test_single_process_multi_gpu_msm():
    """aleo_b200_msm_g1_multi: one process, one host thread per GPU, point ranges, single final combine.  Uses every
    visible GPU (1 on the default box: then it must equal the plain call)."""
    import torch

    ndev = torch.cuda.device_count()
    n = (1 << 18) + 77
    s0, d = o.base_dlogs(n, 8181)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, 8182)
    want = o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, ab.dlog_dot_dev(sc, n, s0, d, 0)))
    hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
    for k in sorted({1, ndev, min(2, ndev)}):
        assert ab.VariableBase.msm_multi(hb, hs, k, 104) == want, k
    assert ab.get_lib().msm_g1_multi(None, None, 10, None, 104, 0) == -1
    assert ab.get_lib().msm_g1_multi(hb.ctypes.data, hb.ctypes.data, n, hs.ctypes.data, 104, ndev + 1) == -3


@pytest.mark.parametrize("mode", ["chunk", "cta", "cta3"])
def test_reduction_level_variants_agree(c_oracle, mode, monkeypatch):
    """bucket reduction with serial chunk levels (a thread per chunk) and with CTA-cooperative levels (a CTA per chunk:
    suffix scan + tree; cta3 = chunks of 8, several levels joining plain sums): plain MSM against the C oracle at 2^15,
    resident-SRS MSM and a batch of commitments against the plain path"""
    import torch
    monkeypatch.setenv("ALEO_B200_MSM_REDUCE", mode)
    n = 1 << 15
    s0, d = o.base_dlogs(n, 6100)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, 6200)
    hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
    want = C.create_string_buffer(144)
    c_oracle.oracle_msm_g1(want, hb.ctypes.data, n, hs.ctypes.data, 104, os.cpu_count() or 1)
    assert ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes() == want.raw
    srs = ab.ResidentSRS.from_device(bases, n, 104)
    try:
        assert srs.msm_dev(sc, n).cpu().numpy().tobytes() == want.raw
        scm = ab.gen_scalars_dev(n, 6200, 0, True)             # the same values in Montgomery form
        one = ab.KZG10.commit_dev(srs, scm, n)
        batch = ab.KZG10.commit_batch_dev(srs, [scm, scm[: n // 2].contiguous(), scm])
        assert torch.equal(batch[0], one) and torch.equal(batch[2], one)
        assert torch.equal(batch[1], ab.KZG10.commit_dev(srs, scm[: n // 2].contiguous(), n // 2))
    finally:
        srs.close()


@pytest.mark.parametrize("levels,k", [(1, 3), (3, 7), (6, 256)])
def test_batch_affine_levels_edge_cases(levels, k, monkeypatch):
    """batch-affine pair-tree levels (csrc/msm_ba.cuh) forced on crowded buckets (c = 5: ~40 entries per bucket): ordinary
    pairs, P + P, P + (-P), infinity bases, odd leftovers, one huge bucket -- against the big-integer oracle"""
    monkeypatch.setenv("ALEO_B200_MSM_C", "5")
    monkeypatch.setenv("ALEO_B200_MSM_BA", str(levels))
    monkeypatch.setenv("ALEO_B200_MSM_BA_K", str(k))
    n = 700
    B = o.synthetic_bases(n, 71)
    s = o.random_fr_vec(n, 72)
    s[0], s[1], s[2] = o.R_MOD - 1, 1, 0

    def check(bases, scalars, stride=104):
        assert _host(bases, scalars, stride) == o.g1_projective_to_bytes(o.msm_pippenger(bases, scalars))

    check(B, s, 104)
    check(B, s, 96)
    check([B[0]] * n, s)
    check([B[0]] * n, [s[0]] * n)
    check([B[i // 2] if i % 2 == 0 else o.g1_neg(B[i // 2]) for i in range(n)], [s[i // 2] for i in range(n)])
    Binf = [None if i % 5 == 0 else B[i] for i in range(n)]
    check(Binf, s, 104)
    check(Binf, s, 96)
    check([None] * n, s)
    check(B, [1] * n)
    check(B, [0 if i % 2 == 0 else (1 if i % 4 == 1 else s[i]) for i in range(n)])


@pytest.mark.parametrize("levels", [0, 2, 5])
@pytest.mark.parametrize("log_n", [16, 18])
def test_batch_affine_levels_match_c_oracle(c_oracle, log_n, levels, monkeypatch):
    """the same MSM with 0 / 2 / 5 batch-affine levels in front of the XYZZ kernel and a 1 MB..64 MB workspace budget
    (several bucket-set groups): bytes equal to the C oracle's"""
    monkeypatch.setenv("ALEO_B200_MSM_BA", str(levels))
    monkeypatch.setenv("ALEO_B200_MSM_BA_MB", "64")
    monkeypatch.setenv("ALEO_B200_MSM_C", "10")
    n = 1 << log_n
    s0, d = o.base_dlogs(n, 7000 + log_n)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    sc = ab.gen_scalars_dev(n, 1999 + log_n)
    got = ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes()
    hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
    out = C.create_string_buffer(144)
    c_oracle.oracle_msm_g1(out, hb.ctypes.data, n, hs.ctypes.data, 104, os.cpu_count() or 1)
    assert got == out.raw
    assert ab.VariableBase.msm(hb, hs, 104) == out.raw      # host ranges accumulate into the same buckets
    if levels:
        monkeypatch.setenv("ALEO_B200_TEST_DENY_BA_ALLOC", "1")   # a device without room for the levels: XYZZ kernel alone
        assert ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes() == out.raw


@pytest.mark.parametrize("stage,ca", [(0, 0), (1, 0), (1, 2), (2, 0), (2, 1)])
def test_batch_affine_operand_staging_variants(c_oracle, stage, ca, monkeypatch):
    """the level kernel's operand paths -- plain loads, cp.async staging above level 0 (the default), staging of level 0
    through the enclosing 16-byte words of stride-104 bases, .cg / .ca copies -- give the same bytes as the C oracle"""
    monkeypatch.setenv("ALEO_B200_MSM_BA", "3")
    monkeypatch.setenv("ALEO_B200_MSM_BA_STAGE", str(stage))
    monkeypatch.setenv("ALEO_B200_MSM_BA_CA", str(ca))
    monkeypatch.setenv("ALEO_B200_MSM_C", "9")
    n = 1 << 16
    s0, d = o.base_dlogs(n, 8100)
    sc = ab.gen_scalars_dev(n, 2100)
    hs = sc.cpu().numpy()
    for stride in (104, 96):
        bases = ab.gen_bases_dev(n, s0, d, 0, stride)
        got = ab.VariableBase.msm_dev(bases, sc, n, stride).cpu().numpy().tobytes()
        hb = bases.cpu().numpy()
        out = C.create_string_buffer(144)
        c_oracle.oracle_msm_g1(out, hb.ctypes.data, n, hs.ctypes.data, stride, os.cpu_count() or 1)
        assert got == out.raw, stride


@pytest.mark.parametrize("glv", ["0", "1"])
def test_glv_split_on_and_off(c_oracle, glv, monkeypatch):
    """proof-sized plain MSMs with and without the GLV split (scalars k = k1 + k2 u^2, 2 n virtual points, half the windows):
    edge scalars against the big-integer oracle, 2^16 / 2^18 against the C oracle, device and host entry points"""
    monkeypatch.setenv("ALEO_B200_MSM_GLV", glv)
    u2 = 0x8508c00000000001 ** 2
    n = 400
    B = o.synthetic_bases(n, 91)
    s = o.random_fr_vec(n, 92)
    s[:12] = [0, 1, u2 - 1, u2, u2 + 1, 2 * u2, (o.R_MOD - 1) // u2 * u2, o.R_MOD - 1, (1 << 126) - 1, 1 << 126, (1 << 127) - 1, (1 << 252) + 12345]
    Binf = [None if i % 3 == 0 else B[i] for i in range(n)]
    for bases in (B, Binf, [B[0]] * n):
        for stride in (104, 96):
            assert _host(bases, s, stride) == o.g1_projective_to_bytes(o.msm_pippenger(bases, s))
    for log_n in (16, 18):
        m = 1 << log_n
        s0, d = o.base_dlogs(m, 9000 + log_n)
        bases = ab.gen_bases_dev(m, s0, d, 0, 104)
        sc = ab.gen_scalars_dev(m, 3000 + log_n)
        got = ab.VariableBase.msm_dev(bases, sc, m, 104).cpu().numpy().tobytes()
        hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
        out = C.create_string_buffer(144)
        c_oracle.oracle_msm_g1(out, hb.ctypes.data, m, hs.ctypes.data, 104, os.cpu_count() or 1)
        assert got == out.raw
        assert ab.VariableBase.msm(hb, hs, 104) == out.raw
