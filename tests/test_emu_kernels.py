"""Kernel-logic checks on the development emulator (tests/emu/emu_runtime.hpp): the SAME kernel
sources as the product, compiled by g++ and run with one OS thread per CUDA thread, compared with
the oracle.  This is development tooling for a GPU-less container -- it proves indexing / carry-chain
logic, not performance, and is never reachable from the aleo_b200 package.  The parity tests proper
are the `-m gpu` tests, which call the product library through the C ABI on a B200."""
import ctypes as C
import json
import os

import pytest

from oracle import bls12_377 as o


def _ntt(lib, vals, log_n, direction, kind):
    buf = C.create_string_buffer(o.fr_vec_to_bytes(vals), max(1, len(vals)) * 32)
    lib.check(lib.ntt_fr_dev(C.cast(buf, C.c_void_p), log_n, 1, direction, kind, None), "ntt")
    return o.fr_vec_from_bytes(buf.raw[: len(vals) * 32])


def _msm(lib, bases, scalars, stride):
    n = min(len(bases), len(scalars))
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(bases[:n], stride), max(1, n * stride))
    sb = C.create_string_buffer(o.fr_vec_to_bytes(scalars[:n], mont=False), max(1, n * 32))
    out = C.create_string_buffer(144)
    lib.check(lib.msm_g1_dev(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(sb, C.c_void_p), stride, None), "msm")
    return out.raw


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 11, 12, 13, 14])
def test_emu_ntt_all_variants(emu_lib, log_n):
    assert b"EMULATOR" in emu_lib.version()
    n = 1 << log_n
    v = o.random_fr_vec(n, 300 + log_n)
    assert _ntt(emu_lib, v, log_n, 0, 0) == o.fft(v)
    assert _ntt(emu_lib, v, log_n, 1, 0) == o.ifft(v)
    assert _ntt(emu_lib, v, log_n, 0, 1) == o.coset_fft(v)
    assert _ntt(emu_lib, v, log_n, 1, 1) == o.coset_ifft(v)


@pytest.mark.parametrize("tile", [10, 11])
def test_emu_ntt_both_tile_sizes(emu_lib, monkeypatch, tile):
    """1024- and 2048-element tiles on the same sizes (the default picks 1024 up to 2^20)"""
    monkeypatch.setenv("ALEO_B200_NTT_TILE", str(tile))
    for log_n in (12, 15):
        v = o.random_fr_vec(1 << log_n, 500 + log_n)
        assert _ntt(emu_lib, v, log_n, 0, 0) == o.fft(v)
        assert _ntt(emu_lib, v, log_n, 1, 1) == o.coset_ifft(v)


def test_emu_ntt_extreme_values(emu_lib):
    """the passes keep semi-reduced values (< 2r) and feed untested sums / differences (< 4r) into raw products:
    vectors made of r - 1, r - 2, 0 and 1 push every intermediate towards those bounds"""
    import random
    log_n = 12
    n = 1 << log_n
    rnd = random.Random(12)
    top = o.R_MOD - 1
    for v in ([top] * n, [top if i & 1 else 0 for i in range(n)], [rnd.choice((0, 1, top, top - 1)) for _ in range(n)]):
        assert _ntt(emu_lib, v, log_n, 0, 0) == o.fft(v)
        assert _ntt(emu_lib, v, log_n, 1, 1) == o.coset_ifft(v)


def test_emu_ntt_three_passes(emu_lib):
    log_n = 17                      # 6 + 6 + 5 bits: exercises the middle-digit reversal
    v = o.random_fr_vec(1 << log_n, 17)
    assert _ntt(emu_lib, v, log_n, 0, 0) == o.fft(v)


def test_emu_msm_golden(emu_lib, golden_dir):
    msm = json.load(open(os.path.join(golden_dir, "msm_golden.json")))
    for name, g in msm.items():
        raw = bytes.fromhex(g["bases104"])
        B = [o.g1_affine_from_bytes(raw[i * 104:(i + 1) * 104]) for i in range(g["n"])]
        s = o.fr_vec_from_bytes(bytes.fromhex(g["scalars"]), mont=False)
        for stride in (104, 96):
            assert _msm(emu_lib, B, s, stride).hex() == g["result"], (name, stride)


def test_emu_msm_edge_cases(emu_lib):
    n = 120
    B = o.synthetic_bases(n, 5)
    s = o.random_fr_vec(n, 6)
    R = o.R_MOD

    def ok(bases, scalars, stride=104):
        return _msm(emu_lib, bases, scalars, stride) == o.g1_projective_to_bytes(o.msm_pippenger(bases, scalars))

    assert _msm(emu_lib, [], [], 104) == o.g1_projective_to_bytes(None)          # empty
    assert ok(B, [0] * n) and ok(B, [1] * n) and ok(B, [R - 1] * n)
    assert ok([B[0]] * n, s)                                                     # all points equal (doubling path)
    assert ok([B[i // 2] if i % 2 == 0 else o.g1_neg(B[i // 2]) for i in range(n)], [s[i // 2] for i in range(n)])
    Binf = [None if i % 7 == 0 else B[i] for i in range(n)]
    assert ok(Binf, s, 104) and ok(Binf, s, 96) and ok([None] * n, s)
    assert ok(B, s[:50])                                                         # ragged: zip semantics


def test_emu_msm_heavy_buckets_are_split(emu_lib):
    n = 2000                        # all scalars equal: one bucket per window -> CTA combine path
    B = o.synthetic_bases(n, 9)
    k = o.random_fr_vec(1, 3)[0]
    assert _msm(emu_lib, B, [k] * n, 104) == o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 9, [k] * n))
    sc = [1 if i % 4 else 0 for i in range(n)]   # witness-like 0 / 1 scalars
    assert _msm(emu_lib, B, sc, 104) == o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 9, sc))


def test_emu_msm_multi_level_reduction(emu_lib):
    n = 1 << 13                     # c = 9: two reduction levels
    B = o.synthetic_bases(n, 21)
    s = o.random_fr_vec(n, 22)
    assert _msm(emu_lib, B, s, 104) == o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 21, s))


def test_emu_generators_and_checks(emu_lib):
    n, seed = 70, 77
    s0, d = o.base_dlogs(n, seed)
    for stride in (104, 96):
        bb = C.create_string_buffer(n * stride)
        emu_lib.check(emu_lib.gen_bases_dev(C.cast(bb, C.c_void_p), n, stride, o.int_to_le_bytes(s0, 32),
                                            o.int_to_le_bytes(d, 32), 5, None), "gen")
        want = [o.g1_mul(o.G1_GEN, (s0 + (5 + i) * d) % o.R_MOD) for i in range(n)]
        assert bb.raw == o.g1_affine_vec_to_bytes(want, stride)
        assert emu_lib.check_on_curve_dev(C.cast(bb, C.c_void_p), n, stride, None) == 1
        bad = bytearray(bb.raw)
        bad[3 * stride + 5] ^= 1
        cb = C.create_string_buffer(bytes(bad), n * stride)
        assert emu_lib.check_on_curve_dev(C.cast(cb, C.c_void_p), n, stride, None) == 0
    sb = C.create_string_buffer(n * 32)
    emu_lib.check(emu_lib.gen_scalars_dev(C.cast(sb, C.c_void_p), n, 1234, 10, 0, None), "scalars")
    sc = o.fr_vec_from_bytes(sb.raw, mont=False)
    assert all(v < (1 << 252) for v in sc) and len(set(sc)) == n
    out = C.create_string_buffer(32)
    emu_lib.check(emu_lib.dlog_dot_dev(C.cast(out, C.c_void_p), C.cast(sb, C.c_void_p), n, o.int_to_le_bytes(s0, 32),
                                       o.int_to_le_bytes(d, 32), 5, None), "dot")
    assert o.le_bytes_to_int(out.raw) == sum(sc[i] * (s0 + (5 + i) * d) for i in range(n)) % o.R_MOD


def test_emu_resident_srs_and_kzg_commit(emu_lib):
    """resident-SRS mode (bases expanded to 2^(cw) P_i, one shared bucket set) and the KZG commit path"""
    n = 700
    B = o.synthetic_bases(n, 61)
    B[5] = None
    for stride in (104, 96):
        bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, stride), n * stride)
        h = C.c_void_p()
        emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), n, stride, None), "srs_create")
        nn, c, w, by = C.c_size_t(), C.c_int(), C.c_int(), C.c_size_t()
        emu_lib.check(emu_lib.srs_info(h, C.byref(nn), C.byref(c), C.byref(w), C.byref(by)), "info")
        assert nn.value == n and c.value == 10 and w.value == 26 and by.value == n * 26 * 96
        for n_used, seed in ((n, 1), (n // 3, 2), (1, 3), (0, 4)):       # any prefix of the SRS
            s = o.random_fr_vec(n_used, 70 + seed)
            if n_used > 10:
                s[3], s[4], s[7] = 0, 1, o.R_MOD - 1
            sb = C.create_string_buffer(o.fr_vec_to_bytes(s, mont=False), max(1, n_used * 32))
            out = C.create_string_buffer(144)
            emu_lib.check(emu_lib.srs_msm_dev(h, C.cast(out, C.c_void_p), C.cast(sb, C.c_void_p), n_used, None), "srs_msm")
            want = o.msm_pippenger(B[:n_used], s) if n_used else None
            assert out.raw == o.g1_projective_to_bytes(want), (stride, n_used)
        # KZG10::commit: Montgomery coefficients in, compressed G1 out
        coeffs = o.random_fr_vec(200, 99)
        cb = C.create_string_buffer(o.fr_vec_to_bytes(coeffs, mont=True), 200 * 32)
        out48 = C.create_string_buffer(48)
        emu_lib.check(emu_lib.kzg_commit_dev(h, C.cast(out48, C.c_void_p), C.cast(cb, C.c_void_p), 200, None), "commit")
        assert out48.raw == o.g1_compress(o.msm_pippenger(B[:200], coeffs))
        emu_lib.check(emu_lib.kzg_commit_dev(h, C.cast(out48, C.c_void_p), C.cast(cb, C.c_void_p), 0, None), "commit0")
        assert out48.raw == o.g1_compress(None)
        emu_lib.check(emu_lib.srs_destroy(h), "destroy")


@pytest.mark.parametrize("chunks", [2, 3])
def test_emu_host_calls_stream_point_ranges(emu_lib, chunks, monkeypatch):
    """host-pointer entry points: the MSM arrives in point ranges whose bucket sums are added up
    (Session::add_chunk) -- same result as one range, for the plain MSM, the resident SRS and KZG commit"""
    monkeypatch.setenv("ALEO_B200_MSM_CHUNKS", str(chunks))
    n = 900
    B = o.synthetic_bases(n, 31)
    B[17] = None
    s = o.random_fr_vec(n, 32)
    s[0], s[1], s[2] = 0, 1, o.R_MOD - 1
    s[400:440] = [s[400]] * 40                                   # a bucket that every range re-opens
    want = o.g1_projective_to_bytes(o.msm_pippenger(B, s))
    for stride in (104, 96):
        bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, stride), n * stride)
        sb = C.create_string_buffer(o.fr_vec_to_bytes(s, mont=False), n * 32)
        out = C.create_string_buffer(144)
        emu_lib.check(emu_lib.msm_g1(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(sb, C.c_void_p), stride), "msm_g1")
        assert out.raw == want, stride
    # all points equal / cancelling pairs across range boundaries
    Beq = [B[3]] * n
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(Beq, 104), n * 104)
    ones = C.create_string_buffer(o.fr_vec_to_bytes([5] * n, mont=False), n * 32)
    out = C.create_string_buffer(144)
    emu_lib.check(emu_lib.msm_g1(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(ones, C.c_void_p), 104), "msm_g1 eq")
    assert out.raw == o.g1_projective_to_bytes(o.g1_mul(B[3], 5 * n))
    Bpm = [B[3] if i < n // 2 else o.g1_neg(B[3]) for i in range(n)]
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(Bpm, 104), n * 104)
    emu_lib.check(emu_lib.msm_g1(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(ones, C.c_void_p), 104), "msm_g1 pm")
    assert out.raw == o.g1_projective_to_bytes(None)
    # resident SRS + commit through the host entry points
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    h = C.c_void_p()
    emu_lib.check(emu_lib.srs_create(C.byref(h), C.cast(bb, C.c_void_p), n, 104), "srs_create")
    for n_used in (n, 601, 9):
        sb = C.create_string_buffer(o.fr_vec_to_bytes(s[:n_used], mont=False), n_used * 32)
        emu_lib.check(emu_lib.srs_msm(h, C.cast(out, C.c_void_p), C.cast(sb, C.c_void_p), n_used), "srs_msm")
        assert out.raw == o.g1_projective_to_bytes(o.msm_pippenger(B[:n_used], s[:n_used])), n_used
        cb = C.create_string_buffer(o.fr_vec_to_bytes(s[:n_used], mont=True), n_used * 32)
        out48 = C.create_string_buffer(48)
        emu_lib.check(emu_lib.kzg_commit(h, C.cast(out48, C.c_void_p), C.cast(cb, C.c_void_p), n_used), "kzg_commit")
        assert out48.raw == o.g1_compress(o.msm_pippenger(B[:n_used], s[:n_used])), n_used
    emu_lib.check(emu_lib.srs_destroy(h), "destroy")


@pytest.mark.parametrize("mode", ["chunk", "cta", "cta2"])   # cta2: CTA levels of 4 elements -> several levels, the later ones join plain sums
@pytest.mark.parametrize("c", [5, 12, 14])
def test_emu_msm_reduction_stages(emu_lib, c, mode, monkeypatch):
    """forced window sizes: c = 5 -> scan stage only, 12 -> one level + scan, 14 -> two levels + scan, each with the
    serial chunk levels (a thread per chunk) and with the CTA-cooperative levels (a CTA per chunk: suffix scan + tree);
    also exercises the quad-cooperative doubling tail over 51 / 22 / 19 windows and the binary-GCD inversion"""
    monkeypatch.setenv("ALEO_B200_MSM_C", str(c))
    monkeypatch.setenv("ALEO_B200_MSM_REDUCE", mode)
    n = 200
    assert emu_lib.msm_window_bits(n) == c
    B = o.synthetic_bases(n, 41)
    s = o.random_fr_vec(n, 42)
    s[0], s[1] = o.R_MOD - 1, 1
    assert _msm(emu_lib, B, s, 104) == o.g1_projective_to_bytes(o.msm_pippenger(B, s))


def _field_op(lib, field, op, a, b=None):
    size = 32 if field == 0 else 48
    n = len(a)
    enc = (lambda v: o.fr_vec_to_bytes(v)) if field == 0 else (lambda v: b"".join(o.int_to_le_bytes(o.fq_to_mont(x), 48) for x in v))
    ab = C.create_string_buffer(enc(a), n * size)
    bb = C.create_string_buffer(enc(b), n * size) if b is not None else None
    out = C.create_string_buffer(n * size)
    lib.check(lib.field_op_dev(field, op, C.cast(out, C.c_void_p), C.cast(ab, C.c_void_p),
                               C.cast(bb, C.c_void_p) if bb is not None else None, n, None), "field_op")
    if field == 0:
        return o.fr_vec_from_bytes(out.raw)
    return [o.fq_from_mont(o.le_bytes_to_int(out.raw[i * 48:(i + 1) * 48])) for i in range(n)]


def field_edge_values(mod, nlimbs):
    """values that stress the carry chains: 0, 1, m-1, all-ones limbs, single high bits, R mod m ..."""
    R = 1 << (32 * nlimbs)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, R % mod, (R * R) % mod, pow(R, -1, mod)]
    vals += [((1 << k) - 1) % mod for k in (31, 32, 33, 63, 64, 65, 32 * nlimbs - 8)]
    vals += [(1 << k) % mod for k in range(0, mod.bit_length(), 29)]
    vals += [v for x in list(vals) for v in ((x * pow(R, -1, mod)) % mod,)]   # Montgomery images with the same shapes
    return vals


@pytest.mark.parametrize("field", [0, 1])
def test_emu_field_ops(emu_lib, field):
    """add / sub / mul / dedicated square / binary-GCD inverse / neg against big-integer arithmetic"""
    import random
    mod, nl = (o.R_MOD, 8) if field == 0 else (o.P_MOD, 12)
    rng = random.Random(1234 + field)
    a = field_edge_values(mod, nl) + [rng.randrange(mod) for _ in range(120)]
    b = list(reversed(field_edge_values(mod, nl))) + [rng.randrange(mod) for _ in range(120)]
    assert _field_op(emu_lib, field, 0, a, b) == [(x + y) % mod for x, y in zip(a, b)]
    assert _field_op(emu_lib, field, 1, a, b) == [(x - y) % mod for x, y in zip(a, b)]
    assert _field_op(emu_lib, field, 2, a, b) == [(x * y) % mod for x, y in zip(a, b)]
    assert _field_op(emu_lib, field, 3, a) == [(x * x) % mod for x in a]
    assert _field_op(emu_lib, field, 4, a) == [pow(x, -1, mod) if x else 0 for x in a]
    assert _field_op(emu_lib, field, 5, a) == [(-x) % mod for x in a]


def test_emu_fp64_multiplier_column_logic(emu_lib):
    """csrc/mont_fp64.cuh on the emulator: the 48-bit limb conversion, the column accumulation with its exponent-field
    biases, the shift-only Montgomery factor and the carry normalisation (the emulator forms the exact hi / lo halves
    with 128-bit integers; the fma_rz split itself is checked on the GPU by tests/test_gpu_field.py)"""
    import random
    mod = o.P_MOD
    rng = random.Random(99)
    a = field_edge_values(mod, 12) + [rng.randrange(mod) for _ in range(200)]
    b = list(reversed(field_edge_values(mod, 12))) + [rng.randrange(mod) for _ in range(200)]
    for _ in range(300):
        limbs = [rng.choice([0, (1 << 48) - 1, 1, 1 << 47, rng.getrandbits(48)]) for _ in range(8)]
        a.append(sum(v << (48 * i) for i, v in enumerate(limbs)) % mod * pow(1 << 384, -1, mod) % mod)
        b.append(rng.randrange(mod))
    enc = lambda v: b"".join(o.int_to_le_bytes(o.fq_to_mont(x), 48) for x in v)  # noqa: E731
    A, B = C.create_string_buffer(enc(a), len(a) * 48), C.create_string_buffer(enc(b), len(b) * 48)
    out = C.create_string_buffer(len(a) * 48)
    dec = lambda raw: [o.fq_from_mont(o.le_bytes_to_int(raw[i:i + 48])) for i in range(0, len(raw), 48)]  # noqa: E731
    emu_lib.check(emu_lib.fq_mul_fp64_dev(C.cast(out, C.c_void_p), C.cast(A, C.c_void_p), C.cast(B, C.c_void_p), len(a), 0, None), "mul")
    assert dec(out.raw) == [(x * y) % mod for x, y in zip(a, b)]
    emu_lib.check(emu_lib.fq_mul_fp64_dev(C.cast(out, C.c_void_p), C.cast(A, C.c_void_p), None, len(a), 1, None), "sqr")
    assert dec(out.raw) == [(x * x) % mod for x in a]


def _bitrev_list(v):
    n = len(v)
    bits = n.bit_length() - 1
    return [v[int(format(i, "0%db" % bits)[::-1], 2)] if bits else v[i] for i in range(n)]


@pytest.mark.parametrize("log_n", [0, 1, 3, 8, 12])
def test_emu_ntt_orders(emu_lib, log_n):
    """FFTOrder IO / OI (bit-reversed output / input) around the in-order transform"""
    n = 1 << log_n
    v = o.random_fr_vec(n, 500 + log_n)

    def run(vals, direction, kind, order):
        buf = C.create_string_buffer(o.fr_vec_to_bytes(vals), n * 32)
        emu_lib.check(emu_lib.ntt_fr_ordered_dev(C.cast(buf, C.c_void_p), log_n, 1, direction, kind, order, None), "ordered")
        return o.fr_vec_from_bytes(buf.raw)

    assert run(v, 0, 0, 0) == o.fft(v)
    assert run(v, 0, 0, 1) == _bitrev_list(o.fft(v))                   # IO
    assert run(_bitrev_list(v), 0, 0, 2) == o.fft(v)                   # OI
    assert run(_bitrev_list(v), 1, 1, 2) == o.coset_ifft(v)            # OI, coset inverse
    assert run(v, 0, 1, 1) == _bitrev_list(o.coset_fft(v))             # IO, coset forward


@pytest.mark.parametrize("n", [1, 2, 7, 256, 2049, 8193, 20000])
def test_emu_poly_helpers(emu_lib, n):
    """distribute_powers, polynomial evaluation and the KZG witness polynomial against the oracle"""
    c = o.random_fr_vec(n, 700 + n)
    g, k, z = o.random_fr_vec(3, 800 + n)
    mont = lambda v: o.int_to_le_bytes(o.fr_to_mont(v), 32)  # noqa: E731
    buf = C.create_string_buffer(o.fr_vec_to_bytes(c), n * 32)
    emu_lib.check(emu_lib.fr_distribute_powers_dev(C.cast(buf, C.c_void_p), n, mont(g), mont(k), None), "distribute")
    assert o.fr_vec_from_bytes(buf.raw) == o.distribute_powers(c, g, k)
    buf = C.create_string_buffer(o.fr_vec_to_bytes(c), n * 32)
    emu_lib.check(emu_lib.fr_distribute_powers_dev(C.cast(buf, C.c_void_p), n, mont(g), None, None), "distribute k=1")
    assert o.fr_vec_from_bytes(buf.raw) == o.distribute_powers(c, g)
    ybuf = C.create_string_buffer(o.fr_vec_to_bytes(c), n * 32)
    xs = o.random_fr_vec(n, 900 + n)
    xbuf = C.create_string_buffer(o.fr_vec_to_bytes(xs), n * 32)
    emu_lib.check(emu_lib.fr_axpy_dev(C.cast(ybuf, C.c_void_p), C.cast(xbuf, C.c_void_p), mont(k), n, None), "axpy")
    assert o.fr_vec_from_bytes(ybuf.raw) == [(y + k * x) % o.R_MOD for y, x in zip(c, xs)]
    src = C.create_string_buffer(o.fr_vec_to_bytes(c), n * 32)
    out = C.create_string_buffer(32)
    emu_lib.check(emu_lib.fr_poly_eval_dev(C.cast(out, C.c_void_p), C.cast(src, C.c_void_p), n, mont(z), None), "eval")
    assert o.fr_vec_from_bytes(out.raw) == [o.poly_eval(c, z)]
    for zz in (z, 0, 1):
        q = C.create_string_buffer(n * 32)
        emu_lib.check(emu_lib.fr_divide_by_linear_dev(C.cast(q, C.c_void_p), C.cast(src, C.c_void_p), n, mont(zz), None), "divide")
        assert o.fr_vec_from_bytes(q.raw) == o.divide_by_linear(c, zz), zz


def test_emu_srs_bucket_range_passes(emu_lib, monkeypatch):
    """resident SRS scatter in 4 bucket-range passes (what 2^24-point handles do so that the live bucket heads fit L2)"""
    monkeypatch.setenv("ALEO_B200_MSM_SRS_PASSES", "4")
    n = 500
    B = o.synthetic_bases(n, 171)
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    h = C.c_void_p()
    emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), n, 104, None), "srs_create")
    s = o.random_fr_vec(n, 172)
    s[0], s[1], s[2] = 0, 1, o.R_MOD - 1
    sb = C.create_string_buffer(o.fr_vec_to_bytes(s, mont=False), n * 32)
    out = C.create_string_buffer(144)
    emu_lib.check(emu_lib.srs_msm_dev(h, C.cast(out, C.c_void_p), C.cast(sb, C.c_void_p), n, None), "srs_msm")
    assert out.raw == o.g1_projective_to_bytes(o.msm_pippenger(B, s))
    emu_lib.check(emu_lib.srs_destroy(h), "destroy")


def test_emu_kzg_open(emu_lib):
    n = 300
    B = o.synthetic_bases(n, 91)
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    h = C.c_void_p()
    emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), n, 104, None), "srs_create")
    c = o.random_fr_vec(n, 92)
    z = o.random_fr_vec(1, 93)[0]
    src = C.create_string_buffer(o.fr_vec_to_bytes(c), n * 32)
    out48 = C.create_string_buffer(48)
    emu_lib.check(emu_lib.kzg_open_dev(h, C.cast(out48, C.c_void_p), C.cast(src, C.c_void_p), n, o.int_to_le_bytes(o.fr_to_mont(z), 32), None), "open")
    q = o.divide_by_linear(c, z)
    assert out48.raw == o.g1_compress(o.msm_pippenger(B[:n - 1], q[:n - 1]))
    emu_lib.check(emu_lib.kzg_open_dev(h, C.cast(out48, C.c_void_p), C.cast(src, C.c_void_p), 1, o.int_to_le_bytes(o.fr_to_mont(z), 32), None), "open1")
    assert out48.raw == o.g1_compress(None)          # a constant polynomial has the zero witness
    emu_lib.check(emu_lib.srs_destroy(h), "destroy")


def _dist_ntt_emulated(lib, x, log_n, world, direction, kind=0):
    """all ranks of aleo_b200_ntt_dist_* in ONE process (the emulator's 'peer memory' is the shared address space):
    returns the natural-order result assembled from the ranks' output blocks"""
    n = 1 << log_n
    kf, kl, npass = C.c_uint32(), C.c_uint32(), C.c_int()
    lib.check(lib.ntt_dist_layout(log_n, world, C.byref(kf), C.byref(kl), C.byref(npass)), "layout")
    r0, rl = 1 << kf.value, 1 << kl.value
    ctxs, handles = [], b""
    for r in range(world):
        h = C.c_void_p()
        lib.check(lib.ntt_dist_create(C.byref(h), log_n, r, world), "create")
        hb = C.create_string_buffer(128)
        lib.check(lib.ntt_dist_handles(h, C.cast(hb, C.c_void_p)), "handles")
        ctxs.append(h)
        handles += hb.raw
    allh = C.create_string_buffer(handles, len(handles))
    for h in ctxs:
        lib.check(lib.ntt_dist_open(h, C.cast(allh, C.c_void_p)), "open")
    wl = rl // world
    ins = []
    for r in range(world):   # column block of the (n / rl) x rl matrix
        loc = [x[a * rl + r * wl + b] for a in range(n // rl) for b in range(wl)]
        ins.append(C.create_string_buffer(o.fr_vec_to_bytes(loc), len(loc) * 32))
    for twice in range(2):   # second round exercises the other receive buffer
        for r in range(world):
            lib.check(lib.ntt_dist_stage1(ctxs[r], C.cast(ins[r], C.c_void_p), direction, kind, None), "stage1")
        outs = []
        for r in range(world):   # (the barrier between the stages is implicit: stage 1 of every rank has returned)
            ob = C.create_string_buffer(n // world * 32)
            lib.check(lib.ntt_dist_stage2(ctxs[r], C.cast(ob, C.c_void_p), direction, kind, None), "stage2")
            outs.append(o.fr_vec_from_bytes(ob.raw))
    for h in ctxs:
        lib.check(lib.ntt_dist_destroy(h), "destroy")
    w0 = r0 // world
    X = [0] * n
    for t in range(world):   # column block of the (n / r0) x r0 matrix
        for a in range(n // r0):
            for b in range(w0):
                X[a * r0 + t * w0 + b] = outs[t][a * w0 + b]
    return X, npass.value


@pytest.mark.parametrize("log_n,world", [(13, 2), (14, 4), (16, 2), (15, 8), (17, 2), (18, 4)])
def test_emu_distributed_ntt_peer_exchange(emu_lib, log_n, world):
    """one NTT over `world` ranks, exchange by stores into the peers' receive buffers: bit-exact with the oracle,
    forward and inverse, 2 / 3 passes, both receive buffers"""
    v = o.random_fr_vec(1 << log_n, 4000 + log_n)
    got, npass = _dist_ntt_emulated(emu_lib, v, log_n, world, 0)
    assert got == o.fft(v), (log_n, world, npass)
    got, _ = _dist_ntt_emulated(emu_lib, v, log_n, world, 1)
    assert got == o.ifft(v)
    if log_n <= 16:
        got, _ = _dist_ntt_emulated(emu_lib, v, log_n, world, 0, 1)
        assert got == o.coset_fft(v)
        got, _ = _dist_ntt_emulated(emu_lib, v, log_n, world, 1, 1)
        assert got == o.coset_ifft(v)


def test_emu_ntt_nine_bit_middle_pass(emu_lib, c_oracle, monkeypatch):
    """forced split 7 + 9 + 5 at 2^21: a 9-bit pass (three radix-8 steps) in the middle position, checked against
    the C oracle (2^17 = 9 + 8 and 2^18 = 9 + 9 cover the first / last positions in test_emu_ntt_*)"""
    import numpy as np
    monkeypatch.setenv("ALEO_B200_NTT_SPLIT", "7,9,5")
    log_n = 21
    n = 1 << log_n
    assert emu_lib.ntt_launches(log_n) == 3
    rng = np.random.default_rng(21)
    x = rng.integers(0, 2**62, size=(n, 4), dtype=np.int64).astype(np.uint64)
    x[:, 3] &= np.uint64((1 << 60) - 1)                      # any value < 2^252 is a valid (Montgomery) residue
    want = x.copy()
    assert c_oracle.oracle_ntt_fr(want.ctypes.data, log_n, 0, 0, os.cpu_count() or 1) == 0
    got = x.copy()
    emu_lib.check(emu_lib.ntt_fr_dev(got.ctypes.data, log_n, 1, 0, 0, None), "ntt")
    assert np.array_equal(got, want)


def test_emu_kzg_commit_batch(emu_lib):
    """several commitments against one resident SRS in one launch sequence: each equals the single commitment"""
    n = 400
    B = o.synthetic_bases(n, 131)
    B[9] = None
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    h = C.c_void_p()
    emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), n, 104, None), "srs_create")
    sizes = [400, 1, 0, 257, 33, 400]
    polys = [o.random_fr_vec(m, 140 + k) for k, m in enumerate(sizes)]
    polys[0][0], polys[0][1], polys[3][5] = 0, o.R_MOD - 1, 1
    bufs = [C.create_string_buffer(o.fr_vec_to_bytes(p), max(1, len(p)) * 32) for p in polys]
    ptrs = (C.c_void_p * len(sizes))(*[C.cast(b, C.c_void_p) for b in bufs])
    lens = (C.c_size_t * len(sizes))(*sizes)
    out = C.create_string_buffer(48 * len(sizes))
    emu_lib.check(emu_lib.kzg_commit_batch_dev(h, C.cast(out, C.c_void_p), ptrs, lens, len(sizes), None), "batch")
    for k, p in enumerate(polys):
        want = o.g1_compress(o.msm_pippenger(B[:len(p)], p) if p else None)
        assert out.raw[48 * k:48 * (k + 1)] == want, k
    emu_lib.check(emu_lib.srs_destroy(h), "destroy")


@pytest.mark.parametrize("c,fb", [(7, 3)])
def test_emu_msm_partitioned_sort(emu_lib, c, fb, monkeypatch):
    """the two-level shared-memory sort (part_count / part_scatter / part_finish) forced on small inputs, with and
    without fine bits and with several window groups; witness-like scalars, infinity bases, and the host path in
    3 point ranges (kept small: the emulator runs every CTA of a synchronising kernel on 256 OS threads)"""
    monkeypatch.setenv("ALEO_B200_MSM_SORT", "p")
    monkeypatch.setenv("ALEO_B200_MSM_C", str(c))
    monkeypatch.setenv("ALEO_B200_MSM_PART_FB", str(fb))
    n = 300
    B = o.synthetic_bases(n, 151)
    B[11] = None
    s = o.random_fr_vec(n, 152)
    for i in range(0, n, 5):
        s[i] = 0
    for i in range(1, n, 7):
        s[i] = 1
    s[3], s[4] = o.R_MOD - 1, o.R_MOD - 2
    want = o.g1_projective_to_bytes(o.msm_pippenger(B, s))
    assert _msm(emu_lib, B, s, 104) == want
    monkeypatch.setenv("ALEO_B200_MSM_CHUNKS", "3")
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 96), n * 96)
    sb = C.create_string_buffer(o.fr_vec_to_bytes(s, mont=False), n * 32)
    out = C.create_string_buffer(144)
    emu_lib.check(emu_lib.msm_g1(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(sb, C.c_void_p), 96), "msm_g1")
    assert out.raw == want


def test_emu_g1_wire_format_on_reference_fixture(emu_lib, golden_dir):
    """compressed G1 <-> affine on the device kernels, pinned by the REFERENCE's own proof string
    (wasm/src/programs/transaction.rs:100): its 12 commitments decompress to the points the oracle finds (all in
    the r-torsion, tests/test_oracle.py) and compress back to the fixture bytes; plus identity / invalid encodings"""
    fx = json.load(open(os.path.join(golden_dir, "proof_fixture.json")))
    _, payload = o.bech32m_decode(fx["proof"])
    chunks = [bytes(payload[off:off + 48]) for off in fx["g1_offsets"]]
    pts = [o.g1_decompress(c) for c in chunks]
    extra = [o.g1_mul(o.G1_GEN, k) for k in (1, 2, 12345)] + [None]
    blob = b"".join(chunks) + b"".join(o.g1_compress(p) for p in extra)
    bad_x = next(x for x in range(2, 50) if o.fq_sqrt((x ** 3 + 1) % o.P_MOD) is None)
    blob += o.int_to_le_bytes(bad_x, 48) + o.int_to_le_bytes(o.P_MOD + 1, 48)          # not on the curve; x >= p
    # on the curve but OUTSIDE the prime-order subgroup (the G1 cofactor is large): deserialize_compressed rejects these
    rogue = o.curve_points_outside_subgroup(2)
    blob += b"".join(o.g1_compress(p) for p in rogue)
    n = len(blob) // 48
    for stride in (104, 96):
        src = C.create_string_buffer(blob, len(blob))
        out = C.create_string_buffer(n * stride)
        nbad = emu_lib.g1_decompress_dev(C.cast(out, C.c_void_p), stride, C.cast(src, C.c_void_p), n, None)
        assert nbad == 4
        want = o.g1_affine_vec_to_bytes(pts + extra + [None, None, None, None], stride)
        assert out.raw == want
        back = C.create_string_buffer(n * 48)
        emu_lib.check(emu_lib.g1_compress_dev(C.cast(back, C.c_void_p), C.cast(out, C.c_void_p), stride, n, None), "compress")
        assert back.raw[:48 * (n - 4)] == blob[:48 * (n - 4)]
        assert back.raw[48 * (n - 4):] == o.g1_compress(None) * 4
        # Validate::No: the rogue points come through as they are
        nbad = emu_lib.g1_decompress_unchecked_dev(C.cast(out, C.c_void_p), stride, C.cast(src, C.c_void_p), n, None)
        assert nbad == 2
        assert out.raw == o.g1_affine_vec_to_bytes(pts + extra + [None, None] + rogue, stride)


def test_emu_kzg_commit_hiding(emu_lib):
    """commitment + blinding commitment against a second SRS (powers_of_beta_times_gamma_g)"""
    n, m = 120, 5
    B1, B2 = o.synthetic_bases(n, 181), o.synthetic_bases(m, 182)
    hs = []
    for B in (B1, B2):
        bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), len(B) * 104)
        h = C.c_void_p()
        emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), len(B), 104, None), "srs_create")
        hs.append(h)
    p, r = o.random_fr_vec(n, 183), o.random_fr_vec(m, 184)
    pb = C.create_string_buffer(o.fr_vec_to_bytes(p), n * 32)
    rb = C.create_string_buffer(o.fr_vec_to_bytes(r), m * 32)
    out = C.create_string_buffer(48)
    emu_lib.check(emu_lib.kzg_commit_hiding_dev(hs[0], hs[1], C.cast(out, C.c_void_p), C.cast(pb, C.c_void_p), n,
                                                C.cast(rb, C.c_void_p), m, None), "hiding")
    assert out.raw == o.g1_compress(o.g1_add(o.msm_pippenger(B1, p), o.msm_pippenger(B2, r)))
    for h in hs:
        emu_lib.check(emu_lib.srs_destroy(h), "destroy")


@pytest.mark.parametrize("log_n", [0, 3, 9, 13])
def test_emu_lagrange_coefficients(emu_lib, log_n):
    n = 1 << log_n
    tau = o.random_fr_vec(1, 7700 + log_n)[0]
    w = o.fr_root_of_unity(log_n)
    for t in (tau, pow(w, 5 % n, o.R_MOD)):                       # generic point; a point of the domain (indicator vector)
        out = C.create_string_buffer(n * 32)
        emu_lib.check(emu_lib.fr_lagrange_coeffs_dev(C.cast(out, C.c_void_p), log_n, o.int_to_le_bytes(o.fr_to_mont(t), 32), None), "lagrange")
        got = o.fr_vec_from_bytes(out.raw)
        assert got == o.lagrange_coefficients(log_n, t)
        assert sum(got) % o.R_MOD == 1                             # the Lagrange basis sums to one


@pytest.mark.parametrize("log_m,log_n", [(0, 0), (4, 4), (6, 2), (9, 0), (10, 7)])
def test_emu_divide_by_vanishing_on_coset(emu_lib, log_m, log_n):
    e = o.random_fr_vec(1 << log_m, 7800 + 16 * log_m + log_n)
    for g in (22, o.random_fr_vec(1, 7900 + log_m)[0]):
        buf = C.create_string_buffer(o.fr_vec_to_bytes(e), (1 << log_m) * 32)
        emu_lib.check(emu_lib.fr_divide_by_vanishing_on_coset_dev(C.cast(buf, C.c_void_p), log_m, log_n,
                                                                   o.int_to_le_bytes(o.fr_to_mont(g), 32), None), "vanishing")
        assert o.fr_vec_from_bytes(buf.raw) == o.divide_by_vanishing_on_coset(e, log_n, g)
    # the quotient identity: p = q (x^n - 1)  =>  coset_fft(p) / Z_n = coset_fft(q)
    if log_m > log_n:
        m, n = 1 << log_m, 1 << log_n
        q = o.random_fr_vec(m - n, 7950 + log_m)
        p = [0] * m
        for i, v in enumerate(q):
            p[i + n] = (p[i + n] + v) % o.R_MOD
            p[i] = (p[i] - v) % o.R_MOD
        assert o.divide_by_vanishing_on_coset(o.coset_fft(p, m), log_n) == o.coset_fft(q, m)


def test_emu_host_ordered_ntt_and_polymul(emu_lib):
    """aleo_b200_ntt_fr_ordered (upstream's snarkvm_ntt argument list on host pointers) and aleo_b200_polymul
    (snarkvm_polymul's shape): out = ifft(prod fft(p_i) * prod e_j)"""
    log_n = 9
    n = 1 << log_n
    v = o.random_fr_vec(n, 811)
    buf = C.create_string_buffer(o.fr_vec_to_bytes(v), n * 32)
    emu_lib.check(emu_lib.ntt_fr_ordered(C.cast(buf, C.c_void_p), log_n, 0, 0, 1), "ordered host IO")
    assert o.fr_vec_from_bytes(buf.raw) == _bitrev_list(o.fft(v))
    buf = C.create_string_buffer(o.fr_vec_to_bytes(_bitrev_list(v)), n * 32)
    emu_lib.check(emu_lib.ntt_fr_ordered(C.cast(buf, C.c_void_p), log_n, 1, 1, 2), "ordered host OI")
    assert o.fr_vec_from_bytes(buf.raw) == o.coset_ifft(v)
    assert emu_lib.ntt_fr_ordered(C.cast(buf, C.c_void_p), log_n, 0, 0, 7) == -1

    # polymul: two short polynomials and one evaluation vector
    a, b = o.random_fr_vec(100, 812), o.random_fr_vec(n // 2 - 3, 813)
    e = o.random_fr_vec(n, 814)
    pad = lambda p: p + [0] * (n - len(p))  # noqa: E731
    fa, fb = o.fft(pad(a)), o.fft(pad(b))

    def polymul(polys, evals):
        pb = [C.create_string_buffer(o.fr_vec_to_bytes(p), max(1, len(p)) * 32) for p in polys]
        eb = [C.create_string_buffer(o.fr_vec_to_bytes(x), n * 32) for x in evals]
        pp = (C.c_void_p * max(1, len(pb)))(*[C.cast(x, C.c_void_p) for x in pb])
        pl = (C.c_size_t * max(1, len(pb)))(*[len(p) for p in polys])
        ep = (C.c_void_p * max(1, len(eb)))(*[C.cast(x, C.c_void_p) for x in eb])
        el = (C.c_size_t * max(1, len(eb)))(*[n for _ in evals])
        out = C.create_string_buffer(n * 32)
        emu_lib.check(emu_lib.polymul(C.cast(out, C.c_void_p), len(pb), pp, pl, len(eb), ep, el, log_n), "polymul")
        out_d = C.create_string_buffer(n * 32)
        emu_lib.check(emu_lib.polymul_dev(C.cast(out_d, C.c_void_p), len(pb), pp, pl, len(eb), ep, el, log_n, None), "polymul_dev")
        assert out.raw == out_d.raw
        return o.fr_vec_from_bytes(out.raw)

    mul = lambda x, y: [p * q % o.R_MOD for p, q in zip(x, y)]  # noqa: E731
    assert polymul([a, b], [e]) == o.ifft(mul(mul(fa, fb), e))
    assert polymul([a, b], []) == o.ifft(mul(fa, fb))                      # = the plain product a * b (degree < n)
    prod = [0] * n
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            prod[i + j] = (prod[i + j] + x * y) % o.R_MOD
    assert polymul([a, b], []) == prod
    assert polymul([a], []) == pad(a)
    assert polymul([], [e]) == o.ifft(e)
    assert polymul([], [e, fa]) == o.ifft(mul(e, fa))
    assert polymul([b], [e, fa]) == o.ifft(mul(mul(fb, e), fa))
    # argument errors: nothing to multiply; polynomial longer than the domain; evaluation vector of the wrong length
    out = C.create_string_buffer(n * 32)
    one = (C.c_void_p * 1)(C.cast(out, C.c_void_p))
    assert emu_lib.polymul(C.cast(out, C.c_void_p), 0, None, None, 0, None, None, log_n) == -1
    assert emu_lib.polymul(C.cast(out, C.c_void_p), 1, one, (C.c_size_t * 1)(n + 1), 0, None, None, log_n) == -1
    assert emu_lib.polymul(C.cast(out, C.c_void_p), 0, None, None, 1, one, (C.c_size_t * 1)(n - 1), log_n) == -1


def test_emu_kzg_open_combinations(emu_lib):
    """m openings of linear combinations in one call == the single-opening path on the combined polynomial"""
    n = 200
    B = o.synthetic_bases(n, 191)
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    h = C.c_void_p()
    emu_lib.check(emu_lib.srs_create_dev(C.byref(h), C.cast(bb, C.c_void_p), n, 104, None), "srs_create")
    polys = [o.random_fr_vec(ln, 300 + ln) for ln in (200, 150, 37, 1)]
    lc = [[1, 5, 0, 7], [0, 0, 3, 0], [o.R_MOD - 1, 2, 2, 2], [0, 0, 0, 9]]
    zs = o.random_fr_vec(4, 77)
    pb = [C.create_string_buffer(o.fr_vec_to_bytes(p), len(p) * 32) for p in polys]
    pp = (C.c_void_p * 4)(*[C.cast(x, C.c_void_p) for x in pb])
    pl = (C.c_size_t * 4)(*[len(p) for p in polys])
    lcb = b"".join(o.int_to_le_bytes(o.fr_to_mont(a), 32) for row in lc for a in row)
    zb = b"".join(o.int_to_le_bytes(o.fr_to_mont(z), 32) for z in zs)
    out = C.create_string_buffer(4 * 48)
    emu_lib.check(emu_lib.kzg_open_combinations_dev(h, C.cast(out, C.c_void_p), pp, pl, 4, lcb, zb, 4, None), "open_combinations")
    for k in range(4):
        comb = [0] * n
        for a, p in zip(lc[k], polys):
            for i, x in enumerate(p):
                comb[i] = (comb[i] + a * x) % o.R_MOD
        ln = max([len(p) for a, p in zip(lc[k], polys) if a] + [0])
        q = o.divide_by_linear(comb[:ln], zs[k]) if ln > 1 else []
        want = o.g1_compress(o.msm_pippenger(B[:len(q) - 1], q[:len(q) - 1]) if len(q) > 1 else None)
        assert out.raw[48 * k:48 * k + 48] == want, k
    emu_lib.check(emu_lib.srs_destroy(h), "destroy")


@pytest.mark.parametrize("levels", [1, 3, 6])
@pytest.mark.parametrize("c", [4, 7])
def test_emu_msm_batch_affine_levels(emu_lib, c, levels, monkeypatch):
    """batch-affine pair-tree levels (msm_ba.cuh) forced in front of the XYZZ accumulation: small windows make crowded
    buckets (c = 4: ~60 entries per bucket, c = 7: ~8), so pairs, odd leftovers, buckets cut by run boundaries and the
    direct XYZZ stage are all hit; K = 3 makes many short batches.  Edge cases: P + P, P + (-P), infinity bases."""
    monkeypatch.setenv("ALEO_B200_MSM_C", str(c))
    monkeypatch.setenv("ALEO_B200_MSM_BA", str(levels))
    monkeypatch.setenv("ALEO_B200_MSM_BA_K", "3")
    n = 360
    B = o.synthetic_bases(n, 51)
    s = o.random_fr_vec(n, 52)
    s[0], s[1], s[2] = o.R_MOD - 1, 1, 0

    def ok(bases, scalars, stride=104):
        return _msm(emu_lib, bases, scalars, stride) == o.g1_projective_to_bytes(o.msm_pippenger(bases, scalars))

    assert ok(B, s, 104) and ok(B, s, 96)
    assert ok([B[0]] * n, s)                                                     # every pair is a doubling
    assert ok([B[i // 2] if i % 2 == 0 else o.g1_neg(B[i // 2]) for i in range(n)], [s[i // 2] for i in range(n)])  # cancellations
    Binf = [None if i % 5 == 0 else B[i] for i in range(n)]
    assert ok(Binf, s, 104) and ok(Binf, s, 96) and ok([None] * n, s)
    assert ok(B, [1] * n) and ok(B, [1 if i % 4 else 0 for i in range(n)])       # one huge bucket: deep tree, then cut pieces
    assert ok(B[:1], s[:1]) and ok(B[:2], [5, 5])


def test_emu_msm_batch_affine_groups_and_ranges(emu_lib, monkeypatch):
    """a workspace budget of 1 MB splits the bucket sets into several groups; three host point ranges accumulate into
    the same buckets (`into`); the resident SRS (one shared set) and a batch of commitments go through the levels too"""
    monkeypatch.setenv("ALEO_B200_MSM_C", "6")
    monkeypatch.setenv("ALEO_B200_MSM_BA", "2")
    monkeypatch.setenv("ALEO_B200_MSM_BA_K", "5")
    n = 600
    B = o.synthetic_bases(n, 61)
    s = o.random_fr_vec(n, 62)
    want = o.g1_projective_to_bytes(o.msm_pippenger(B, s))
    assert _msm(emu_lib, B, s, 104) == want
    monkeypatch.setenv("ALEO_B200_MSM_BA_MB", "1")     # 1 MB / 100 B = 10 485 entries: 17 windows of 600 entries per group
    assert _msm(emu_lib, B, s, 104) == want
    monkeypatch.setenv("ALEO_B200_TEST_DENY_BA_ALLOC", "1")   # no room for the levels' workspace: the XYZZ kernel alone
    assert _msm(emu_lib, B, s, 104) == want
    monkeypatch.delenv("ALEO_B200_TEST_DENY_BA_ALLOC")
    monkeypatch.setenv("ALEO_B200_MSM_CHUNKS", "3")
    bb = C.create_string_buffer(o.g1_affine_vec_to_bytes(B, 104), n * 104)
    sb = C.create_string_buffer(o.fr_vec_to_bytes(s, mont=False), n * 32)
    out = C.create_string_buffer(144)
    emu_lib.check(emu_lib.msm_g1(C.cast(out, C.c_void_p), C.cast(bb, C.c_void_p), n, C.cast(sb, C.c_void_p), 104), "msm_g1")
    assert out.raw == want


@pytest.mark.parametrize("glv", ["0", "1"])
def test_emu_msm_with_and_without_glv_split(emu_lib, glv, monkeypatch):
    """plain MSMs below the batch-affine threshold split their scalars k = k1 + k2 u^2 and run over 2 n virtual points
    ((beta x, -y) for the second half) with half the windows; ALEO_B200_MSM_GLV = 0 keeps the 253-bit path.  Scalars at
    the split's edges: multiples of u^2, u^2 - 1, r - 1, values just below / above 2^126."""
    monkeypatch.setenv("ALEO_B200_MSM_GLV", glv)
    u2 = 0x8508c00000000001 ** 2
    n = 300
    B = o.synthetic_bases(n, 81)
    s = o.random_fr_vec(n, 82)
    s[:12] = [0, 1, u2 - 1, u2, u2 + 1, 2 * u2, (o.R_MOD - 1) // u2 * u2, o.R_MOD - 1, (1 << 126) - 1, 1 << 126, (1 << 127) - 1, (1 << 252) + 12345]
    want = o.g1_projective_to_bytes(o.msm_pippenger(B, s))
    for stride in (104, 96):
        assert _msm(emu_lib, B, s, stride) == want, stride
    Binf = [None if i % 3 == 0 else B[i] for i in range(n)]
    assert _msm(emu_lib, Binf, s, 104) == o.g1_projective_to_bytes(o.msm_pippenger(Binf, s))
    assert _msm(emu_lib, [B[0]] * n, s, 104) == o.g1_projective_to_bytes(o.msm_pippenger([B[0]] * n, s))
    assert emu_lib.msm_window_bits(1 << 18) == 13 and emu_lib.msm_window_bits(1 << 12) == (8 if glv == "1" else 8)
