"""Parity of the sm_100a NTT (through the C ABI / the EvaluationDomain mirror) with the oracle.
Bit-exact: every output element must equal the oracle's, as 32-byte Montgomery images."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import aleo_b200 as ab
from oracle import bls12_377 as o

pytestmark = pytest.mark.gpu
VARIANTS = [("fft", 0, 0, o.fft), ("ifft", 1, 0, o.ifft), ("coset_fft", 0, 1, o.coset_fft), ("coset_ifft", 1, 1, o.coset_ifft)]


def _dev_tensor(raw):
    import torch
    return torch.from_numpy(np.frombuffer(raw, dtype=np.int64).reshape(-1, 4).copy()).cuda()


def test_library_is_the_cuda_build():
    assert b"sm_100a" in ab.get_lib().version()
    ab.get_lib().check(ab.get_lib().init(0), "init")


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 7, 10, 11, 12, 13, 14, 15, 16, 17])
def test_parity_with_python_oracle(log_n):
    n = 1 << log_n
    v = o.random_fr_vec(n, 900 + log_n)
    raw = o.fr_vec_to_bytes(v)
    dom = ab.EvaluationDomain.new(n)
    for name, _, _, ref in VARIANTS:
        want = o.fr_vec_to_bytes(ref(v))
        assert getattr(dom, name)(raw) == want, (name, "host")            # host-pointer entry point
        t = _dev_tensor(raw)
        getattr(dom, name + "_in_place_dev")(t)
        assert t.cpu().numpy().tobytes() == want, (name, "device")        # device-resident entry point


@pytest.mark.parametrize("log_n", [12, 15, 17])
def test_extreme_values(log_n):
    """the passes keep semi-reduced values (< 2r) and feed untested sums / differences (< 4r) into raw products:
    vectors made of r - 1, r - 2, 0 and 1 push every intermediate towards those bounds (2 and 3 passes, both tiles)"""
    import random
    n = 1 << log_n
    rnd = random.Random(log_n)
    top = o.R_MOD - 1
    dom = ab.EvaluationDomain.new(n)
    for v in ([top] * n, [top if i & 1 else 0 for i in range(n)], [rnd.choice((0, 1, top, top - 1)) for _ in range(n)]):
        raw = o.fr_vec_to_bytes(v)
        for name, _, _, ref in VARIANTS:
            assert getattr(dom, name)(raw) == o.fr_vec_to_bytes(ref(v)), name


def test_golden_vectors(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ntt_golden.json")))
    for key in ("0", "1", "3", "6", "10"):
        dom = ab.EvaluationDomain.new(1 << int(key))
        raw = bytes.fromhex(g[key]["input"])
        for name, _, _, _ in VARIANTS:
            assert getattr(dom, name)(raw).hex() == g[key][name], (key, name)
    # fft_in_place resizes (zero-pads) to the domain first: 5 coefficients in a domain of 8
    dom = ab.EvaluationDomain.new(5)
    assert dom.size == 8
    assert dom.fft(bytes.fromhex(g["pad5to8"]["input"])).hex() == g["pad5to8"]["fft"]


@pytest.mark.parametrize("log_n", [18, 20, 23, 25])
def test_parity_with_c_oracle_large(c_oracle, log_n):
    """sizes the Python oracle cannot reach: 2^18 (3 passes) .. 2^25 (4 passes)"""
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 4242 + log_n, 0, True)
    host = x.cpu().numpy()
    for name, inv, coset, _ in (VARIANTS if log_n <= 20 else VARIANTS[:1] + VARIANTS[3:]):
        y = x.clone()
        getattr(dom, name + "_in_place_dev")(y)
        ref = host.copy()
        assert c_oracle.oracle_ntt_fr(ref.ctypes.data, log_n, inv, coset, os.cpu_count() or 1) == 0
        assert np.array_equal(y.cpu().numpy(), ref), name


@pytest.mark.parametrize("tile,log_n", [(11, 13), (11, 18), (10, 21), (10, 25)])
def test_both_tile_sizes_at_sizes_that_default_to_the_other(c_oracle, monkeypatch, tile, log_n):
    """pass kernels come in two tile sizes (1024 elements up to 2^20, 2048 above): force the one a size would not pick"""
    monkeypatch.setenv("ALEO_B200_NTT_TILE", str(tile))
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 777 + log_n, 0, True)
    host = x.cpu().numpy()
    for name, inv, coset, _ in VARIANTS:
        y = x.clone()
        getattr(dom, name + "_in_place_dev")(y)
        ref = host.copy()
        assert c_oracle.oracle_ntt_fr(ref.ctypes.data, log_n, inv, coset, os.cpu_count() or 1) == 0
        assert np.array_equal(y.cpu().numpy(), ref), name


@pytest.mark.parametrize("log_n", [22, 24, 26])
def test_round_trip_and_linearity_at_full_size(log_n):
    import torch

    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 7, 0, True)
    y = x.clone()
    dom.fft_in_place_dev(y)
    assert not torch.equal(x, y)
    dom.ifft_in_place_dev(y)
    assert torch.equal(x, y)                      # ifft(fft(x)) == x, bit for bit
    dom.coset_fft_in_place_dev(y)
    dom.coset_ifft_in_place_dev(y)
    assert torch.equal(x, y)
    # a delta at index 1 transforms to the powers of omega: X[k] = w^k  (checks the output ORDER)
    z = torch.zeros_like(x)
    one = np.frombuffer(o.int_to_le_bytes(o.fr_to_mont(1), 32), dtype=np.int64)
    z[1] = torch.from_numpy(one.copy()).cuda()
    dom.fft_in_place_dev(z)
    w = o.fr_root_of_unity(log_n)
    zs = z.cpu().numpy()
    for k in (0, 1, 2, 3, 5, n // 2 + 1, n - 1, 12345 % n):
        assert zs[k].tobytes() == o.int_to_le_bytes(o.fr_to_mont(pow(w, k, o.R_MOD)), 32), k


def test_batch_of_transforms():
    log_n, batch = 12, 5
    n = 1 << log_n
    vs = [o.random_fr_vec(n, 60 + b) for b in range(batch)]
    t = _dev_tensor(b"".join(o.fr_vec_to_bytes(v) for v in vs))
    ab.EvaluationDomain.new(n).fft_in_place_dev(t, batch=batch)
    got = t.cpu().numpy().tobytes()
    for b in range(batch):
        assert got[b * n * 32:(b + 1) * n * 32] == o.fr_vec_to_bytes(o.fft(vs[b]))
    # small-kernel batch path (n <= 2^11)
    n = 1 << 6
    vs = [o.random_fr_vec(n, 80 + b) for b in range(7)]
    t = _dev_tensor(b"".join(o.fr_vec_to_bytes(v) for v in vs))
    ab.EvaluationDomain.new(n).coset_ifft_in_place_dev(t, batch=7)
    got = t.cpu().numpy().tobytes()
    for b in range(7):
        assert got[b * n * 32:(b + 1) * n * 32] == o.fr_vec_to_bytes(o.coset_ifft(vs[b]))


def test_error_codes():
    lib = ab.get_lib()
    assert lib.ntt_fr(None, 4, 0, 0) == ab._lib.EINVAL
    buf = C.create_string_buffer(64)
    assert lib.ntt_fr(C.cast(buf, C.c_void_p), 33, 0, 0) == ab._lib.ETOOLARGE
    assert lib.ntt_fr_dev(C.cast(buf, C.c_void_p), 1, 1, 2, 0, None) == ab._lib.EINVAL
    assert ab.EvaluationDomain.new((1 << 47) + 1) is None


def test_twiddle_matrix_step():
    """aleo_b200_ntt_twiddle_dev: data[r][c] *= w_N^((r + row0)(c + col0)), forward and inverse roots"""
    from aleo_b200.dist import CudaBackend

    log_n, rows, cols, row0, col0 = 14, 6, 10, 37, 3
    v = o.random_fr_vec(rows * cols, 31)
    for inverse in (False, True):
        w = o.fr_root_of_unity(log_n)
        if inverse:
            w = pow(w, -1, o.R_MOD)
        t = _dev_tensor(o.fr_vec_to_bytes(v))
        CudaBackend().twiddle(t, log_n, inverse, rows, cols, row0, col0)
        want = [v[r * cols + c] * pow(w, (r + row0) * (c + col0), o.R_MOD) % o.R_MOD for r in range(rows) for c in range(cols)]
        assert t.cpu().numpy().tobytes() == o.fr_vec_to_bytes(want)


@pytest.mark.parametrize("log_n", [16, 17, 24])
def test_four_step_schedule_matches_single_kernel_path(log_n):
    """the multi-GPU schedule (aleo_b200/dist.py) run on one GPU must reproduce the direct transform bit
    for bit -- batched local transforms, twiddle step, transposes and output ordering"""
    import torch

    from aleo_b200 import dist as adist

    n = 1 << log_n
    x = ab.gen_scalars_dev(n, 606 + log_n, 0, True)
    for inverse in (False, True):
        direct = x.clone()
        dom = ab.EvaluationDomain.new(n)
        (dom.ifft_in_place_dev if inverse else dom.fft_in_place_dev)(direct)
        out = adist.ntt_four_step(adist.column_block(x, log_n, 0, 1), log_n, inverse=inverse)
        assert torch.equal(adist.gather_natural([out], log_n), direct)


def _bitrev_perm(log_n):
    n = 1 << log_n
    idx = np.arange(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for b in range(log_n):
        rev |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(log_n - 1 - b)
    return rev.astype(np.int64)


@pytest.mark.parametrize("log_n", [1, 4, 11, 12, 16, 20])
def test_fft_order_variants(log_n):
    """FFTOrder::IO / OI (snarkVM fft_helper_in_place_with_pc & co.): bit-reversed output /
    input around the in-order transform, all four transform kinds, batch of 2 at the small sizes"""
    import torch
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    rev = torch.from_numpy(_bitrev_perm(log_n)).cuda()
    x = ab.gen_scalars_dev(n, 6000 + log_n, 0, True)
    for direction, kind in ((0, 0), (1, 0), (0, 1), (1, 1)):
        ref = dom._run_dev(x.clone(), direction, kind)                                   # II, checked against the oracle above
        io = dom._run_dev_ordered(x.clone(), direction, kind, 1)
        assert torch.equal(io, ref[rev]), ("IO", direction, kind)
        oi = dom._run_dev_ordered(x[rev].contiguous(), direction, kind, 2)
        assert torch.equal(oi, ref), ("OI", direction, kind)
    assert torch.equal(dom.out_order_fft_in_place_with_pc_dev(x.clone(), pc=object()), dom.fft_in_place_dev(x.clone())[rev])
    if log_n <= 12:
        two = torch.cat([x, x.flip(0)]).contiguous()
        want = torch.cat([dom.fft_in_place_dev(x.clone())[rev], dom.fft_in_place_dev(x.flip(0).contiguous())[rev]])
        assert torch.equal(dom._run_dev_ordered(two, 0, 0, 1, batch=2), want)
    assert ab.get_lib().ntt_fr_ordered_dev(x.data_ptr(), log_n, 1, 0, 0, 7, None) == -1


@pytest.mark.parametrize("log_n", [12, 16, 20, 24])
def test_peer_ntt_single_rank_equals_plain_transform(log_n):
    """aleo_b200_ntt_dist_* with world = 1: the same passes, the 'exchange' pass stores into this rank's own receive
    buffer -- must equal the in-place transform (the multi-rank schedule is checked on the emulator and, under
    torchrun, by tools/dist_ntt_check.py and bench.py)"""
    import torch
    from aleo_b200.dist import PeerNTT

    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 8100 + log_n, 0, True)
    p = PeerNTT(log_n)
    try:
        assert p.layout() == (n >> p.log_r_last, 1 << p.log_r_last, n >> p.log_r_first, 1 << p.log_r_first)
        for inverse in (False, True):
            for coset in (False, True):
                want = dom._run_dev(x.clone(), 1 if inverse else 0, 1 if coset else 0)
                for _ in range(2):                     # both receive buffers
                    got = p.transform(x, inverse=inverse, coset=coset)
                    assert torch.equal(got.reshape(-1, 4), want)
    finally:
        p.close()


def test_host_call_on_pageable_memory_is_staged():
    """aleo_b200_ntt_fr on a 64 MB pageable numpy buffer (what a Rust Vec is): both directions of the copy go through
    the pinned staging buffers of the helper threads; result equals the device-resident transform"""
    log_n = 21
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 9100, 0, True)
    host = x.cpu().numpy().copy()                       # pageable
    want = dom.coset_fft_in_place_dev(x.clone()).cpu().numpy()
    dom.ntt_host_buffer(host, 0, 1)
    assert np.array_equal(host, want)
    dom.ntt_host_buffer(host, 1, 1)                     # and back
    assert np.array_equal(host, x.cpu().numpy())


@pytest.mark.parametrize("log_n", [3, 12, 16, 21])
def test_host_pointer_ordered_transform(log_n):
    """aleo_b200_ntt_fr_ordered: upstream's snarkvm_ntt argument list (order included) on HOST memory -- against the
    oracle at small sizes, against the device-resident ordered path above that; 2^21 is a pageable 64 MB buffer (staged)"""
    import torch
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    perm = _bitrev_perm(log_n)
    if log_n <= 12:
        v = o.random_fr_vec(n, 4400 + log_n)
        rev = [v[i] for i in perm]
        assert bytes(dom.out_order_fft_in_place_with_pc(bytearray(o.fr_vec_to_bytes(v)))) == o.fr_vec_to_bytes([o.fft(v)[i] for i in perm])
        assert bytes(dom.ifft_helper_in_place_with_pc(bytearray(o.fr_vec_to_bytes(rev)), 2)) == o.fr_vec_to_bytes(o.ifft(v))
        assert bytes(dom.fft_helper_in_place_with_pc(bytearray(o.fr_vec_to_bytes(v)), 0, pc=object())) == o.fr_vec_to_bytes(o.fft(v))
    x = ab.gen_scalars_dev(n, 4500 + log_n, 0, True)
    for direction, kind, order in ((0, 0, 1), (1, 1, 2), (0, 1, 1), (1, 0, 0)):
        host = x.cpu().numpy().copy()
        dom.ntt_host_buffer(host, direction, kind, order)
        want = dom._run_dev_ordered(x.clone(), direction, kind, order)
        assert np.array_equal(host, want.cpu().numpy()), (direction, kind, order)
    assert ab.get_lib().ntt_fr_ordered(x.cpu().numpy().ctypes.data, log_n, 0, 0, 9) == -1


@pytest.mark.parametrize("log_n", [10, 20])
def test_polymul_matches_the_transform_pipeline(log_n):
    """aleo_b200_polymul[_dev] (snarkvm_polymul's shape) = ifft(prod fft(p_i) * prod e_j): against the oracle at 2^10, against
    the same pipeline built from the already-pinned entry points at 2^20; host form on pageable memory"""
    import torch
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    a = ab.gen_scalars_dev(n // 2, 5100 + log_n, 0, True)
    b = ab.gen_scalars_dev(n // 2 - 5, 5200 + log_n, 0, True)
    e = ab.gen_scalars_dev(n, 5300 + log_n, 0, True)

    def padded(t):
        z = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
        z[: t.shape[0]] = t
        return z

    fa, fb = dom.fft_in_place_dev(padded(a)), dom.fft_in_place_dev(padded(b))
    want = dom.ifft_in_place_dev(ab.Evaluations.mul(ab.Evaluations.mul(fa, fb), e))
    got = ab.PolyMultiplier().add_polynomial(a).add_polynomial(b).add_evaluation(e).multiply_dev(log_n)
    assert torch.equal(got, want)
    host = ab.PolyMultiplier().add_polynomial(a.cpu().numpy().copy()).add_polynomial(b.cpu().numpy().copy()) \
        .add_evaluation(e.cpu().numpy().copy()).multiply(log_n)
    assert host == want.cpu().numpy().tobytes()
    only_e = ab.PolyMultiplier().add_evaluation(e).multiply_dev(log_n)
    assert torch.equal(only_e, dom.ifft_in_place_dev(e.clone()))
    if log_n == 10:
        av = o.fr_vec_from_bytes(a.cpu().numpy().tobytes())
        bv = o.fr_vec_from_bytes(b.cpu().numpy().tobytes())
        prod = [0] * n
        for i, x in enumerate(av):
            for j, y in enumerate(bv):
                prod[i + j] = (prod[i + j] + x * y) % o.R_MOD
        two = ab.PolyMultiplier().add_polynomial(a).add_polynomial(b).multiply_dev(log_n)
        assert two.cpu().numpy().tobytes() == o.fr_vec_to_bytes(prod)
    lib = ab.get_lib()
    assert lib.polymul_dev(got.data_ptr(), 0, None, None, 0, None, None, log_n, None) == -1
