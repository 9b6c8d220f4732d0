"""Parity of the polynomial helpers (distribute_powers, evaluate, KZG witness polynomial, kzg_open) with the oracle:
bit-exact at sizes the oracle reaches, algebraic identities at 2^20."""
import ctypes as C

import numpy as np
import pytest
import torch

import aleo_b200 as ab
from aleo_b200 import poly
from oracle import bls12_377 as o

pytestmark = pytest.mark.gpu


def _dev(vals):
    return torch.from_numpy(np.frombuffer(o.fr_vec_to_bytes(vals), dtype=np.int64).reshape(-1, 4).copy()).cuda()


def _host(t):
    return o.fr_vec_from_bytes(t.cpu().numpy().tobytes())


@pytest.mark.parametrize("n", [1, 2, 31, 256, 2047, 2049, 8192, 8193, 70000])
def test_helpers_match_oracle(n):
    c = o.random_fr_vec(n, 1700 + n)
    g, k, z = o.random_fr_vec(3, 1800 + n)
    assert _host(poly.distribute_powers_dev(_dev(c), g, k)) == o.distribute_powers(c, g, k)
    assert _host(poly.distribute_powers_dev(_dev(c), g)) == o.distribute_powers(c, g)
    assert _host(poly.distribute_powers_dev(_dev(c), 22)) == o.distribute_powers(c, 22)
    xs = o.random_fr_vec(n, 1900 + n)
    assert _host(poly.axpy_dev(_dev(c), _dev(xs), k)) == [(y + k * x) % o.R_MOD for y, x in zip(c, xs)]
    assert poly.evaluate_dev(_dev(c), z) == o.poly_eval(c, z)
    assert poly.evaluate_dev(_dev(c), 0) == c[0]
    for zz in (z, 0, 1, o.R_MOD - 1):
        assert _host(poly.divide_by_linear_dev(_dev(c), zz)) == o.divide_by_linear(c, zz), zz


def test_coset_fft_is_distribute_powers_then_fft():
    """EvaluationDomain::coset_fft = distribute_powers(g = 22) followed by fft"""
    n = 1 << 14
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 321, 0, True)
    a = dom.coset_fft_in_place_dev(x.clone())
    b = dom.fft_in_place_dev(poly.distribute_powers_dev(x.clone(), 22))
    assert torch.equal(a, b)


def test_witness_polynomial_identity_at_scale():
    """p(x) - p(z) == q(x) (x - z) at a random x, 2^20 + 5 coefficients"""
    n = (1 << 20) + 5
    p = ab.gen_scalars_dev(n, 99, 0, True)
    z, x = o.random_fr_vec(2, 4455)
    q = poly.divide_by_linear_dev(p, z)
    assert _host(q[-1:]) == [0]
    lhs = (poly.evaluate_dev(p, x) - poly.evaluate_dev(p, z)) % o.R_MOD
    assert lhs == poly.evaluate_dev(q, x) * (x - z) % o.R_MOD


def test_kzg_open_matches_oracle():
    n = 500
    B = o.synthetic_bases(n, 191)
    srs = ab.ResidentSRS.from_host(o.g1_affine_vec_to_bytes(B, 104), 104)
    try:
        c = o.random_fr_vec(n, 192)
        z = o.random_fr_vec(1, 193)[0]
        out = torch.empty(48, dtype=torch.uint8, device="cuda")
        lib = ab.get_lib()
        lib.check(lib.kzg_open_dev(srs._h, out.data_ptr(), _dev(c).data_ptr(), n, o.int_to_le_bytes(o.fr_to_mont(z), 32), None), "open")
        torch.cuda.synchronize()
        q = o.divide_by_linear(c, z)
        assert out.cpu().numpy().tobytes() == o.g1_compress(o.msm_pippenger(B[:n - 1], q[:n - 1]))
    finally:
        srs.close()


@pytest.mark.parametrize("log_n", [0, 4, 13, 18])
def test_lagrange_coefficients(log_n):
    """EvaluationDomain::evaluate_all_lagrange_coefficients: against the oracle at small sizes, by identities at 2^18
    (sum_i L_i(tau) = 1 and sum_i L_i(tau) p(w^i) = p(tau) for a polynomial of degree < n)"""
    n = 1 << log_n
    tau = o.random_fr_vec(1, 8800 + log_n)[0]
    L = poly.lagrange_coeffs_dev(log_n, tau)
    if log_n <= 13:
        assert _host(L) == o.lagrange_coefficients(log_n, tau)
        w = o.fr_root_of_unity(log_n)
        assert _host(poly.lagrange_coeffs_dev(log_n, pow(w, 3 % n, o.R_MOD))) == o.lagrange_coefficients(log_n, pow(w, 3 % n, o.R_MOD))
    else:
        ones = _dev([1])          # <L, 1> via evaluate at z = 1
        assert poly.evaluate_dev(L, 1) == 1
        p = ab.gen_scalars_dev(n, 8900, 0, True)
        evals = ab.EvaluationDomain.new(n).fft_in_place_dev(p.clone())
        prod = ab.Evaluations.mul(L, evals)
        assert poly.evaluate_dev(prod, 1) == poly.evaluate_dev(p, tau)


@pytest.mark.parametrize("log_m,log_n", [(0, 0), (5, 5), (6, 2), (10, 7), (12, 0), (13, 12)])
def test_divide_by_vanishing_on_coset_matches_oracle(log_m, log_n):
    e = o.random_fr_vec(1 << log_m, 9100 + 16 * log_m + log_n)
    assert _host(poly.divide_by_vanishing_on_coset_dev(_dev(e), log_n)) == o.divide_by_vanishing_on_coset(e, log_n)
    g = o.random_fr_vec(1, 9200 + log_m)[0]
    assert _host(poly.divide_by_vanishing_on_coset_dev(_dev(e), log_n, g)) == o.divide_by_vanishing_on_coset(e, log_n, g)


def test_quotient_by_vanishing_polynomial_round_trip_at_scale():
    """p = q * (x^n - 1) with deg q < m - n: coset_fft(p) / Z_n on the coset, coset_ifft -> q (the shape of the prover's
    quotient polynomials: constraint domain n = 2^18 inside a multiplication domain m = 2^20)"""
    log_m, log_n = 20, 18
    m, n = 1 << log_m, 1 << log_n
    q = ab.gen_scalars_dev(m - n, 9300, 0, True)
    p = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
    p[n:] = q                                                    # q * x^n ...
    p[:m - n] = ab.Evaluations.sub(p[:m - n].contiguous(), q)    # ... - q
    dom = ab.EvaluationDomain.new(m)
    e = dom.coset_fft_in_place_dev(p)
    poly.divide_by_vanishing_on_coset_dev(e, log_n)
    back = dom.coset_ifft_in_place_dev(e)
    assert torch.equal(back[:m - n], q)
    assert not back[m - n:].any()


def test_kzg_open_and_commit_lagrange_wrappers():
    """KZG10.open_dev is the C call of test_kzg_open_matches_oracle; commit_lagrange over a handle built from the
    Lagrange basis commits the evaluations to the same point as commit over the monomial basis commits the coefficients
    (here with bases whose discrete logs are powers of a known tau, so that both bases can be built)"""
    log_n = 6
    n = 1 << log_n
    tau = o.random_fr_vec(1, 9400)[0]
    G = o.G1_GEN
    powers = [o.g1_mul(G, pow(tau, i, o.R_MOD)) for i in range(n)]
    lag = [o.g1_mul(G, v) for v in o.lagrange_coefficients(log_n, tau)]
    srs = ab.ResidentSRS.from_host(o.g1_affine_vec_to_bytes(powers, 104), 104)
    srs_l = ab.ResidentSRS.from_host(o.g1_affine_vec_to_bytes(lag, 104), 104)
    try:
        c = o.random_fr_vec(n, 9401)
        evals = o.fft(c, n)
        a = ab.KZG10.commit_dev(srs, _dev(c), n).cpu().numpy().tobytes()
        b = ab.KZG10.commit_lagrange_dev(srs_l, _dev(evals), n).cpu().numpy().tobytes()
        assert a == b == o.g1_compress(o.g1_mul(G, o.poly_eval(c, tau)))
        z = o.random_fr_vec(1, 9402)[0]
        w = ab.KZG10.open_dev(srs, _dev(c), z).cpu().numpy().tobytes()
        assert w == o.g1_compress(o.g1_mul(G, o.poly_eval(o.divide_by_linear(c, z), tau)))
    finally:
        srs.close()
        srs_l.close()


def test_kzg_open_combinations_equals_single_openings():
    """aleo_b200_kzg_open_combinations_dev (SonicKZG10::open_combinations / batch_open shape): m openings of linear
    combinations in one launch sequence == KZG10::open of each combined polynomial; first against the oracle at n = 300,
    then at proof scale (2^16-point SRS, 6 polynomials, 5 openings) against the single-opening path"""
    from aleo_b200.kzg import KZG10
    n = 300
    B = o.synthetic_bases(n, 291)
    srs = ab.ResidentSRS.from_host(o.g1_affine_vec_to_bytes(B, 104), 104)
    try:
        polys = [o.random_fr_vec(ln, 700 + ln) for ln in (300, 128, 1)]
        lc = [[3, 0, 1], [1, o.R_MOD - 2, 0], [0, 0, 5]]
        zs = o.random_fr_vec(3, 78)
        out = KZG10.open_combinations_dev(srs, [_dev(p) for p in polys], lc, zs).cpu().numpy().tobytes()
        for k in range(3):
            comb = [0] * n
            for a, p in zip(lc[k], polys):
                for i, x in enumerate(p):
                    comb[i] = (comb[i] + a * x) % o.R_MOD
            ln = max([len(p) for a, p in zip(lc[k], polys) if a] + [0])
            q = o.divide_by_linear(comb[:ln], zs[k]) if ln > 1 else []
            assert out[48 * k:48 * k + 48] == o.g1_compress(o.msm_pippenger(B[:len(q) - 1], q[:len(q) - 1]) if len(q) > 1 else None), k
    finally:
        srs.close()
    n = 1 << 16
    s0, d = o.base_dlogs(n, 3131)
    srs = ab.ResidentSRS.from_device(ab.gen_bases_dev(n, s0, d, 0, 104), n, 104)
    try:
        polys = [ab.gen_scalars_dev(ln, 900 + i, 0, True) for i, ln in enumerate((n, n, n // 2, n // 2 + 7, 4096, 1))]
        lc = [[1, 7, 0, 0, 0, 0], [0, 0, 1, 11, 13, 17], [5, 4, 3, 2, 1, 1], [0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 3]]
        zs = o.random_fr_vec(5, 79)
        got = KZG10.open_combinations_dev(srs, polys, lc, zs)
        for k in range(5):
            comb = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
            for a, p in zip(lc[k], polys):
                if a:
                    poly.axpy_dev(comb[: p.shape[0]], p, a)
            ln = max(p.shape[0] for a, p in zip(lc[k], polys) if a)
            assert torch.equal(got[k], KZG10.open_dev(srs, comb[:ln].contiguous(), zs[k])), k
    finally:
        srs.close()
