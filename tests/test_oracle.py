"""The Python oracle against (a) the parameters re-derived from the BLS12 family parameter, (b) the
one proof fixture the reference pins (wasm/src/programs/transaction.rs:100), (c) mathematical
definitions (Horner evaluation, naive sum of scalar multiples), (d) the committed golden vectors."""
import json
import os

import pytest

from oracle import bls12_377 as o


def test_parameters_match_survey_appendix_a():
    x = 0x8508C00000000001
    assert o.R_MOD == x**4 - x**2 + 1 == 8444461749428370424248824938781546531375899335154063827935233455917409239041
    assert o.P_MOD == (x - 1) ** 2 * o.R_MOD // 3 + x
    assert o.FR_INV64 == 0x0A117FFFFFFFFFFF and o.FQ_INV64 == 0x8508BFFFFFFFFFFF
    limbs = [0x7D1C7FFFFFFFFFF3, 0x7257F50F6FFFFFF2, 0x16D81575512C0FEE, 0x0D4BDA322BBB9A9D]
    assert o.FR_R == sum(l << (64 * i) for i, l in enumerate(limbs))
    qlimbs = [0x02CDFFFFFFFFFF68, 0x51409F837FFFFFB1, 0x9F7DB3A98A7D3FF2, 0x7B4E97B76E7C6305, 0x4CF495BF803C84E8, 0x008D6661E2FDF49A]
    assert o.FQ_R == sum(l << (64 * i) for i, l in enumerate(qlimbs))
    assert (o.R_MOD - 1) % (1 << 47) == 0 and (o.R_MOD - 1) % (1 << 48) != 0
    assert o.FR_TWO_ADIC_ROOT == 8065159656716812877374967518403273466521432693661810619979959746626482506078
    assert pow(o.FR_TWO_ADIC_ROOT, 1 << 47, o.R_MOD) == 1 and pow(o.FR_TWO_ADIC_ROOT, 1 << 46, o.R_MOD) != 1
    assert o.fr_root_of_unity(16) == 2952730183556248050906521597191761902300071509314809436739482176595887578715
    assert o.fr_root_of_unity(24) == 5421008228431423068756151873766382091821790348027112600470148769114500057419
    assert o.g1_is_on_curve(o.G1_GEN) and o.g1_mul(o.G1_GEN, o.R_MOD) is None
    assert o.G1_COFACTOR == 30631250834960419227450344600217059328


def test_reference_proof_fixture_decodes_to_subgroup_points(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "proof_fixture.json")))
    hrp, payload = o.bech32m_decode(fx["proof"])
    assert hrp == "proof" and len(payload) == fx["payload_len"] == 901
    assert payload[0] == 0                                       # version
    assert payload[1:17] == (1).to_bytes(8, "little") * 2        # batch_sizes = [1]
    assert payload[161] == 1                                     # mask_poly: Some
    for off in fx["g1_offsets"]:
        chunk = payload[off:off + 48]
        pt = o.g1_decompress(chunk)
        assert pt is not None and o.g1_is_on_curve(pt)
        assert o.g1_mul(pt, o.R_MOD) is None                     # r-torsion
        assert o.g1_compress(pt) == chunk                        # wire format round trip
    for off in fx["fr_offsets"]:
        assert int.from_bytes(payload[off:off + 32], "little") < o.R_MOD


def test_reference_proof_fixture_is_unchanged_in_reference(golden_dir):
    ref = "/root/reference/wasm/src/programs/transaction.rs"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present (GPU box)")
    fx = json.load(open(os.path.join(golden_dir, "proof_fixture.json")))
    assert fx["proof"] in open(ref).read()


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_ntt_is_polynomial_evaluation(log_n):
    n = 1 << log_n
    v = o.random_fr_vec(n, 10 + log_n)
    w = o.fr_root_of_unity(log_n)
    ev = o.fft(v)
    cev = o.coset_fft(v)
    for k in range(0, n, max(1, n // 16)):
        assert ev[k] == o.poly_eval(v, pow(w, k, o.R_MOD))
        assert cev[k] == o.poly_eval(v, o.FR_GENERATOR * pow(w, k, o.R_MOD) % o.R_MOD)
    assert o.ifft(ev) == v and o.coset_ifft(cev) == v


def test_fft_zero_pads_and_truncates_like_resize():
    v = o.random_fr_vec(5, 3)
    assert o.fft(v) == o.fft(v + [0, 0, 0])
    assert o.fft(v + [7] * 10, 8) == o.fft((v + [7] * 10)[:8])
    with pytest.raises(ValueError):
        o.domain_size((1 << 47) + 1)


def test_msm_definitions_agree():
    for n in (1, 2, 17, 64):
        B = o.synthetic_bases(n, 100 + n)
        s = o.random_fr_vec(n, 200 + n)
        want = o.msm_naive(B, s)
        assert want == o.msm_pippenger(B, s, c=4) == o.msm_pippenger(B, s, c=7) == o.msm_expected_from_dlogs(n, 100 + n, s)
    assert o.msm_naive([], []) is None
    B = o.synthetic_bases(4, 9)
    assert o.msm_naive(B, [0, 0, 0, 0]) is None
    assert o.msm_naive(B, [1, 2]) == o.msm_naive(B[:2], [1, 2, 3, 4][:2])      # zip semantics
    assert o.msm_naive([B[0], o.g1_neg(B[0])], [5, 5]) is None


def test_memory_images_round_trip():
    B = o.synthetic_bases(3, 1) + [None]
    for stride in (104, 96):
        raw = o.g1_affine_vec_to_bytes(B, stride)
        assert len(raw) == 4 * stride
        assert [o.g1_affine_from_bytes(raw[i * stride:(i + 1) * stride], stride) for i in range(4)] == B
    assert o.g1_affine_to_bytes(None, 104)[96] == 1
    for pt in B:
        assert o.g1_projective_from_bytes(o.g1_projective_to_bytes(pt)) == pt
    v = o.random_fr_vec(4, 2)
    assert o.fr_vec_from_bytes(o.fr_vec_to_bytes(v)) == v
    assert o.fr_vec_from_bytes(o.fr_vec_to_bytes(v, mont=False), mont=False) == v


def test_golden_vectors_match_oracle(golden_dir):
    ntt = json.load(open(os.path.join(golden_dir, "ntt_golden.json")))
    for key in ("0", "3", "6"):
        v = o.fr_vec_from_bytes(bytes.fromhex(ntt[key]["input"]))
        assert o.fr_vec_to_bytes(o.fft(v)).hex() == ntt[key]["fft"]
        assert o.fr_vec_to_bytes(o.coset_ifft(v)).hex() == ntt[key]["coset_ifft"]
    msm = json.load(open(os.path.join(golden_dir, "msm_golden.json")))
    for name in ("uniform64", "edge_infinity_cancel_repeat", "all_r_minus_1"):
        g = msm[name]
        raw = bytes.fromhex(g["bases104"])
        B = [o.g1_affine_from_bytes(raw[i * 104:(i + 1) * 104]) for i in range(g["n"])]
        s = o.fr_vec_from_bytes(bytes.fromhex(g["scalars"]), mont=False)
        assert o.g1_projective_to_bytes(o.msm_pippenger(B, s)).hex() == g["result"]
