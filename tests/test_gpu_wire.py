"""Compressed G1 wire format on the GPU, pinned by the REFERENCE's own data: the proof string of
wasm/src/programs/transaction.rs:100 (tests/golden/proof_fixture.json) holds 12 KZG commitments."""
import json
import os

import numpy as np
import pytest
import torch

import aleo_b200 as ab
from aleo_b200 import wire
from oracle import bls12_377 as o

pytestmark = pytest.mark.gpu


def test_reference_proof_commitments_decompress_and_recompress(golden_dir):
    fx = json.load(open(os.path.join(golden_dir, "proof_fixture.json")))
    _, payload = o.bech32m_decode(fx["proof"])
    chunks = [bytes(payload[off:off + 48]) for off in fx["g1_offsets"]]
    blob = b"".join(chunks)
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    for stride in (104, 96):
        aff, bad = wire.g1_decompress_dev(t, stride)
        assert bad == 0
        assert aff.cpu().numpy().tobytes() == o.g1_affine_vec_to_bytes([o.g1_decompress(c) for c in chunks], stride)
        assert ab.check_on_curve_dev(aff, len(chunks), stride)
        assert wire.g1_compress_dev(aff, len(chunks), stride).cpu().numpy().tobytes() == blob


def test_round_trip_at_scale_and_invalid_encodings():
    n = 50000
    s0, d = o.base_dlogs(n, 909)
    bases = ab.gen_bases_dev(n, s0, d, 0, 104)
    comp = wire.g1_compress_dev(bases, n, 104)
    for i in (0, 1, n - 1):          # spot-check the compressed bytes against the oracle
        assert comp[48 * i:48 * i + 48].cpu().numpy().tobytes() == o.g1_compress(o.g1_mul(o.G1_GEN, (s0 + i * d) % o.R_MOD))
    back, bad = wire.g1_decompress_dev(comp, 104)
    assert bad == 0 and torch.equal(back, bases[: n * 104])
    # flip the sign flag of every point: still valid, decompresses to -P
    neg = comp.clone().view(-1, 48)
    neg[:, 47] ^= 0x80
    nb, bad = wire.g1_decompress_dev(neg.reshape(-1), 104)
    assert bad == 0 and not torch.equal(nb, bases[: n * 104])
    assert torch.equal(wire.g1_compress_dev(nb, n, 104).view(-1, 48)[:, :47], comp.view(-1, 48)[:, :47])
    # identity and invalid encodings
    bad_x = next(x for x in range(2, 50) if o.fq_sqrt((x ** 3 + 1) % o.P_MOD) is None)
    blob = o.g1_compress(None) + o.int_to_le_bytes(bad_x, 48) + o.int_to_le_bytes(o.P_MOD + 5, 48) + o.g1_compress(o.G1_GEN)
    aff, bad = wire.g1_decompress_dev(torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(), 96)
    assert bad == 2
    assert aff.cpu().numpy().tobytes() == o.g1_affine_vec_to_bytes([None, None, None, o.G1_GEN], 96)


def test_reference_proof_string_through_the_device_codec(golden_dir):
    """Proof::from_str -> all 12 group elements decompressed on the device in one batch (curve + subgroup validated) ->
    compressed on the device -> Proof::to_string: byte-identical to the reference's 1454-character string"""
    fx = json.load(open(os.path.join(golden_dir, "proof_fixture.json")))
    p = wire.Proof.from_str(fx["proof"])
    for stride in (104, 96):
        aff, bad = p.commitments_affine_dev(stride)
        assert bad == 0
        assert aff.cpu().numpy().tobytes() == o.g1_affine_vec_to_bytes([o.g1_decompress(c) for c in p.commitments()], stride)
        q = p.with_commitments_from_affine_dev(aff, stride)
        assert q.to_string() == fx["proof"] and len(q.to_string()) == 1454
    # a commitment swapped for an honest different point changes the string and survives the round trip
    other = wire.g1_compress_dev(ab.gen_bases_dev(1, 5, 1, 0, 104), 1, 104).cpu().numpy().tobytes()
    comms = p.commitments()
    comms[3] = other
    p2 = p.with_commitments(comms)
    aff, bad = p2.commitments_affine_dev(104)
    assert bad == 0 and p2.with_commitments_from_affine_dev(aff, 104).to_string() == p2.to_string() != fx["proof"]


def test_subgroup_check_rejects_curve_points_outside_g1():
    """deserialize_compressed (Validate::Yes) runs is_in_correct_subgroup_assuming_on_curve: BLS12-377's G1 cofactor is
    large, so points on y^2 = x^3 + 1 outside the r-torsion exist and must be counted invalid; the unchecked variant
    (Validate::No) lets them through"""
    rogue = o.curve_points_outside_subgroup(5)
    honest = [o.g1_mul(o.G1_GEN, k) for k in (1, 7, o.R_MOD - 1)]
    blob = b"".join(o.g1_compress(pt) for pt in rogue[:2] + honest + rogue[2:])
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    aff, bad = wire.g1_decompress_dev(t, 104)
    assert bad == 5
    assert aff.cpu().numpy().tobytes() == o.g1_affine_vec_to_bytes([None, None] + honest + [None] * 3, 104)
    aff, bad = wire.g1_decompress_dev(t, 104, validate=False)
    assert bad == 0
    assert aff.cpu().numpy().tobytes() == o.g1_affine_vec_to_bytes(rogue[:2] + honest + rogue[2:], 104)
    # at scale every honest point passes: 20000 multiples of the generator
    n = 20000
    s0, d = o.base_dlogs(n, 4242)
    bases = ab.gen_bases_dev(n, s0, d, 0, 96)
    back, bad = wire.g1_decompress_dev(wire.g1_compress_dev(bases, n, 96), 96)
    assert bad == 0 and torch.equal(back, bases[: n * 96])
