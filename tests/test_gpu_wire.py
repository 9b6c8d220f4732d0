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
