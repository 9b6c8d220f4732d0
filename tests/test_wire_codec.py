"""Product-side Proof codec (aleo_b200/wire.py) against the REFERENCE's own proof string
(wasm/src/programs/transaction.rs:100 -> tests/golden/proof_fixture.json): decode -> re-encode must be byte-identical
(1454 characters, 901 bytes).  No GPU needed; the device half (batch decompress / compress of the 12 commitments)
is tests/test_gpu_wire.py."""
import json
import os

import pytest

from aleo_b200 import wire
from oracle import bls12_377 as o          # checker only


@pytest.fixture(scope="module")
def fx(golden_dir):
    return json.load(open(os.path.join(golden_dir, "proof_fixture.json")))


def test_bech32m_round_trip_matches_the_oracle_decoder(fx):
    hrp, payload = wire.bech32m_decode(fx["proof"])
    assert (hrp, payload) == o.bech32m_decode(fx["proof"])
    assert hrp == "proof" and len(payload) == fx["payload_len"] == 901
    assert wire.bech32m_encode(hrp, payload) == fx["proof"] and len(fx["proof"]) == 1454
    for n in range(0, 70):                       # every padding length
        blob = bytes((7 * i + n) & 0xFF for i in range(n))
        assert wire.bech32m_decode(wire.bech32m_encode("proof", blob)) == ("proof", blob)


def test_bech32m_rejects_corruption(fx):
    s = fx["proof"]
    bad = s[:100] + ("q" if s[100] != "q" else "p") + s[101:]
    with pytest.raises(ValueError):
        wire.bech32m_decode(bad)
    with pytest.raises(ValueError):
        wire.bech32m_decode(s[:50] + s[50].upper() + s[51:])
    with pytest.raises(ValueError):
        wire.Proof.from_str("proof1" + s[6:-1])


def test_reference_proof_string_decodes_and_reencodes_byte_identically(fx):
    p = wire.Proof.from_str(fx["proof"])
    _, payload = o.bech32m_decode(fx["proof"])
    assert p.version == 0 and p.batch_sizes == [1]
    assert p.to_bytes() == payload
    assert p.to_string() == fx["proof"] and str(p) == fx["proof"]
    # the group elements sit where SURVEY App. B / the fixture say, in serialisation order
    assert p.commitments() == [bytes(payload[off:off + 48]) for off in fx["g1_offsets"]]
    assert len(p.commitments()) == 12 and p.mask_poly is not None and len(p.pc_proofs) == 2
    assert p.pc_proofs[0].random_v is not None and p.pc_proofs[1].random_v is None and p.pc_evaluations is None
    # every Fr of the layout is canonical (< r) and sits at the fixture's offsets
    frs = p.z_b_evals + [p.g_1_eval] + p.g_a_evals + p.g_b_evals + p.g_c_evals + p.sums + [p.pc_proofs[0].random_v]
    assert frs == [bytes(payload[off:off + 32]) for off in fx["fr_offsets"]]
    assert all(int.from_bytes(f, "little") < o.R_MOD for f in frs)
    # with_commitments is the inverse of commitments
    assert p.with_commitments(p.commitments()).to_string() == fx["proof"]
    swapped = p.with_commitments(list(reversed(p.commitments())))
    assert swapped.to_string() != fx["proof"] and swapped.with_commitments(p.commitments()).to_bytes() == payload


def test_truncated_and_padded_payloads_are_rejected(fx):
    _, payload = o.bech32m_decode(fx["proof"])
    with pytest.raises(ValueError):
        wire.Proof.from_bytes(payload[:-1])
    with pytest.raises(ValueError):
        wire.Proof.from_bytes(payload + b"\x00")
    with pytest.raises(ValueError):
        wire.Proof.from_bytes(payload[:161] + b"\x02" + payload[162:])     # Option tag of mask_poly


def test_multi_circuit_layout_round_trips():
    """[U] generalisation (several circuits / instances): self-consistency only -- the reference pins batch_sizes = [1]"""
    g = lambda i: bytes([i]) * 48   # noqa: E731
    f = lambda i: bytes([i]) * 32   # noqa: E731
    p = wire.Proof(0, [2, 1], [g(i) for i in range(9)], None, g(20), g(21), [g(22), g(23)], [g(24), g(25)], [g(26), g(27)], g(28),
                   [f(1), f(2), f(3)], f(4), [f(5), f(6)], [f(7), f(8)], [f(9), f(10)], [f(i) for i in range(11, 17)],
                   [wire.KZGProof(g(30), None), wire.KZGProof(g(31), f(40))], [f(50)])
    q = wire.Proof.from_bytes(p.to_bytes())
    assert q == p and wire.Proof.from_str(p.to_string()) == p
    assert len(p.commitments()) == 9 + 2 + 6 + 1 + 2
