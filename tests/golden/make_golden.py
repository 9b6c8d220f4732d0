#!/usr/bin/env python3
"""Regenerates the golden fixtures in this directory from the Python big-integer oracle
(oracle/bls12_377.py).  The reference repository holds no MSM / NTT vectors, so these
are oracle outputs on seeded inputs; the one reference-pinned fixture is proof_fixture.json, the
proof string of /root/reference/wasm/src/programs/transaction.rs:100 with its decoded layout.
Run from the repo root:  python tests/golden/make_golden.py [/root/reference]"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bls12_377 as o  # noqa: E402


def hexs(b):
    return b.hex()


def main():
    ntt = {}
    for log_n in (0, 1, 3, 6, 10):
        n = 1 << log_n
        v = o.random_fr_vec(n, 0xA1E0B200 + log_n)
        ntt[str(log_n)] = {
            "input": hexs(o.fr_vec_to_bytes(v)),
            "fft": hexs(o.fr_vec_to_bytes(o.fft(v))),
            "ifft": hexs(o.fr_vec_to_bytes(o.ifft(v))),
            "coset_fft": hexs(o.fr_vec_to_bytes(o.coset_fft(v))),
            "coset_ifft": hexs(o.fr_vec_to_bytes(o.coset_ifft(v))),
        }
    # zero-padding semantics: 5 coefficients in a domain of 8
    v5 = o.random_fr_vec(5, 77)
    ntt["pad5to8"] = {"input": hexs(o.fr_vec_to_bytes(v5)), "fft": hexs(o.fr_vec_to_bytes(o.fft(v5, 8)))}
    json.dump(ntt, open(os.path.join(HERE, "ntt_golden.json"), "w"))

    msm = {}
    R = o.R_MOD
    for name, n, scal in (
        ("uniform64", 64, None),
        ("uniform257", 257, None),
        ("witness_like", 96, lambda i, s: 0 if i % 2 == 0 else (1 if i % 4 == 1 else s)),
        ("all_r_minus_1", 33, lambda i, s: R - 1),
        ("single", 1, None),
    ):
        seed = 0xB200 + n
        bases = o.synthetic_bases(n, seed)
        s = o.random_fr_vec(n, seed + 1)
        if scal:
            s = [scal(i, s[i]) for i in range(n)]
        naive = o.msm_naive(bases, s)
        assert naive == o.msm_pippenger(bases, s) == o.msm_expected_from_dlogs(n, seed, s)
        msm[name] = {
            "n": n, "seed": seed,
            "bases104": hexs(o.g1_affine_vec_to_bytes(bases, 104)),
            "scalars": hexs(o.fr_vec_to_bytes(s, mont=False)),
            "result": hexs(o.g1_projective_to_bytes(naive)),
        }
    # identity-heavy case: infinity bases and cancelling pairs
    n = 40
    bases = o.synthetic_bases(n, 0xC0DE)
    s = o.random_fr_vec(n, 0xC0DF)
    for i in range(0, n, 5):
        bases[i] = None
    bases[1], s[1] = o.g1_neg(bases[2]), s[2]          # P and -P with the same scalar cancel
    bases[3], bases[4] = bases[6], bases[6]            # repeated point
    msm["edge_infinity_cancel_repeat"] = {
        "n": n, "seed": 0xC0DE,
        "bases104": hexs(o.g1_affine_vec_to_bytes(bases, 104)),
        "scalars": hexs(o.fr_vec_to_bytes(s, mont=False)),
        "result": hexs(o.g1_projective_to_bytes(o.msm_naive(bases, s))),
    }
    json.dump(msm, open(os.path.join(HERE, "msm_golden.json"), "w"))

    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = open(os.path.join(ref, "wasm/src/programs/transaction.rs")).read()
    proof = re.search(r"proof1[0-9a-z]+", src).group(0)
    hrp, payload = o.bech32m_decode(proof)
    g1_offsets = [17, 65, 113, 162] + [210 + 48 * i for i in range(6)] + [770, 851]
    fr_offsets = [498 + 32 * i for i in range(5)] + [666 + 32 * i for i in range(3)] + [819]
    json.dump({"source": "wasm/src/programs/transaction.rs:100", "proof": proof, "hrp": hrp, "payload_len": len(payload),
               "g1_offsets": g1_offsets, "fr_offsets": fr_offsets}, open(os.path.join(HERE, "proof_fixture.json"), "w"))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
