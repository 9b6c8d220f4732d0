"""Wire formats of the proving path, product side.

* Compressed G1 on the device (include/aleo_b200.h ``aleo_b200_g1_{de,}compress_dev``): snarkVM's CanonicalSerialize /
  CanonicalDeserialize of ``G1Affine`` -- the 48-byte points inside proofs and key files.
* ``Proof``: the ``proof1...`` string of an Aleo transition / fee (bech32m, hrp ``proof``) and the byte layout under it
  (decoded from the reference's own fixture wasm/src/programs/transaction.rs:100): what
  ``Proof::<N>::from_str`` / ``to_string`` do upstream (snarkvm-synthesizer-snark 0.14.5 ``Proof`` = version byte +
  snarkvm-algorithms ``snark::marlin::Proof``).  The commitments of a proof go through the device codec in one batch
  (``Proof.commitments_affine_dev`` / ``Proof.with_commitments_from_affine_dev``).

The single-circuit layout (batch_sizes = [1]) is pinned byte for byte by the reference fixture; the generalisation to
several circuits / instances follows the field order of that struct and is marked [U] in SURVEY.md App. B.
``.prover`` / ``.usrs`` key-file readers are deliberately absent: the files cannot be fetched here.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import List, Optional

from . import _lib

G1_BYTES = 48
FR_BYTES = 32


def g1_decompress_dev(compressed_t, affine_stride: int = 104, validate: bool = True):
    """(n * 48) uint8 CUDA tensor -> ((n * stride) uint8 CUDA tensor of affine points, number of invalid encodings).
    validate=True is deserialize_compressed (curve AND prime-order-subgroup check); False is
    deserialize_compressed_unchecked (no subgroup check: trusted SRS / key files)."""
    import torch

    lib = _lib.get_lib()
    n = compressed_t.numel() * compressed_t.element_size() // G1_BYTES
    out = torch.empty(max(n, 1) * affine_stride, dtype=torch.uint8, device=compressed_t.device)
    fn, name = (lib.g1_decompress_dev, "aleo_b200_g1_decompress_dev") if validate else \
        (lib.g1_decompress_unchecked_dev, "aleo_b200_g1_decompress_unchecked_dev")
    with torch.cuda.device(compressed_t.device):
        bad = lib.check(fn(out.data_ptr(), affine_stride, compressed_t.data_ptr(), n, torch.cuda.current_stream().cuda_stream), name)
    return out[: n * affine_stride], bad


def g1_compress_dev(affine_t, n: int, affine_stride: int = 104):
    """n affine points (stride 104 / 96) -> (n * 48) uint8 CUDA tensor"""
    import torch

    lib = _lib.get_lib()
    out = torch.empty(max(n, 1) * G1_BYTES, dtype=torch.uint8, device=affine_t.device)
    with torch.cuda.device(affine_t.device):
        lib.check(lib.g1_compress_dev(out.data_ptr(), affine_t.data_ptr(), affine_stride, n,
                                      torch.cuda.current_stream().cuda_stream), "aleo_b200_g1_compress_dev")
    return out[: n * G1_BYTES]


# ---- bech32m (BIP-350), as the bech32 crate does for snarkVM's string types ------------------------------------------
_CHARSET = "qpzry9x8gf2tvdw0s3jn54khce6mua7l"
_BECH32M = 0x2BC830A3
_GEN = (0x3B6A57B2, 0x26508E6D, 0x1EA119FA, 0x3D4233DD, 0x2A1462B3)


def _polymod(values) -> int:
    chk = 1
    for v in values:
        top = chk >> 25
        chk = ((chk & 0x1FFFFFF) << 5) ^ v
        for i, g in enumerate(_GEN):
            if (top >> i) & 1:
                chk ^= g
    return chk


def _hrp_expand(hrp: str):
    return [ord(c) >> 5 for c in hrp] + [0] + [ord(c) & 31 for c in hrp]


def bech32m_decode(s: str):
    """-> (hrp, payload bytes); ValueError on a bad character, mixed case, checksum or padding"""
    if s.lower() != s and s.upper() != s:
        raise ValueError("mixed-case bech32m string")
    s = s.lower()
    pos = s.rfind("1")
    if pos < 1 or pos + 7 > len(s):
        raise ValueError("no separator / too short")
    hrp = s[:pos]
    try:
        data = [_CHARSET.index(c) for c in s[pos + 1:]]
    except ValueError:
        raise ValueError("invalid bech32m character") from None
    if _polymod(_hrp_expand(hrp) + data) != _BECH32M:
        raise ValueError("bad bech32m checksum")
    acc = bits = 0
    out = bytearray()
    for v in data[:-6]:
        acc = ((acc << 5) | v) & 0xFFF
        bits += 5
        if bits >= 8:
            bits -= 8
            out.append((acc >> bits) & 0xFF)
    if bits >= 5 or (acc & ((1 << bits) - 1)):
        raise ValueError("non-zero padding")
    return hrp, bytes(out)


def bech32m_encode(hrp: str, payload: bytes) -> str:
    acc = bits = 0
    data = []
    for b in payload:
        acc = ((acc << 8) | b) & 0xFFF
        bits += 8
        while bits >= 5:
            bits -= 5
            data.append((acc >> bits) & 31)
    if bits:
        data.append((acc << (5 - bits)) & 31)
    pm = _polymod(_hrp_expand(hrp) + data + [0] * 6) ^ _BECH32M
    data += [(pm >> (5 * (5 - i))) & 31 for i in range(6)]
    return hrp + "1" + "".join(_CHARSET[d] for d in data)


# ---- Proof ------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.buf, self.pos = buf, 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.buf):
            raise ValueError("proof truncated at byte %d" % self.pos)
        out = self.buf[self.pos:self.pos + n]
        self.pos += n
        return out

    def u8(self) -> int:
        return self.take(1)[0]

    def u64(self) -> int:
        return int.from_bytes(self.take(8), "little")

    def option(self, n: int) -> Optional[bytes]:
        tag = self.u8()
        if tag > 1:
            raise ValueError("bad Option tag %d at byte %d" % (tag, self.pos - 1))
        return self.take(n) if tag else None


def _u64(v: int) -> bytes:
    return int(v).to_bytes(8, "little")


def _option(v: Optional[bytes]) -> bytes:
    return b"\x00" if v is None else b"\x01" + v


@dataclass
class KZGProof:
    """kzg10::KZGProof: witness commitment w (compressed G1) and the optional hiding evaluation random_v (Fr, canonical LE)"""
    w: bytes
    random_v: Optional[bytes]


@dataclass
class Proof:
    """``Proof<N>`` of snarkVM 0.14.5 as serialised inside ``proof1...`` strings.  All group elements stay 48-byte
    compressed G1, all field elements 32-byte canonical little-endian Fr: the codec is byte-exact by construction."""
    version: int
    batch_sizes: List[int]
    witness_commitments: List[bytes]          # 3 per instance: w, z_a, z_b
    mask_poly: Optional[bytes]
    g_1: bytes
    h_1: bytes
    g_a: List[bytes]                          # one per circuit
    g_b: List[bytes]
    g_c: List[bytes]
    h_2: bytes
    z_b_evals: List[bytes]                    # one per instance
    g_1_eval: bytes
    g_a_evals: List[bytes]                    # one per circuit
    g_b_evals: List[bytes]
    g_c_evals: List[bytes]
    sums: List[bytes] = field(default_factory=list)      # prover third message: 3 Fr (sum_a, sum_b, sum_c) per circuit
    pc_proofs: List[KZGProof] = field(default_factory=list)
    pc_evaluations: Optional[List[bytes]] = None

    HRP = "proof"

    # -- bytes ----------------------------------------------------------------------------------------------------
    @classmethod
    def from_bytes(cls, payload: bytes) -> "Proof":
        r = _Reader(bytes(payload))
        version = r.u8()
        batch_sizes = [r.u64() for _ in range(r.u64())]
        circuits, instances = len(batch_sizes), sum(batch_sizes)
        wit = [r.take(G1_BYTES) for _ in range(3 * instances)]
        mask = r.option(G1_BYTES)
        g_1, h_1 = r.take(G1_BYTES), r.take(G1_BYTES)
        g_a = [r.take(G1_BYTES) for _ in range(circuits)]
        g_b = [r.take(G1_BYTES) for _ in range(circuits)]
        g_c = [r.take(G1_BYTES) for _ in range(circuits)]
        h_2 = r.take(G1_BYTES)
        z_b = [r.take(FR_BYTES) for _ in range(instances)]
        g_1_eval = r.take(FR_BYTES)
        ga_e = [r.take(FR_BYTES) for _ in range(circuits)]
        gb_e = [r.take(FR_BYTES) for _ in range(circuits)]
        gc_e = [r.take(FR_BYTES) for _ in range(circuits)]
        nsums = r.u64()
        if nsums != circuits:
            raise ValueError("prover message holds %d sum triples for %d circuits" % (nsums, circuits))
        sums = [r.take(FR_BYTES) for _ in range(3 * nsums)]
        pcs = []
        for _ in range(r.u64()):
            w = r.take(G1_BYTES)
            pcs.append(KZGProof(w, r.option(FR_BYTES)))
        evs = None
        if r.u8():
            evs = [r.take(FR_BYTES) for _ in range(r.u64())]
        if r.pos != len(r.buf):
            raise ValueError("%d trailing bytes after the proof" % (len(r.buf) - r.pos))
        return cls(version, batch_sizes, wit, mask, g_1, h_1, g_a, g_b, g_c, h_2, z_b, g_1_eval, ga_e, gb_e, gc_e, sums, pcs, evs)

    def to_bytes(self) -> bytes:
        out = bytearray([self.version]) + _u64(len(self.batch_sizes))
        for b in self.batch_sizes:
            out += _u64(b)
        out += b"".join(self.witness_commitments) + _option(self.mask_poly) + self.g_1 + self.h_1
        out += b"".join(self.g_a) + b"".join(self.g_b) + b"".join(self.g_c) + self.h_2
        out += b"".join(self.z_b_evals) + self.g_1_eval + b"".join(self.g_a_evals) + b"".join(self.g_b_evals) + b"".join(self.g_c_evals)
        out += _u64(len(self.sums) // 3) + b"".join(self.sums)
        out += _u64(len(self.pc_proofs))
        for p in self.pc_proofs:
            out += p.w + _option(p.random_v)
        if self.pc_evaluations is None:
            out += b"\x00"
        else:
            out += b"\x01" + _u64(len(self.pc_evaluations)) + b"".join(self.pc_evaluations)
        return bytes(out)

    # -- strings ----------------------------------------------------------------------------------------------------
    @classmethod
    def from_str(cls, s: str) -> "Proof":
        hrp, payload = bech32m_decode(s)
        if hrp != cls.HRP:
            raise ValueError("expected a '%s1...' string, got hrp %r" % (cls.HRP, hrp))
        return cls.from_bytes(payload)

    def to_string(self) -> str:
        return bech32m_encode(self.HRP, self.to_bytes())

    __str__ = to_string

    # -- the group elements ------------------------------------------------------------------------------------------
    def commitments(self) -> List[bytes]:
        """every compressed G1 of the proof in serialisation order: the polynomial commitments, then the opening
        witnesses -- one G1 MSM each on the prover side (11 + 2 for a single-circuit proof)"""
        out = list(self.witness_commitments)
        if self.mask_poly is not None:
            out.append(self.mask_poly)
        out += [self.g_1, self.h_1] + self.g_a + self.g_b + self.g_c + [self.h_2]
        return out + [p.w for p in self.pc_proofs]

    def with_commitments(self, comms: List[bytes]) -> "Proof":
        """the same proof with its group elements replaced (same order as ``commitments``)"""
        comms = [bytes(c) for c in comms]
        if len(comms) != len(self.commitments()) or any(len(c) != G1_BYTES for c in comms):
            raise ValueError("expected %d compressed points of 48 bytes" % len(self.commitments()))
        it = iter(comms)
        nw, nc = len(self.witness_commitments), len(self.g_a)
        wit = [next(it) for _ in range(nw)]
        mask = next(it) if self.mask_poly is not None else None
        g_1, h_1 = next(it), next(it)
        g_a, g_b, g_c = ([next(it) for _ in range(nc)] for _ in range(3))
        h_2 = next(it)
        pcs = [KZGProof(next(it), p.random_v) for p in self.pc_proofs]
        return replace(self, witness_commitments=wit, mask_poly=mask, g_1=g_1, h_1=h_1, g_a=g_a, g_b=g_b, g_c=g_c, h_2=h_2, pc_proofs=pcs)

    def commitments_affine_dev(self, affine_stride: int = 104, device=None, validate: bool = True):
        """all group elements decompressed on the device in one batch: ((count * stride) uint8 CUDA tensor, invalid count)"""
        import torch

        blob = b"".join(self.commitments())
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
        return g1_decompress_dev(t, affine_stride, validate)

    def with_commitments_from_affine_dev(self, affine_t, affine_stride: int = 104) -> "Proof":
        """re-encodes device-resident affine points (e.g. fresh MSM results) into the proof: device compression, one batch"""
        count = len(self.commitments())
        blob = g1_compress_dev(affine_t, count, affine_stride).cpu().numpy().tobytes()
        return self.with_commitments([blob[i * G1_BYTES:(i + 1) * G1_BYTES] for i in range(count)])
