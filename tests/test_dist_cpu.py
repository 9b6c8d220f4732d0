"""The N > 1 host logic (aleo_b200/dist.py) on CPU: two gloo ranks, the CUDA backend replaced by a
test-only backend that uses the oracle for the local transforms / local MSMs.  What is under test is
the decomposition and the data movement (point ranges, column blocks, twiddle indices, all-to-all
transpose, output ordering) -- the kernels themselves are covered by the gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aleo_b200 import dist as adist  # noqa: E402
from oracle import bls12_377 as o  # noqa: E402


def _fr_rows(t):
    return o.fr_vec_from_bytes(t.contiguous().numpy().tobytes())


def _to_tensor(vals):
    return torch.from_numpy(np.frombuffer(o.fr_vec_to_bytes(vals), dtype=np.int64).reshape(-1, 4).copy())


class OracleBackend:
    """test-only stand-in for CudaBackend (same method contracts)"""

    def ntt_batch(self, t, log_n, batch, inverse):
        n = 1 << log_n
        flat = t.reshape(batch * n, 4)
        vals = _fr_rows(flat)
        out = []
        for b in range(batch):
            seg = vals[b * n:(b + 1) * n]
            out += o.ifft(seg) if inverse else o.fft(seg)
        t.copy_(_to_tensor(out).reshape(t.shape))
        return t

    def twiddle(self, t, log_n_global, inverse, rows, cols, row0, col0):
        w = o.fr_root_of_unity(log_n_global)
        if inverse:
            w = pow(w, -1, o.R_MOD)
        vals = _fr_rows(t.reshape(rows * cols, 4))
        out = [vals[r * cols + c] * pow(w, (r + row0) * (c + col0), o.R_MOD) % o.R_MOD for r in range(rows) for c in range(cols)]
        t.copy_(_to_tensor(out).reshape(t.shape))
        return t

    def msm(self, bases, scalars, n, stride):
        raw = bases.numpy().tobytes()
        B = [o.g1_affine_from_bytes(raw[i * stride:(i + 1) * stride], stride) for i in range(n)]
        s = o.fr_vec_from_bytes(scalars.numpy().tobytes(), mont=False)[:n]
        return torch.frombuffer(bytearray(o.g1_projective_to_bytes(o.msm_pippenger(B, s))), dtype=torch.uint8)

    def sum_partials(self, parts, count):
        acc = None
        raw = parts.numpy().tobytes()
        for i in range(count):
            acc = o.g1_add(acc, o.g1_projective_from_bytes(raw[i * 144:(i + 1) * 144]))
        return torch.frombuffer(bytearray(o.g1_projective_to_bytes(acc)), dtype=torch.uint8)


def _worker(rank, world, port, log_n, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        be = OracleBackend()
        n = 1 << log_n
        x = o.random_fr_vec(n, 4321)
        full = _to_tensor(x)
        for inverse in (False, True):
            blk = adist.column_block(full, log_n, rank, world)
            out = adist.ntt_four_step(blk, log_n, inverse=inverse, backend=be)
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            nat = adist.gather_natural(gathered, log_n)
            want = o.ifft(x) if inverse else o.fft(x)
            results["ntt_inv%d_r%d" % (inverse, rank)] = _fr_rows(nat) == want
        # MSM: contiguous point ranges + single combine
        total = 37
        bases = o.synthetic_bases(total, 99)
        scal = o.random_fr_vec(total, 98)
        first, cnt = adist.point_range(total, rank, world)
        b = torch.frombuffer(bytearray(o.g1_affine_vec_to_bytes(bases[first:first + cnt], 104)), dtype=torch.uint8)
        s = torch.from_numpy(np.frombuffer(o.fr_vec_to_bytes(scal[first:first + cnt], mont=False), dtype=np.int64).reshape(-1, 4).copy())
        got = adist.msm_sharded(b, s, cnt, 104, backend=be)
        results["msm_r%d" % rank] = got.numpy().tobytes() == o.g1_projective_to_bytes(o.msm_expected_from_dlogs(total, 99, scal))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("log_n", [6, 7])
def test_two_rank_four_step_ntt_and_sharded_msm(log_n):
    world = 2
    port = 29500 + (os.getpid() + log_n) % 2000
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, log_n, results), nprocs=world, join=True)
    res = dict(results)
    assert len(res) == 3 * world and all(res.values()), res


def test_point_ranges_cover_exactly():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [adist.point_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_single_process_four_step_matches_oracle():
    """world size 1 (no process group): the schedule itself, odd and even log_n"""
    be = OracleBackend()
    for log_n in (4, 5):
        x = o.random_fr_vec(1 << log_n, 7)
        out = adist.ntt_four_step(adist.column_block(_to_tensor(x), log_n, 0, 1), log_n, backend=be)
        assert _fr_rows(adist.gather_natural([out], log_n)) == o.fft(x)
