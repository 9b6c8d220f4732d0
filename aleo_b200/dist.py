"""Multi-GPU decomposition of the hot path, one process per GPU over
torch.distributed (NCCL over NVLink on the box; the same code runs over gloo in the CPU tests with an
injected backend).

* MSM shards by contiguous point range: every rank runs the full Pippenger pipeline on its slice
  (bases stay resident per GPU); the single final combine is an all-gather of the 144-byte partial
  results followed by ``aleo_b200_g1_sum_dev``.
* NTT uses the four-step (Bailey) decomposition N = n1 * n2 with exactly one all-to-all:
      input  : rank j holds the column block  A_j[i1][c] = x[i1 * n2 + j * n2/g + c]      (n1 x n2/g)
      step 1 : n2/g local transforms of length n1 over i1          (batched kernel launch)
      step 2 : multiply by w_N^(i2 * k1)                           (aleo_b200_ntt_twiddle_dev)
      step 3 : all-to-all transpose over NVLink
      step 4 : n1/g local transforms of length n2 over i2
      output : rank t holds the row block     B_t[k][k2] = X[(t * n1/g + k) + n1 * k2]    (n1/g x n2)
  i.e. the output is in transposed (digit-swapped) order, distributed -- what the next transform of a
  prover round consumes with the roles of n1 and n2 exchanged; ``gather_natural`` restores natural
  order when a caller needs it.  The inverse transform is the same schedule with inverse local
  transforms (their n1^-1 * n2^-1 scalings multiply to N^-1) and the inverse twiddle.
* Batches of independent proofs: replicas only, no collective (nothing to implement here).
"""
from __future__ import annotations

from . import _lib


class CudaBackend:
    """the product backend: libaleo_b200.so on torch CUDA tensors, current stream"""

    def _stream(self):
        import torch
        return torch.cuda.current_stream().cuda_stream

    def ntt_batch(self, t, log_n: int, batch: int, inverse: bool):
        import torch
        lib = _lib.get_lib()
        with torch.cuda.device(t.device):
            lib.check(lib.ntt_fr_dev(t.data_ptr(), log_n, batch, _lib.NTT_INVERSE if inverse else _lib.NTT_FORWARD,
                                     _lib.NTT_STANDARD, self._stream()), "aleo_b200_ntt_fr_dev")
        return t

    def twiddle(self, t, log_n_global: int, inverse: bool, rows: int, cols: int, row0: int, col0: int):
        import torch
        lib = _lib.get_lib()
        with torch.cuda.device(t.device):
            lib.check(lib.ntt_twiddle_dev(t.data_ptr(), log_n_global, _lib.NTT_INVERSE if inverse else _lib.NTT_FORWARD,
                                          rows, cols, row0, col0, self._stream()), "aleo_b200_ntt_twiddle_dev")
        return t

    def msm(self, bases, scalars, n: int, stride: int):
        from .msm import VariableBase
        return VariableBase.msm_dev(bases, scalars, n, stride)

    def sum_partials(self, parts, count: int):
        from .msm import VariableBase
        return VariableBase.sum_partials_dev(parts, count)


def _world(group=None):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ---- MSM ------------------------------------------------------------------------------------------
def point_range(n_total: int, rank: int, world: int):
    """contiguous point range [first, first + count) of rank `rank` (remainder spread over low ranks)"""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def msm_sharded(bases_local, scalars_local, n_local: int, stride: int = 104, group=None, backend=None):
    """every rank passes ITS point range; every rank gets the full result (144-byte tensor)"""
    import torch
    import torch.distributed as dist

    backend = backend or CudaBackend()
    rank, world = _world(group)
    partial = backend.msm(bases_local, scalars_local, n_local, stride)
    if world == 1:
        return partial
    gathered = torch.empty(144 * world, dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(gathered, partial.contiguous(), group=group)
    return backend.sum_partials(gathered, world)


# ---- NTT ------------------------------------------------------------------------------------------
def four_step_shape(log_n: int):
    """(log n1, log n2) with n1 >= n2"""
    l1 = (log_n + 1) // 2
    return l1, log_n - l1


def column_block(x_full, log_n: int, rank: int, world: int):
    """helper: natural-order vector (N, 4) -> this rank's input block A_j (n1, n2/g, 4)"""
    l1, l2 = four_step_shape(log_n)
    n1, n2 = 1 << l1, 1 << l2
    w = n2 // world
    return x_full.reshape(n1, n2, 4)[:, rank * w:(rank + 1) * w, :].contiguous()


def gather_natural(blocks, log_n: int):
    """helper: list over ranks of output blocks B_t (n1/g, n2, 4) -> natural-order vector (N, 4)"""
    import torch
    l1, l2 = four_step_shape(log_n)
    n1, n2 = 1 << l1, 1 << l2
    b = torch.cat(list(blocks), dim=0)                  # (n1, n2, 4) indexed [k1][k2] = X[k1 + n1 k2]
    return b.reshape(n1, n2, 4).transpose(0, 1).contiguous().reshape(n1 * n2, 4)


def ntt_four_step(block, log_n: int, inverse: bool = False, group=None, backend=None):
    """distributed radix-2 NTT of size 2^log_n over Fr; `block` is this rank's (n1, n2/g, 4) int64 column
    block (Montgomery limbs); returns this rank's (n1/g, n2, 4) block of the transposed-order output."""
    import torch
    import torch.distributed as dist

    backend = backend or CudaBackend()
    rank, world = _world(group)
    l1, l2 = four_step_shape(log_n)
    n1, n2 = 1 << l1, 1 << l2
    if n2 % world or n1 % world:
        raise ValueError("world size %d must divide n1 = %d and n2 = %d" % (world, n1, n2))
    wc, wr = n2 // world, n1 // world
    if tuple(block.shape) != (n1, wc, 4):
        raise ValueError("expected a (%d, %d, 4) column block, got %r" % (n1, wc, tuple(block.shape)))
    cols = block.transpose(0, 1).contiguous()                                    # (n2/g, n1, 4): columns contiguous
    backend.ntt_batch(cols, l1, wc, inverse)                                     # step 1
    backend.twiddle(cols, log_n, inverse, wc, n1, rank * wc, 0)                  # step 2
    if world > 1:                                                                # step 3
        send = cols.reshape(wc, world, wr, 4).permute(1, 0, 2, 3).contiguous()   # (g, n2/g, n1/g, 4)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        rows = recv.reshape(n2, wr, 4)                                           # [i2][k1']
    else:
        rows = cols
    rows = rows.transpose(0, 1).contiguous()                                     # (n1/g, n2, 4)
    backend.ntt_batch(rows, l2, wr, inverse)                                     # step 4
    return rows


# ---- NTT with the exchange fused into the transform (peer-memory stores over NVLink) ---------------------------------
class PeerNTT:
    """One Fr NTT of size 2^log_n spread over the GPUs of a node, one process per GPU (include/aleo_b200.h
    ``aleo_b200_ntt_dist_*``).  The pass split is the single-GPU one; the last-but-one pass stores every result
    element straight into the receive buffer of the rank that needs it (CUDA IPC mapped peer memory), a
    flag per source rank in the same peer memory (release store by the exchange pass, acquire spin by the last pass)
    orders the ranks -- no collective per transform --, and the last pass runs from the receive buffer.  Input: this rank's column block of the (N / R_last) x R_last matrix of x; output: its column block
    of the (N / R_first) x R_first matrix of X (natural order) -- see ``layout``."""

    def __init__(self, log_n: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist

        self._lib = _lib.get_lib()
        self.log_n = log_n
        self.group = group
        self.rank, self.world = _world(group)
        kf, kl, npass = C.c_uint32(), C.c_uint32(), C.c_int()
        self._lib.check(self._lib.ntt_dist_layout(log_n, self.world, C.byref(kf), C.byref(kl), C.byref(npass)),
                        "aleo_b200_ntt_dist_layout")
        self.log_r_first, self.log_r_last, self.passes = kf.value, kl.value, npass.value
        self._h = C.c_void_p()
        self._lib.check(self._lib.ntt_dist_create(C.byref(self._h), log_n, self.rank, self.world), "aleo_b200_ntt_dist_create")
        mine = C.create_string_buffer(128)
        self._lib.check(self._lib.ntt_dist_handles(self._h, C.cast(mine, C.c_void_p)), "aleo_b200_ntt_dist_handles")
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device())
            t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).to(dev)
            allh = torch.empty(128 * self.world, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, t, group=group)
            raw = allh.cpu().numpy().tobytes()
        else:
            raw = mine.raw
        buf = C.create_string_buffer(raw, len(raw))
        self._lib.check(self._lib.ntt_dist_open(self._h, C.cast(buf, C.c_void_p)), "aleo_b200_ntt_dist_open")

    def layout(self):
        """(rows_in, cols_in, rows_out, cols_out) of this rank's input / output column blocks"""
        n = 1 << self.log_n
        rl, rf = 1 << self.log_r_last, 1 << self.log_r_first
        return n // rl, rl // self.world, n // rf, rf // self.world

    def input_block(self, x_full):
        """helper: natural-order (N, 4) tensor -> this rank's input block"""
        rows, cols, _, _ = self.layout()
        return x_full.reshape(rows, cols * self.world, 4)[:, self.rank * cols:(self.rank + 1) * cols, :].contiguous()

    def output_block_of(self, X_full):
        """helper: natural-order (N, 4) result -> the block this rank ends up with"""
        _, _, rows, cols = self.layout()
        return X_full.reshape(rows, cols * self.world, 4)[:, self.rank * cols:(self.rank + 1) * cols, :].contiguous()

    def transform(self, block, out=None, inverse: bool = False, coset: bool = False):
        """one distributed transform; no collective: the ranks order themselves through flags in peer memory"""
        import torch

        if out is None:
            out = torch.empty_like(block).reshape(-1, 4)
        direction = _lib.NTT_INVERSE if inverse else _lib.NTT_FORWARD
        kind = _lib.NTT_COSET if coset else _lib.NTT_STANDARD
        with torch.cuda.device(block.device):
            stream = torch.cuda.current_stream().cuda_stream
            self._lib.check(self._lib.ntt_dist_transform(self._h, block.data_ptr(), out.data_ptr(), direction, kind, stream),
                            "aleo_b200_ntt_dist_transform")
        return out

    def stage_times(self, block, out, inverse: bool = False, coset: bool = False, reps: int = 5):
        """per-stage device times of one transform, median of `reps`: dict of ms (local passes before the exchange pass,
        exchange pass = butterflies + stores into the peers, wait for every peer's flag, last pass)"""
        import ctypes as C

        import torch
        direction = _lib.NTT_INVERSE if inverse else _lib.NTT_FORWARD
        kind = _lib.NTT_COSET if coset else _lib.NTT_STANDARD
        rows = []
        ms = (C.c_float * 4)()
        with torch.cuda.device(block.device):
            stream = torch.cuda.current_stream().cuda_stream
            for _ in range(reps):
                self._lib.check(self._lib.ntt_dist_profile(self._h, block.data_ptr(), out.data_ptr(), direction, kind, stream, ms),
                                "aleo_b200_ntt_dist_profile")
                rows.append([ms[i] for i in range(4)])
        med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(4)]
        return {"local_passes": round(med[0], 4), "exchange_pass": round(med[1], 4), "wait_for_peers": round(med[2], 4),
                "last_pass": round(med[3], 4)}

    def close(self):
        if self._h:
            self._lib.ntt_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
