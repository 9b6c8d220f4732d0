#!/usr/bin/env python3
"""Multi-GPU check of the peer-memory NTT (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_ntt_check.py [log_n ...]
Every rank computes the full single-GPU transform of the same seeded vector and compares its own output block with
what the distributed transform left it; then times the transform (device events, max over ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import aleo_b200 as ab  # noqa: E402
from aleo_b200.dist import PeerNTT, ntt_four_step, four_step_shape  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ab.get_lib().check(ab.get_lib().init(local), "init")
ok_all = True
for log_n in [int(a) for a in sys.argv[1:]] or [16, 20, 24]:
    n = 1 << log_n
    x = ab.gen_scalars_dev(n, 555 + log_n, 0, True, device=dev)          # same vector on every rank
    dom = ab.EvaluationDomain.new(n)
    p = PeerNTT(log_n)
    oks = []
    for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
        want = dom._run_dev(x.clone(), 1 if inverse else 0, 1 if coset else 0)
        blk = p.input_block(x)
        for _ in range(3):
            got = p.transform(blk, inverse=inverse, coset=coset)
        oks.append(bool(torch.equal(got.reshape(-1), p.output_block_of(want).reshape(-1))))
    blk = p.input_block(x)
    out = torch.empty_like(blk).reshape(-1, 4)
    for _ in range(3):
        p.transform(blk, out=out)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        p.transform(blk, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    stages = p.stage_times(blk, out)
    flags = [None] * world
    dist.all_gather_object(flags, oks)
    good = all(all(f) for f in flags)
    ok_all = ok_all and good
    if rank == 0:
        print("log_n=%d world=%d passes=%d R_first=2^%d R_last=2^%d parity(fwd,inv,coset fwd,coset inv) per rank=%s  %.3f ms -> %.0f Melem/s  rank-0 stages (ms) %s" %
              (log_n, world, p.passes, p.log_r_first, p.log_r_last, flags, ms.item(), n / ms.item() / 1e3, stages), flush=True)
    p.close()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
