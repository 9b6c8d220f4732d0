#!/usr/bin/env python3
"""End-to-end MSM from PAGEABLE host memory with N ranks on one host (run under torchrun): whole-box Mpts/s of
aleo_b200_msm_g1 for the staging configuration in the environment (ALEO_B200_FEED_SLICE_KB, ALEO_B200_FEED_THREADS).
  ALEO_B200_FEED_SLICE_KB=1024 python -m torch.distributed.run --nproc-per-node 8 ... tools/e2e_feed_sweep.py [log_n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import aleo_b200 as ab  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ab.get_lib().check(ab.get_lib().init(local), "init")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
bases = ab.gen_bases_dev(n, 12345, 67891, rank * n, 104, device=dev)
sc = ab.gen_scalars_dev(n, 1, rank * n, False, device=dev)
pb, ps = bases.cpu().numpy().copy(), sc.cpu().numpy().copy()        # pageable
want = ab.VariableBase.msm_dev(bases, sc, n, 104).cpu().numpy().tobytes()
del bases, sc
ok = ab.VariableBase.msm(pb, ps, 104) == want


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


reps = 4
barrier()
t0 = time.perf_counter()
for _ in range(reps):
    ab.VariableBase.msm(pb, ps, 104)
barrier()
secs = torch.tensor([time.perf_counter() - t0], device=dev)
if world > 1:
    dist.all_reduce(secs, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world=%d log_n=%d slice_kb=%s threads=%s: %.1f Mpts/s whole box (%.1f ms per call), ok=%s" %
          (world, log_n, os.environ.get("ALEO_B200_FEED_SLICE_KB", "default"), os.environ.get("ALEO_B200_FEED_THREADS", "default"),
           world * n * reps / secs.item() / 1e6, secs.item() / reps * 1e3, ok), flush=True)
if world > 1:
    dist.destroy_process_group()
