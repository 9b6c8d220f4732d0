#!/usr/bin/env python3
"""Do the phases of two MSMs overlap when they are submitted on two streams?  (Is the sort of one hidden under the
accumulation of the other?)  Decides whether pipelining window groups inside ONE MSM can pay.
  python tools/msm_overlap_probe.py [log_n]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

ab.get_lib().check(ab.get_lib().init(0), "init")
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
scs = [ab.gen_scalars_dev(n, 1 + k) for k in range(2)]
outs = [torch.empty(144, dtype=torch.uint8, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
reps = 6


def work(k, count):
    torch.cuda.set_device(0)
    with torch.cuda.stream(streams[k]):
        for _ in range(count):
            ab.VariableBase.msm_dev(bases, scs[k], n, 104, out=outs[k])
    streams[k].synchronize()


work(0, 2)
work(1, 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
work(0, reps)
work(1, reps)
seq = time.perf_counter() - t0
ts = [threading.Thread(target=work, args=(k, reps)) for k in range(2)]
t0 = time.perf_counter()
for t in ts:
    t.start()
for t in ts:
    t.join()
torch.cuda.synchronize()
par = time.perf_counter() - t0
print("2^%d: %d MSMs one after the other %.2f ms each; two streams concurrently %.2f ms each (%.1f %% faster)" %
      (log_n, 2 * reps, seq / (2 * reps) * 1e3, par / (2 * reps) * 1e3, 100 * (seq / par - 1)), flush=True)
