#!/usr/bin/env python3
"""Short profiling target: a few MSMs and NTTs of one size on resident data.
  python tools/ncu_target.py [log_n] [reps]      (run plain first, then under ncu; see profiles/README.md)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << log_n
ab.get_lib().check(ab.get_lib().init(0), "init")
bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
sc = ab.gen_scalars_dev(n, 1)
for _ in range(reps):
    out = ab.VariableBase.msm_dev(bases, sc, n, 104)
x = ab.gen_scalars_dev(n, 2, 0, True)
dom = ab.EvaluationDomain.new(n)
for _ in range(reps):
    dom.fft_in_place_dev(x)
    dom.coset_ifft_in_place_dev(x)
torch.cuda.synchronize()
print("ok", out.cpu().numpy()[:8])
