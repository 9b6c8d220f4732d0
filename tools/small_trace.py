#!/usr/bin/env python3
"""Where does a proof-sized MSM (2^15 .. 2^18 points: SURVEY.md 8a row 14) spend its time?  Runs the plain MSM, the
resident-SRS MSM and the batched KZG commit at the given sizes with ALEO_B200_MSM_TRACE=1 (per-launch event timings on
stderr) and prints the totals.  Development tool.
  python tools/small_trace.py [LOG_N ...]"""
import os
import sys

os.environ["ALEO_B200_MSM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

lib = ab.get_lib()
lib.check(lib.init(0), "init")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for log_n in [int(a) for a in sys.argv[1:]] or [16, 18]:
    n = 1 << log_n
    bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
    sc = ab.gen_scalars_dev(n, 1)
    scm = ab.gen_scalars_dev(n, 1, 0, True)
    out = torch.empty(144, dtype=torch.uint8, device="cuda")
    print("==== plain MSM 2^%d (c = %d, %d launches) ====" % (log_n, ab.VariableBase.window_bits(n), ab.VariableBase.launches(n)), file=sys.stderr, flush=True)
    ab.VariableBase.msm_dev(bases, sc, n, 104, out=out)
    torch.cuda.synchronize()
    srs = ab.ResidentSRS.from_device(bases, n, 104)
    print("==== resident-SRS MSM 2^%d (%r) ====" % (log_n, srs.info()), file=sys.stderr, flush=True)
    srs.msm_dev(sc, n, out=out)
    torch.cuda.synchronize()
    os.environ.pop("ALEO_B200_MSM_TRACE")
    t_plain = timed(lambda: ab.VariableBase.msm_dev(bases, sc, n, 104, out=out))
    t_srs = timed(lambda: srs.msm_dev(sc, n, out=out))
    o48 = torch.empty((13, 48), dtype=torch.uint8, device="cuda")
    t_one = timed(lambda: ab.KZG10.commit_dev(srs, scm, n, out=o48[0]))
    t_b13 = timed(lambda: ab.KZG10.commit_batch_dev(srs, [scm] * 13, out=o48))
    print("2^%d: plain %.3f ms | resident SRS %.3f ms | kzg_commit %.3f ms | 13 commits batched %.3f ms (%.3f each)" %
          (log_n, t_plain, t_srs, t_one, t_b13, t_b13 / 13), flush=True)
    os.environ["ALEO_B200_MSM_TRACE"] = "1"
    srs.close()
