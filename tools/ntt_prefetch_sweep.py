#!/usr/bin/env python3
"""A/B of ALEO_B200_NTT_PREFETCH (L2 prefetch of a later tile's inputs during the store phase of an NTT pass):
forward transform time per size and setting, result compared with the setting 0.  Development tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

ab.get_lib().check(ab.get_lib().init(0), "init")
for log_n in [int(a) for a in sys.argv[1:]] or [22, 24, 26]:
    n = 1 << log_n
    dom = ab.EvaluationDomain.new(n)
    x0 = ab.gen_scalars_dev(n, 7, 0, True)
    ref = None
    for ahead in (0, 74, 148, 296, 592, 1184, 0):
        os.environ["ALEO_B200_NTT_PREFETCH"] = str(ahead)
        x = x0.clone()
        dom.fft_in_place_dev(x)
        ref = x.clone() if ref is None else ref
        same = bool(torch.equal(x, ref))
        for _ in range(3):
            dom.fft_in_place_dev(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            dom.fft_in_place_dev(x)
        e1.record()
        torch.cuda.synchronize()
        print("2^%d prefetch %4d tiles ahead: %.4f ms  same_result=%s" % (log_n, ahead, e0.elapsed_time(e1) / reps, same), flush=True)
