#!/usr/bin/env python3
"""Resident-SRS MSM timing per size (development tool): python tools/srs_sweep.py LOG_A LOG_B"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

lib = ab.get_lib()
lib.check(lib.init(0), "init")
for log_n in range(int(sys.argv[1]), int(sys.argv[2]) + 1):
    n = 1 << log_n
    bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
    sc = ab.gen_scalars_dev(n, 1)
    srs = ab.ResidentSRS.from_device(bases, n, 104)
    info = srs.info()
    out = torch.empty(144, dtype=torch.uint8, device="cuda")
    ph = (C.c_float * 3)()
    res = []
    for _ in range(4):
        lib.check(lib.srs_msm_dev_profile(srs._h, out.data_ptr(), sc.data_ptr(), n, torch.cuda.current_stream().cuda_stream, ph), "p")
        res.append((ph[0], ph[1], ph[2]))
    best = min(res[1:], key=sum)
    print("log_n=%d c=%d W=%d sort/acc/tail ms: %.3f/%.3f/%.3f total %.3f -> %.1f Mpts/s" %
          (log_n, info["window_bits"], info["windows"], best[0], best[1], best[2], sum(best), n / sum(best) / 1e3), flush=True)
    srs.close()
    del bases, sc
    torch.cuda.empty_cache()
