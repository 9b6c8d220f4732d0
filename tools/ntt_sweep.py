#!/usr/bin/env python3
"""NTT timing per size on one B200 (development tool): python tools/ntt_sweep.py LOG_A LOG_B"""
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
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 7, 0, True)
    x0 = x.clone()
    dom.fft_in_place_dev(x)
    dom.ifft_in_place_dev(x)
    ok = torch.equal(x, x0)
    res = {}
    for name, fn in (("fft", dom.fft_in_place_dev), ("coset_ifft", dom.coset_ifft_in_place_dev)):
        fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            fn(x)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
    pm = (C.c_float * 4)()
    lib.check(lib.ntt_fr_dev_profile(x.data_ptr(), log_n, 0, 0, torch.cuda.current_stream().cuda_stream, pm), "p")
    print("log_n=%d fft %.3f ms (%.0f Melem/s) coset_ifft %.3f ms passes %s roundtrip=%s" %
          (log_n, res["fft"], n / res["fft"] / 1e3, res["coset_ifft"], " ".join("%.3f" % pm[i] for i in range(dom.launches())), ok), flush=True)
    del x, x0
    torch.cuda.empty_cache()
