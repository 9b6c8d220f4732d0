#!/usr/bin/env python3
"""MSM tuning sweeps on one B200 (development tool; prints one line per configuration).

  python tools/msm_sweep.py phases LOG_N [C ...]     resident MSM phases (sort / accumulate / tail) per window size c
  python tools/msm_sweep.py host LOG_N [K ...]       host-pointer aleo_b200_msm_g1 wall time per chunk count K (1..3)
  python tools/msm_sweep.py sizes LOG_A LOG_B        resident MSM total time per size with the default window choice
  python tools/msm_sweep.py split LOG_N W,W,.. [...]  host-pointer call per explicit point-range weights (ALEO_B200_MSM_SPLIT)
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402

mode = sys.argv[1]
lib = ab.get_lib()
lib.check(lib.init(0), "init")
stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


STRIDE = int(os.environ.get("MSM_SWEEP_STRIDE", "104"))  # 104: Rust G1Affine layout, 96: packed (modes ba / phases / sizes)


def setup(log_n):
    n = 1 << log_n
    return n, ab.gen_bases_dev(n, 12345, 67891, 0, STRIDE), ab.gen_scalars_dev(n, 1)


if mode == "phases":
    log_n = int(sys.argv[2])
    n, bases, sc = setup(log_n)
    out = torch.empty(144, dtype=torch.uint8, device="cuda")
    ph = (C.c_float * 3)()
    ref = None
    for c in [int(x) for x in sys.argv[3:]] or [0]:
        if c:
            os.environ["ALEO_B200_MSM_C"] = str(c)
        res = []
        for _ in range(4):
            lib.check(lib.msm_g1_dev_profile(out.data_ptr(), bases.data_ptr(), n, sc.data_ptr(), STRIDE, stream(), ph), "p")
            res.append((ph[0], ph[1], ph[2]))
        raw = out.cpu().numpy().tobytes()
        ref = ref or raw
        best = min(res[1:], key=sum)
        print("log_n=%d c=%d sort/acc/tail ms: %.2f/%.2f/%.2f total %.2f same_result=%s" %
              (log_n, lib.msm_window_bits(n), best[0], best[1], best[2], sum(best), raw == ref), flush=True)
elif mode == "ba":  # ba LOG_N  SPEC ...   SPEC = levels[:k[:c[:pf]]] (batch-affine levels, additions per inversion, window bits, prefetch distance); 0 = off
    log_n = int(sys.argv[2])
    n, bases, sc = setup(log_n)
    out = torch.empty(144, dtype=torch.uint8, device="cuda")
    ph = (C.c_float * 3)()
    ref = None
    for spec in sys.argv[3:]:
        f = spec.split(":")
        os.environ["ALEO_B200_MSM_BA"] = f[0]
        os.environ.pop("ALEO_B200_MSM_BA_K", None)
        os.environ.pop("ALEO_B200_MSM_C", None)
        if len(f) > 1 and f[1]:
            os.environ["ALEO_B200_MSM_BA_K"] = f[1]
        if len(f) > 2 and f[2]:
            os.environ["ALEO_B200_MSM_C"] = f[2]
        os.environ.pop("ALEO_B200_MSM_BA_PF", None)
        if len(f) > 3 and f[3]:
            os.environ["ALEO_B200_MSM_BA_PF"] = f[3]
        res = []
        for _ in range(4):
            lib.check(lib.msm_g1_dev_profile(out.data_ptr(), bases.data_ptr(), n, sc.data_ptr(), STRIDE, stream(), ph), "p")
            res.append((ph[0], ph[1], ph[2]))
        raw = out.cpu().numpy().tobytes()
        ref = ref or raw
        best = min(res[1:], key=sum)
        print("log_n=%d ba=%s c=%d sort/acc/tail ms: %.2f/%.2f/%.2f total %.2f -> %.1f Mpts/s same_result=%s" %
              (log_n, spec, lib.msm_window_bits(n), best[0], best[1], best[2], sum(best), n / sum(best) / 1e3, raw == ref), flush=True)
elif mode == "host":
    log_n = int(sys.argv[2])
    n, bases, sc = setup(log_n)
    hb = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
    hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hb.copy_(bases)
    hs.copy_(sc)
    torch.cuda.synchronize()
    ref = None
    for k in [int(x) for x in sys.argv[3:]] or [1, 2, 3]:
        os.environ["ALEO_B200_MSM_CHUNKS"] = str(k)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            raw = ab.VariableBase.msm(hb, hs, 104)
            ts.append((time.perf_counter() - t0) * 1e3)
        ref = ref or raw
        print("log_n=%d chunks=%d host-call ms: %s  -> %.1f Mpts/s same_result=%s" %
              (log_n, k, " ".join("%.2f" % t for t in ts), n / min(ts[1:]) / 1e3, raw == ref), flush=True)
elif mode == "split":
    log_n = int(sys.argv[2])
    n, bases, sc = setup(log_n)
    hb = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
    hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hb.copy_(bases)
    hs.copy_(sc)
    torch.cuda.synchronize()
    ref = None
    for w in sys.argv[3:]:
        if w == "default":
            os.environ.pop("ALEO_B200_MSM_SPLIT", None)
        else:
            os.environ["ALEO_B200_MSM_SPLIT"] = w
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            raw = ab.VariableBase.msm(hb, hs, 104)
            ts.append((time.perf_counter() - t0) * 1e3)
        ref = ref or raw
        print("log_n=%d split=%s host-call ms: %s  -> %.1f Mpts/s same_result=%s" %
              (log_n, w, " ".join("%.2f" % t for t in ts), n / min(ts[1:]) / 1e3, raw == ref), flush=True)
elif mode == "pageable":  # one process per ALEO_B200_FEED_THREADS value (read once per process)
    log_n = int(sys.argv[2])
    n, bases, sc = setup(log_n)
    hb = bases.cpu().numpy().copy()   # plain (pageable) numpy memory, like a Rust Vec
    hs = sc.cpu().numpy().copy()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        raw = ab.VariableBase.msm(hb, hs, 104)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("log_n=%d feed_threads=%s pageable host-call ms: %s  -> %.1f Mpts/s" %
          (log_n, os.environ.get("ALEO_B200_FEED_THREADS", "default"), " ".join("%.2f" % t for t in ts), n / min(ts[1:]) / 1e3), flush=True)
elif mode == "sizes":
    for log_n in range(int(sys.argv[2]), int(sys.argv[3]) + 1):
        n, bases, sc = setup(log_n)
        out = torch.empty(144, dtype=torch.uint8, device="cuda")
        ph = (C.c_float * 3)()
        res = []
        for _ in range(4):
            lib.check(lib.msm_g1_dev_profile(out.data_ptr(), bases.data_ptr(), n, sc.data_ptr(), STRIDE, stream(), ph), "p")
            res.append((ph[0], ph[1], ph[2]))
        best = min(res[1:], key=sum)
        print("log_n=%d c=%d sort/acc/tail ms: %.3f/%.3f/%.3f total %.3f -> %.1f Mpts/s" %
              (log_n, lib.msm_window_bits(n), best[0], best[1], best[2], sum(best), n / sum(best) / 1e3), flush=True)
        del bases, sc
        torch.cuda.empty_cache()
