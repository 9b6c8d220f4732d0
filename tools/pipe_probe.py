#!/usr/bin/env python3
"""Issue rates of the integer and FP64 pipes on this GPU (aleo_b200_bench_imad kinds 0..5); prints T ops/s."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aleo_b200 as ab  # noqa: E402

lib = ab.get_lib()
lib.check(lib.init(0), "init")
names = ["IMAD (mad.lo)", "IMAD.WIDE 32x32+64", "IMAD.WIDE.U32.X carry chains", "dependent Fq products (wide MACs)",
         "DFMA", "DFMA + IMAD.WIDE interleaved 1:1 (pairs)", "dependent Fq products on the FP64 pipe (as wide MACs)",
         "dependent Fq squares on the FP64 pipe (as wide MACs)", "dependent Fq squares, dedicated IMAD.WIDE square (as wide MACs)"]
for kind, name in enumerate(names):
    ms, ops = C.c_double(), C.c_double()
    lib.check(lib.bench_imad(kind, 4096, C.byref(ms), C.byref(ops)), "bench_imad")
    print("kind %d %-45s %8.3f ms  %7.3f T/s" % (kind, name, ms.value, ops.value / ms.value / 1e9), flush=True)
