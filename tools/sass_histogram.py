#!/usr/bin/env python3
"""Opcode histogram of the field-product probe kernels and of the hot kernels' callees in the built library
(cuobjdump -sass; no GPU needed).  Writes Markdown to stdout:  python tools/sass_histogram.py > profiles/r02_sass_histogram.md"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_sass_shape import EXPECT, sass_histogram  # noqa: E402

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "aleo_b200", "libaleo_b200.so")
print("# SASS opcode histogram of the field products (sm_100a, `cuobjdump -sass`)\n")
print("One-product probe kernels (`util_kernels.cuh aleo_probe_*`); heavy FMA pipe = IMAD.WIDE* + IMAD.HI*.\n")
for fun in sorted(EXPECT):
    ops = sass_histogram(LIB, fun)
    wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in ops.items() if k.startswith("IMAD.HI"))
    print("## %s -- %d IMAD.WIDE + %d IMAD.HI (expected %d..%d incl. 3 address computations), %d instructions\n" %
          (fun, wide, hi, EXPECT[fun][0], EXPECT[fun][1], sum(ops.values())))
    print("| opcode | count |\n|---|---|")
    for k, v in ops.most_common(14):
        print("| %s | %d |" % (k, v))
    print()
