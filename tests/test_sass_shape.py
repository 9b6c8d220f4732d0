"""The roofline arithmetic of DESIGN.md / bench.py rests on how ptxas compiles the field products: every
(mad.lo.cc, madc.hi.cc) pair of mont.cuh must fuse into ONE IMAD.WIDE.U32(.X) on the heavy FMA pipe -- Fq product 276,
dedicated Fq square ~210, Fr raw product 120.  The carry chains are separate `asm volatile` statements sharing CC.CF
implicitly, so nothing but the generated SASS says whether that still holds.  This test disassembles three one-product
probe kernels of the BUILT library (util_kernels.cuh aleo_probe_*) with cuobjdump -- no GPU needed -- and checks the
heavy-pipe instruction counts; the histogram is committed under profiles/ by tools/sass_histogram.py."""
import collections
import re
import shutil
import subprocess

import pytest

# heavy-pipe instructions one product may issue: the multiplier rows + the probe's three address computations
EXPECT = {"aleo_probe_fq_mul": (276, 280), "aleo_probe_fq_sqr": (205, 215), "aleo_probe_fr_mul_raw": (120, 124)}


def sass_histogram(lib_path, fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib_path], capture_output=True, text=True, timeout=300).stdout
    ops = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    return ops


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="CUDA toolkit (cuobjdump) not on PATH")
@pytest.mark.parametrize("fun", sorted(EXPECT))
def test_field_products_compile_to_fused_wide_multiplies(product_lib_path, fun):
    ops = sass_histogram(product_lib_path, fun)
    assert ops, "probe kernel %s not found in %s" % (fun, product_lib_path)
    wide = sum(v for k, v in ops.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in ops.items() if k.startswith("IMAD.HI"))
    lo = sum(v for k, v in ops.items() if k == "IMAD" or k.startswith("IMAD.U32") or k.startswith("IMAD.X"))
    lo_hi = EXPECT[fun]
    assert lo_hi[0] <= wide + hi <= lo_hi[1], "%s: %d IMAD.WIDE + %d IMAD.HI on the heavy pipe, expected %r (%r)" % (fun, wide, hi, lo_hi, dict(ops))
    # an unfused pair shows up as separate low / high multiplies: none in the two products; the dedicated square keeps
    # 12 (ptxas leaves one pair per row of the doubled operand unfused -- 12 of ~213 heavy-pipe instructions, in 2 of the
    # 10 field operations of a mixed addition; an opaque register copy did not change it)
    allowed = 12 if fun == "aleo_probe_fq_sqr" else 0
    assert lo <= allowed + 2 and hi <= allowed, "%s: unfused multiply halves: %d IMAD(lo), %d IMAD.HI" % (fun, lo, hi)
    assert wide >= lo_hi[0] - allowed
