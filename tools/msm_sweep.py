#!/usr/bin/env python3
"""Times the resident MSM phases at one size (used with ALEO_B200_MSM_WAVES=... to tune the run count)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aleo_b200 as ab
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
lib = ab.get_lib(); lib.check(lib.init(0), "init")
bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
sc = ab.gen_scalars_dev(n, 1)
out = torch.empty(144, dtype=torch.uint8, device="cuda")
ph = (C.c_float * 3)()
res = []
for i in range(4):
    lib.check(lib.msm_g1_dev_profile(out.data_ptr(), bases.data_ptr(), n, sc.data_ptr(), 104, torch.cuda.current_stream().cuda_stream, ph), "p")
    res.append((ph[0], ph[1], ph[2]))
print("waves=%s log_n=%d phases(ms) sort/acc/tail: %s" % (os.environ.get("ALEO_B200_MSM_WAVES", "default"), log_n,
      " | ".join("%.2f/%.2f/%.2f" % r for r in res[1:])))
