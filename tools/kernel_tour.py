#!/usr/bin/env python3
"""Small tour of every kernel family with self-checks (was meant for compute-sanitizer, which is closed on this pool): python tools/kernel_tour.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402
from aleo_b200 import poly, wire  # noqa: E402
from aleo_b200.dist import PeerNTT  # noqa: E402

lib = ab.get_lib()
lib.check(lib.init(0), "init")
n = 20000
bases = ab.gen_bases_dev(n, 12345, 67891, 0, 104)
sc = ab.gen_scalars_dev(n, 1)
sc[::3] = 0
out = ab.VariableBase.msm_dev(bases, sc, n, 104)
hb, hs = bases.cpu().numpy(), sc.cpu().numpy()
os.environ["ALEO_B200_MSM_CHUNKS"] = "3"
host = ab.VariableBase.msm(hb, hs, 104)
assert host == out.cpu().numpy().tobytes()
srs = ab.ResidentSRS.from_device(bases, n, 104)
assert srs.msm_dev(sc, n).cpu().numpy().tobytes() == host
polys = [ab.gen_scalars_dev(m, 7 + m, 0, True) for m in (n, 5000, 1, 777)]
batch = ab.KZG10.commit_batch_dev(srs, polys)
single = torch.stack([ab.KZG10.commit_dev(srs, p, p.shape[0]) for p in polys])
assert torch.equal(batch, single)
srs.close()
for log_n in (5, 12, 13, 17, 18, 19):
    dom = ab.EvaluationDomain.new(1 << log_n)
    x = ab.gen_scalars_dev(1 << log_n, log_n, 0, True)
    y = dom.coset_ifft_in_place_dev(dom.coset_fft_in_place_dev(x.clone()))
    assert torch.equal(x, y)
    dom._run_dev_ordered(x.clone(), 0, 0, 1)
p = PeerNTT(17)
x = ab.gen_scalars_dev(1 << 17, 3, 0, True)
assert torch.equal(p.transform(x).reshape(-1, 4), ab.EvaluationDomain.new(1 << 17).fft_in_place_dev(x.clone()))
p.close()
c = ab.gen_scalars_dev(9000, 4, 0, True)
poly.distribute_powers_dev(c.clone(), 22, 5)
poly.evaluate_dev(c, 12345)
poly.divide_by_linear_dev(c, 777)
poly.lagrange_coeffs_dev(10, 999)
poly.field_op_dev(poly.FR, poly.INV, c)
comp = wire.g1_compress_dev(bases, n, 104)
back, bad = wire.g1_decompress_dev(comp, 104)
assert bad == 0 and torch.equal(back, bases[: n * 104])
torch.cuda.synchronize()
print("kernel tour ok")
