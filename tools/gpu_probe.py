#!/usr/bin/env python3
"""First-contact GPU probe: integer-pipe microbenchmark, parity spot checks and raw timings.
Run on the GPU box:  python tools/gpu_probe.py [--quick] > gpurun_out/probe.log
(results also land in gpurun_out/probe.json).  Uses the oracle only as a checker."""
import argparse
import ctypes as C
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import aleo_b200 as ab  # noqa: E402
from oracle import bls12_377 as o  # noqa: E402

OUT = {}


def section(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                OUT[name] = fn(*a, **k)
            except Exception as e:  # keep going: partial results are still useful
                OUT[name] = {"error": repr(e), "trace": traceback.format_exc()}
            print("== %s (%.1fs): %s" % (name, time.time() - t, json.dumps(OUT[name])[:2000]), flush=True)
        return run
    return deco


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2], times[0]


@section("imad")
def imad():
    lib = ab.get_lib()
    res = {}
    for kind, name in enumerate(["imad_lo", "imad_wide", "imad_wide_x_chain", "fq_mul_chain"]):
        ms, ops = C.c_double(), C.c_double()
        iters = 4096 if kind < 3 else 512
        lib.check(lib.bench_imad(kind, iters, C.byref(ms), C.byref(ops)), "bench_imad")
        res[name] = {"ms": ms.value, "gops_per_s": ops.value / ms.value / 1e6}
    return res


def fr_tensor(vals):
    import numpy as np
    return torch.from_numpy(np.frombuffer(o.fr_vec_to_bytes(vals), dtype=np.int64).reshape(-1, 4).copy()).cuda()


@section("ntt_parity")
def ntt_parity(sizes):
    res = {}
    for log_n in sizes:
        n = 1 << log_n
        v = o.random_fr_vec(n, 500 + log_n)
        dom = ab.EvaluationDomain.new(n)
        want = {"fft": o.fft(v), "ifft": o.ifft(v), "coset_fft": o.coset_fft(v), "coset_ifft": o.coset_ifft(v)}
        for name, ref in want.items():
            t = fr_tensor(v)
            getattr(dom, name + "_in_place_dev")(t)
            got = o.fr_vec_from_bytes(t.cpu().numpy().tobytes())
            res["%s_2^%d" % (name, log_n)] = bool(got == ref)
        # host-pointer entry point
        host = dom.fft(o.fr_vec_to_bytes(v))
        res["host_fft_2^%d" % log_n] = bool(o.fr_vec_from_bytes(host) == want["fft"])
    res["all_ok"] = all(res.values())
    return res


@section("ntt_roundtrip_large")
def ntt_roundtrip(sizes):
    res = {}
    for log_n in sizes:
        n = 1 << log_n
        dom = ab.EvaluationDomain.new(n)
        x = ab.gen_scalars_dev(n, 99, 0, True)
        y = x.clone()
        dom.fft_in_place_dev(y)
        changed = not torch.equal(x, y)
        dom.ifft_in_place_dev(y)
        ok1 = torch.equal(x, y)
        dom.coset_fft_in_place_dev(y)
        dom.coset_ifft_in_place_dev(y)
        ok2 = torch.equal(x, y)
        res["2^%d" % log_n] = {"changed": changed, "ifft_fft": bool(ok1), "coset_roundtrip": bool(ok2)}
        del x, y
    return res


@section("msm_parity_small")
def msm_small():
    res = {}
    for n in (0, 1, 5, 300, 3000):
        B = o.synthetic_bases(n, 40 + n)
        s = o.random_fr_vec(n, 41 + n)
        exp = o.g1_projective_to_bytes(o.msm_expected_from_dlogs(n, 40 + n, s) if n else None)
        for stride in (104, 96):
            got = ab.VariableBase.msm(o.g1_affine_vec_to_bytes(B, stride), o.fr_vec_to_bytes(s, mont=False), stride)
            res["n%d_s%d" % (n, stride)] = bool(got == exp)
    res["all_ok"] = all(res.values())
    return res


@section("msm_dlog")
def msm_dlog(sizes):
    res = {}
    for log_n in sizes:
        n = 1 << log_n
        s0, d = o.base_dlogs(n, 7000 + log_n)
        t0 = time.time()
        bases = ab.gen_bases_dev(n, s0, d, 0, 104)
        torch.cuda.synchronize()
        tgen = time.time() - t0
        oncurve = ab.check_on_curve_dev(bases, n, 104)
        # spot-check a few generated points against the oracle
        raw = bases.view(-1, 104)[[0, 1, n // 2, n - 1]].cpu().numpy().tobytes()
        spot = all(o.g1_affine_from_bytes(raw[k * 104:(k + 1) * 104]) == o.g1_mul(o.G1_GEN, (s0 + i * d) % o.R_MOD)
                   for k, i in enumerate([0, 1, n // 2, n - 1]))
        sc = ab.gen_scalars_dev(n, 31337 + log_n)
        k = ab.dlog_dot_dev(sc, n, s0, d)
        exp = o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, k))
        out = ab.VariableBase.msm_dev(bases, sc, n, 104)
        got = out.cpu().numpy().tobytes()
        med, best = ev_time(lambda: ab.VariableBase.msm_dev(bases, sc, n, 104, out), iters=3, warm=1)
        res["2^%d" % log_n] = {"ok": bool(got == exp), "on_curve": bool(oncurve), "spot": bool(spot), "gen_s": tgen,
                               "c": ab.VariableBase.window_bits(n), "ms_median": med, "ms_best": best,
                               "mpts_per_s": n / med / 1e3}
        del bases, sc
    return res


@section("ntt_timing")
def ntt_timing(sizes):
    res = {}
    for log_n in sizes:
        n = 1 << log_n
        dom = ab.EvaluationDomain.new(n)
        x = ab.gen_scalars_dev(n, 5, 0, True)
        r = {}
        for name in ("fft", "ifft", "coset_fft", "coset_ifft"):
            med, best = ev_time(lambda: getattr(dom, name + "_in_place_dev")(x), iters=5, warm=2)
            r[name] = {"ms": med, "best": best, "melem_per_s": n / med / 1e3}
        res["2^%d" % log_n] = r
        del x
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = ab.get_lib()
    lib.check(lib.init(0), "init")
    print(lib.version().decode(), torch.cuda.get_device_name(0), flush=True)
    imad()
    ntt_parity([1, 4, 10, 11, 12, 13] if args.quick else [1, 4, 10, 11, 12, 13, 15, 16, 17])
    msm_small()
    ntt_roundtrip([18, 20] if args.quick else [18, 20, 22, 24, 25, 26])
    msm_dlog([12, 16] if args.quick else [12, 16, 18, 20, 22, 24])
    ntt_timing([16, 20] if args.quick else [12, 16, 18, 20, 22, 24, 26])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
