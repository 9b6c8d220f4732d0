#!/usr/bin/env python3
"""bench.py -- BLS12-377 G1 MSM (headline) and Fr NTT (secondary) on B200, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n 24]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...        # CPU arm: the C oracle port on the host cores

A step = one G1 MSM over 2^log_n points per GPU (the rank's point range of an N*2^log_n-point MSM,
BASELINE.json config 4's decomposition) followed, for N > 1, by the single final combine (all-gather
of the 144-byte partials + g1_sum).  `value` = total points / max-over-ranks device time, inputs
resident in HBM.  `e2e` = the same work through the host-pointer C-ABI entry point
aleo_b200_msm_g1 on PAGEABLE host buffers (what a Rust Vec is; H2D of bases + scalars and D2H of the
result inside the timed region), with the same call on pinned buffers beside it.  The JSON line also
carries the NTT figures, both rooflines, BASELINE configs[3] (fixed 2^26 MSM + NTT over the N GPUs)
and the CPU baseline; `roofline.secondary` mirrors the secondary results compactly.
The oracle (oracle/) is used here only as the checker and as the timed CPU baseline.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MACS_PER_MIXED_ADD = 3000          # 10 Fq products x (2*12^2 + 12) 32x32->64 MACs (algorithmic)
ISSUED_MACS_PER_MIXED_ADD = 8 * 276 + 2 * 210   # IMAD.WIDE actually issued: 8 products + 2 dedicated squares (cuobjdump)
ISSUED_MACS_PER_AFFINE_ADD = 5 * 276 + 210      # batch-affine addition (csrc/msm_ba.cuh): 5 products + 1 square, inversion shared
NTT_BYTES_PER_ELEM_PER_PASS = 64   # 32 B read + 32 B write


_JSON_OUT = None  # the process's original stdout (main() points file descriptor 1 at stderr)

def load_c_oracle():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "oracle"], cwd=ROOT, check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.oracle_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    lib.oracle_gen_bases.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.oracle_ntt_fr.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    return lib


def cpu_msm_baseline(log_n_sample, steps, warmup=0, seed=0xA1E0B200 + 2, budget_s=None):
    """C oracle port (Pippenger, pthreads over windows) on all host cores; returns (Mpts/s, info).
    `steps` timed runs after `warmup` untimed ones; with budget_s the timed runs stop early (never below 3) once the
    budget is spent, and info says how many were timed."""
    import numpy as np

    from oracle import bls12_377 as o

    lib = load_c_oracle()
    cores = os.cpu_count() or 1
    n = 1 << log_n_sample
    s0, d = o.base_dlogs(n, seed)
    bases = np.zeros(n * 104, dtype=np.uint8)
    lib.oracle_gen_bases(bases.ctypes.data, n, 104, o.int_to_le_bytes(s0, 32), o.int_to_le_bytes(d, 32), 0, cores)
    rng = np.random.default_rng(seed)
    scalars = rng.integers(0, 2**63, size=(n, 4), dtype=np.int64).astype(np.uint64)
    scalars |= rng.integers(0, 2, size=(n, 4), dtype=np.uint64) << np.uint64(63)
    scalars[:, 3] &= np.uint64((1 << 60) - 1)          # 252 bits < r
    out = C.create_string_buffer(144)
    t_begin = time.perf_counter()
    for _ in range(max(0, warmup)):
        lib.oracle_msm_g1(out, bases.ctypes.data, n, scalars.ctypes.data, 104, cores)
    times = []
    for _ in range(max(1, steps)):
        t = time.perf_counter()
        lib.oracle_msm_g1(out, bases.ctypes.data, n, scalars.ctypes.data, 104, cores)
        times.append(time.perf_counter() - t)
        if budget_s is not None and len(times) >= 3 and time.perf_counter() - t_begin + times[-1] > budget_s:
            break
    # correctness of the timed code itself (known discrete logs) on a prefix the big-integer oracle finishes quickly
    sc = [int.from_bytes(scalars[i].tobytes(), "little") for i in range(min(n, 1 << 12))]
    chk = C.create_string_buffer(144)
    lib.oracle_msm_g1(chk, bases.ctypes.data, len(sc), scalars.ctypes.data, 104, cores)
    assert chk.raw == o.g1_projective_to_bytes(o.msm_expected_from_dlogs(len(sc), seed, sc)), "CPU baseline produced a wrong result"
    med = sorted(times)[len(times) // 2]
    return n / med / 1e6, {"cores": cores, "sample": "G1 MSM n=2^%d, uniform 252-bit scalars, median of %d timed run(s) after %d warm-up"
                                                     % (log_n_sample, len(times), max(0, warmup)),
                           "seconds_per_run": med, "timed_runs": len(times), "log_n": log_n_sample}


def cpu_ntt_baseline(log_n_sample, runs=3):
    """forward Fr NTT of the C oracle port on all host cores: one untimed call builds (and caches) the root table --
    snarkVM's prover hands its cached FFTPrecomputation to the transform as well --, then the median of `runs`"""
    import numpy as np

    lib = load_c_oracle()
    cores = os.cpu_count() or 1
    n = 1 << log_n_sample
    data = np.random.default_rng(5).integers(0, 2**60, size=(n, 4), dtype=np.int64).astype(np.uint64)
    lib.oracle_ntt_fr(data.ctypes.data, log_n_sample, 0, 0, cores)
    times = []
    for _ in range(max(1, runs)):
        t = time.perf_counter()
        lib.oracle_ntt_fr(data.ctypes.data, log_n_sample, 0, 0, cores)
        times.append(time.perf_counter() - t)
    med = sorted(times)[len(times) // 2]
    return n / med / 1e6, {"cores": cores, "sample": "Fr NTT n=2^%d forward, root table cached, median of %d runs" % (log_n_sample, len(times)),
                           "seconds_per_run": med}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.strip().lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(log_n, world):
    """the `config` object shared by both arms (BASELINE.json configs[1]: the sweep's largest single-GPU size)"""
    return {"workload": "BLS12-377 G1 MSM n=2^%d points per GPU (point range of an %d*2^%d-point MSM; bases (s0+i*d)G 104-byte "
                        "stride, uniform 252-bit scalars), bases+scalars resident in HBM; N>1 adds the single final combine"
                        % (log_n, world, log_n),
            "log_n": log_n, "cache": "inputs (%.1f GB per GPU) larger than L2" % ((1 << log_n) * 136 / 1e9),
            "parallelism": "point-range shard x%d" % world}


def run_reference(args):
    """CPU arm: the C restatement (oracle/oracle.c) on all host cores at the SAME workload as the GPU arm (one step = one
    G1 MSM over the 2^log_n points of a GPU's point range).  A step takes ~10-20 s at 2^24, so the timed steps stop once
    --cpu-budget-s is spent (never fewer than 3); `steps` reports how many were timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    log_n = args.log_n if args.cpu_log_n <= 0 else min(args.log_n, args.cpu_log_n)
    val, info = cpu_msm_baseline(log_n, args.steps, warmup=min(args.warmup, 1), budget_s=args.cpu_budget_s)
    ntt_val, ntt_info = cpu_ntt_baseline(min(log_n, 24))
    world = max(1, args.gpus)
    sample = info["sample"] + (" (the full per-GPU workload)" if log_n == args.log_n else " (BOUNDED SAMPLE of the 2^%d workload)" % args.log_n)
    if world > 1:
        sample += "; N > 1: the CPU runs ONE rank's point range, Mpts/s is per whole host"
    line = {
        "impl": "reference", "metric": "bls12_377_g1_msm_mpts_per_s", "value": val, "unit": "Mpts/s", "n_gpus": args.gpus,
        "steps": info["timed_runs"], "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": info["seconds_per_run"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (384-bit Montgomery Fq)",
        "data": "synthetic", "config": workload_config(args.log_n, world),
        "cpu_baseline": {"value": val, "unit": "Mpts/s", "cores": info["cores"], "kind": "port", "sample": sample,
                         "note": "C restatement of the Pippenger algorithm class (oracle/oracle.c, windows in parallel like "
                                 "snarkVM's standard::msm); the snarkVM Rust binary cannot be built here (no Rust toolchain)"},
        "ntt": {"value": ntt_val, "unit": "Melem/s", **ntt_info},
        "e2e": {"value": val, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def proof_shaped_throughput(ab, o, torch, dev, world, dist, args):
    """SYNTHETIC stand-in for BASELINE.json configs 3 / 5 (end-to-end Varuna proofs need snarkVM's Rust, absent here;
    SURVEY.md 8d).  One "proof-shaped unit" = the MSM / NTT schedule of a 1-circuit proof as decoded from the
    reference's fixture (13 large MSMs) at proof-realistic sizes, all operands device resident:
    13 KZG commitments against a resident 2^18-point SRS (3 x 2^18, 4 x 2^17, 6 x 2^16 coefficients), 8 ifft 2^17,
    8 coset_fft 2^18, 8 pointwise products 2^18, 4 coset_ifft 2^18.  `threads` host threads with one CUDA stream
    each submit units concurrently (snarkVM commits on a rayon ExecutionPool; the library is re-entrant), every GPU
    runs its own replica stream (no collective).  Not a proof: no witness synthesis, no transcript."""
    import threading

    n_srs = 1 << 18
    s0, d = o.base_dlogs(n_srs, 777)
    srs_bases = ab.gen_bases_dev(n_srs, s0, d, 0, 104, device=dev)
    srs = ab.ResidentSRS.from_device(srs_bases, n_srs, 104)
    del srs_bases
    threads = max(1, args.proof_threads)
    units_per_thread = 2
    commit_sizes = [18] * 3 + [17] * 4 + [16] * 6
    d17, d18 = ab.EvaluationDomain.new(1 << 17), ab.EvaluationDomain.new(1 << 18)

    class Work:
        def __init__(self, seed):
            self.stream = torch.cuda.Stream(device=dev)
            self.polys = {ln: ab.gen_scalars_dev(1 << ln, seed + ln, 0, True, device=dev) for ln in (16, 17, 18)}
            self.b17 = ab.gen_scalars_dev(8 << 17, seed + 1, 0, True, device=dev)
            self.b18 = ab.gen_scalars_dev(8 << 18, seed + 2, 0, True, device=dev)
            self.c18 = ab.gen_scalars_dev(8 << 18, seed + 3, 0, True, device=dev)
            self.out = torch.empty((13, 48), dtype=torch.uint8, device=dev)

        def unit(self):
            with torch.cuda.stream(self.stream):
                d17.ifft_in_place_dev(self.b17, batch=8)
                d18.coset_fft_in_place_dev(self.b18, batch=8)
                ab.Evaluations.mul(self.b18, self.c18, out=self.b18)
                d18.coset_ifft_in_place_dev(self.b18[: 4 << 18], batch=4)
                if batched:
                    ab.KZG10.commit_batch_dev(srs, [self.polys[ln] for ln in commit_sizes], out=self.out)
                else:
                    for k, ln in enumerate(commit_sizes):
                        ab.KZG10.commit_dev(srs, self.polys[ln], 1 << ln, out=self.out[k])

    batched = True
    works = [Work(1000 * (t + 1)) for t in range(threads)]
    # one commitment checked against the oracle through known discrete logs: sum_i c_i (s0 + i d) G, compressed
    w0 = works[0]
    canon = o.fr_vec_from_bytes(w0.polys[16][:256].cpu().numpy().tobytes())
    small = ab.KZG10.commit_dev(srs, w0.polys[16], 256).cpu().numpy().tobytes()
    ok = small == o.g1_compress(o.g1_mul(o.G1_GEN, sum(c * (s0 + i * d) for i, c in enumerate(canon)) % o.R_MOD))

    def run(nunits):
        def body(w):
            torch.cuda.set_device(dev)
            for _ in range(nunits):
                w.unit()
            w.stream.synchronize()
        ts = [threading.Thread(target=body, args=(w,)) for w in works]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(1)                                   # warm-up: plans, pool growth
    secs = run(units_per_thread)
    serial = None
    if world == 1:
        t0 = time.perf_counter()
        for _ in range(2):
            works[0].unit()
        works[0].stream.synchronize()
        serial = (time.perf_counter() - t0) / 2
    one_by_one = None
    if world == 1:                           # the same stream of units with 13 separate commit calls per unit
        batched = False
        run(1)
        one_by_one = threads * units_per_thread / run(units_per_thread)
        batched = True
    # the batched commitments must equal the one-by-one ones
    w0.unit()
    w0.stream.synchronize()
    ref = torch.stack([ab.KZG10.commit_dev(srs, w0.polys[ln], 1 << ln) for ln in commit_sizes])
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(ref, w0.out))
    secs_t = torch.tensor([secs], device=dev)
    if dist is not None:
        dist.all_reduce(secs_t, op=dist.ReduceOp.MAX)
    units = world * threads * units_per_thread
    srs.close()
    return {"label": "SYNTHETIC proof-shaped units (13 resident-SRS KZG commits 2^16..2^18 + 20 NTTs 2^17..2^18 + 8 pointwise "
                     "products); NOT Varuna proofs -- BASELINE configs 3 / 5 need snarkVM's Rust toolchain",
            "value": units / secs_t.item(), "unit": "units/s", "units": units, "host_threads_per_gpu": threads,
            "ms_per_unit_one_stream": None if serial is None else serial * 1e3,
            "commits": "aleo_b200_kzg_commit_batch_dev: the 13 commitments of a unit in one launch sequence",
            "units_per_s_with_13_separate_commit_calls": one_by_one,
            "commit_checked_against_oracle": bool(ok), "replicas": world}


def ntt_spot_check(ab, o, torch, x_in, X_out, log_n, picks=(1, 0x2F3A7, -5)):
    """size-independent parity of a forward transform: X[k] must equal the polynomial x evaluated at w^k
    (aleo_b200_fr_poly_eval_dev, an independent kernel that tests/test_gpu_poly.py pins against the oracle)"""
    n = 1 << log_n
    w = o.fr_root_of_unity(log_n)
    ok = True
    for k in picks:
        k %= n
        got = int.from_bytes(X_out[k].cpu().numpy().tobytes(), "little")
        ok = ok and got == o.fr_to_mont(ab.poly.evaluate_dev(x_in, pow(w, k, o.R_MOD)))
    return bool(ok)


def dist_ntt_block(ab, o, torch, dist, dev, rank, world, glog, args, barrier, scaling):
    """ONE Fr NTT of 2^glog over `world` GPUs (aleo_b200_ntt_dist_*: the exchange is fused into the last-but-one pass).
    Parity: every rank builds the same seeded 2^glog vector, runs the single-GPU transform on it (itself spot-checked
    against polynomial evaluation) and compares ITS WHOLE output block of the distributed transform with it, for the
    forward and the coset-inverse kind; asserted."""
    from aleo_b200 import dist as adist

    n_g = 1 << glog
    x_full = ab.gen_scalars_dev(n_g, 78, 0, True, device=dev)          # same vector on every rank
    dom = ab.EvaluationDomain.new(n_g)
    peer = adist.PeerNTT(glog)
    blk = peer.input_block(x_full)
    oks = {}
    for name, inverse, coset in (("forward", False, False), ("coset_inverse", True, True)):
        want = dom._run_dev(x_full.clone(), 1 if inverse else 0, 1 if coset else 0)
        if name == "forward":
            oks["single_gpu_vs_poly_eval"] = ntt_spot_check(ab, o, torch, x_full, want, glog)
        got = peer.transform(blk, inverse=inverse, coset=coset)
        oks[name] = bool(torch.equal(got.reshape(-1), peer.output_block_of(want).reshape(-1)))
        del want, got
    del x_full
    flags = [None] * world
    dist.all_gather_object(flags, oks)
    parity = all(all(f.values()) for f in flags)
    assert parity, "distributed NTT: a rank's block differs from the single-GPU transform: %r" % (flags,)
    xo = torch.empty_like(blk).reshape(-1, 4)
    for _ in range(args.warmup):
        peer.transform(blk, out=xo)
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d0.record()
    for _ in range(args.steps):
        peer.transform(blk, out=xo)
    d1.record()
    barrier()
    dms = torch.tensor([d0.elapsed_time(d1)], device=dev)
    dist.all_reduce(dms, op=dist.ReduceOp.MAX)
    dstep = dms.item() / args.steps
    stage_ms = peer.stage_times(blk, xo) if hasattr(peer, "stage_times") else None
    peer_passes, peer_rf, peer_rl = peer.passes, peer.log_r_first, peer.log_r_last
    peer.close()
    del xo
    # the same transform with the exchange as a separate NCCL all-to-all (aleo_b200/dist.py ntt_four_step)
    l1, l2 = adist.four_step_shape(glog)
    g1, g2 = 1 << l1, 1 << l2
    xb = blk.reshape(-1)[: (g1 * (g2 // world)) * 4].reshape(g1, g2 // world, 4).contiguous()
    del blk
    for _ in range(args.warmup):
        adist.ntt_four_step(xb, glog)
    barrier()
    d0.record()
    for _ in range(args.steps):
        adist.ntt_four_step(xb, glog)
    d1.record()
    barrier()
    nms = torch.tensor([d0.elapsed_time(d1)], device=dev)
    dist.all_reduce(nms, op=dist.ReduceOp.MAX)
    nstep = nms.item() / args.steps
    del xb
    torch.cuda.empty_cache()
    return {"metric": "bls12_377_fr_ntt_melem_per_s", "value": n_g / dstep / 1e3, "unit": "Melem/s", "ms_per_step": dstep,
            "log_n_global": glog, "passes": peer_passes, "scaling": scaling,
            "schedule": "single-GPU pass split %d passes (R_first 2^%d, R_last 2^%d); the last-but-one pass stores into the peers' "
                        "receive buffers (CUDA IPC peer memory over NVLink); the last pass runs from the receive buffer"
                        % (peer_passes, peer_rf, peer_rl),
            "exchange_bytes_per_rank": n_g // world * 32 * (world - 1) // world,
            "parity": parity, "parity_what": "every rank's whole output block == single-GPU transform (forward and coset inverse); "
                                             "single-GPU transform spot-checked against polynomial evaluation",
            "stage_ms": stage_ms,
            "nccl_all_to_all_schedule": {"value": n_g / nstep / 1e3, "unit": "Melem/s", "ms_per_step": nstep,
                                         "what": "four-step %dx%d with torch transposes around one all_to_all_single (NCCL)" % (g1, g2)}}


def config4_fixed_size(ab, o, torch, dist, dev, rank, world, args, barrier):
    """BASELINE.json configs[3] as written (strong scaling): ONE G1 MSM of 2^26 points sharded by contiguous point range
    (2^26 / N per GPU, single final combine) and ONE Fr NTT of 2^26 elements over the N GPUs (N = 1: the single-GPU
    kernels).  Both checked: the MSM through known discrete logs, the NTT block by block against the single-GPU
    transform, which is spot-checked against polynomial evaluation."""
    lt = args.config4_log_n
    n_tot = 1 << lt
    n = n_tot // world
    first = rank * n
    seed = 0xA1E0B200 + 4
    steps = max(1, min(args.steps, 5))
    s0, d = o.base_dlogs(n_tot, seed)
    bases = ab.gen_bases_dev(n, s0, d, first, 104, device=dev)
    scalars = ab.gen_scalars_dev(n, seed, first, False, device=dev)
    partial = torch.empty(144, dtype=torch.uint8, device=dev)
    gathered = torch.empty(144 * world, dtype=torch.uint8, device=dev)
    result = torch.empty(144, dtype=torch.uint8, device=dev)

    def step():
        ab.VariableBase.msm_dev(bases, scalars, n, 104, out=partial)
        if world > 1:
            dist.all_gather_into_tensor(gathered, partial)
            ab.VariableBase.sum_partials_dev(gathered, world, out=result)
            return result
        return partial

    ks = [ab.dlog_dot_dev(scalars, n, s0, d, first)]
    if world > 1:
        mine = ks[0]
        ks = [None] * world
        dist.all_gather_object(ks, mine)
    msm_ok = step().cpu().numpy().tobytes() == o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, sum(ks) % o.R_MOD))
    assert msm_ok, "config 4: sharded MSM result does not match the oracle"
    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    msm_ms = ms.item() / steps
    c_bits = ab.VariableBase.window_bits(n)
    del bases, scalars
    torch.cuda.empty_cache()
    out = {"what": "BASELINE.json configs[3]: fixed-size 2^%d MSM by point range + 2^%d NTT over N GPUs (strong scaling)" % (lt, lt),
           "n_gpus": world, "scaling": "strong",
           "msm": {"log_n_total": lt, "points_per_gpu": n, "window_bits": c_bits, "ms_per_step": msm_ms, "value": n_tot / msm_ms / 1e3,
                   "unit": "Mpts/s", "checked_against_oracle": bool(msm_ok)}}
    if world == 1:
        dom = ab.EvaluationDomain.new(n_tot)
        x = ab.gen_scalars_dev(n_tot, 78, 0, True, device=dev)
        X = x.clone()
        dom.fft_in_place_dev(X)
        ntt_ok = ntt_spot_check(ab, o, torch, x, X, lt)
        dom.ifft_in_place_dev(X)
        ntt_ok = ntt_ok and bool(torch.equal(X, x))
        assert ntt_ok, "config 4: 2^%d NTT failed its polynomial-evaluation / round-trip check" % lt
        del X
        dom.fft_in_place_dev(x)
        barrier()
        e0.record()
        for _ in range(steps):
            dom.fft_in_place_dev(x)
        e1.record()
        barrier()
        ntt_ms = e0.elapsed_time(e1) / steps
        del x
        out["ntt"] = {"log_n_total": lt, "ms_per_step": ntt_ms, "value": n_tot / ntt_ms / 1e3, "unit": "Melem/s", "passes": dom.launches(),
                      "parity": bool(ntt_ok), "parity_what": "3 outputs == polynomial evaluation at w^k, and ifft(fft(x)) == x"}
    elif (world & (world - 1)) == 0 and world <= 8:
        nd = dist_ntt_block(ab, o, torch, dist, dev, rank, world, lt, args, barrier, scaling="strong")
        out["ntt"] = {"log_n_total": lt, "ms_per_step": nd["ms_per_step"], "value": nd["value"], "unit": "Melem/s", "passes": nd["passes"],
                      "parity": nd["parity"], "parity_what": nd["parity_what"], "stage_ms": nd["stage_ms"],
                      "nccl_all_to_all_ms": nd["nccl_all_to_all_schedule"]["ms_per_step"]}
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import aleo_b200 as ab
    from oracle import bls12_377 as o          # checker only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ab.get_lib()                         # raises if the CUDA library was not built: no fallback
    lib.check(lib.init(local), "aleo_b200_init")

    log_n = args.log_n
    n = 1 << log_n
    seed = 0xA1E0B200 + 2
    s0, d = o.base_dlogs(n * world, seed)
    first = rank * n
    bases = ab.gen_bases_dev(n, s0, d, first, 104, device=dev)
    scalars = ab.gen_scalars_dev(n, seed, first, False, device=dev)
    stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    partial = torch.empty(144, dtype=torch.uint8, device=dev)
    gathered = torch.empty(144 * world, dtype=torch.uint8, device=dev)
    result = torch.empty(144, dtype=torch.uint8, device=dev)

    def step():
        ab.VariableBase.msm_dev(bases, scalars, n, 104, out=partial)
        if world > 1:
            dist.all_gather_into_tensor(gathered, partial)
            ab.VariableBase.sum_partials_dev(gathered, world, out=result)
            return result
        return partial

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of exactly what is timed: known discrete logs ------------------------------
    assert ab.check_on_curve_dev(bases, n, 104), "generated bases are not on the curve"
    k_local = ab.dlog_dot_dev(scalars, n, s0, d, first)
    ks = [k_local]
    if world > 1:
        ks = [None] * world
        dist.all_gather_object(ks, k_local)
    got = step().cpu().numpy().tobytes()
    want = o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, sum(ks) % o.R_MOD))
    checked = bool(got == want)
    assert checked, "MSM result does not match the oracle"

    # ---- timed region: device-resident ------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = world * n / ms_step / 1e3               # Mpts/s, whole job

    # ---- per-kernel timing for the roofline (same kernels, events around the phases) --------------
    ph = (C.c_float * 3)()
    acc_ms, sort_ms, tail_ms = [], [], []
    for _ in range(max(2, args.steps)):
        lib.check(lib.msm_g1_dev_profile(partial.data_ptr(), bases.data_ptr(), n, scalars.data_ptr(), 104, stream(), ph), "profile")
        sort_ms.append(ph[0]); acc_ms.append(ph[1]); tail_ms.append(ph[2])
    acc = sum(acc_ms) / len(acc_ms)
    c_bits = ab.VariableBase.window_bits(n)
    windows = 253 // c_bits + 1
    entry_windows = -(-252 // c_bits)      # half-range scalars have 252 bits: the window above only holds the recoding carry
    ba_levels = ab.VariableBase.batch_affine_levels(n)
    ba_share = 1.0 - 0.5 ** ba_levels       # additions done by the batch-affine levels (uniform scalars)
    issued_per_add = ba_share * ISSUED_MACS_PER_AFFINE_ADD + (1.0 - ba_share) * ISSUED_MACS_PER_MIXED_ADD
    macs = n * entry_windows * MACS_PER_MIXED_ADD
    ms_i, ops_i = C.c_double(), C.c_double()
    lib.check(lib.bench_imad(2, 4096, C.byref(ms_i), C.byref(ops_i)), "bench_imad")
    peak_gmac = ops_i.value / ms_i.value / 1e6          # IMAD.WIDE.U32.X issue rate, measured live
    prof = {}
    ppath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(ppath):
        prof = json.load(open(ppath))
    roofline = {"kernel": "MSM accumulation phase: msm::ba::level_kernel x %d + msm::accumulate_kernel" % ba_levels if ba_levels else "msm::accumulate_kernel",
                "bound": "int", "achieved": macs / acc / 1e6, "peak": peak_gmac, "unit": "GMAC/s",
                "frac": macs / acc / 1e6 / peak_gmac,
                "pipe_frac": n * entry_windows * issued_per_add / acc / 1e6 / peak_gmac,
                "issued_macs_per_mixed_add": ISSUED_MACS_PER_MIXED_ADD, "issued_macs_per_affine_add": ISSUED_MACS_PER_AFFINE_ADD,
                "batch_affine_levels": ba_levels, "batch_affine_share_of_additions": ba_share, "issued_macs_per_addition": issued_per_add,
                "traffic": (prof["msm_accumulate_dram_bytes_per_entry"] * n * entry_windows) if "msm_accumulate_dram_bytes_per_entry" in prof else None,
                "traffic_source": prof.get("msm_accumulate_source", "ncu --set full at n=2^22 via profiles/ncu_traffic.json: DRAM bytes per sorted entry x n x windows with entries"),
                "algorithmic_bytes_per_launch": n * entry_windows * 108,
                "algorithmic_macs_per_launch": macs, "window_bits": c_bits, "windows": windows, "windows_with_entries": entry_windows,
                "kernel_ms": acc, "kernel_share_of_step": acc / (sum(sort_ms) / len(sort_ms) + acc + sum(tail_ms) / len(tail_ms)),
                "phases_ms": {"recode_sort_plan": sum(sort_ms) / len(sort_ms), "accumulate": acc, "combine_reduce_final": sum(tail_ms) / len(tail_ms)},
                "peak_source": "measured live: carry-chained IMAD.WIDE.U32.X microbenchmark in this library (MEASURED_PEAKS.json "
                               "has no integer-pipe figure); 32-bit MAC = one IMAD.WIDE",
                "note": "MSM is integer-pipe bound (SURVEY.md 8d); the schema's hbm/tensor bounds do not apply to this kernel. "
                        "frac counts SURVEY's ALGORITHMIC 3000 MACs per bucket addition (n x windows that receive entries); the "
                        "kernels issue fewer (batch-affine additions with a shared inversion: 1590; XYZZ mixed additions with a "
                        "dedicated squaring: 2628), so frac exceeds the share of the pipe they occupy: pipe_frac"}

    # ---- skewed scalars (real KZG witnesses are full of 0 / 1 / small and "negative small" values) -------
    witness = None
    if world == 1:
        ws = scalars.clone()
        idx = torch.arange(n, device=dev)
        ws[idx % 2 == 0] = 0                                                     # 50 % zero
        ws[idx % 4 == 1] = torch.tensor([1, 0, 0, 0], dtype=torch.int64, device=dev)   # 25 % one
        minus = np.frombuffer(o.int_to_le_bytes(o.R_MOD - 2, 32), dtype=np.int64).copy()
        ws[idx % 16 == 3] = torch.from_numpy(minus).to(dev)                       # 6 % "-2" (r - 2); the rest uniform
        kw = ab.dlog_dot_dev(ws, n, s0, d, first)
        w_ok = ab.VariableBase.msm_dev(bases, ws, n, 104).cpu().numpy().tobytes() == o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, kw))
        torch.cuda.synchronize()
        w0e, w1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0e.record()
        for _ in range(args.steps):
            ab.VariableBase.msm_dev(bases, ws, n, 104, out=partial)
        w1e.record()
        torch.cuda.synchronize()
        wms = w0e.elapsed_time(w1e) / args.steps
        witness = {"what": "same MSM with witness-like scalars: 50 % zero, 25 % one, 6 % r - 2, rest uniform (hot buckets; the flat "
                           "run scheme keeps every lane equally loaded)", "value": n / wms / 1e3, "unit": "Mpts/s", "ms_per_step": wms,
                   "nonzero_fraction": 0.5, "checked_against_oracle": bool(w_ok)}
        del ws, idx

    # ---- resident SRS (KZG10::commit shape): bases expanded once, one shared bucket set ----------------
    srs_obj = None
    if not args.no_srs:
        t_exp = time.perf_counter()
        srs = ab.ResidentSRS.from_device(bases, n, 104)
        torch.cuda.synchronize()
        t_exp = time.perf_counter() - t_exp
        sinfo = srs.info()
        srs_ok = srs.msm_dev(scalars, n).cpu().numpy().tobytes() == o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, k_local))
        for _ in range(args.warmup):
            srs.msm_dev(scalars, n, out=partial)
        barrier()
        s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0e.record()
        for _ in range(args.steps):
            srs.msm_dev(scalars, n, out=partial)
        s1e.record()
        barrier()
        sms = torch.tensor([s0e.elapsed_time(s1e)], device=dev)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        sstep = sms.item() / args.steps
        sph = (C.c_float * 3)()
        lib.check(lib.srs_msm_dev_profile(srs._h, partial.data_ptr(), scalars.data_ptr(), n, stream(), sph), "srs profile")
        hs_pin = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        hs_pin.copy_(scalars)
        torch.cuda.synchronize()
        srs.msm(hs_pin)
        barrier()
        t0s = time.perf_counter()
        for _ in range(args.steps):
            srs.msm(hs_pin)                      # H2D of the scalars + MSM + D2H, bases stay resident
        barrier()
        e2e_srs = torch.tensor([time.perf_counter() - t0s], device=dev)
        if world > 1:
            dist.all_reduce(e2e_srs, op=dist.ReduceOp.MAX)
        wins = sinfo["windows"]
        srs_obj = {"what": "same MSM over a resident SRS handle (aleo_b200_srs_*): bases expanded once to 2^(cw)P_i, shared buckets; "
                           "this is KZG10::commit's shape (fixed powers_of_beta_g)", "value": world * n / sstep / 1e3, "unit": "Mpts/s",
                   "ms_per_step": sstep, "window_bits": sinfo["window_bits"], "windows": wins, "device_bytes": sinfo["device_bytes"],
                   "expand_seconds_once": t_exp, "checked_against_oracle": bool(srs_ok),
                   "phases_ms": {"recode_sort_plan": sph[0], "accumulate": sph[1], "combine_reduce_final": sph[2]},
                   "accumulate_int_frac": n * wins * MACS_PER_MIXED_ADD / sph[1] / 1e6 / peak_gmac,
                   "e2e_scalars_only": {"value": world * n * args.steps / e2e_srs.item() / 1e6, "unit": "Mpts/s",
                                        "h2d_bytes_per_step": world * n * 32, "d2h_bytes_per_step": world * 144,
                                        "api": "aleo_b200_srs_msm (host scalars, pinned)"},
                   "gpu_launches_per_step": srs.launches(n)}
        srs.close()
        del srs, hs_pin
        torch.cuda.empty_cache()

    # ---- e2e through the host-pointer C-ABI call (H2D of bases + scalars and D2H of the result inside the timed region).
    #      Headline = PAGEABLE caller memory (what a Rust Vec is: the library stages it through pinned buffers with helper
    #      threads); the same call on pinned buffers is reported beside it ------------------------------------------------
    hb = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
    hs = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hb.copy_(bases)
    hs.copy_(scalars)
    torch.cuda.synchronize()
    host_parts = torch.empty(144 * world, dtype=torch.uint8, device=dev)

    def e2e_step(b, sc):
        raw = ab.VariableBase.msm(b, sc, 104)              # H2D + MSM + D2H inside
        if world > 1:
            mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
            dist.all_gather_into_tensor(host_parts, mine)
            return ab.VariableBase.sum_partials_dev(host_parts, world).cpu().numpy().tobytes()
        return raw

    def e2e_time(b, sc):
        ok = e2e_step(b, sc) == want                       # warm-up + check
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(b, sc)
        barrier()
        secs = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        return world * n * args.steps / secs.item() / 1e6, bool(ok)

    e2e_pinned, pin_ok = e2e_time(hb, hs)
    pb, ps = hb.numpy().copy(), hs.numpy().copy()          # plain malloc'ed (pageable) numpy buffers
    del hb, hs
    e2e_val, pg_ok = e2e_time(pb, ps)
    e2e_ok = pin_ok and pg_ok
    del pb, ps
    host_plan = ab.VariableBase.host_plan(n)

    # ---- secondary: Fr NTT of the same size, resident ---------------------------------------------
    del bases
    torch.cuda.empty_cache()
    dom = ab.EvaluationDomain.new(n)
    x = ab.gen_scalars_dev(n, 77, first, True, device=dev)
    x0 = x.clone()
    dom.fft_in_place_dev(x)
    dom.ifft_in_place_dev(x)
    ntt_ok = bool(torch.equal(x, x0))
    for _ in range(args.warmup):
        dom.fft_in_place_dev(x)
    barrier()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record()
    for _ in range(args.steps):
        dom.fft_in_place_dev(x)
    n1.record()
    barrier()
    ntt_ms = torch.tensor([n0.elapsed_time(n1)], device=dev)
    if world > 1:
        dist.all_reduce(ntt_ms, op=dist.ReduceOp.MAX)
    ntt_step = ntt_ms.item() / args.steps
    pm = (C.c_float * 4)()
    passes = []
    for _ in range(max(2, args.steps)):
        lib.check(lib.ntt_fr_dev_profile(x.data_ptr(), log_n, 0, 0, stream(), pm), "ntt profile")
        passes.append([pm[i] for i in range(dom.launches())])
    pass_avg = sum(sum(p) for p in passes) / (len(passes) * dom.launches())
    peaks = {}
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        peaks = json.load(open(mp))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ntt_bytes = NTT_BYTES_PER_ELEM_PER_PASS * n
    # DESIGN.md "NTT arithmetic": 3.25 products per element and pass inside the tile; a pass boundary costs 1 (apply) when
    # its twiddles come from a direct table, 2 (two-level table product + apply) otherwise
    npass = dom.launches()
    k_bits = [log_n // npass + (1 if i < log_n % npass else 0) for i in range(npass)]
    mults_per_elem, log_cur = 3.25 * npass, log_n
    for i in range(npass - 1):
        direct = (12 <= log_n <= 24) if i == 0 else (log_cur <= 18)      # ntt_host.cuh: DIRECT0_* / DIRECT_TW_MAX_LOG
        mults_per_elem += 1 if direct else 2
        log_cur -= k_bits[i]
    ntt = {"metric": "bls12_377_fr_ntt_melem_per_s", "value": world * n / ntt_step / 1e3, "unit": "Melem/s", "ms_per_step": ntt_step,
           "log_n": log_n, "passes": dom.launches(), "round_trip_ok": ntt_ok,
           "roofline": {"kernel": "ntt::pass_kernel", "bound": "hbm", "achieved": ntt_bytes / pass_avg / 1e6, "peak": hbm_peak, "unit": "GB/s",
                        "frac": ntt_bytes / pass_avg / 1e6 / hbm_peak, "traffic": (prof["ntt_pass_dram_bytes_at_2p22"] * n / (1 << 22)) if "ntt_pass_dram_bytes_at_2p22" in prof else None,
                        "algorithmic_bytes_per_launch": ntt_bytes, "kernel_ms": pass_avg,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)",
                        "int_frac": n * mults_per_elem * 120 / (ntt_step * 1e-3) / 1e9 / peak_gmac,
                        "note": "the 253-bit NTT is integer-pipe bound (about %.1f Fr products of 120 IMAD.WIDE per element); int_frac is "
                                "measured against the same live IMAD.WIDE peak as the MSM" % mults_per_elem}}

    # the same transform through the host-pointer entry point (aleo_b200_ntt_fr on a pinned host buffer: H2D + 3 passes +
    # D2H inside the timed region) -- PCIe bound, which is why the prover-side helpers keep polynomials resident
    if world == 1:
        hx = torch.empty((n, 4), dtype=torch.int64).pin_memory()
        hx.copy_(x)
        torch.cuda.synchronize()
        want_h = x.clone()
        dom.fft_in_place_dev(want_h)
        dom.ntt_host_buffer(hx, 0, 0)                      # warm-up + check
        h_ok = bool(torch.equal(hx, want_h.cpu()))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dom.ntt_host_buffer(hx, 0, 0)
        h_s = (time.perf_counter() - t0) / args.steps
        ntt["e2e"] = {"value": n / h_s / 1e6, "unit": "Melem/s", "ms_per_step": h_s * 1e3, "h2d_bytes_per_step": n * 32,
                      "d2h_bytes_per_step": n * 32, "api": "aleo_b200_ntt_fr (host pointer, pinned)", "matches_resident_path": h_ok}
        del hx, want_h

    # ---- N > 1: ONE NTT of 2^(log_n + log2 N) over the N GPUs: exchange fused into the transform (peer-memory stores),
    #      with the NCCL all-to-all schedule timed beside it; every rank's block is compared with the single-GPU transform ----
    ntt_dist = None
    del x, x0
    torch.cuda.empty_cache()
    if world > 1 and (world & (world - 1)) == 0 and world <= 8:
        ntt_dist = dist_ntt_block(ab, o, torch, dist, dev, rank, world, log_n + world.bit_length() - 1, args, barrier, scaling="weak")

    # ---- BASELINE.json configs[3] as written: a FIXED 2^26 MSM sharded by point range and a FIXED 2^26 NTT over the N GPUs
    config4 = None
    if not args.no_config4:
        config4 = config4_fixed_size(ab, o, torch, dist, dev, rank, world, args, barrier)

    # ---- size sweep 2^16 .. 2^22 (BASELINE.json metric range / config 2), single-GPU run only -----------------
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        for ln in range(16, min(log_n, 23), 2):
            m = 1 << ln
            sb = ab.gen_bases_dev(m, s0, d, 0, 104, device=dev)
            ss = ab.gen_scalars_dev(m, seed + ln, 0, False, device=dev)
            want_s = o.g1_projective_to_bytes(o.g1_mul(o.G1_GEN, ab.dlog_dot_dev(ss, m, s0, d, 0)))
            ok_s = ab.VariableBase.msm_dev(sb, ss, m, 104).cpu().numpy().tobytes() == want_s
            reps = 10
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                ab.VariableBase.msm_dev(sb, ss, m, 104, out=partial)
            a1.record()
            torch.cuda.synchronize()
            msm_ms = a0.elapsed_time(a1) / reps
            sd = ab.EvaluationDomain.new(m)
            sx = ab.gen_scalars_dev(m, 5 + ln, 0, True, device=dev)
            sd.fft_in_place_dev(sx)
            torch.cuda.synchronize()
            a0.record()
            for _ in range(reps):
                sd.fft_in_place_dev(sx)
            a1.record()
            torch.cuda.synchronize()
            ntt_ms_s = a0.elapsed_time(a1) / reps
            sweep.append({"log_n": ln, "msm_ms": msm_ms, "msm_mpts_per_s": m / msm_ms / 1e3, "msm_window_bits": ab.VariableBase.window_bits(m),
                          "msm_checked_against_oracle": bool(ok_s), "ntt_ms": ntt_ms_s, "ntt_melem_per_s": m / ntt_ms_s / 1e3})
            del sb, ss, sx

    # ---- synthetic proof-shaped stream (stand-in for BASELINE configs 3 / 5, which need snarkVM's Rust) ----------
    proof_shaped = None
    if not args.no_proof_shape:
        proof_shaped = proof_shaped_throughput(ab, o, torch, dev, world, dist if world > 1 else None, args)

    # ---- CPU baseline (rank 0, single-GPU run only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, info = cpu_msm_baseline(min(log_n, args.cpu_sample_log_n), 3, warmup=1)
        cpu = {"value": val, "unit": "Mpts/s", "cores": info["cores"], "kind": "port",
               "sample": info["sample"] + " (BOUNDED SAMPLE of the 2^%d workload; `--impl reference` times the full size)" % log_n}

    if rank == 0:
        # the integer roof without trusting the probe: SMs x 32 IMAD.WIDE lanes per clock x the SM clock sampled under load
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        roofline["peak_theoretical"] = sms * 32 * mhz * 1e-3                      # GMAC/s
        roofline["peak_theoretical_what"] = "%d SMs x 32 IMAD.WIDE/clk/SM (half-rate FMA-heavy pipe) x %.0f MHz (clocks.sm_mhz)" % (sms, mhz)
        roofline["frac_of_theoretical"] = roofline["achieved"] / roofline["peak_theoretical"]
        roofline["pipe_frac_of_theoretical"] = roofline["pipe_frac"] * roofline["peak"] / roofline["peak_theoretical"]
        # compact mirror of the secondary results (the driver's record keeps `roofline` whole)
        r3 = lambda v: None if v is None else round(v, 4)  # noqa: E731
        sec = {"ntt_2p%d" % log_n: {"ms": r3(ntt_step), "melem_s": r3(ntt["value"]), "int_frac": r3(ntt["roofline"]["int_frac"]),
                                     "hbm_frac": r3(ntt["roofline"]["frac"]), "round_trip_ok": ntt_ok}}
        if ntt_dist:
            sec["ntt_dist_weak"] = {"log_n": ntt_dist["log_n_global"], "ms": r3(ntt_dist["ms_per_step"]), "melem_s": r3(ntt_dist["value"]),
                                    "parity": ntt_dist["parity"], "nccl_a2a_ms": r3(ntt_dist["nccl_all_to_all_schedule"]["ms_per_step"]),
                                    "stage_ms": ntt_dist["stage_ms"]}
        if config4:
            sec["config4_msm_2p%d" % args.config4_log_n] = {"ms": r3(config4["msm"]["ms_per_step"]), "mpts_s": r3(config4["msm"]["value"]),
                                                             "ok": config4["msm"]["checked_against_oracle"], "scaling": "strong"}
            if "ntt" in config4:
                sec["config4_ntt_2p%d" % args.config4_log_n] = {"ms": r3(config4["ntt"]["ms_per_step"]), "melem_s": r3(config4["ntt"]["value"]),
                                                                 "parity": config4["ntt"]["parity"], "scaling": "strong"}
        if sweep:
            sec["sweep_ms"] = {"2p%d" % e["log_n"]: [r3(e["msm_ms"]), r3(e["ntt_ms"]), e["msm_checked_against_oracle"]] for e in sweep}
            sec["sweep_ms_what"] = "[msm ms, ntt ms, msm checked]"
        sec["e2e_pinned_mpts_s"] = r3(e2e_pinned)
        if srs_obj:
            sec["srs_resident_mpts_s"] = r3(srs_obj["value"])
        roofline["secondary"] = sec
        line = {
            "metric": "bls12_377_g1_msm_mpts_per_s", "value": value, "unit": "Mpts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (384-bit Montgomery Fq, 256-bit Fr)", "data": "synthetic",
            "config": workload_config(log_n, world),
            "checked_against_oracle": checked and e2e_ok,
            "e2e": {"value": e2e_val, "unit": "Mpts/s", "h2d_bytes_per_step": world * n * 136, "d2h_bytes_per_step": world * 144,
                    "api": "aleo_b200_msm_g1 (host pointers, PAGEABLE caller memory as a Rust Vec is)", "host_memory": "pageable",
                    "pinned": {"value": e2e_pinned, "unit": "Mpts/s", "what": "the same call on cudaHostAlloc'ed buffers"},
                    "point_ranges": host_plan["ranges"], "window_bits": host_plan["window_bits"],
                    "note": "the call stages 4 MB slices through pinned double buffers with helper threads and copies / accumulates "
                            "point range by point range: H2D of range k+1 overlaps range k"},
            "gpu_launches": (ab.VariableBase.launches(n) + (1 if world > 1 else 0)) * args.steps,
            "roofline": roofline, "witness_like_scalars": witness, "srs_resident": srs_obj, "ntt": ntt, "ntt_distributed": ntt_dist, "config4": config4, "sweep": sweep, "proof_shaped": proof_shaped, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=24)
    ap.add_argument("--cpu-log-n", type=int, default=0, help="--impl reference: cap the CPU arm's size (0 = the GPU arm's own --log-n)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=22, help="GPU arm: size of its bounded cpu_baseline sample")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: stop timing further steps after this many seconds (>= 3 steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-srs", action="store_true", help="skip the resident-SRS (KZG commit) measurement")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 2^16..2^22 size sweep")
    ap.add_argument("--no-config4", action="store_true", help="skip BASELINE configs[3] (fixed 2^26 MSM + NTT over the N GPUs)")
    ap.add_argument("--config4-log-n", type=int, default=26)
    ap.add_argument("--no-proof-shape", action="store_true", help="skip the synthetic proof-shaped stream")
    ap.add_argument("--proof-threads", type=int, default=4, help="host threads (one CUDA stream each) submitting proof-shaped work")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints to file descriptor 1 on the way (NCCL's version
    # banner under NCCL_DEBUG=VERSION / INFO does) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global _JSON_OUT
    _JSON_OUT = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
