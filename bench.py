#!/usr/bin/env python
"""bench.py — NGT hot-path benchmark (BASELINE.json metric: NGT nodes eliminated/s, FP64 Schur TFLOP/s).

    python bench.py --gpus N --steps K --warmup W          our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...          the reference's own CPU NGT on this box's host cores

A "step" is one pass of the hot path over one network: ingest kernel (rates -> P, tau) + phase-1 elimination of every
intermediate node + phase 2.  Workload at N=1: BASELINE config 2, the synthetic 2-D energy landscape with ~10^5 minima
and ~10^6 transition states, |A|=|B|=1 (kmc_rates_b200/synth.py: landscape_2d(316)).  At N>1 every rank eliminates its
own network of that shape (independent networks of a sweep; no data-path collective) -> weak scaling.

  value     nodes eliminated per second, rate constants already resident in HBM, symbolic analysis done once outside
            the timed region (the "analyse once, re-factor per temperature" use); device time from CUDA events on the
            engine's stream, max over ranks.
  e2e       the same metric through the public call a user makes (NGT.from_coo(host arrays) + compute_rates() + rate
            getters): host ordering / symbolic analysis, H2D of every array, kernels, D2H of the results, wall clock
            with device syncs on both sides.  Bytes are counted by the library (every cudaMemcpy it issues).
  roofline  the Schur-update DMMA kernel (dominant by time): flops its launches perform / their CUDA-event time
            (measured in a single-stream pass, so the sum can never exceed the pass).
  sharded   at every N: the configs BASELINE quotes at 1/2/4/8 GPUs, run over the N ranks with real NCCL collectives
            issued by the C++ engine -- config 5 (4096 networks, strong scaling), config 4 (phase 2 sharded), config 3
            (one dense network, 1-D block-cyclic columns) -- each with its single-GPU time from the same build and the
            max relative error against the single-GPU result.
  dense / other_configs / cpu_baseline / same_size: N=1 only (see the functions below).
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NGT nodes eliminated/s"
UNIT = "nodes/s"
SIDE = 316            # 316^2 = 99 856 minima, ~9.9e5 transition states
TEMPERATURE = 0.3
CPU_SAMPLE_SIDE = 70  # cpu_baseline leg: reference CPU NGT on 4 900-minima landscapes, ~20 s per network per core
REF_BUDGET_S = 240.0  # --impl reference: total wall-clock budget; the sample size per step is derived from it
# Measured on this pool's host cores: the reference takes ~3.1 s at 2 500 minima and grows ~ n^2.1
# (4.8 s -> 87 s from 2 500 to 10 000 minima); used only to size the bounded sample.
REF_T0_S, REF_N0, REF_EXP = 3.1, 2500.0, 2.1
C5_NB, C5_N = 4096, 256
C4_SIDE, C4_GROUP = 141, 512


def reference_sample_side(steps, warmup):
    """largest landscape side whose estimated reference time fits REF_BUDGET_S over steps+warmup steps (>= 50)"""
    per_step = REF_BUDGET_S / max(1, steps + warmup)
    n = REF_N0 * (per_step / REF_T0_S) ** (1.0 / REF_EXP)
    return int(max(50, min(100, n ** 0.5)))


def workload_config(n_gpus, side):
    return {
        "workload": "synthetic 2-D energy landscape (BASELINE config 2): %d minima, ~%.2e transition states, |A|=|B|=1, "
                    "T=%.2f; one network per GPU" % (side * side, 9.93 * side * side, TEMPERATURE),
        "generator": "kmc_rates_b200.synth.landscape_2d(side=%d, T=%.2f, seed=rank)" % (side, TEMPERATURE),
        "l2": "inputs larger than L2: the ~0.7 GB front pool is re-zeroed and rewritten every step",
        "parallelism": "independent networks, 1 per GPU x %d" % n_gpus,
    }


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own C++ NGT (oracle/_ref, compiled from /root/reference) or, if that
# library did not travel, the oracle port.  One network per host core, all cores busy.
# ------------------------------------------------------------------------------------------------------------------
def _preload_reference():
    """load the CPU engine in THIS process before any worker is forked, so that whoever watches which libraries the
    benchmark process maps sees the reference engine (oracle/_ref/libngt_ref.so) and not just its children"""
    from oracle import cpu_ngt
    kind = cpu_ngt.best_kind()
    cpu_ngt._lib(kind)
    return kind


def _kind_text(kind):
    return ("unmodified reference C++ NGT compiled from /root/reference (oracle/_ref/libngt_ref.so)" if kind == "reference"
            else "oracle C++ port (oracle/libngt_oracle.so; the compiled reference did not travel)")


def _cpu_worker(args):
    side, seed, kind = args
    from kmc_rates_b200 import synth
    from oracle.cpu_ngt import CpuNGT
    n, src, dst, k, A, B = synth.landscape_2d(side, T=TEMPERATURE, seed=seed)
    o = CpuNGT(n, src, dst, k, A, B, kind=kind)   # ctor = the reference's graph build (ngt.hpp:105-181)
    t0 = time.perf_counter()
    o.compute_rates()
    t1 = time.perf_counter()
    return n - 2, t1 - t0, [float(x) for x in o.rates()]


def cpu_reference_step(side, cores, kind, seed0=0):
    """one network per core, concurrently; returns (nodes eliminated, seconds of the slowest worker, rates per network)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(side, seed0 + i, kind) for i in range(cores)])
    nodes = sum(r[0] for r in res)
    solve = max(r[1] for r in res)
    return nodes, solve, [r[2] for r in res]


def cpu_baseline_object(kind, cores, side=CPU_SAMPLE_SIDE):
    nodes, solve, rates = cpu_reference_step(side, cores, kind)
    scale = (99856.0 / (side * side)) ** 1.1
    obj = {"value": nodes / solve, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": "%d independent %d-minima landscapes of the workload's generator (seeds 0..%d; one per host core, "
                     "concurrently); compute_rates() of the %s timed per worker (steady clock), slowest worker %.2f s. "
                     "Bounded sample: the reference's cost grows ~n^2.1 (BASELINE.md §2), so at the full 99 856-minima "
                     "size it would be ~%.0fx slower per node than measured here"
                     % (cores, side * side, cores - 1, _kind_text(kind), solve, scale),
           "extrapolated_full_size_value": nodes / solve / scale}
    return obj, rates


def _config_point_worker(task):
    """one CPU timing point of BASELINE.md §3 (runs in a forked worker; numpy + the CPU engine only)"""
    name, kind, arg = task
    from kmc_rates_b200 import synth
    from oracle.cpu_ngt import CpuNGT
    if name == "c1":
        s, d, k = synth.complete_random(100, seed=0)
        best = 1e30
        for _ in range(3):
            o = CpuNGT(100, s, d, k, [0], [1], kind=kind)
            t0 = time.perf_counter(); o.compute_rates(); best = min(best, time.perf_counter() - t0)
        return name, arg, best, [float(x) for x in o.rates()]
    if name == "c3":
        n = arg
        s, d, k = synth.complete_random(n, seed=n)
        o = CpuNGT(n, s, d, k, [0], [1], kind=kind)
        t0 = time.perf_counter(); o.compute_rates()
        return name, arg, time.perf_counter() - t0, [float(x) for x in o.rates()]
    if name == "c4":
        g, n = arg, 192
        s, d, k = synth.complete_random(n, seed=g)
        A, B = list(range(g)), list(range(n - g, n))
        o1 = CpuNGT(n, s, d, k, A, B, kind=kind)
        t0 = time.perf_counter(); o1.phase_one(); t1 = time.perf_counter() - t0
        o2 = CpuNGT(n, s, d, k, A, B, kind=kind)
        t0 = time.perf_counter(); o2.compute_rates(); t2 = time.perf_counter() - t0
        return name, arg, max(t2 - t1, 1e-9), [float(x) for x in o2.rates()]
    if name == "c5":
        b, K = arg
        n = K.shape[0]
        i, j = np.nonzero(~np.eye(n, dtype=bool))
        o = CpuNGT(n, i, j, K[i, j], [0], [n - 1], kind=kind)
        t0 = time.perf_counter(); o.compute_rates()
        return name, b, time.perf_counter() - t0, [float(x) for x in o.rates()]
    raise ValueError(name)


def _fit_exponent(xs, ts):
    lx, lt = np.log(np.asarray(xs, float)), np.log(np.asarray(ts, float))
    return float(np.polyfit(lx, lt, 1)[0])


def cpu_config_points(kind, cores, c5_sample):
    """CPU timing points for the other BASELINE configs at reduced sizes, with labelled extrapolations.
    c5_sample: list of (network index, host rate matrix) taken from the GPU sweep's own inputs."""
    import multiprocessing as mp
    tasks = [("c4", kind, 64), ("c3", kind, 800), ("c4", kind, 32), ("c3", kind, 400), ("c3", kind, 200), ("c4", kind, 16),
             ("c1", kind, 100)] + [("c5", kind, s) for s in c5_sample]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_config_point_worker, tasks, chunksize=1)
    wall = time.perf_counter() - t0
    out = {"wall_s": wall, "cores": cores, "kind": kind}
    c1 = [r for r in res if r[0] == "c1"][0]
    out["config1_complete100"] = {"seconds": c1[2], "nodes_per_s": 98 / c1[2], "rates": c1[3]}
    c3 = sorted([r for r in res if r[0] == "c3"], key=lambda r: r[1])
    e3 = _fit_exponent([r[1] for r in c3], [r[2] for r in c3])
    out["config3_dense"] = {
        "points": {str(r[1]): r[2] for r in c3}, "fitted_exponent": e3,
        "extrapolated_16384_seconds": c3[-1][2] * (16384.0 / c3[-1][1]) ** 3,
        "extrapolation": "EXTRAPOLATED from the largest measured point with the N^3 law of BASELINE.md §3 (one core; "
                         "the reference is single-threaded)"}
    c4 = sorted([r for r in res if r[0] == "c4"], key=lambda r: r[1])
    e4 = _fit_exponent([r[1] for r in c4], [r[2] for r in c4])
    out["config4_phase2"] = {
        "points": {str(r[1]): r[2] for r in c4}, "fitted_exponent": e4,
        "graph": "complete 192-node graph, |A|=|B| as listed; phase 2 = compute_rates() - phase_one() seconds",
        "extrapolated_512_seconds": c4[-1][2] * (512.0 / c4[-1][1]) ** 4,
        "extrapolation": "EXTRAPOLATED from |A|=|B|=64 with the |A|^4 law of BASELINE.md §3 (one core)"}
    c5 = [r for r in res if r[0] == "c5"]
    if c5:
        tot = sum(r[2] for r in c5)
        out["config5_sweep"] = {
            "networks_timed": len(c5), "cpu_seconds_sum": tot, "seconds_per_network": tot / len(c5),
            "extrapolated_4096_seconds_all_cores": tot / len(c5) * C5_NB / cores,
            "extrapolation": "EXTRAPOLATED linearly in the number of networks, %d cores busy" % cores}
    return out, {"c1": c1[3], "c5": {r[1]: r[3] for r in c5}}


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = _preload_reference()
    cores = os.cpu_count() or 1
    side = reference_sample_side(args.steps, args.warmup)
    for _ in range(args.warmup):
        cpu_reference_step(side, cores, kind)
    tot_nodes, tot_t = 0, 0.0
    for i in range(args.steps):
        nodes, solve, _r = cpu_reference_step(side, cores, kind, seed0=100 * i)
        tot_nodes += nodes
        tot_t += solve
    value = tot_nodes / tot_t
    scale = (99856.0 / (side * side)) ** 1.1
    sample = ("each step: %d independent %d-minima landscapes (the GPU arm's generator at a bounded size, so that "
              "steps+warmup fit ~%d s), one per host core concurrently, compute_rates() of the %s; the reference's cost "
              "grows ~n^2.1, so this bounded sample overstates its full-size (99 856 minima) rate by ~%.0fx" %
              (cores, side * side, int(REF_BUDGET_S), _kind_text(kind), scale))
    cfg = workload_config(args.gpus, SIDE)
    # what this arm actually ran: a bounded sample of the workload named above, NOT the full-size network
    cfg["measured_on"] = "%d x %d-minima landscapes per step (bounded sample of the workload; see cpu_baseline.sample)" % (cores, side * side)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "sample_minima": side * side, "extrapolated_full_size_value": value / scale},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def measure_dgemm_peak(torch, n=8192, reps=6):
    """cuBLAS DGEMM on this box, the FP64 tensor roofline denominator (MEASURED_PEAKS.json has no FP64 figure)."""
    a = torch.rand(n, n, dtype=torch.float64, device="cuda")
    b = torch.rand(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def build_fingerprint():
    """hash of the CUDA sources (kernels + launch schedule) this library was built from: ncu-derived numbers are only
    quoted for the same kernels and schedule (the host analysis is not part of it: tests pin its output, the fronts)"""
    h = hashlib.sha256()
    for f in ("kernels.cu", "kernels.cuh", "engine.cu"):
        with open(os.path.join(ROOT, "kmc_rates_b200", "csrc", f), "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel_key):
    """DRAM bytes from the committed ncu launch list -- only if it was captured from THIS build (fingerprint match)"""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
    except Exception:
        return None, "no profiles/r02_traffic.json"
    if t.get("build_fingerprint") != build_fingerprint():
        return None, "profiles/r02_traffic.json was captured from another build (%s)" % t.get("build_fingerprint")
    return t.get(kernel_key), "profiles/r02_traffic.json (ncu launch list of this build, %s)" % t.get("command", "")


def c5_rate_matrices(torch, lo, hi):
    """BASELINE config 5 inputs for networks [lo, hi) of the 4096-temperature sweep, generated on the device; every
    rank draws the same energies (fixed seed) and takes its slice of the temperatures"""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    E = torch.rand(C5_N, generator=gen, dtype=torch.float64, device="cuda")
    bar = torch.rand(C5_N, C5_N, generator=gen, dtype=torch.float64, device="cuda") * 0.9 + 0.1
    bar = torch.maximum(bar, bar.T)
    temps = torch.logspace(np.log10(0.05), 0.0, C5_NB, dtype=torch.float64, device="cuda")[lo:hi]
    K = torch.exp(-((torch.maximum(E[:, None], E[None, :]) + bar)[None] - E[None, :, None]) / temps[:, None, None])
    K.diagonal(dim1=1, dim2=2).zero_()
    return K


def run_c5(torch, Handle, lo, hi, reps=3):
    K = c5_rate_matrices(torch, lo, hi)
    h = Handle.from_dense(K, [0], [C5_N - 1])
    h.timed_steps(1)
    ms = h.timed_steps(reps) / reps
    rates = np.array(h.rates()).reshape(hi - lo, 4)
    h.close()
    return ms, rates, K


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def sharded_object(torch, Handle, comm, dist, rank, world, dense_n, keep=None):
    """The BASELINE configs quoted at 1/2/4/8 GPUs, over `world` ranks with the engine's own NCCL collectives.
    Every rank also runs the single-GPU path on the same inputs: that gives the N=1 time of the same build and the
    reference result for the in-run parity (max relative error over ranks)."""
    from kmc_rates_b200 import synth
    from kmc_rates_b200.distributed import shard_range

    def vmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    out = {"world": world}
    # ---- config 5: 4096 independent 256-node networks, contiguous blocks per rank, no data-path collective ----------
    ms1, rates1, K1 = run_c5(torch, Handle, 0, C5_NB)
    if keep is not None:
        keep["c5_K"] = K1
        keep["c5_rates"] = rates1
    lo, hi = shard_range(C5_NB, rank, world)
    sync()
    msN, ratesN, _K = run_c5(torch, Handle, lo, hi) if world > 1 else (ms1, rates1, None)
    del _K
    full = torch.zeros(C5_NB, 4, dtype=torch.float64, device="cuda")
    full[lo:hi] = torch.from_numpy(ratesN).cuda()
    if dist is not None:
        dist.all_reduce(full)                                  # gather of 4 numbers per network (results only)
    msN = vmax(msN)
    out["config5_4096x256_sweep"] = {
        "partition": "contiguous blocks of %d networks per rank" % (hi - lo), "scaling": "strong",
        "ms": msN, "ms_1gpu_same_build": ms1, "speedup": ms1 / msN, "efficiency": ms1 / msN / world,
        "networks_per_s": C5_NB / (msN * 1e-3), "nodes_eliminated_per_s": C5_NB * (C5_N - 2) / (msN * 1e-3),
        "alg_tflops": C5_NB * (2.0 / 3.0) * C5_N ** 3 / (msN * 1e-3) / 1e12,
        "collectives": "none on the data path; all-reduce of the %d-byte result table afterwards" % (C5_NB * 32),
        "max_rel_err_vs_1gpu": vmax(rel(full.cpu().numpy()[:, :2], rates1[:, :2])),
    }
    if keep is None:
        del K1
    torch.cuda.empty_cache()

    # ---- config 4: 2*10^4-node landscape, |A|=|B|=512: phase 1 replicated, phase 2 sharded ---------------------------
    n, src, dst, k, A, B = synth.landscape_groups(C4_SIDE, C4_GROUP, C4_GROUP, T=TEMPERATURE, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B, device=torch.cuda.current_device())
    h.compute_rates()
    h.phase_one(); h.phase_two()
    st1 = h.stats()
    r1, (om1, tau1) = np.array(h.rates()), h.final()
    m = len(A) + len(B)
    h.phase_one()
    h.phase_two_sharded(comm, broadcast_root=True)          # warm-up (plan build, NCCL channels)
    sync()
    h.phase_one()
    ms2 = vmax(h.phase_two_sharded(comm, broadcast_root=True))
    rN, (omN, tauN) = np.array(h.rates()), h.final()
    root_doubles = h.reduced_block_dev()[1]
    out["config4_19881_nodes_A512_B512"] = {
        "partition": "phase 1 replicated; every rank takes a contiguous 1/%d of A and of B in the phase-2 elimination tree" % world,
        "phase1_ms_1gpu": st1["ms_phase1"], "phase2_ms": ms2, "phase2_ms_1gpu_same_build": st1["ms_phase2"],
        "phase2_speedup": st1["ms_phase2"] / ms2, "phase2_efficiency": st1["ms_phase2"] / ms2 / world,
        "per_a_results_per_s": m / (ms2 * 1e-3),
        "collectives": "ncclBroadcast of the reduced block (%d B) + 2 x ncclAllReduce(sum) of %d B, on the engine's stream"
                       % (root_doubles * 8, m * 8),
        "max_rel_err_vs_1gpu": vmax(max(rel(rN, r1), rel(omN, om1), rel(tauN, tau1))),
        "k_AB": float(rN[0]), "k_BA": float(rN[1]),
        "note": "phase 2 is a binary tree of eliminations (4/3 |A|^3 flops; the reference's per-a loop is |A|^4): the top "
                "of the tree is shared by all ranks, so sharding it buys little -- the single-GPU time is what matters",
    }
    if keep is not None:
        keep["c4"] = (n, src, dst, k, A, B, r1)
    h.close()

    # ---- config 3: ONE dense network over the ranks (1-D block-cyclic columns, panel broadcast) ----------------------
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    K = torch.rand(dense_n, dense_n, generator=gen, dtype=torch.float64, device="cuda")
    K.fill_diagonal_(0.0)
    h1 = Handle.from_dense(K, [0], [1])
    h1.timed_steps(1)
    ms_single = h1.timed_steps(2) / 2
    rs, alg_flops = np.array(h1.rates()), h1.stats()["alg_flops"]
    Ps, ts = h1.reduced()
    h1.close()
    hd = Handle.from_dense(K, [0], [1])
    hd.dist_phase_one(comm)                                   # warm-up
    sync()
    ms_d = hd.dist_phase_one(comm)
    hd.phase_two()
    ms_d = vmax(ms_d + hd.stats()["ms_phase2"])
    rd = np.array(hd.rates())
    Pd, td = hd.reduced()
    nblk = (dense_n - 2 + 255) // 256
    bytes_bcast = sum((2048 + (dense_n - 256 * b) * 264) * 8 for b in range(nblk))
    out["config3_dense_%d" % dense_n] = {
        "partition": "1-D block-cyclic columns (block 256) over %d ranks; panel owner factors, NCCL broadcast, every rank "
                     "updates its own columns; look-ahead on a second stream" % world,
        "ms": ms_d, "ms_1gpu_same_build": ms_single, "speedup": ms_single / ms_d, "efficiency": ms_single / ms_d / world,
        "alg_tflops_aggregate": alg_flops / (ms_d * 1e-3) / 1e12,
        "collectives": "per block: ncclAllReduce(sum) 2048 B + ncclBroadcast of the factored panel; total broadcast %.3f GB; "
                       "final ncclAllReduce of the reduced block" % (bytes_bcast / 1e9),
        "host_syncs_per_block": 0,
        "max_rel_err_vs_1gpu": vmax(max(rel(rd, rs), rel(Pd, Ps), rel(td, ts))),
        "k_AB": float(rd[0]),
    }
    hd.close()
    del K
    torch.cuda.empty_cache()
    return out


def dense_object(torch, Handle, dgemm_tflops, n=16384, steps=2):
    """BASELINE config 3 shape at 1 GPU: complete graph of n nodes generated on the device, A={0}, B={1}."""
    K = torch.rand(n, n, dtype=torch.float64, device="cuda")
    K.fill_diagonal_(0.0)
    h = Handle.from_dense(K, [0], [1])
    h.timed_steps(1)
    ms = h.timed_steps(steps) / steps
    st = h.stats()
    pms, pfl, pl = h.profile_schur()
    pass_ms = h.stats()["ms_profiled_pass"]
    h.timed_steps(1)
    kab = float(h.rates()[0])
    # independent check: mean first-passage time into B = {1} by a dense LU solve on the device (cuSOLVER; a checker)
    M = K.clone()
    M -= torch.diag(K.sum(dim=1))
    keepm = torch.ones(n, dtype=torch.bool, device="cuda")
    keepm[1] = False
    t = torch.linalg.solve(M[keepm][:, keepm], -torch.ones(n - 1, dtype=torch.float64, device="cuda"))
    kab_solve = 1.0 / float(t[0])
    del M, t
    out = {
        "workload": "complete graph, %d nodes, k_ij ~ U(0,1) generated on device, A={0}, B={1} (BASELINE config 3, 1 GPU)" % n,
        "ms_per_step": ms, "nodes_per_s": (n - 2) / (ms * 1e-3),
        "alg_tflops": st["alg_flops"] / (ms * 1e-3) / 1e12,
        "frac_of_dgemm_peak_whole_step": st["alg_flops"] / (ms * 1e-3) / 1e12 / dgemm_tflops,
        "schur_kernel": {"ms_in_single_stream_pass": pms, "pass_ms": pass_ms, "share_of_pass": pms / pass_ms,
                         "tflops": pfl / (pms * 1e-3) / 1e12, "frac_of_dgemm_peak": pfl / (pms * 1e-3) / 1e12 / dgemm_tflops,
                         "launches": pl,
                         "note": "event pairs around every Schur launch of a pass run on ONE stream (no look-ahead), so "
                                 "the sum is a true share of that pass; the timed steps above use the look-ahead stream"},
        "k_AB": kab, "max_rel_err_vs_dense_LU_solve": abs(kab - kab_solve) / kab_solve,
        "launches_per_step": st["n_launches"],
    }
    h.close()
    del K
    torch.cuda.empty_cache()
    return out


def config1_object():
    """config 1: README-style random complete graph, 100 nodes, through the drop-in GraphReduction API"""
    from kmc_rates_b200 import GraphReduction, synth
    rates = synth.to_dict(*synth.complete_random(100, seed=0))      # the network the CPU point of config 1 uses
    GraphReduction(rates, [0], [1]).compute_rates()
    t0 = time.perf_counter()
    red = GraphReduction(rates, [0], [1])
    red.compute_rates()
    kab, kba = red.get_rate_AB(), red.get_rate_BA()
    return {"ms_end_to_end": 1e3 * (time.perf_counter() - t0), "k_AB": kab, "k_BA": kba}


def same_size_object(torch, Handle, cores, side, ref_value, ref_rates, device):
    """Both arms on the IDENTICAL networks: the cpu_baseline leg's bounded sample (one landscape per host core) solved
    end to end on the GPU -- host arrays in, analysis, kernels, rates out, one network after the other."""
    from kmc_rates_b200 import synth
    nets = [synth.landscape_2d(side, T=TEMPERATURE, seed=s) for s in range(cores)]
    n, src, dst, k, A, B = nets[0]
    Handle.from_coo(n, src, dst, k, A, B, device=device).close()        # warm the host workspace for this size
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got, each = [], []
    for (n, src, dst, k, A, B) in nets:
        ta = time.perf_counter()
        g = Handle.from_coo(n, src, dst, k, A, B, device=device)
        tb = time.perf_counter()
        g.compute_rates()
        got.append(np.array(g.rates()))
        tc = time.perf_counter()
        g.close()
        each.append([round(1e3 * (tb - ta), 2), round(1e3 * (tc - tb), 2), round(1e3 * (time.perf_counter() - tc), 2)])
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    nodes = sum(x[0] - 2 for x in nets)
    return {"networks": cores, "minima_per_network": side * side,
            "gpu_e2e_nodes_per_s": nodes / sec, "gpu_e2e_ms_per_network": 1e3 * sec / cores,
            "reference_nodes_per_s": ref_value, "ratio_gpu_e2e_over_reference": nodes / sec / ref_value,
            "max_rel_err_vs_reference": rel(np.array(got), np.array(ref_rates)),
            "ms_create_compute_close_each": each,
            "note": "same generator, seeds and sizes in both arms; the reference runs its %d networks concurrently on %d "
                    "cores, the GPU arm runs them one after the other through the public end-to-end call" % (cores, cores)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=SIDE)
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--dense-n", type=int, default=16384)
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: everything else any library prints there (NCCL announces its
    # version on stdout when a communicator is created) is sent to stderr; the record goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    from kmc_rates_b200 import synth, _capi
    from kmc_rates_b200._capi import Handle

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # several ranks share the host: each analyses its network with its share of the cores (the native host pipeline
    # reads NGT_B200_HOST_THREADS once, before the first handle is created)
    if world > 1:
        os.environ.setdefault("NGT_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — kmc_rates_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    side = args.side
    n, src, dst, k, A, B = synth.landscape_2d(side, T=TEMPERATURE, seed=rank)
    # pinned host copies: the e2e leg reads its inputs from pinned memory
    tsrc = torch.from_numpy(src).pin_memory()
    tdst = torch.from_numpy(dst).pin_memory()
    tk = torch.from_numpy(k).pin_memory()

    # ---- HBM-resident leg: analysis once, K timed numeric steps -----------------------------------------------
    h = Handle.from_coo(n, tsrc.numpy(), tdst.numpy(), tk.numpy(), A, B, device=local)
    for _ in range(warmup):
        h.timed_steps(1)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ms_total = h.timed_steps(args.steps)
    barrier()
    rates = h.rates()
    # per-phase breakdown from one extra step outside the CUDA graph (events between the phases), not timed above
    h.phase_one()
    h.phase_two()
    st = h.stats()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    nodes_per_step = st["n_eliminated"]
    value = world * nodes_per_step * args.steps / (ms_max * 1e-3)

    # ---- e2e leg: host arrays -> public API -> host results, every step ------------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    # one untimed end-to-end call first: the timed handle `h` above still holds its device buffers, so the first
    # create here would pay cudaMalloc for a full set (the library caches freed buffers for the next handle)
    g = Handle.from_coo(n, tsrc.numpy(), tdst.numpy(), tk.numpy(), A, B, device=local)
    g.compute_rates()
    g.close()
    barrier()
    b0 = h.stats()
    t0 = time.perf_counter()
    kab = 0.0
    for _ in range(e2e_steps):
        g = Handle.from_coo(n, tsrc.numpy(), tdst.numpy(), tk.numpy(), A, B, device=local)
        g.compute_rates()
        kab = float(g.rates()[0])
        g.close()
    barrier()
    e2e_s = time.perf_counter() - t0
    b1 = h.stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * nodes_per_step * e2e_steps / float(te.item())
    clocks = sampler.stop()
    nnz = int(src.size)
    # counted by the library: every host<->device copy it issued during the e2e steps (process-wide counters)
    h2d = (b1["h2d_bytes_total"] - b0["h2d_bytes_total"]) // e2e_steps
    d2h = (b1["d2h_bytes_total"] - b0["d2h_bytes_total"]) // e2e_steps
    # reuse leg (set_rates on an analysed handle): what a temperature sweep pays per network
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.set_rates(tk.numpy())
        h.compute_rates()
        _ = h.rates()
    barrier()
    reuse_s = time.perf_counter() - t0
    reuse_value = world * nodes_per_step * e2e_steps / reuse_s

    # ---- the sharded configs, all ranks --------------------------------------------------------------------------------
    keep = {} if (world == 1 and not args.no_cpu) else None
    sharded = None
    if not args.no_sharded:
        try:
            if world > 1:
                comm = _capi.Comm.from_torch(device=local)
            else:
                comm = _capi.Comm(local, 0, 1, lambda b: b)
            sharded = sharded_object(torch, Handle, comm, dist, rank, world, args.dense_n, keep)
            comm.close()
        except Exception as e:   # report, do not hide
            sharded = {"error": repr(e)}
            if dist is not None:
                raise

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        dgemm = measure_dgemm_peak(torch)
        pms, pfl, pl = h.profile_schur()
        pass_ms = h.stats()["ms_profiled_pass"]   # the same profiled pass: consistent event overheads
        traffic, traffic_src = ncu_traffic("schur_bytes_per_launch")
        step_bytes, _src = ncu_traffic("step_dram_bytes")
        hbm = peaks.get("hbm_gbs")
        roofline = {
            "kernel": "update_dmma_kernel (Schur update P_RR += P_RX U', FP64 DMMA m8n8k4)",
            "bound": "tensor", "achieved": pfl / (pms * 1e-3) / 1e12, "peak": dgemm, "unit": "TFLOP/s",
            "frac": pfl / (pms * 1e-3) / 1e12 / dgemm, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (best of 6, CUDA events); "
                           "MEASURED_PEAKS.json carries no FP64 figure (hbm_gbs=%s)" % hbm,
            "flops_per_step": pfl, "launches_per_step": pl, "kernel_ms_per_step": pms,
            "share_of_step": pms / pass_ms,
            "share_note": "kernel CUDA-event time / duration of the same profiled (event-bracketed, graph-free, "
                          "single-stream) pass; ncu launch list of the same step: profiles/r02_launches_c2.md",
            "whole_step": {
                "alg_flops": st["alg_flops"], "tflops": st["alg_flops"] / (ms_per_step * 1e-3) / 1e12,
                "frac_of_dgemm": st["alg_flops"] / (ms_per_step * 1e-3) / 1e12 / dgemm,
                "dram_bytes_measured": step_bytes,
                "hbm_GBs": (step_bytes / (ms_per_step * 1e-3) / 1e9) if step_bytes else None,
                "hbm_frac": (step_bytes / (ms_per_step * 1e-3) / 1e9 / hbm) if (step_bytes and hbm) else None,
                "note": "DRAM bytes = sum over the step's launches in the ncu list of this build (cold-cache, an upper "
                        "bound for the graph replay) + the pool memset; the sparse phase is blocked into dense fronts, so "
                        "SURVEY §8(d)'s per-node byte model does not describe its traffic and is not reported",
            },
        }
        cpu = dense = other = same = None
        if world == 1 and not args.no_cpu:
            kind = _preload_reference()
            cores = os.cpu_count() or 1
            cpu, ref_rates = cpu_baseline_object(kind, cores)
            try:
                same = same_size_object(torch, Handle, cores, CPU_SAMPLE_SIDE, cpu["value"], ref_rates, local)
            except Exception as e:
                same = {"error": repr(e)}
            try:
                sample = []
                if keep and "c5_K" in keep:
                    for b in np.linspace(0, C5_NB - 1, 64).astype(int):
                        sample.append((int(b), keep["c5_K"][int(b)].cpu().numpy()))
                pts, ref = cpu_config_points(kind, cores, sample)
                parity = {}
                c1 = config1_object()
                parity["config1_vs_reference"] = rel([c1["k_AB"], c1["k_BA"]], ref["c1"][:2])
                if keep and "c5_rates" in keep and ref["c5"]:
                    parity["config5_vs_reference_64_networks"] = max(
                        rel(keep["c5_rates"][b][:2], r[:2]) for b, r in ref["c5"].items())
                if keep and "c4" in keep:
                    from oracle.py_linalg import TwoStateRatesCOO
                    n4, s4, d4, k4, A4, B4, r4 = keep["c4"]
                    la = TwoStateRatesCOO(n4, s4, d4, k4, A4, B4)
                    la.compute_rates()
                    parity["config4_kAB_vs_sparse_linear_solve"] = abs(r4[0] - la.get_rate_AB()) / la.get_rate_AB()
                from oracle.py_linalg import TwoStateRatesCOO
                la = TwoStateRatesCOO(n, src, dst, k, A, B)
                la.compute_rates()
                parity["config2_kAB_vs_sparse_linear_solve"] = abs(float(rates[0]) - la.get_rate_AB()) / la.get_rate_AB()
                cpu["configs"] = pts
                cpu["max_rel_err"] = parity
                cpu["max_rel_err_note"] = ("GPU results of this run against the CPU reference where it can run (configs 1, 5) and "
                                           "against the independent sparse linear solve (scipy SuperLU, oracle/py_linalg.py) at the "
                                           "sizes it cannot (configs 2, 4); config 3: see dense.max_rel_err_vs_dense_LU_solve; "
                                           "same-size landscapes: same_size.max_rel_err_vs_reference")
                other = {"config1_complete100_GraphReduction_api": c1}
            except Exception as e:
                cpu["configs"] = {"error": repr(e)}
        if keep:
            keep.clear()
        torch.cuda.empty_cache()
        if world == 1 and not args.no_dense:
            try:
                dense = dense_object(torch, Handle, dgemm, n=args.dense_n)
            except Exception as e:   # report, do not hide
                dense = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, side),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "bytes_note": "counted by the library: the sum of every cudaMemcpy it issued during the e2e steps / steps",
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
                    "includes": "host nested-dissection ordering + symbolic analysis, every H2D copy, kernels, D2H of rates",
                    "reuse_value": reuse_value,
                    "reuse_note": "set_rates(host k) + compute_rates() on an already analysed handle (temperature sweep)"},
            "gpu_launches": int(st["n_launches"]) * args.steps,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "same_size": same,
            "sharded": sharded,
            "dense": dense,
            "other_configs": other,
            "stats": {"nodes_eliminated_per_step": nodes_per_step, "rates": nnz, "fronts": st["n_fronts"],
                      "levels": st["n_levels"], "max_front": st["max_front"], "pool_GB": st["pool_doubles"] * 8 / 1e9,
                      "alg_flops": st["alg_flops"], "ms_symbolic_host": st["ms_symbolic"], "ms_ingest": st["ms_ingest"],
                      "ms_phase1": st["ms_phase1"], "ms_phase2": st["ms_phase2"],
                      "cuda_graph": os.environ.get("NGT_B200_NO_GRAPH") is None, "k_AB": float(rates[0]),
                      "k_BA": float(rates[1]), "k_AB_e2e": kab, "build_fingerprint": build_fingerprint()},
        }
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == "__main__":
    main()
