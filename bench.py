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
            with device syncs on both sides.
  roofline  the Schur-update DMMA kernel (dominant by time): flops its launches perform / their CUDA-event time.
  dense     extra object: BASELINE config 3 shape (complete graph, one front) — FP64 Schur TFLOP/s.
"""
import argparse
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
def _cpu_worker(args):
    side, seed, kind = args
    from kmc_rates_b200 import synth
    from oracle.cpu_ngt import CpuNGT
    n, src, dst, k, A, B = synth.landscape_2d(side, T=TEMPERATURE, seed=seed)
    o = CpuNGT(n, src, dst, k, A, B, kind=kind)   # ctor = the reference's graph build (ngt.hpp:105-181)
    t0 = time.perf_counter()
    o.compute_rates()
    t1 = time.perf_counter()
    return n - 2, t1 - t0, float(o.rates()[0])


def cpu_reference_step(side, cores, kind, seed0=0):
    """one network per core, concurrently; returns (nodes eliminated, wall seconds of the slowest worker incl. pool)"""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(side, seed0 + i, kind) for i in range(cores)])
    wall = time.perf_counter() - t0
    nodes = sum(r[0] for r in res)
    solve = max(r[1] for r in res)
    return nodes, solve, wall


def cpu_baseline_object(cores=None, side=CPU_SAMPLE_SIDE):
    from oracle import cpu_ngt
    kind = cpu_ngt.best_kind()
    cores = cores or (os.cpu_count() or 1)
    nodes, solve, wall = cpu_reference_step(side, cores, kind)
    return {"value": nodes / solve, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d independent %d-minima landscapes of the same generator (one per host core, concurrently); "
                      "compute_rates() of the %s timed per worker (steady clock), slowest worker %.2f s. Bounded sample: "
                      "the reference's cost grows ~n^2.1 (BASELINE.md), so at the full 99 856-minima size it would be "
                      "~%.0fx slower per node than measured here"
                      % (cores, side * side, "reference C++ NGT compiled from /root/reference (oracle/_ref)"
                         if kind == "reference" else "oracle C++ port", solve, (99856.0 / (side * side)) ** 1.1),
            "extrapolated_full_size_value": nodes / solve / (99856.0 / (side * side)) ** 1.1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_ngt
    kind = cpu_ngt.best_kind()
    cores = os.cpu_count() or 1
    side = reference_sample_side(args.steps, args.warmup)
    for _ in range(args.warmup):
        cpu_reference_step(side, cores, kind)
    tot_nodes, tot_t = 0, 0.0
    for i in range(args.steps):
        nodes, solve, _wall = cpu_reference_step(side, cores, kind, seed0=100 * i)
        tot_nodes += nodes
        tot_t += solve
    value = tot_nodes / tot_t
    sample = ("each step: %d independent %d-minima landscapes (same generator as the GPU arm, bounded so that "
              "steps+warmup fit ~%d s), one per host core concurrently, compute_rates() of the %s; the reference's cost "
              "grows ~n^2.1, so this bounded sample overstates its full-size (99 856 minima) rate by ~%.0fx" %
              (cores, side * side, int(REF_BUDGET_S),
               "unmodified reference C++ NGT (oracle/_ref)" if kind == "reference" else "oracle C++ port",
               (99856.0 / (side * side)) ** 1.1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, SIDE),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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


def dense_object(torch, Handle, dgemm_tflops, n=16384, steps=2):
    """BASELINE config 3 shape at 1 GPU: complete graph of n nodes generated on the device, A={0}, B={1}."""
    K = torch.rand(n, n, dtype=torch.float64, device="cuda")
    K.fill_diagonal_(0.0)
    h = Handle.from_dense(K, [0], [1])
    h.timed_steps(1)
    ms = h.timed_steps(steps) / steps
    st = h.stats()
    pms, pfl, pl = h.profile_schur()
    kab = float(h.rates()[0]) if False else None
    h.timed_steps(1)
    kab = float(h.rates()[0])
    out = {
        "workload": "complete graph, %d nodes, k_ij ~ U(0,1) generated on device, A={0}, B={1} (BASELINE config 3, 1 GPU)" % n,
        "ms_per_step": ms, "nodes_per_s": (n - 2) / (ms * 1e-3),
        "alg_tflops": st["alg_flops"] / (ms * 1e-3) / 1e12,
        "schur_kernel_tflops": pfl / (pms * 1e-3) / 1e12, "schur_kernel_ms": pms, "schur_launches": pl,
        "schur_share_of_step": pms / ms,
        "frac_of_dgemm_peak_whole_step": st["alg_flops"] / (ms * 1e-3) / 1e12 / dgemm_tflops,
        "frac_of_dgemm_peak_schur_kernel": pfl / (pms * 1e-3) / 1e12 / dgemm_tflops,
        "k_AB": kab, "launches_per_step": st["n_launches"],
    }
    h.close()
    del K
    torch.cuda.empty_cache()
    return out


def other_configs_object(torch, Handle):
    """the remaining BASELINE configs at 1 GPU, each a short measurement (details: DESIGN.md §5, profiles/)"""
    from kmc_rates_b200 import GraphReduction, synth
    out = {}
    # config 1: README-style random complete graph, 100 nodes, through the drop-in GraphReduction API
    rng = np.random.default_rng(0)
    rates = {(i, j): float(rng.random()) for i in range(100) for j in range(100) if i != j}
    GraphReduction(rates, [0], [1]).compute_rates()
    t0 = time.perf_counter()
    red = GraphReduction(rates, [0], [1])
    red.compute_rates()
    kab = red.get_rate_AB()
    out["config1_complete100_GraphReduction_api"] = {"ms_end_to_end": 1e3 * (time.perf_counter() - t0), "k_AB": kab}
    # config 4: 2*10^4-node landscape, |A| = |B| = 512 (phase 2 = 1024 independent per-a eliminations)
    n, src, dst, k, A, B = synth.landscape_groups(141, 512, 512, T=TEMPERATURE, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B)
    h.compute_rates()
    h.phase_one(); h.phase_two()
    st = h.stats()
    out["config4_19881_nodes_A512_B512"] = {"phase1_ms": st["ms_phase1"], "phase2_ms": st["ms_phase2"],
                                            "per_a_reductions_per_s": 1024 / (st["ms_phase2"] * 1e-3),
                                            "k_AB": float(h.rates()[0]), "k_BA": float(h.rates()[1])}
    h.close()
    # config 5: 4096 independent 256-node networks (temperature sweep), one batched handle
    NB, N = 4096, 256
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    E = torch.rand(N, generator=gen, dtype=torch.float64, device="cuda")
    bar = torch.rand(N, N, generator=gen, dtype=torch.float64, device="cuda") * 0.9 + 0.1
    bar = torch.maximum(bar, bar.T)
    temps = torch.logspace(np.log10(0.05), 0.0, NB, dtype=torch.float64, device="cuda")
    K = torch.exp(-((torch.maximum(E[:, None], E[None, :]) + bar)[None] - E[None, :, None]) / temps[:, None, None])
    K.diagonal(dim1=1, dim2=2).zero_()
    h = Handle.from_dense(K, [0], [N - 1])
    h.timed_steps(1)
    ms = h.timed_steps(3) / 3
    out["config5_4096x256_sweep"] = {"ms_per_sweep": ms, "networks_per_s": NB / (ms * 1e-3),
                                     "nodes_eliminated_per_s": NB * (N - 2) / (ms * 1e-3)}
    h.close()
    del K
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=SIDE)
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dense-n", type=int, default=16384)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from kmc_rates_b200 import synth
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
    barrier()
    t0 = time.perf_counter()
    kab = 0.0
    for _ in range(e2e_steps):
        g = Handle.from_coo(n, tsrc.numpy(), tdst.numpy(), tk.numpy(), A, B, device=local)
        g.compute_rates()
        kab = float(g.rates()[0])
        g.close()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * nodes_per_step * e2e_steps / float(te.item())
    clocks = sampler.stop()
    nnz = int(src.size)
    # bytes the engine uploads per create: k (8) + slot table (8) + source index (4) per rate, 2 x 8 + 4 per node
    # (row pointers, tau slots, node ids), front index lists and maps (~4 per front entry)
    h2d = nnz * (8 + 8 + 4) + n * 20 + 4 * int(st["n_fronts"]) * 32
    d2h = (2 * 3 + 2 * 2) * 8 + n * 8
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
        step_ms_single = h.stats()["ms_profiled_pass"]   # the same profiled pass: consistent event overheads
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "schur_dram_bytes_per_launch.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("c2_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "kernel": "update_dmma_kernel (Schur update P_RR += P_RX U', FP64 DMMA m8n8k4)",
            "bound": "tensor", "achieved": pfl / (pms * 1e-3) / 1e12, "peak": dgemm, "unit": "TFLOP/s",
            "frac": pfl / (pms * 1e-3) / 1e12 / dgemm, "traffic": traffic,
            "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (best of 6, CUDA events); "
                           "MEASURED_PEAKS.json carries no FP64 figure (hbm_gbs=%s)" % peaks.get("hbm_gbs"),
            "flops_per_step": pfl, "launches_per_step": pl, "kernel_ms_per_step": pms,
            "share_of_step": pms / step_ms_single,
            "share_note": "kernel CUDA-event time / duration of the same profiled (event-bracketed, graph-free) pass; "
                          "ncu launch list of the same step: profiles/r01_launches_c2.md (panel_kernel, a 32-step "
                          "Gauss-Jordan dependency chain per launch, is co-dominant there)",
            "hbm_model": {
                "note": "SURVEY §8(d) sparse-phase model: bytes a one-node-at-a-time elimination would stream",
                "alg_bytes_per_step": st["alg_bytes"], "achieved_GBs": st["alg_bytes"] / (ms_per_step * 1e-3) / 1e9,
                "peak_GBs": peaks.get("hbm_gbs"),
                "frac": (st["alg_bytes"] / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
            },
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline_object()
        dense = None
        if world == 1 and not args.no_dense:
            try:
                dense = dense_object(torch, Handle, dgemm, n=args.dense_n)
            except Exception as e:   # report, do not hide
                dense = {"error": repr(e)}
        other = None
        if world == 1 and not args.no_dense:
            try:
                other = other_configs_object(torch, Handle)
            except Exception as e:
                other = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, side),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(te.item()) / e2e_steps,
                    "includes": "host nested-dissection ordering + symbolic analysis, every H2D copy, kernels, D2H of rates",
                    "reuse_value": reuse_value,
                    "reuse_note": "set_rates(host k) + compute_rates() on an already analysed handle (temperature sweep)"},
            "gpu_launches": int(st["n_launches"]) * args.steps,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "dense": dense,
            "other_configs": other,
            "stats": {"nodes_eliminated_per_step": nodes_per_step, "rates": nnz, "fronts": st["n_fronts"],
                      "levels": st["n_levels"], "max_front": st["max_front"], "pool_GB": st["pool_doubles"] * 8 / 1e9,
                      "alg_flops": st["alg_flops"], "ms_symbolic_host": st["ms_symbolic"], "ms_ingest": st["ms_ingest"],
                      "ms_phase1": st["ms_phase1"], "ms_phase2": st["ms_phase2"],
                      "cuda_graph": os.environ.get("NGT_B200_NO_GRAPH") is None, "k_AB": float(rates[0]),
                      "k_BA": float(rates[1]), "k_AB_e2e": kab},
        }
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
