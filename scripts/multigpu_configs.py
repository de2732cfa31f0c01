"""Multi-GPU runs of the two BASELINE configs that shard (run under torchrun, one rank per GPU):

  config 5  batched temperature sweep: 4096 independent 256-node networks; contiguous blocks of networks per rank,
            no data-path collective; rates gathered at the end.
  config 4  2*10^4-node landscape, |A| = |B| = 512: phase 1 replicated, per-a reductions of phase 2 sharded over the
            ranks, one all-reduce of the (1-P'_aa, tau'_a) pairs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multigpu_configs.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from kmc_rates_b200 import synth
from kmc_rates_b200._capi import Handle
from kmc_rates_b200.distributed import shard_range, gather_rates, compute_rates_sharded, all_reduce_sum

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def tmax(x):
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"n_gpus": world}
# ---------------- config 5 ----------------
NB, N = int(os.environ.get("C5_NETS", "4096")), 256
lo, hi = shard_range(NB, rank, world)
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
E = torch.rand(N, generator=gen, dtype=torch.float64, device="cuda")
bar = torch.rand(N, N, generator=gen, dtype=torch.float64, device="cuda") * 0.9 + 0.1
bar = torch.maximum(bar, bar.T)
Ets = torch.maximum(E[:, None], E[None, :]) + bar
temps = torch.logspace(np.log10(0.05), 0.0, NB, dtype=torch.float64, device="cuda")[lo:hi]
K = torch.exp(-(Ets[None] - E[None, :, None]) / temps[:, None, None])
K.diagonal(dim1=1, dim2=2).zero_()
h = Handle.from_dense(K, [0], [N - 1], device=local)
h.timed_steps(1)
barrier()
ms = h.timed_steps(3) / 3
ms = tmax(ms)
r = h.rates().reshape(-1, 4)
full = gather_rates(r, NB, rank, world) if world > 1 else r
out["config5"] = {"networks": NB, "nodes_per_network": N, "ms_per_sweep": ms, "networks_per_s": NB / (ms * 1e-3),
                  "nodes_eliminated_per_s": NB * (N - 2) / (ms * 1e-3), "checksum_kAB": float(np.sum(np.log(full[:, 0]))),
                  "k_AB_lowT_highT": [float(full[0, 0]), float(full[-1, 0])]}
h.close(); del K
# ---------------- config 4 ----------------
side, g = int(os.environ.get("C4_SIDE", "141")), int(os.environ.get("C4_GROUP", "512"))
n, src, dst, k, A, B = synth.landscape_groups(side, g, g, T=0.3, seed=0)
h = Handle.from_coo(n, src, dst, k, A, B, device=local)
for rep in range(3):
    barrier(); t0 = time.perf_counter()
    h.phase_one()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if world > 1:
        compute_rates_sharded.__globals__["_dist"]  # noqa: keep import alive
        from kmc_rates_b200.distributed import phase_two_sharded
        phase_two_sharded(h, rank, world, lambda a: all_reduce_sum(a))
    else:
        h.phase_two()
    barrier(); t2 = time.perf_counter()
p1, p2 = tmax(t1 - t0), tmax(t2 - t1)
rates = h.rates()
out["config4"] = {"nodes": n, "group": g, "phase1_ms_wall": 1e3 * p1, "phase2_sharded_ms_wall": 1e3 * p2,
                  "per_a_reductions_per_s": 2 * g / p2, "kAB": float(rates[0]), "kBA": float(rates[1]),
                  "kAB_SS": float(rates[2]), "kBA_SS": float(rates[3])}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
