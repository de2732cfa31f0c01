"""BASELINE config 3 over N GPUs: one complete graph, 1-D block-cyclic columns, NCCL panel broadcast.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dense_multigpu.py [n]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from kmc_rates_b200._capi import Handle
from kmc_rates_b200.distributed import dense_phase_one_distributed

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
K = torch.rand(n, n, generator=gen, dtype=torch.float64, device="cuda"); K.fill_diagonal_(0.0)
dist.broadcast(K, src=0)          # every rank starts from the same rate matrix
h = Handle.from_dense(K, [0], [1], device=local)
times = []
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    dense_phase_one_distributed(h)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    times.append(t1 - t0)
h.phase_two()
r = h.rates()
t = torch.tensor([min(times[1:])], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = {"n": n, "n_gpus": world, "phase1_s": float(t.item()), "alg_tflops": (2.0 / 3.0) * n ** 3 / float(t.item()) / 1e12,
       "nodes_per_s": (n - 2) / float(t.item()), "kAB": float(r[0]), "kBA": float(r[1])}
if rank == 0:
    ref = Handle.from_dense(K, [0], [1], device=local)
    ref.timed_steps(1); ms = ref.timed_steps(2) / 2
    out["single_gpu_ms"] = ms; out["single_gpu_kAB"] = float(ref.rates()[0])
    out["rel_diff_vs_single_gpu"] = abs(out["kAB"] - out["single_gpu_kAB"]) / out["single_gpu_kAB"]
    print(json.dumps(out))
dist.destroy_process_group()
