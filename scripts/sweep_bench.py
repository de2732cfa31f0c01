"""Timing of the device level-structure sweep (csrc/ordering.cu) on the 10^5-minima landscape; run on the GPU box with
NGT_B200_TIMING=1 to see upload / kernel / download per call."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kmc_rates_b200 import synth, _capi

side = int(sys.argv[1]) if len(sys.argv) > 1 else 316
n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0)
u = np.concatenate([src, dst]); v = np.concatenate([dst, src])
e = np.unique(np.stack([u, v], 1), axis=0)
ptr = np.zeros(n + 1, np.int64); np.add.at(ptr, e[:, 0] + 1, 1)
ptr = np.cumsum(ptr).astype(np.int32); adj = e[:, 1].astype(np.int32)
torch.zeros(1, device="cuda")
def clocks():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print("clocks idle:", clocks(), flush=True)
for rep in range(6):
    t0 = time.perf_counter(); r = _capi.level_sweep(ptr, adj, -1); t1 = time.perf_counter()
    print("idle rep %d: %.2f ms, %d levels" % (rep, 1e3 * (t1 - t0), r[2]), flush=True)
print("clocks after:", clocks(), flush=True)
# keep the GPU busy on another stream
a = torch.randn(8192, 8192, device="cuda", dtype=torch.float16)
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    for _ in range(400):
        a @ a
time.sleep(0.05)
print("clocks busy:", clocks(), flush=True)
for rep in range(4):
    t0 = time.perf_counter(); r = _capi.level_sweep(ptr, adj, -1); t1 = time.perf_counter()
    print("busy rep %d: %.2f ms" % (rep, 1e3 * (t1 - t0)), flush=True)
torch.cuda.synchronize()
