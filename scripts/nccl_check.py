"""Real multi-rank NCCL check of the sharded paths (launched by torchrun; one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/nccl_check.py

Every rank compares (a) ngt_dist_phase_one (one dense network over the ranks) and (b) ngt_phase_two_sharded (with and
without the broadcast of rank 0's reduced block) with the same network computed on its own GPU alone; rank 0 prints
NCCL_CHECK_OK when every rank agreed to 1e-10 (single-GPU vs CPU oracle is covered by the -m gpu tests)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dense-n", type=int, default=1100)
    ap.add_argument("--side", type=int, default=40)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    from kmc_rates_b200.distributed import communicator, compute_rates_sharded, dense_phase_one_distributed

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = communicator()
    worst = 0.0

    # (a) one dense network over the ranks
    n = a.dense_n
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234)                      # same matrix on every rank
    K = torch.rand(n, n, generator=gen, dtype=torch.float64, device="cuda")
    K.fill_diagonal_(0.0)
    A, B = [0, 1], [n - 3, n - 2, n - 1]
    ref = Handle.from_dense(K, A, B)
    ref.compute_rates()
    h = Handle.from_dense(K, A, B)
    ms = dense_phase_one_distributed(h, comm=comm)
    h.phase_two()
    e = float(np.max(np.abs(h.rates() - ref.rates()) / np.abs(ref.rates())))
    P, t = h.reduced()
    Pr, tr = ref.reduced()
    e = max(e, float(np.max(np.abs(P - Pr) / np.maximum(np.abs(Pr), 1e-300))), float(np.max(np.abs(t - tr) / tr)))
    worst = max(worst, e)
    print("rank %d: dense n=%d over %d ranks: %.2f ms, max rel err vs single GPU %.2e" % (rank, n, world, ms, e), flush=True)

    # (b) sharded phase 2 on a landscape with multi-member A and B
    n2, src, dst, k, A2, B2 = synth.landscape_groups(a.side, 37, 29, T=0.3, seed=5)
    ref2 = Handle.from_coo(n2, src, dst, k, A2, B2, device=local)
    ref2.compute_rates()
    for bc in (False, True):
        h2 = Handle.from_coo(n2, src, dst, k, A2, B2, device=local)
        ms2 = compute_rates_sharded(h2, comm=comm, broadcast=bc)
        e2 = float(np.max(np.abs(h2.rates() - ref2.rates()) / np.abs(ref2.rates())))
        om, tau = h2.final()
        omr, taur = ref2.final()
        e2 = max(e2, float(np.max(np.abs(om - omr) / omr)), float(np.max(np.abs(tau - taur) / taur)))
        worst = max(worst, e2)
        print("rank %d: phase 2 sharded over %d ranks (broadcast=%s): %.3f ms, max rel err %.2e" % (rank, world, bc, ms2, e2), flush=True)

    t = torch.tensor([worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t.item()) < 1e-10
    if rank == 0:
        print("NCCL_CHECK_OK" if ok else "NCCL_CHECK_FAILED", "max rel err over ranks %.2e" % float(t.item()), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
