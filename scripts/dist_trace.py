"""Where does a block's time go in the multi-GPU dense loop?  (torchrun; NGT_B200_DIST_TRACE=1 prints per-stage sums)
    NGT_B200_DIST_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dist_trace.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from kmc_rates_b200._capi import Handle
from kmc_rates_b200.distributed import communicator
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
K = torch.rand(n, n, generator=gen, dtype=torch.float64, device="cuda"); K.fill_diagonal_(0.0)
h = Handle.from_dense(K, [0], [1])
comm = communicator()
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier()
    ms = h.dist_phase_one(comm)
    if rank == 0:
        print("rep %d: %.2f ms" % (rep, ms), flush=True)
dist.barrier(); dist.destroy_process_group()
