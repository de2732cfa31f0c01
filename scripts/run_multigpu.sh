# Multi-GPU batch (gpurun --gpus N): bench.py under torchrun at N ranks + scripts/nccl_check.py.  Outputs under gpurun_out/final_nN/.
set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
O=gpurun_out/final_n${N}; mkdir -p $O
nvidia-smi -L > $O/smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_$N.json 2> $O/bench_$N.err; echo "bench rc=$?"
tail -c 600 $O/bench_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_check.py > $O/nccl_check.log 2>&1; echo "nccl_check rc=$?"; tail -2 $O/nccl_check.log
echo done
