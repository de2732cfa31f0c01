set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
N=${1:-2}
timeout 600 python -m pytest tests/test_distributed.py tests/test_utils.py -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_check.py > $O/nccl_check.log 2>&1; echo "nccl_check rc=$?"; tail -3 $O/nccl_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_$N.json 2> $O/bench_$N.err; echo "bench rc=$?"
NGT_B200_DIST_PANELWISE_APPLY=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_${N}_oldapply.json 2> $O/bench_${N}_oldapply.err
echo done
