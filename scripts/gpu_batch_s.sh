cd $GRAFT_REPO_ROOT
N=${1:-2}
O=gpurun_out/split_n$N; mkdir -p $O
timeout 300 python -m pytest tests/test_distributed.py -m gpu -q > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_check.py > $O/nccl_check.log 2>&1; echo "nccl_check rc=$?"; tail -2 $O/nccl_check.log
NGT_B200_DIST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/dist_trace.py > $O/trace.log 2>&1
grep "^rep" $O/trace.log; grep "rank 0 of" $O/trace.log | tail -1
