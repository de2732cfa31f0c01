cd $GRAFT_REPO_ROOT
N=${1:-8}
O=gpurun_out/r2_trace$N; mkdir -p $O
NGT_B200_DIST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_trace.py > $O/trace.log 2>&1
grep "rep\|dist trace" $O/trace.log | tail -12
