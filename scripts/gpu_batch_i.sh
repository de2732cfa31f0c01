set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
N=${1:-2}
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python scripts/profile_step.py dist 16384 > $O/dist_new.log 2>&1; tail -n 1 $O/dist_new.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_dist.csv python scripts/profile_step.py dist 8192 --warmup 0 > $O/ncu_dist.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_$N.json 2> $O/bench_$N.err; echo "bench rc=$?"
echo done
