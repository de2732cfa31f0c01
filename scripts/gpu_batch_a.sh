set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a/bench.err
NGT_B200_TIMING=1 timeout 120 python scripts/e2e_breakdown.py > gpurun_out/r2a/e2e.log 2>&1
timeout 200 python scripts/kernel_shares.py > gpurun_out/r2a/shares.log 2>&1
for w in c2 c4 c5; do
  NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2a/launches_$w.csv python scripts/profile_step.py $w > gpurun_out/r2a/ncu_$w.log 2>&1
done
echo done
