set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -25 $O/pytest.log
if ! grep -q "pytest rc=0" $O/pytest.log; then
  NGT_B200_NO_DENSE_NET=1 timeout 600 python -m pytest tests -m gpu -q --maxfail=12 > $O/pytest_nodn.log 2>&1; tail -5 $O/pytest_nodn.log
  NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_oldgj.so timeout 600 python -m pytest tests -m gpu -q --maxfail=12 > $O/pytest_oldgj.log 2>&1; tail -5 $O/pytest_oldgj.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 800 $O/bench.err
NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_oldgj.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-dense --no-sharded > $O/bench_oldgj.json 2> $O/bench_oldgj.err
NGT_B200_TIMING=1 timeout 120 python scripts/e2e_breakdown.py > $O/e2e.log 2>&1
timeout 200 python scripts/kernel_shares.py > $O/shares.log 2>&1
for w in c2 c5; do
  NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_$w.csv python scripts/profile_step.py $w > $O/ncu_$w.log 2>&1
done
echo done
