set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f; mkdir -p $O
./build/panel_bench 1000 600 > $O/panel_bench_1000.log 2>&1; head -4 $O/panel_bench_1000.log; grep cluster $O/panel_bench_1000.log
NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_dnclk.so timeout 120 python scripts/profile_step.py c5 > $O/dnclk.log 2>&1; tail -3 $O/dnclk.log
for t in 1 4 16 32; do NGT_B200_HOST_THREADS=$t NGT_B200_TIMING=1 ./build/analysis_bench /dev/null 1 > /dev/null 2>&1; done
python scripts/dump_graph.py 316 /tmp/g316.bin > /dev/null
for t in 1 4 8 16 32; do echo "threads $t" >> $O/analysis.log; NGT_B200_HOST_THREADS=$t NGT_B200_TIMING=1 ./build/analysis_bench /tmp/g316.bin 3 2>&1 | grep "nested dissection\|rep 2" >> $O/analysis.log; done
cat $O/analysis.log
NGT_B200_TIMING=1 timeout 120 python scripts/e2e_breakdown.py > $O/e2e.log 2>&1; grep "^rep" $O/e2e.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 200 python scripts/kernel_shares.py > $O/shares.log 2>&1; cat $O/shares.log
echo done
