set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c; mkdir -p $O
for f in 1000 600 300; do ./build/panel_bench $f $((f*6/10)) > $O/panel_bench_$f.log 2>&1; ./build/panel_bench_oldgj $f $((f*6/10)) > $O/panel_bench_oldgj_$f.log 2>&1; done
cat $O/panel_bench_1000.log
timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_shapes.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -q --maxfail=8 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 600 $O/bench.err
NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_c5.csv python scripts/profile_step.py c5 > $O/ncu_c5.log 2>&1
NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_c2.csv python scripts/profile_step.py c2 > $O/ncu_c2.log 2>&1
echo done
