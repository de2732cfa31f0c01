# One-GPU evidence batch (run on the B200 box through gpurun): GPU tests, smoke, the ncu launch list of the config 2 step of
# THIS build -> profiles/r02_traffic.json, then the bench line that quotes it.  Outputs under gpurun_out/final4/.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/final4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 300 python scripts/profile_step.py c2 > $O/plain_c2.log 2>&1 && NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_c2.csv python scripts/profile_step.py c2 > $O/ncu_c2.log 2>&1
python scripts/make_traffic_json.py $O/launches_c2.csv 2 754850048 "NGT_B200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv python scripts/profile_step.py c2" > $O/traffic.log 2>&1; cp profiles/r02_traffic.json $O/r02_traffic.json
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
echo done
