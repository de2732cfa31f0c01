set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/final1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,driver_version --format=csv > $O/smi.txt; lscpu | head -20 > $O/lscpu.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
/usr/bin/time -v timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; grep "Elapsed" $O/bench_n1.err
/usr/bin/time -v timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; grep "Elapsed" $O/bench_ref.err
for w in c2 c4 c5; do
  NGT_B200_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file $O/launches_$w.csv python scripts/profile_step.py $w > $O/ncu_$w.log 2>&1
done
NGT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^panel_kernel -s 150 -c 1 -o $O/full_panel_c2 python scripts/profile_step.py c2 > $O/ncu_full_panel.log 2>&1
NGT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:update_dmma_kernel -s 100 -c 1 -o $O/full_update_c2 python scripts/profile_step.py c2 > $O/ncu_full_update.log 2>&1
NGT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_network -s 1 -c 1 -o $O/full_dense_network_c5 python scripts/profile_step.py c5 > $O/ncu_full_dn.log 2>&1
NGT_B200_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:update_dmma_tmap -s 8 -c 1 -o $O/full_tmap_dense8192 python scripts/profile_step.py dense 8192 > $O/ncu_full_tmap.log 2>&1
ls -la $O
echo done
