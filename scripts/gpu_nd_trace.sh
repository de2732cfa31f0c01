cd $GRAFT_REPO_ROOT
O=gpurun_out/ndtrace2; mkdir -p $O
python scripts/dump_graph.py 316 /tmp/g316.bin > /dev/null
NGT_B200_TIMING=1 NGT_B200_ND_TRACE=1 ./build/analysis_bench /tmp/g316.bin 3 > $O/trace.log 2>&1
grep "nested dissection\|rep 2" $O/trace.log | tail -3
NGT_B200_TIMING=1 timeout 120 python scripts/e2e_breakdown.py > $O/e2e.log 2>&1; grep "^rep\|nested dissection" $O/e2e.log
