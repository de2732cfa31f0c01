set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python scripts/profile_step.py dist 16384 > $O/dist_new.log 2>&1; tail -1 $O/dist_new.log
NGT_B200_DIST_RECT_FACTOR=1 NGT_B200_DIST_PANELWISE_APPLY=1 timeout 300 python scripts/profile_step.py dist 16384 > $O/dist_old.log 2>&1; tail -1 $O/dist_old.log
NGT_B200_NO_LOOKAHEAD=1 timeout 300 python scripts/profile_step.py dist 16384 > $O/dist_new_nola.log 2>&1; tail -1 $O/dist_new_nola.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_dist.csv python scripts/profile_step.py dist 8192 --warmup 0 > $O/ncu_dist.log 2>&1
echo done
