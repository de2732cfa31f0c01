cd $GRAFT_REPO_ROOT
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_dnclk.so timeout 300 python scripts/profile_step.py dist 8192 --warmup 0 > $O/clk.log 2>&1; head -4 $O/clk.log; tail -n 2 $O/clk.log
timeout 300 python scripts/profile_step.py dist 16384 > $O/dist_new.log 2>&1; tail -n 1 $O/dist_new.log
