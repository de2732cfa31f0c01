set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e; mkdir -p $O
NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_dnclk.so timeout 120 python scripts/profile_step.py c5 > $O/dnclk.log 2>&1; tail -5 $O/dnclk.log
NGT_B200_LIB=$PWD/kmc_rates_b200/csrc/libngt_b200_dnclk.so timeout 120 python scripts/profile_step.py c5 256 148 >> $O/dnclk.log 2>&1; tail -3 $O/dnclk.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo done
