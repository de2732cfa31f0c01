mkdir -p gpurun_out/sw
timeout 300 python -m pytest tests/test_gpu_ordering.py -x -q > gpurun_out/sw/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/sw/pytest.log
NGT_B200_TIMING=1 NGT_B200_ND_TRACE=1 timeout 200 python scripts/e2e_breakdown.py > gpurun_out/sw/e2e.log 2>&1
NGT_B200_TIMING=1 timeout 300 python scripts/big_sparse.py > gpurun_out/sw/big_dev.log 2>&1
HOST_SWEEPS=1 NGT_B200_TIMING=1 timeout 300 python scripts/big_sparse.py > gpurun_out/sw/big_host.log 2>&1
