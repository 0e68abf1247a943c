cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_families_gpu.py tests/test_main_driver.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_11_pytest.log
timeout 600 python scripts/c3_run.py tf32x3 2>&1 | grep "C3" > gpurun_out/r2_11_c3.log
timeout 600 python scripts/c1_run.py 2>&1 | tail -6 > gpurun_out/r2_11_c1.log
