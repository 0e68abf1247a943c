cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_icnn_wide_gpu.py tests/test_main_driver.py tests/test_model_gpu.py -m gpu -q -s 2>&1 | grep -v "^    \|^$" > gpurun_out/r2_06_pytest.log
timeout 300 python scripts/wide_check.py > gpurun_out/r2_06_wide_check.log 2>&1
