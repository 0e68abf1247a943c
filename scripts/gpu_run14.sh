cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_families_gpu.py -m gpu -q 2>&1 | tail -2 > gpurun_out/r2_14_pytest.log
timeout 600 python scripts/c3_run.py tf32x3 2>&1 | grep "C3 train" > gpurun_out/r2_14_c3.log
