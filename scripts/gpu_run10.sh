cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -60 > gpurun_out/r2_10_pytest.log
timeout 300 python tests/tools/tc_check.py --fwd-only 2>&1 | grep "prec=3\|TIMING tf32x3" > gpurun_out/r2_10_tc.log
timeout 600 python scripts/c3_run.py tf32x3 2>&1 | grep "C3" > gpurun_out/r2_10_c3.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err; echo "bench rc $?" >> gpurun_out/r2_10_bench.err
