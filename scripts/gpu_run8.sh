cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -60 > gpurun_out/r2_08_pytest.log
timeout 600 python scripts/c3_run.py tf32x3 > gpurun_out/r2_08_c3.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_08_bench.json 2> gpurun_out/r2_08_bench.err; echo "bench rc $?" >> gpurun_out/r2_08_bench.err
