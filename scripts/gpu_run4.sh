set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_04_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_04_bench.json 2> gpurun_out/r2_04_bench.err; echo "bench rc $?" >> gpurun_out/r2_04_bench.err
tail -5 gpurun_out/r2_04_pytest.log
