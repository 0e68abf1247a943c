set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_01_smi.txt
for kc in 0 256 128; do B200VAE_KCHUNK=$kc timeout 300 python tests/tools/tc_check.py --fwd-only > gpurun_out/r2_01_tc_kchunk$kc.log 2>&1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_01_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_01_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; echo "bench rc $?" >> gpurun_out/r2_01_bench.err
tail -3 gpurun_out/r2_01_pytest.log
