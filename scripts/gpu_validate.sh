# Round-end style validation on a GPU box: gpurun --timeout 3000 -- "bash scripts/gpu_validate.sh" (outputs under gpurun_out/)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -40 > gpurun_out/validate_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/validate_smoke.log
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/validate_bench_ref.json 2> gpurun_out/validate_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
