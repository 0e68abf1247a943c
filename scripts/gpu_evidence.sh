# Evidence pass of the final code (one B200): gpurun --timeout 2400 -- "bash scripts/gpu_evidence.sh"; outputs under gpurun_out/ev_*
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -12 > gpurun_out/ev_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/ev_smoke.log
timeout 600 python tests/tools/tc_check.py > gpurun_out/ev_tc_check.log 2>&1
timeout 600 python scripts/c3_run.py f16x3 > gpurun_out/ev_c3_run.log 2>&1
( timeout 300 python scripts/wide_check.py 256 tf32x3,tf32; timeout 300 python scripts/wide_check.py 8192 tf32x3,tf32 ) > gpurun_out/ev_wide_check.log 2>&1
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 2 > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err
# ncu: launch list of one eager step, then --set full of the three FP16-mode kernels (training variants) and the decode-only forward
python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ev_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ev_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ev_ncu_launch.log 2>&1
python scripts/profile_decode.py --precision f16x3 --H 1024 --bwd --iters 2 > gpurun_out/ev_dec_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"icnn_tc3_fwd|icnn_tc3_bwd_kernel|icnn_tc3_dP0" -c 6 -o gpurun_out/ev_prof_f16 -f python scripts/profile_decode.py --precision f16x3 --H 1024 --bwd --iters 2 > gpurun_out/ev_ncu_f16.log 2>&1
ncu --set full --clock-control none -k regex:"icnn_tc3_fwd" -c 2 -o gpurun_out/ev_prof_f16_decode -f python scripts/profile_decode.py --precision f16x3 --H 1024 --iters 2 > gpurun_out/ev_ncu_f16_dec.log 2>&1
