# f16x3 profiling pass: gpurun --timeout 1800 -- "bash scripts/gpu_profile_f16.sh"
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/c3_run.py f16x3 > gpurun_out/c3_run_f16.log 2>&1
python scripts/profile_decode.py --precision f16x3 --H 1024 --bwd > gpurun_out/profile_dec_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"icnn_tc3_fwd|icnn_tc3_bwd_kernel|icnn_tc3_dP0" -c 6 -o gpurun_out/prof_r2_f16 -f python scripts/profile_decode.py --precision f16x3 --H 1024 --bwd --iters 2 > gpurun_out/profile_ncu_f16.log 2>&1
python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/profile_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/profile_launches_bench_f16.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/profile_ncu3.log 2>&1
