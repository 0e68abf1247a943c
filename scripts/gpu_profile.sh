# ncu captures (launch list + --set full) used for profiles/: gpurun --timeout 1800 -- "bash scripts/gpu_profile.sh" (round-2 tf32x3 / all-pairs captures; the f16x3 evidence pass is scripts/gpu_evidence.sh)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/profile_allpairs.py 50000 > gpurun_out/profile_ap_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:allpairs -c 2 -o gpurun_out/prof_r2_allpairs -f python scripts/profile_allpairs.py 50000 > gpurun_out/profile_ncu.log 2>&1
python scripts/profile_decode.py --precision tf32x3 --H 1024 > gpurun_out/profile_dec_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:icnn_tc3_fwd -c 2 -o gpurun_out/prof_r2_tc3_fwd_x3 -f python scripts/profile_decode.py --precision tf32x3 --H 1024 > gpurun_out/profile_ncu2.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['extra']['lipschitz_estimator']))" > gpurun_out/profile_ap_bench.json
STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/profile_launches_c3_warm.csv python scripts/profile_c3.py > gpurun_out/profile_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/profile_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/profile_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/profile_ncu3.log 2>&1
