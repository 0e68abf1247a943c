cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 800 --csv --log-file gpurun_out/r2_15_launches_c3_warm.csv python scripts/profile_c3.py > gpurun_out/r2_15_ncu.log 2>&1
