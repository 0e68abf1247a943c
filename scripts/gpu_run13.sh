cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/profile_c3.py > gpurun_out/r2_13_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_13_launches_c3.csv python scripts/profile_c3.py > gpurun_out/r2_13_ncu.log 2>&1
