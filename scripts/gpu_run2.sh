set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "0 2000" "256 2000" "256 500" "256 100" "256 20" "512 2000" "512 100"; do set -- $cfg; B200VAE_KCHUNK=$1 B200VAE_PARK_NS=$2 timeout 300 python tests/tools/tc_check.py --fwd-only 2>&1 | grep "TIMING\|H=1024 B=512 default prec=3" > gpurun_out/r2_02_tc_k$1_p$2.log; done
head -50 gpurun_out/r2_02_tc_k*.log
