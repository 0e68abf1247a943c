set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for kc in 0 512 256 128; do B200VAE_KCHUNK=$kc timeout 300 python tests/tools/tc_check.py --fwd-only 2>&1 | grep "TIMING tf32x3\|H=1024 B=512 default prec=3" > gpurun_out/r2_03_tc_k$kc.log; done
head -50 gpurun_out/r2_03_tc_k*.log
