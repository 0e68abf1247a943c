cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_multigpu.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_mg_peer.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "rc $?" >> gpurun_out/r2_bench2.err
