cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_07_bench2.json 2> gpurun_out/r2_07_bench2.err; echo "rc $?" >> gpurun_out/r2_07_bench2.err
timeout 600 python -m pytest tests/test_peer_multigpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_07_peer.log
