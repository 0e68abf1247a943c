"""Where the end-to-end step differs from the device-resident step (1 GPU): per-step CUDA events around variants of the loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vae_song_b200 import model, train

dev = torch.device("cuda:0")
B = 65536
m = bench.trained_like_(model.LIDVAE(precision="f16x3", **bench.MODEL_KW)).to(dev).train()
tr = train.DataParallelTrainer(m, lr=1e-3)
rng = np.random.default_rng(100)
host = [torch.from_numpy(bench.chessboard(B, rng)).pin_memory() for _ in range(4)]
devp = [h.to(dev) for h in host]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
tr.step(devp[0]); tr.capture(devp[0])
ring = torch.zeros(2, dtype=torch.float32).pin_memory()
dsink = torch.zeros(1, device=dev)

def timed(fn, K=40, W=5):
    for i in range(W): fn(i)
    torch.cuda.synchronize()
    evs = []
    for i in range(K):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / K

def v0(i): tr.step_graphed(devp[i % 4])
def v1(i):
    t, _, _ = tr.step_graphed(devp[i % 4]); ring[i & 1:(i & 1) + 1].copy_(t.reshape(1), non_blocking=True)
def v1b(i):
    t, _, _ = tr.step_graphed(devp[i % 4]); dsink.copy_(t.reshape(1), non_blocking=True)
def v2(i): tr.step_graphed(host[i % 4], next_x=host[(i + 1) % 4])
def v2b(i): tr.step_graphed(host[i % 4])
def v3(i):
    t, _, _ = tr.step_graphed(host[i % 4], next_x=host[(i + 1) % 4]); ring[i & 1:(i & 1) + 1].copy_(t.reshape(1), non_blocking=True)
def v4(i):
    t, _, _ = tr.step_graphed(host[i % 4], next_x=host[(i + 1) % 4]); ring[i & 1:(i & 1) + 1].copy_(t.reshape(1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
for name, fn in (("resident", v0), ("resident + loss D2H (pinned, async)", v1), ("resident + loss D2D", v1b), ("host x, prefetched", v2),
                 ("host x, direct H2D", v2b), ("host x prefetched + loss D2H", v3), ("... + host sync each step", v4), ("resident again", v0)):
    print(f"{name:45s} {timed(fn):.4f} ms", flush=True)
