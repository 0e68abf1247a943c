"""torchrun --nproc-per-node N scripts/dp_check.py : N-GPU batch-sharded training must equal single-GPU
full-batch training (same seeds, eps sliced from the global draw, SyncBatchNorm, flat-buffer all-reduce)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vae_song_b200 import model, train

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
solo = dist.new_group(ranks=[0])

def make():
    torch.manual_seed(0)
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[128, 256], hidden_channels=[16, 8], inverse_lipschitz=0.2, beta=0.5)
    rng = np.random.default_rng(3)
    with torch.no_grad():
        for ic in (m.decoder[0], m.decoder[1]):
            H = ic.hidden_channel
            ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
            ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
            ic.A[0].bias.copy_(torch.tensor(rng.normal(-0.3, 1.0, (H,)), dtype=torch.float32))
    return m.to(dev).train()

Bg = 1024
g = torch.Generator(device="cpu").manual_seed(5)
X = [torch.randn(Bg, 2, generator=g) for _ in range(3)]
E = [torch.randn(Bg, 2, generator=g) for _ in range(3)]
tr = train.DataParallelTrainer(make(), lr=1e-3)
lo, hi = train.shard_rows(Bg, rank, world)
losses = []
for x, e in zip(X, E):
    total, rec, reg = tr.step(x[lo:hi].to(dev), e[lo:hi].to(dev))
    losses.append(float(tr.global_losses(total)[0]))
ok = True
if rank == 0:
    ref = train.DataParallelTrainer(make(), lr=1e-3, process_group=solo)
    ref.world, ref.rank = 1, 0
    rl = [float(ref.step(x.to(dev), e.to(dev))[0]) for x, e in zip(X, E)]
    a, b = tr.fp.flat, ref.fp.flat
    err = float((a - b).abs().max() / b.abs().max())
    print(f"world={world} losses sharded {losses} single {rl}  max param rel diff {err:.2e}")
    ok = err < 5e-4 and all(abs(p - q) <= 2e-4 * abs(q) for p, q in zip(losses, rl))
    print("DP_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
