"""BASELINE configs[4] (configs/config_shapenet_setlrvae.yaml: attention Set LR-VAE on 2048-point clouds, latent 128):
train-step time on synthetic x ~ N(0, I3) [B, 2048, 3], batch-sharded with train.DataParallelTrainer.

  python scripts/c5_run.py                                       # 1 GPU, batch 16 (the config's batch)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/c5_run.py

Per run: (a) weak scaling, 16 clouds per GPU; (b) the config's global batch 16 split over the ranks (strong scaling);
each launched eagerly and replayed as one CUDA graph (gradient all-reduce + Adam inside the graph: the peer-memory kernel of
csrc/peer.cu when CUDA IPC mapping works, else NCCL).  On 1 GPU additionally: the same step with the reference's Chamfer
formulation (torch.cdist materialising [B,2048,2048], model.py:896-912) in place of the tiled nearest-neighbour kernel.
The transformer encoder / decoder stacks are stock torch.nn (out of scope); ours on this path are the Chamfer kernel, the fused
reparam / KL / latent-recon loss kernel and the exchange + Adam kernel.  Timing: CUDA events, max over ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from vae_song_b200 import main as M, model, train

rank, world, lrank = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = M.load_config(os.path.join(ROOT, "configs", "config_shapenet_setlrvae.yaml"))
NP = cfg["model_params"]["num_points"]
GB = cfg["common_params"]["batch_size"]


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def build():
    torch.manual_seed(0)
    tag, m, kw = next(M.iter_models(cfg))
    m = m.to(dev).train()
    m.wu_alpha = 1.0
    return tag, m, kw


def timed(fn, n, warm=5):
    for i in range(warm):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        out = fn(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), out


def run(per_gpu, label):
    g = torch.Generator(device="cpu").manual_seed(42 + rank)
    pool = torch.randn(8, per_gpu, NP, 3, generator=g).to(dev)
    tag, m, kw = build()
    tr = train.DataParallelTrainer(m, lr=1e-2, grad_clip=kw["grad_clip"])
    ms_e, out = timed(lambda i: tr.step(pool[i % 8]), 20)
    tr.capture(pool[0])
    ms_g, outg = timed(lambda i: tr.step_graphed(pool[i % 8]), 50)
    tr.check()
    tot = per_gpu * world
    comm = "none" if world == 1 else ("peer-memory kernel" if tr.peer is not None else "NCCL")
    say(f"C5 {tag} {label}: {world} GPU x {per_gpu} clouds (global batch {tot}), exchange {comm}: eager {ms_e:.3f} ms/step "
        f"({tot / ms_e * 1e3:.0f} clouds/s), one CUDA graph {ms_g:.3f} ms/step ({tot / ms_g * 1e3:.0f} clouds/s = "
        f"{tot * NP / ms_g / 1e3:.2f} M points/s), loss {float(out[0]):.4f}")
    return ms_g


run(GB, "weak")
if world > 1 and GB % world == 0:
    run(GB // world, "strong")
if world == 1:
    def cdist_chamfer(pred, gt):                                   # the reference's formulation
        d = torch.cdist(pred, gt, p=2) ** 2
        return (d.min(dim=2)[0].mean(dim=1) + d.min(dim=1)[0].mean(dim=1)).mean()
    ours = model.chamfer_distance
    model.chamfer_distance = cdist_chamfer
    try:
        run(GB, "weak, Chamfer through torch.cdist (reference formulation)")
    finally:
        model.chamfer_distance = ours
    npar = sum(p.numel() for p in build()[1].parameters())
    say(f"C5 model: {npar} parameters")
if world > 1:
    dist.destroy_process_group()
