"""torchrun --nproc-per-node N scripts/dp_check.py [--time]: N-GPU batch-sharded training must equal single-GPU
full-batch training (same seeds, eps sliced from the global draw, cross-rank BatchNorm, gradient all-reduce), for both
exchange back-ends: comm="nccl" (torch.distributed collectives) and comm="peer" (kernels over NVLink peer memory,
csrc/peer.cuh).  Also checks the raw peer all-gather, CUDA-graph replay of the peer step, and (with --time) times the
two back-ends on the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vae_song_b200 import model, train, peer as peer_mod

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
solo = dist.new_group(ranks=[0])
ok = True


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def make(icnn=(128, 256), hidden=(16, 8), precision="fp32"):
    torch.manual_seed(0)
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=list(icnn), hidden_channels=list(hidden), inverse_lipschitz=0.2,
                     beta=0.5, precision=precision)
    rng = np.random.default_rng(3)
    with torch.no_grad():
        for ic in (m.decoder[0], m.decoder[1]):
            H = ic.hidden_channel
            ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
            ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
            ic.A[0].bias.copy_(torch.tensor(rng.normal(-0.3, 1.0, (H,)), dtype=torch.float32))
    return m.to(dev).train()


# ---- 1. raw peer all-gather ---------------------------------------------------------------------------------------
comm = peer_mod.PeerComm()
for n in (1, 7, 385):
    x = torch.arange(n, device=dev, dtype=torch.float32) + 1000.0 * rank
    for it in range(5):                                       # repeated use of one slot (parity double buffer)
        out = comm.allgather(x + it, slot=3)
        want = torch.stack([torch.arange(n, device=dev, dtype=torch.float32) + 1000.0 * r + it for r in range(world)])
        if not torch.equal(out, want):
            ok = False
            say(f"peer allgather n={n} it={it} MISMATCH")
comm.check()
t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = torch.randn(385, device=dev)
outs = [torch.empty_like(x) for _ in range(world)]
for name, fn in (("peer", lambda: comm.allgather(x, slot=4)), ("nccl", lambda: dist.all_gather(outs, x))):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t[0].record()
    for _ in range(200):
        fn()
    t[1].record(); torch.cuda.synchronize()
    say(f"allgather of 385 floats, {world} ranks, {name}: {t[0].elapsed_time(t[1]) / 200 * 1e3:.1f} us per call (launch-inclusive)")
say("PEER_ALLGATHER", "PASS" if ok else "FAIL")

# ---- 2. sharded == single process, both back-ends, eager and graphed ---------------------------------------------------
Bg = 1024
g = torch.Generator(device="cpu").manual_seed(5)
X = [torch.randn(Bg, 2, generator=g) for _ in range(4)]
E = [torch.randn(Bg, 2, generator=g) for _ in range(4)]
lo, hi = train.shard_rows(Bg, rank, world)
rl = ref_flat = None
if rank == 0:
    ref = train.DataParallelTrainer(make(), lr=1e-3, process_group=solo)
    ref.world, ref.rank = 1, 0
    rl = [float(ref.step(x.to(dev), e.to(dev))[0]) for x, e in zip(X, E)]
    ref_flat = ref.fp.flat.clone()
for mode, graphed in (("nccl", False), ("peer", False), ("peer", True)):
    tr = train.DataParallelTrainer(make(), lr=1e-3, comm=mode)
    if graphed:
        tr.capture(X[0][lo:hi].to(dev), E[0][lo:hi].to(dev))
    losses = []
    for x, e in zip(X, E):
        stepf = tr.step_graphed if graphed else tr.step
        total = stepf(x[lo:hi].to(dev), e[lo:hi].to(dev))[0]
        losses.append(float(tr.global_losses(total)[0]))
    if tr.peer is not None:
        tr.peer.check()
    # replicas must be bit-identical across ranks
    mine = tr.fp.flat[:tr.fp.numel].clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(mine, ref0))
    flags = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        err = float((mine - ref_flat[:tr.fp.numel]).abs().max() / ref_flat.abs().max())
        good = err < 5e-4 and all(abs(p - q) <= 2e-4 * abs(q) for p, q in zip(losses, rl)) and int(flags) == 1
        ok = ok and good
        print(f"world={world} comm={mode} graphed={graphed}: losses {['%.6f' % v for v in losses]} single {['%.6f' % v for v in rl]} "
              f"max param rel diff {err:.2e} replicas identical {int(flags) == 1} -> {'PASS' if good else 'FAIL'}", flush=True)
    del tr
say("DP_CHECK", "PASS" if ok else "FAIL")

# ---- 2b. all-pairs Lipschitz estimator: pair tiles sharded over the ranks == single rank -----------------------------------
from vae_song_b200 import utils as vutils
mm = make().eval()
torch.manual_seed(9)
Xl = torch.randn(3000, 2, device=dev)
dist.broadcast(Xl, src=0)
sharded = vutils.estimate_lipschitz_allpairs(mm.decode, Xl)                      # default group: tiles split by range
single = vutils.estimate_lipschitz_allpairs(mm.decode, Xl, process_group=solo) if rank == 0 else None
if rank == 0:
    good = (sharded["count"] == single["count"] == 3000 * 2999 // 2 and sharded["max"] == single["max"]
            and sharded["min"] == single["min"] and abs(sharded["mean"] - single["mean"]) <= 1e-6 * abs(single["mean"]))   # fp32 per-CTA partial sums: order differs
    ok = ok and good
    print(f"all-pairs Lipschitz sharded over {world} ranks: max {sharded['max']:.6g} min {sharded['min']:.6g} mean "
          f"{sharded['mean']:.9g} (single {single['mean']:.9g}) -> {'PASS' if good else 'FAIL'}", flush=True)

# ---- 3. timing on the bench workload --------------------------------------------------------------------------------------
if "--time" in sys.argv:
    B = 65536
    for prec in ("f16x3", "tf32x3", "tf32"):
        for mode in ("nccl", "peer"):
            torch.manual_seed(1)
            m = model.LIDVAE(dataset="pinwheel", inverse_lipschitz=0.2, beta=1.0, precision=prec).to(dev).train()
            tr = train.DataParallelTrainer(m, lr=1e-3, comm=mode)
            x = torch.randn(B, 2, device=dev); e = torch.randn(B, 2, device=dev)
            tr.capture(x, e)
            for _ in range(5):
                tr.step_graphed(x, e)
            torch.cuda.synchronize(); dist.barrier()
            ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev[0].record()
            for _ in range(30):
                tr.step_graphed(x, e)
            ev[1].record(); torch.cuda.synchronize()
            ms = torch.tensor([ev[0].elapsed_time(ev[1]) / 30], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if tr.peer is not None:
                tr.peer.check()
            say(f"train step {prec} comm={mode} world={world} B/GPU={B}: {float(ms):.3f} ms -> {world * B / float(ms) * 1e3 / 1e6:.1f} M samples/s")
            del tr, m

dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)      # no NCCL communicator teardown: it can block for minutes with live CUDA graphs
