"""BASELINE configs[2] end to end (the reference's README.md:41-59 run: K=8 corner_heavy mixture, hidden_channels
128 64 64 32 16 8 4 2, batch 256, Adam 1e-3, then the per-cell KL / Lipschitz sweeps of lipschitz.py steps 4-6):
times the batch-256 train step (eager train_model loop vs the whole-step CUDA graph) and the evaluation, batched
(vae_song_b200.lipschitz) vs one estimator call per cell like the reference's loops."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_song_b200 import lipschitz as L, model, train
from vae_song_b200.utils import estimate_local_lipschitz, reparameterize

dev = "cuda"
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(20); np.random.seed(20)
K, Kz, B = 8, 16, 256
ds = L.GaussianMixture2D(8, 20000, center_range=K, stds=0.3, pattern="corner_heavy", seed=20)
loader = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=True, drop_last=True, generator=torch.Generator().manual_seed(20))
m = model.LIDVAE(inverse_lipschitz=0.2, beta=0.001, dataset="pinwheel", hidden_channels=[128, 64, 64, 32, 16, 8, 4, 2], precision=prec).to(dev)
rng = np.random.default_rng(1)
from vae_song_b200 import utils as vutils
for ic in (m.decoder[0], m.decoder[1]):
    vutils.trained_like_icnn_(ic, rng)

def sync(): torch.cuda.synchronize()
train.train_model(m, [b for _, b in zip(range(8), loader)], 1, 1e-3, dev)     # warm-up: module loading, allocator, cuBLAS
sync(); t0 = time.perf_counter()
train.train_model(m, loader, 1, 1e-3, dev)
sync(); t_eager = (time.perf_counter() - t0) / len(loader)
tr = train.DataParallelTrainer(m, lr=1e-3)
xb = next(iter(loader))[0].to(dev)
tr.capture(xb)
sync(); t0 = time.perf_counter()
for X, _ in loader:
    tr.step_graphed(X.to(dev, non_blocking=True))
sync(); t_graph_loader = (time.perf_counter() - t0) / len(loader)
# the step itself: device-resident batches, CUDA events around 200 replays (the DataLoader above costs more host time per
# batch of 256 than the replayed step runs)
staged = [b.to(dev) for b, _ in zip((x for x, _ in loader), range(16))]
for i in range(20):
    tr.step_graphed(staged[i % 16])
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200):
    tr.step_graphed(staged[i % 16])
e1.record(); sync()
t_graph = e0.elapsed_time(e1) / 200 * 1e-3
print(f"C3 whole-step CUDA graph fed by the DataLoader (host-bound): {t_graph_loader * 1e3:.3f} ms/step")
# the UNMODIFIED reference (oracle/_ref, populated by oracle/fetch_ref.py) on the same GPU, same loader, same loop
t_ref = None
try:
    from oracle import fetch_ref
    ns = fetch_ref.import_ref()
    torch.manual_seed(20)
    mr = ns.model.LIDVAE(inverse_lipschitz=0.2, beta=0.001, dataset="pinwheel", hidden_channels=[128, 64, 64, 32, 16, 8, 4, 2]).to(dev)
    rr = np.random.default_rng(1)
    for ic in (mr.decoder[0], mr.decoder[1]):
        H = ic.A0.weight.shape[0]
        with torch.no_grad():
            ic.W[0].param.copy_(torch.tensor(rr.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
            ic.W[1].param.copy_(torch.tensor(rr.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
            ic.A[0].bias.copy_(torch.tensor(rr.normal(-0.3, 1.0, (H,)), dtype=torch.float32))
    import contextlib, io
    with contextlib.redirect_stderr(io.StringIO()):
        ns.lipschitz.train_model(mr, [b for _, b in zip(range(8), loader)], 1, 1e-3, dev)
        sync(); t0 = time.perf_counter()
        ns.lipschitz.train_model(mr, loader, 1, 1e-3, dev)
    sync(); t_ref = (time.perf_counter() - t0) / len(loader)
    print(f"C3 train step, batch {B}: UNMODIFIED reference (oracle/_ref, stock PyTorch eager FP32) on this GPU {t_ref * 1e3:.3f} ms/step "
          f"({B / t_ref / 1e3:.1f} k samples/s)")
    del mr
except Exception as exc:
    print("reference on this GPU unavailable:", exc)
print(f"C3 train step, batch {B}, precision {prec}: eager loop {t_eager * 1e3:.3f} ms/step ({B / t_eager / 1e3:.1f} k samples/s), "
      f"whole-step CUDA graph {t_graph * 1e3:.3f} ms/step ({B / t_graph / 1e3:.1f} k samples/s)"
      + (f" = {t_ref / t_graph:.1f}x the unmodified reference" if t_ref else ""))

m.eval()
L.evaluate(m, ds, K, Kz, -3.0, 3.0, 2, dev)          # warm-up (allocator growth for the 512k-row batched decodes)
sync(); t0 = time.perf_counter()
res = L.evaluate(m, ds, K, Kz, -3.0, 3.0, 2, dev)
sync(); t_batched = time.perf_counter() - t0

def per_cell_loop():        # the reference's structure: one encode / reparameterize / estimator call (with its host sync) per cell
    out = []
    with torch.no_grad():
        for c in range(K * K):
            Xc = ds.X[ds.y == c].to(dev)
            if Xc.size(0) < 2:
                continue
            mu, lv = m.encode(Xc)
            z = reparameterize(mu, lv, nsamples=10).reshape(-1, 2)
            out.append(estimate_local_lipschitz(m.decode, z, num_pairs=2000))
        cx = np.linspace(-3, 3, Kz)
        for yi in range(Kz):
            for xi in range(Kz):
                z = torch.tensor([cx[xi], cx[yi]], dtype=torch.float32, device=dev).repeat(100, 1) + torch.randn(100, 2, device=dev) * 0.1
                mu, lv = m.encode(m.decode(z))
                out.append(estimate_local_lipschitz(m.decode, z, num_pairs=2000))
        mu, lv = m.encode(ds.X.to(dev))
        idx = torch.randperm(ds.X.size(0))[:5000].to(dev)
        out.append(estimate_local_lipschitz(m.decode, reparameterize(mu[idx], lv[idx], 1).squeeze(1), num_pairs=5000))
    return out
per_cell_loop(); sync(); t0 = time.perf_counter()
n_calls = len(per_cell_loop())
sync(); t_loop = time.perf_counter() - t0
print(f"C3 evaluation (K={K} X-cells, K_z={Kz} Z-cells, data-based): batched drivers {t_batched * 1e3:.1f} ms vs {n_calls} sequential "
      f"estimator calls {t_loop * 1e3:.1f} ms -> {t_loop / t_batched:.1f}x; data-based L = {res['data_lips']:.4g}, KL = {res['data_kl']:.4g}")
