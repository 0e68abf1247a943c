"""BASELINE configs[0] (configs/config_pinwheel.yaml: pinwheel LR-VAE, 12 x 16 MLP encoder / decoder, batch 1024, staged
backward of main.py:255-292): train-step time with the fused MLP layer kernels (opt-in, FlexibleVAE.fused_mlp) vs the same
nn.Modules run by PyTorch (default).  The eager staged backward is host bound: the stock modules win here."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_song_b200 import main as M
cfg = M.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "config_pinwheel.yaml"))
tr, _ = M.synthetic_dataset("pinwheel", 16384, 1024)
X = tr.tensors[0].cuda()
for fused in (True, False):
    torch.manual_seed(0)
    tag, m, kw = next(M.iter_models(cfg))
    m = m.cuda().train(); m.fused_mlp = fused; m.wu_alpha = 1.0
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    def step(i):
        return M.train_step(m, X[(i % 16) * 1024:(i % 16 + 1) * 1024], opt, None, kw["num_mc_samples"], kw["grad_clip"])
    for i in range(10): step(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(50): out = step(i)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 50 * 1e3
    print(f"C1 {tag} batch 1024, {'fused MLP kernels' if fused else 'stock nn.Modules'}: {ms:.3f} ms/step ({1024 / ms:.0f} k samples/s), loss {float(out[0]):.4f}")

# the same staged step as ONE CUDA graph (DataParallelTrainer(staged_backward=True)): no host work per step
from vae_song_b200 import train
for fused in (True, False):
    torch.manual_seed(0)
    tag, m, kw = next(M.iter_models(cfg))
    m = m.cuda().train(); m.fused_mlp = fused; m.wu_alpha = 1.0
    tr = train.DataParallelTrainer(m, lr=1e-2, staged_backward=True, grad_clip=kw["grad_clip"])
    tr.capture(X[:1024])
    for i in range(10): tr.step_graphed(X[(i % 16) * 1024:(i % 16 + 1) * 1024])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(200): out = tr.step_graphed(X[(i % 16) * 1024:(i % 16 + 1) * 1024])
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 200 * 1e3
    print(f"C1 {tag} batch 1024, whole staged step as one CUDA graph, {'fused MLP kernels' if fused else 'stock nn.Modules'}: "
          f"{ms:.3f} ms/step ({1024 / ms:.0f} k samples/s), loss {float(out[0]):.4f}")
