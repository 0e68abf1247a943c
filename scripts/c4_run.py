"""BASELINE configs[3] (configs/config_mnist.yaml: MNIST-shaped 1x28x28 data, batch 256): train-step time of the three
models SURVEY.md 8(d) lists under C4 --
  * conv LR-VAE (LRVAE, L = 4 Monte-Carlo samples, staged backward of main.py:262-284) and conv Beta-VAE (VanillaVAE):
    stock conv stacks (out of scope) around the fused reparam / KL / reconstruction / latent-recon loss kernel;
  * LIDVAE(dataset='mnist'): conv encoder + the ICNN(32,512) -> eye(784,32) -> ICNN(784,1024) Brenier-map decoder through
    the wide-input tcgen05 kernels, against the reference's own formulation of the same step (torch ops +
    autograd.grad(create_graph=True) + autograd double-backward + torch.optim.Adam, FP32) run by stock PyTorch here.
Each step is timed launched eagerly from Python and replayed as ONE CUDA graph (train.DataParallelTrainer.capture)."""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import eager_icnn_potential
import numpy as np, torch, torch.nn.functional as F
from vae_song_b200 import main as M, model, ops, train, utils as vutils

dev, B = "cuda", 256
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
tr_set, _ = M.synthetic_dataset("mnist", 16 * B, B)
X = tr_set.tensors[0].to(dev)
batch = lambda i: X[(i % 16) * B:(i % 16 + 1) * B]


def timed(fn, n, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        out = fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


# ------------------------------------------------------------------ conv LR-VAE / Beta-VAE through the main.py driver's step
cfg = M.load_config(os.path.join(ROOT, "configs", "config_mnist.yaml"))
for exp in ("lrvae", "vae"):
    c = copy.deepcopy(cfg); c["experiment_type"] = exp
    torch.manual_seed(0)
    tag, m, kw = next(M.iter_models(c))
    m = m.to(dev).train()
    if hasattr(m, "wu_alpha"):
        m.wu_alpha = 1.0
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    ms_e, out = timed(lambda i: M.train_step(m, batch(i), opt, None, kw["num_mc_samples"], kw["grad_clip"]), 50)
    torch.manual_seed(0)
    tag, m, kw = next(M.iter_models(c))
    m = m.to(dev).train()
    if hasattr(m, "wu_alpha"):
        m.wu_alpha = 1.0
    t = train.DataParallelTrainer(m, lr=1e-2, staged_backward=True, grad_clip=kw["grad_clip"],
                                  forward_kwargs={"L": kw["num_mc_samples"]})
    t.capture(batch(0))
    ms_g, outg = timed(lambda i: t.step_graphed(batch(i)), 200)
    npar = sum(p.numel() for p in m.parameters())
    print(f"C4 {tag} ({type(m).__name__}, {npar} params, L={kw['num_mc_samples']}) batch {B}: eager loop {ms_e:.3f} ms/step "
          f"({B / ms_e:.1f} k samples/s), whole step as one CUDA graph {ms_g:.3f} ms/step ({B / ms_g:.1f} k samples/s), "
          f"loss {float(out[0]):.4f} / {float(outg[0]):.4f}")

# ------------------------------------------------------------------ LIDVAE with the MNIST-shaped ICNN decoder
rng = np.random.default_rng(1)
for prec in ("tf32x3", "tf32", "fp32"):
    torch.manual_seed(0)
    m = model.LIDVAE(dataset="mnist", inverse_lipschitz=0.2, beta=0.001, precision=prec).to(dev).train()
    for ic in (m.decoder[0], m.decoder[1]):
        vutils.trained_like_icnn_(ic, rng)
    ref = copy.deepcopy(m) if prec == "tf32x3" else None
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)

    def step(i):
        x = batch(i)
        opt.zero_grad()
        recon, mu, lv, z_in, z_rec = m(x)
        total = m.loss(x, recon, mu, lv, z_in, z_rec)[0]
        total.backward()
        opt.step()
        return total.detach()          # (a live autograd graph would pin default-stream AccumulateGrad nodes into the capture)
    ms_e, out = timed(step, 30)
    t = train.DataParallelTrainer(m, lr=1e-3)
    t.capture(batch(0))
    ms_g, outg = timed(lambda i: t.step_graphed(batch(i)), 100)
    # dense algorithmic flops of the decoder's train step per sample (F_train(d,H) = 8H^2 + 22dH)
    flop = sum(8 * H * H + 22 * d * H for d, H in ((32, 512), (784, 1024)))
    print(f"C4 LIDVAE(mnist) precision {prec} batch {B}: eager loop {ms_e:.3f} ms/step ({B / ms_e:.1f} k samples/s), one CUDA "
          f"graph {ms_g:.3f} ms/step ({B / ms_g:.1f} k samples/s; decoder {B * flop / ms_g / 1e9:.1f} TFLOP/s dense-algorithmic), "
          f"loss {float(out):.4f} / {float(outg[0]):.4f}")
    if ref is not None:
        ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
        kappa = ref.il_factor

        def brenier(ic, zz):
            psi = eager_icnn_potential(zz, ic._mode(), *ic._flat_params()) + kappa * zz.pow(2).sum(1, keepdim=True)
            return torch.autograd.grad(psi, [zz], torch.ones_like(psi), create_graph=True)[0]

        def rstep(i):
            x = batch(i)
            ropt.zero_grad()
            ret = ref.encoder(x)
            mu, var = ret.split(ret.shape[1] // 2, 1)
            lv = F.softplus(var)
            z = mu + torch.randn_like(mu) * torch.exp(0.5 * lv)
            y = brenier(ref.decoder[1], F.linear(brenier(ref.decoder[0], z), ref.B)).view_as(x)
            rec = ((x - y) ** 2).mean(0).sum()
            kl = (-0.5 * (1 + lv - mu ** 2 - lv.exp())).mean(0).sum()
            total = rec + ref.beta * kl
            total.backward()
            ropt.step()
            return total.detach()
        ms_r, outr = timed(rstep, 30)
        print(f"C4 LIDVAE(mnist) reference formulation, stock PyTorch eager FP32 on this GPU, batch {B}: {ms_r:.3f} ms/step "
              f"({B / ms_r:.1f} k samples/s), loss {float(outr):.4f}")
