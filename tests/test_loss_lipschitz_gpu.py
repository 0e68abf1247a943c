"""Loss / reparam / Lipschitz / Adam kernels against the reference-generated goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo

from conftest import GOLDEN
from helpers import close_report

pytestmark = pytest.mark.gpu
T = lambda a, rg=False: torch.tensor(np.asarray(a, dtype=np.float32), device="cuda", requires_grad=rg)


@pytest.mark.parametrize("name", ["l1", "l4", "l4_wide"])
def test_fused_loss_vs_reference(name):
    from vae_song_b200 import ops
    G = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    pre = f"lrvae_{name}/"
    alpha, beta, wu = G[pre + "hyper"]
    xh, mu, lv, zr = T(G[pre + "xh"], True), T(G[pre + "mu"], True), T(G[pre + "lv"], True), T(G[pre + "zrec"], True)
    rec, reg, lr = ops.VaeLossFn.apply(T(G[pre + "x"]), xh, mu, lv, T(G[pre + "zin"]), zr, False)
    total = rec + reg * beta + lr * alpha * wu
    total.backward()
    np.testing.assert_allclose([float(total), float(rec), float(reg * beta), float(lr * alpha * wu)], G[pre + "out"], rtol=1e-5)
    for t, k in ((xh, "g_xh"), (mu, "g_mu"), (lv, "g_lv"), (zr, "g_zrec")):
        close_report(t.grad.cpu().numpy(), G[pre + k], 1e-5, k)


def test_logmse_vs_reference():
    from vae_song_b200 import ops
    G = np.load(os.path.join(GOLDEN, "loss_cases.npz"))
    pre = "lid_logmse/"
    xh, mu, lv = T(G[pre + "xh"], True), T(G[pre + "mu"], True), T(G[pre + "lv"], True)
    rec, reg, _ = ops.VaeLossFn.apply(T(G[pre + "x"]), xh, mu, lv, None, None, True)
    (rec + 0.4 * reg).backward()
    np.testing.assert_allclose([float(rec + 0.4 * reg), float(rec), float(reg)], G[pre + "out"], rtol=1e-5)
    close_report(xh.grad.cpu().numpy(), G[pre + "g_xh"], 1e-5, "g_xh")
    close_report(mu.grad.cpu().numpy(), G[pre + "g_mu"], 1e-5, "g_mu")
    close_report(lv.grad.cpu().numpy(), G[pre + "g_lv"], 1e-5, "g_lv")


@pytest.mark.parametrize("shape", [(1, 5, 3), (4, 33, 2), (3, 1000, 28)])
def test_reparam_fwd_bwd(shape):
    from vae_song_b200 import ops
    L, B, D = shape
    rng = np.random.default_rng(3)
    mu, lv, eps, gz = rng.normal(0, 1, (B, D)), rng.normal(-1, 1, (B, D)), rng.normal(0, 1, shape), rng.normal(0, 1, shape)
    mut, lvt = T(mu, True), T(lv, True)
    z = ops.ReparamFn.apply(mut, lvt, T(eps))
    (z * T(gz)).sum().backward()
    f = lambda a: np.asarray(a, np.float32).astype(np.float64)
    close_report(z.detach().cpu().numpy(), lo.reparam(f(mu), f(lv), f(eps)), 1e-6, "z")
    sd = np.exp(0.5 * f(lv))
    close_report(mut.grad.cpu().numpy(), f(gz).sum(0), 1e-5, "dmu")
    close_report(lvt.grad.cpu().numpy(), (f(gz) * f(eps) * 0.5 * sd).sum(0), 1e-5, "dlv")
    z2 = ops.ReparamFn.apply(mut.detach(), lvt.detach(), T(eps[0]))          # [B,D] eps (LIDVAE path)
    close_report(z2.cpu().numpy(), lo.reparam(f(mu), f(lv), f(eps[0])), 1e-6, "z 2d")


def test_loss_large_and_deterministic():
    """BASELINE-size reduction (65536 x 2 and 4096 x 784): matches fp64 oracle, bit-identical rerun."""
    from vae_song_b200 import ops
    rng = np.random.default_rng(4)
    for B, Dx, D in ((65536, 2, 2), (4096, 784, 32)):
        x, xh = rng.normal(0, 1, (B, Dx)).astype(np.float32), rng.normal(0, 1, (B, Dx)).astype(np.float32)
        mu, lv = rng.normal(0, 1, (B, D)).astype(np.float32), rng.normal(0, 1, (B, D)).astype(np.float32)
        a = ops.VaeLossFn.apply(T(x), T(xh), T(mu), T(lv), None, None, False)
        b = ops.VaeLossFn.apply(T(x), T(xh), T(mu), T(lv), None, None, False)
        assert float(a[0]) == float(b[0]) and float(a[1]) == float(b[1])
        np.testing.assert_allclose(float(a[0]), lo.recon_mse(x.astype(np.float64), xh.astype(np.float64)), rtol=2e-6)
        np.testing.assert_allclose(float(a[1]), lo.kl(mu.astype(np.float64), lv.astype(np.float64)), rtol=2e-6)


def test_lipschitz_pairs_vs_reference():
    from vae_song_b200 import ops
    G = np.load(os.path.join(GOLDEN, "lipschitz_cases.npz"))
    for name in ("n500_p2000", "n100_p100", "n5_p64_dups"):
        Z, Y, i1, i2, res = (G[f"{name}/{k}"] for k in ("Z", "Y", "i1", "i2", "res"))
        r = ops.lipschitz_pair_ratios(T(Z), T(Y), torch.tensor(i1, device="cuda"), torch.tensor(i2, device="cuda"))
        ref = lo.lipschitz_ratios(Z.astype(np.float32).astype(np.float64), Y.astype(np.float32).astype(np.float64), i1, i2)
        close_report(r.cpu().numpy(), ref, 1e-5, "ratios")
        got = lo.lipschitz_from_ratios(r.cpu().numpy().astype(np.float64))
        np.testing.assert_allclose(got, res, rtol=1e-4)
        if name.endswith("dups"):
            assert (r.cpu().numpy()[i1 == i2] == 1.0).all()     # both clamps -> exactly 1.0


def test_estimate_local_lipschitz_api():
    """Same call as lipschitz.py:184 on our LIDVAE; on CUDA the seed-0 generator is Philox (Appendix B.7),
    so compare against the oracle evaluated on the very pairs the call drew."""
    from vae_song_b200 import model, utils
    torch.manual_seed(0)
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[64, 128], hidden_channels=[8]).cuda().eval()
    Z = torch.randn(300, 2, device="cuda")
    res = utils.estimate_local_lipschitz(m.decode, Z, num_pairs=500, use_grad=True)
    gen = torch.Generator(device="cuda").manual_seed(0)
    i1 = torch.randint(0, 300, (500,), device="cuda", generator=gen)
    i2 = torch.randint(0, 300, (500,), device="cuda", generator=gen)
    with torch.no_grad():
        Y = m.decode(Z)
    r = lo.lipschitz_ratios(Z.double().cpu().numpy(), Y.double().cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy())
    np.testing.assert_allclose(res, lo.lipschitz_from_ratios(r), rtol=1e-3)
    assert utils.estimate_local_lipschitz(m.decode, Z[:1]) == (0.0, 0.0, 0.0)


@pytest.mark.parametrize("N,dx,dy", [(2, 2, 2), (63, 2, 2), (200, 2, 2), (333, 3, 17), (5000, 2, 2)])
def test_allpairs_vs_bruteforce(N, dx, dy):
    from vae_song_b200 import ops
    rng = np.random.default_rng(N)
    X, Y = rng.normal(0, 1, (N, dx)).astype(np.float32), rng.normal(0, 2, (N, dy)).astype(np.float32)
    if N > 10:
        X[5] = X[3]; Y[5] = Y[3]                       # coincident pair -> ratio 1.0 via both clamps
    stats, hist = ops.lipschitz_allpairs(T(X), T(Y), 1e-3, nbins=64, hist_lo=-16.0, hist_hi=16.0)
    mx, mn, sm, cnt = stats.tolist()
    if N <= 400:
        ref = lo.lipschitz_allpairs(X, Y)
        assert cnt == ref["count"] == N * (N - 1) // 2
        np.testing.assert_allclose([mx, mn, sm], [ref["max"], ref["min"], ref["sum"]], rtol=2e-5)
        assert int(hist.sum()) == ref["count"]
    else:   # 12.5 M pairs: torch brute force on the GPU in fp64
        Xd, Yd = T(X).double(), T(Y).double()
        r = torch.cdist(Yd, Yd).clamp(min=1e-3) / torch.cdist(Xd, Xd).clamp(min=1e-3)
        iu = torch.triu_indices(N, N, 1, device="cuda")
        rr = r[iu[0], iu[1]]
        assert cnt == rr.numel()
        np.testing.assert_allclose([mx, mn, sm], [float(rr.max()), float(rr.min()), float(rr.sum())], rtol=1e-4)
    # sharded tiles combine to the same answer (the multi-GPU decomposition, run on one device)
    nt = ops.lipschitz_num_tiles(N)
    parts = [ops.lipschitz_allpairs(T(X), T(Y), 1e-3, lo_, hi_)[0].tolist()
             for lo_, hi_ in ((0, nt // 3), (nt // 3, nt // 2), (nt // 2, nt))]
    assert max(p[0] for p in parts) == mx and min(p[1] for p in parts if p[3] > 0) == mn
    assert sum(p[3] for p in parts) == cnt


def test_fused_adam_matches_torch():
    from vae_song_b200 import ops
    torch.manual_seed(0)
    p = torch.randn(100003, device="cuda")
    q = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([q], lr=1e-3, weight_decay=0.01)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 6):
        g = torch.randn_like(p)
        q.grad = g.clone()
        opt.step()
        ops.adam_step_(p, g * 4.0, m, v, t, lr=1e-3, weight_decay=0.01, grad_scale=0.25)
    close_report(p.cpu().numpy(), q.detach().cpu().numpy(), 1e-6, "adam")


def test_batched_cell_estimator_equals_per_cell_calls():
    """lipschitz.py:48-154 evaluates up to 512 cells one by one; the batched helper must reproduce every cell."""
    from vae_song_b200 import model, utils
    torch.manual_seed(1)
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[64, 128], hidden_channels=[8]).cuda().eval()
    cells = [torch.randn(n, 2, device="cuda") * 0.3 + i for i, n in enumerate([40, 2, 1, 0, 333, 100])]
    got = utils.estimate_local_lipschitz_batched(m.decode, cells, num_pairs=500)
    for i, X in enumerate(cells):
        ref = utils.estimate_local_lipschitz(m.decode, X, num_pairs=500) if X.size(0) >= 2 else (0.0, 0.0, 0.0)
        np.testing.assert_allclose(got[i], ref, rtol=1e-5)


def test_utils_reparameterize_and_kld_match_reference_formulas():
    """utils.reparameterize draws eps exactly like utils.py:40-47 (randn_like on the expanded [B,ns,nz] std), so the
    same seed gives the reference's samples; utils.kld is model.py:884 / utils.py:140-141."""
    from vae_song_b200 import utils
    g = torch.Generator(device="cuda").manual_seed(9)
    mu = torch.randn(50, 3, device="cuda", generator=g)
    lv = torch.randn(50, 3, device="cuda", generator=g) * 0.5
    torch.manual_seed(123)
    z = utils.reparameterize(mu, lv, nsamples=7)
    torch.manual_seed(123)
    std = lv.mul(0.5).exp().unsqueeze(1).expand(50, 7, 3)
    ref = mu.unsqueeze(1).expand(50, 7, 3) + torch.randn_like(std) * std
    assert z.shape == (50, 7, 3)
    close_report(z.cpu().numpy(), ref.cpu().numpy(), 1e-6, "reparameterize")
    ref_kl = (-0.5 * (1 + lv - mu ** 2 - lv.exp())).mean(dim=0).sum().item()
    assert abs(utils.kld(mu, lv) - ref_kl) <= 1e-5 * abs(ref_kl)


def test_train_model_loop_runs_and_learns():
    """train.train_model == lipschitz.py:23-44 (same signature): a few epochs on a toy mixture reduce the loss."""
    from vae_song_b200 import model, train
    torch.manual_seed(0)
    X = torch.cat([torch.randn(512, 2) * 0.3 + 2.0, torch.randn(512, 2) * 0.3 - 2.0])
    ds = torch.utils.data.TensorDataset(X, torch.zeros(1024, dtype=torch.long))
    loader = torch.utils.data.DataLoader(ds, batch_size=256, shuffle=True, drop_last=True,
                                         generator=torch.Generator().manual_seed(0))
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[64, 64], hidden_channels=[16, 8], inverse_lipschitz=0.1, beta=0.01)
    rng = np.random.default_rng(0)
    with torch.no_grad():
        for ic in (m.decoder[0], m.decoder[1]):
            ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / 64), 0.5, (64, 64)), dtype=torch.float32))
            ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / 64), 0.5, (1, 64)), dtype=torch.float32))

    def eval_loss():
        m.eval()
        with torch.no_grad():
            x = X.cuda()
            recon, mu, lv, z, _ = m(x, latent_rand_sampling=False)
            return float(m.loss(x, recon, mu, lv, z, None)[0])
    m.cuda()
    before = eval_loss()
    train.train_model(m, loader, epochs=15, lr=2e-3, device="cuda", grad_clip={"enabled": True, "clip_type": "norm", "max_norm": 5.0})
    after = eval_loss()
    assert np.isfinite(after) and after < 0.7 * before, (before, after)
