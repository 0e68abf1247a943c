"""End-of-run latent metrics (utils.py:49-164 of the reference) -- calc_mi through the tiled all-pairs log-density kernel
(csrc/metrics.cu) against the reference's own [B,B,nz] formulation evaluated in fp64 on the same samples."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_calc_mi(mu, logvar, z):
    """utils.py:87-107 with the sampled z given (fp64)."""
    mu, logvar, z = mu.double(), logvar.double(), z.double()
    B, nz = mu.shape
    neg_entropy = (-0.5 * nz * math.log(2 * math.pi) - 0.5 * (1 + logvar).sum(-1)).mean()
    dev = z.unsqueeze(1) - mu.unsqueeze(0)
    log_density = -0.5 * ((dev ** 2) / logvar.exp().unsqueeze(0)).sum(-1) - 0.5 * (nz * math.log(2 * math.pi) + logvar.sum(-1)).unsqueeze(0)
    log_qz = torch.logsumexp(log_density, dim=1) - math.log(B)
    return (neg_entropy - log_qz.mean()).item()


@pytest.mark.parametrize("B,nz", [(1, 2), (37, 2), (256, 32), (1000, 5), (300, 128)])
def test_calc_mi_matches_reference_formula(B, nz):
    from vae_song_b200 import utils
    g = torch.Generator(device="cpu").manual_seed(B + nz)
    mu = (torch.randn(B, nz, generator=g) * 1.5).cuda()
    lv = (torch.randn(B, nz, generator=g) * 0.7 - 0.5).cuda()
    eps = torch.randn(B, nz, generator=g).cuda()
    z = mu + eps * (0.5 * lv).exp()
    got = utils.calc_mi(mu, lv, eps=eps)
    want = ref_calc_mi(mu, lv, z)
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (got, want)


def ref_nll_iw(mu, log_var, z, loss_rec):
    """utils.py:109-120 with the sampled z [B,ns,nz] given (fp64)."""
    mu, log_var, z = mu.double(), log_var.double(), z.double()
    nz, ns = z.size(2), z.size(1)
    log_comp = torch.distributions.normal.Normal(torch.zeros(nz, dtype=torch.float64, device=z.device),
                                                 torch.ones(nz, dtype=torch.float64, device=z.device)).log_prob(z).sum(-1) - loss_rec
    m, lv = mu.unsqueeze(1), log_var.unsqueeze(1)
    log_inf = -0.5 * (((z - m) ** 2) / lv.exp()).sum(-1) - 0.5 * (nz * math.log(2 * math.pi) + lv.sum(-1))
    return -(torch.logsumexp((log_comp - log_inf).reshape(-1), 0) - math.log(ns)).item()


@pytest.mark.parametrize("B,ns,nz", [(1, 1, 2), (37, 100, 2), (256, 100, 28), (1000, 7, 5)])
def test_nll_iw_matches_reference_formula(B, ns, nz):
    from vae_song_b200 import utils
    g = torch.Generator(device="cpu").manual_seed(B + ns + nz)
    mu = (torch.randn(B, nz, generator=g) * 1.5).cuda()
    lv = (torch.randn(B, nz, generator=g) * 0.7 - 0.5).cuda()
    eps = torch.randn(B, ns, nz, generator=g).cuda()
    z = mu.unsqueeze(1) + eps * (0.5 * lv).exp().unsqueeze(1)
    got = utils.nll_iw(mu, lv, 3.25, nsamples=ns, eps=eps)
    want = ref_nll_iw(mu, lv, z, 3.25)
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (got, want)
    # the random stream is utils.reparameterize's: same seed, same estimate as the injected draw
    torch.manual_seed(5)
    e2 = torch.randn_like(mu.unsqueeze(1).expand(B, ns, nz))
    torch.manual_seed(5)
    assert utils.nll_iw(mu, lv, torch.tensor(3.25, device="cuda"), nsamples=ns) == utils.nll_iw(mu, lv, 3.25, nsamples=ns, eps=e2)


def test_measure_pc_runmodel_runs():
    from vae_song_b200 import model, utils
    torch.manual_seed(0)
    m = model.VanillaVAE(beta=0.5, dataset="pinwheel", hidden_channels=[16, 16], encoder_type="mlp", decoder_type="mlp").cuda().eval()
    x = torch.randn(128, 2)
    loader = [(x, torch.zeros(128))]
    au, kl, mi, nll, var = utils.measure_pc_runmodel(m, loader, "cuda")
    assert all(np.isfinite(v) for v in (au, kl, mi, nll, var)) and 0.0 <= au <= 1.0 and kl >= 0.0
    mu, lv = m.encode(x.cuda())
    assert abs(utils.calc_au_per_batch(mu) - float((mu.var(0, unbiased=False) >= 0.01).float().mean())) < 1e-6
