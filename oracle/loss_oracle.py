"""CPU oracle for reparameterisation, Gaussian KL, reconstruction and latent-reconstruction
losses, and for the random-pair / all-pairs Lipschitz estimator.  TEST INFRASTRUCTURE ONLY
(see oracle/icnn_oracle.py header for who may import this and how parity is pinned).

Reference lines restated:
  * reparam           model.py:843 ; model.py:423-424 ([L,B,D]) ; utils.py:40-47
  * KL                model.py:550/606/884 ; utils.py:140-141 ; per-sample lipschitz.py:62
  * recon (MSE)       model.py:542/589/870     ((x-xhat)^2).mean(0).sum()
  * recon (log-MSE)   model.py:872-882 (LIDVAE: reshape(B,-1).mean(1)) ; :544-548 (.mean(1)x3)
  * latent recon      model.py:551/603         ((z_in-z_rec)^2).mean(dim=0).sum(), dim 0 is L
  * estimator         utils.py:532-567
"""
from __future__ import annotations

import numpy as np


def reparam(mu, lv, eps):
    """z = mu + eps*exp(0.5*lv); eps [B,D] or [L,B,D] (broadcast over L)."""
    return mu + eps * np.exp(lv * lv.dtype.type(0.5))


def kl(mu, lv):
    """(-0.5*(1+lv-mu^2-exp(lv))).mean(0).sum()"""
    t = -0.5 * (1.0 + lv - mu * mu - np.exp(lv))
    return t.mean(axis=0).sum()


def kl_per_sample(mu, lv):
    return -0.5 * (1.0 + lv - mu * mu - np.exp(lv)).sum(axis=1)


def kl_grad(mu, lv, g=1.0):
    B = mu.shape[0]
    return g * mu / B, g * (-0.5 * (1.0 - np.exp(lv))) / B


def recon_mse(x, xh):
    B = x.shape[0]
    d = (x - xh).reshape(B, -1)
    return (d * d).mean(axis=0).sum()


def recon_mse_grad(x, xh, g=1.0):
    """d/dxhat"""
    return g * (-2.0 / x.shape[0]) * (x - xh)


def recon_logmse(x, xh):
    B = x.shape[0]
    d = (x - xh).reshape(B, -1)
    Dx = d.shape[1]
    mse = (d * d).mean(axis=1)
    return (0.5 * Dx * (np.log(2.0 * np.pi * mse + 1e-5) + 1.0)).mean()


def recon_logmse_grad(x, xh, g=1.0):
    B = x.shape[0]
    d = (x - xh).reshape(B, -1)
    Dx = d.shape[1]
    mse = (d * d).mean(axis=1)
    coef = 0.5 * Dx * (2.0 * np.pi) / (2.0 * np.pi * mse + 1e-5) / B      # dL/dmse_b
    return (g * coef[:, None] * (-2.0 / Dx) * d).reshape(x.shape)


def latent_recon(z_in, z_rec):
    """((z_in - z_rec)**2).mean(dim=0).sum()  -- dim 0 is L for [L,B,D] stacks (quirk B.2)."""
    d = z_in - z_rec
    return (d * d).mean(axis=0).sum()


def latent_recon_grad(z_in, z_rec, g=1.0):
    """d/dz_rec"""
    return g * (-2.0 / z_in.shape[0]) * (z_in - z_rec)


# ------------------------------------------------------------------------------ Lipschitz
def torch_quantile_linear(r, q):
    """torch.quantile(r, q) default 'linear' interpolation on a 1-D array."""
    s = np.sort(r)
    pos = q * (s.shape[0] - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, s.shape[0] - 1)
    w = pos - lo
    return s[lo] + (s[hi] - s[lo]) * s.dtype.type(w)


def lipschitz_ratios(X, Y, i1, i2, eps=1e-3):
    """utils.py:548-562: r = clamp(|y1-y2|,eps)/clamp(|x1-x2|,eps), rows flattened, p=2."""
    X2, Y2 = X.reshape(X.shape[0], -1), Y.reshape(Y.shape[0], -1)
    dy = np.maximum(np.sqrt(((Y2[i1] - Y2[i2]) ** 2).sum(1)), X.dtype.type(eps))
    dx = np.maximum(np.sqrt(((X2[i1] - X2[i2]) ** 2).sum(1)), X.dtype.type(eps))
    return dy / dx


def lipschitz_from_ratios(r, quantile=0.05, eps=1e-3):
    """utils.py:563-567 -> (1/A, B, max(1/A, B))."""
    A = max(torch_quantile_linear(r, quantile), r.dtype.type(eps))
    Bq = torch_quantile_linear(r, 1.0 - quantile)
    invA = r.dtype.type(1.0) / A
    return float(invA), float(Bq), float(max(invA, Bq))


def lipschitz_allpairs(X, Y, eps=1e-3):
    """Brute-force all unordered pairs i<j (north_star kernel 4; an extension, not in the reference).
    Returns dict(max, min, sum, count)."""
    X2 = X.reshape(X.shape[0], -1).astype(np.float64)
    Y2 = Y.reshape(Y.shape[0], -1).astype(np.float64)
    N = X2.shape[0]
    iu, ju = np.triu_indices(N, k=1)
    dy = np.maximum(np.sqrt(((Y2[iu] - Y2[ju]) ** 2).sum(1)), eps)
    dx = np.maximum(np.sqrt(((X2[iu] - X2[ju]) ** 2).sum(1)), eps)
    r = dy / dx
    return dict(max=float(r.max()), min=float(r.min()), sum=float(r.sum()), count=int(r.shape[0]), ratios=r)
