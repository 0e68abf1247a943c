"""Independent checks of the CPU oracle (oracle/icnn_oracle.py) that need neither the reference nor a GPU: its analytic
Brenier map and double-backward against central finite differences of its OWN potential in fp64,
and the properties the construction promises -- psi convex in z, the map kappa-strongly monotone (the inverse-Lipschitz
bound of model.py:692: <xhat(a) - xhat(b), a - b> >= 2 kappa |a - b|^2).  Together with tests/test_oracle_golden.py (the
oracle against outputs of the unmodified reference) this pins the checker the GPU parity tests rely on."""
import numpy as np
import pytest

from oracle import icnn_oracle as io

CASES = [(1, 16, "mixed", io.MODE_EXP), (2, 32, "mixed", io.MODE_EXP), (3, 24, "clampy", io.MODE_CLAMP), (6, 20, "mixed", io.MODE_EXP)]
IDS = [f"d{c[0]}_h{c[1]}_{c[2]}" for c in CASES]


def _setup(case, B=40, seed=0):
    d, H, regime, mode = case
    rng = np.random.default_rng(seed + 17 * d + H)
    p = io.random_params(rng, d, H, np.float64, regime)
    z = rng.normal(0, 1, (B, d))
    return d, H, mode, p, z, rng


def _away_from_kinks(rows_ok, frac=0.9):
    # a finite-difference stencil that straddles a LeakyReLU kink is legitimately off; the maps are piecewise smooth, so
    # all but a few random points must agree
    assert rows_ok.mean() >= frac, f"only {rows_ok.mean():.2f} of the points agree with finite differences"


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_brenier_map_is_the_gradient_of_the_potential(case):
    d, H, mode, p, z, _ = _setup(case)
    kappa, h = 0.15, 1e-6
    psi, xhat, _ = io.icnn_brenier(z, p, mode, kappa)
    fd = np.empty_like(z)
    for j in range(d):
        e = np.zeros(d); e[j] = h
        pp, _, _ = io.icnn_brenier(z + e, p, mode, 0.0)
        pm, _, _ = io.icnn_brenier(z - e, p, mode, 0.0)
        fd[:, j] = (pp - pm) / (2 * h) + 2 * kappa * z[:, j]
    ok = np.abs(fd - xhat).max(1) <= 1e-5 * (1 + np.abs(xhat).max(1))
    _away_from_kinks(ok)


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_double_backward_matches_finite_differences(case):
    d, H, mode, p, z, rng = _setup(case, B=24, seed=3)
    kappa, h = 0.1, 1e-6
    v = rng.normal(0, 1, z.shape)
    gpsi = rng.normal(0, 1, z.shape[0])

    def loss(zz, pp):
        psi, xhat, _ = io.icnn_brenier(zz, pp, mode, kappa)
        return float((v * xhat).sum() + (gpsi * psi).sum())

    dz, g = io.icnn_brenier_backward(z, v, p, mode, kappa, gpsi=gpsi)
    # d/dz, row by row (L is a sum over independent rows)
    fd = np.empty_like(z)
    for j in range(d):
        e = np.zeros_like(z); e[:, j] = h
        for b in range(z.shape[0]):
            eb = np.zeros_like(z); eb[b] = e[b]
            fd[b, j] = (loss(z + eb, p) - loss(z - eb, p)) / (2 * h)
    ok = np.abs(fd - dz).max(1) <= 2e-5 * (1 + np.abs(dz).max(1))
    _away_from_kinks(ok, 0.8)
    # d/dparam on a few random entries of every tensor
    n_ok = n_all = 0
    for k in io.PARAM_KEYS:
        flat = p[k].reshape(-1)
        for idx in rng.choice(flat.size, size=min(6, flat.size), replace=False):
            pp, pm = {q: a.copy() for q, a in p.items()}, {q: a.copy() for q, a in p.items()}
            pp[k].reshape(-1)[idx] += h; pm[k].reshape(-1)[idx] -= h
            num = (loss(z, pp) - loss(z, pm)) / (2 * h)
            ana = g[k].reshape(-1)[idx]
            n_all += 1
            n_ok += abs(num - ana) <= 2e-5 * (1 + abs(ana)) + 1e-7
    assert n_ok >= 0.85 * n_all, f"{n_ok}/{n_all} parameter entries agree with finite differences"


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_potential_is_convex_and_map_strongly_monotone(case):
    d, H, mode, p, z, rng = _setup(case, B=200, seed=5)
    kappa = 0.2
    a, b = z[:100], z[100:]
    t = rng.uniform(0.05, 0.95, (100, 1))
    psi_a, xa, _ = io.icnn_brenier(a, p, mode, kappa)
    psi_b, xb, _ = io.icnn_brenier(b, p, mode, kappa)
    psi_m, _, _ = io.icnn_brenier(t * a + (1 - t) * b, p, mode, kappa)
    tt = t[:, 0]
    assert (psi_m <= tt * psi_a + (1 - tt) * psi_b + 1e-9 * (1 + np.abs(psi_a) + np.abs(psi_b))).all()
    lhs = ((xa - xb) * (a - b)).sum(1)
    rhs = 2 * kappa * ((a - b) ** 2).sum(1)
    assert (lhs >= rhs * (1 - 1e-9) - 1e-12).all()


def test_backward_is_linear_in_v_and_additive_over_the_batch():
    d, H, mode, p, z, rng = _setup((2, 32, "mixed", io.MODE_EXP), B=64, seed=9)
    v1, v2 = rng.normal(0, 1, z.shape), rng.normal(0, 1, z.shape)
    dz1, g1 = io.icnn_brenier_backward(z, v1, p, mode, 0.1)
    dz2, g2 = io.icnn_brenier_backward(z, v2, p, mode, 0.1)
    dz3, g3 = io.icnn_brenier_backward(z, v1 + 2 * v2, p, mode, 0.1)
    np.testing.assert_allclose(dz1 + 2 * dz2, dz3, rtol=1e-10, atol=1e-12)
    _, ga = io.icnn_brenier_backward(z[:20], v1[:20], p, mode, 0.1)
    _, gb = io.icnn_brenier_backward(z[20:], v1[20:], p, mode, 0.1)
    for k in io.PARAM_KEYS:
        np.testing.assert_allclose(g1[k] + 2 * g2[k], g3[k], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(ga[k] + gb[k], g1[k], rtol=1e-9, atol=1e-11)
    # the <v, xhat> path never touches the biases of A.0 / A.1 (sigma'' = 0): exact zeros, as the reference's autograd gives
    assert not g1["A1b"].any() and not g1["A2b"].any()


# ---------------------------------------------------------------------------------------------- loss oracle
def _fd(f, x, h=1e-6):
    g = np.empty_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += h; xm[i] -= h
        g[i] = (f(xp) - f(xm)) / (2 * h)
    return g


def test_loss_gradients_match_finite_differences():
    from oracle import loss_oracle as lo
    rng = np.random.default_rng(2)
    B, D, Dx = 7, 3, 5
    mu, lv = rng.normal(0, 1, (B, D)), rng.normal(0, 0.5, (B, D))
    x, xh = rng.normal(0, 1, (B, Dx)), rng.normal(0, 1, (B, Dx))
    dmu, dlv = lo.kl_grad(mu, lv, 0.7)
    np.testing.assert_allclose(dmu, _fd(lambda m: 0.7 * lo.kl(m, lv), mu), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(dlv, _fd(lambda l: 0.7 * lo.kl(mu, l), lv), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(lo.recon_mse_grad(x, xh, 1.3), _fd(lambda y: 1.3 * lo.recon_mse(x, y), xh), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(lo.recon_logmse_grad(x, xh, 1.3), _fd(lambda y: 1.3 * lo.recon_logmse(x, y), xh), rtol=1e-5, atol=1e-8)
    zi, zr = rng.normal(0, 1, (4, B, D)), rng.normal(0, 1, (4, B, D))          # [L,B,D]: mean over L (Appendix B.2)
    np.testing.assert_allclose(lo.latent_recon_grad(zi, zr, 0.4), _fd(lambda y: 0.4 * lo.latent_recon(zi, y), zr), rtol=1e-6, atol=1e-8)
    # KL of the standard normal posterior is zero, and it is the batch mean of the per-sample KL
    assert abs(lo.kl(np.zeros((B, D)), np.zeros((B, D)))) == 0.0
    np.testing.assert_allclose(lo.kl(mu, lv), lo.kl_per_sample(mu, lv).mean(), rtol=1e-12)


def test_lipschitz_estimator_properties():
    from oracle import loss_oracle as lo
    rng = np.random.default_rng(4)
    X = rng.normal(0, 1, (300, 2))
    i1, i2 = rng.integers(0, 300, 2000), rng.integers(0, 300, 2000)
    # an isometry has ratio 1 on every pair with |dx| >= eps (coincident pairs clamp to eps/eps = 1 as well)
    c, s = np.cos(0.7), np.sin(0.7)
    r = lo.lipschitz_ratios(X, X @ np.array([[c, -s], [s, c]]), i1, i2)
    np.testing.assert_allclose(r, 1.0, rtol=1e-12)
    # a uniform scaling by a has ratio a where the clamps are inactive, and the estimate is (1/a, a, max)
    a = 3.0
    keep = np.sqrt(((X[i1] - X[i2]) ** 2).sum(1)) > 1e-2
    r = lo.lipschitz_ratios(X, a * X, i1[keep], i2[keep])
    np.testing.assert_allclose(r, a, rtol=1e-12)
    inv, bi, both = lo.lipschitz_from_ratios(r)
    np.testing.assert_allclose([inv, bi, both], [1 / a, a, a], rtol=1e-12)
    # torch.quantile's linear interpolation: exact on a ramp
    assert lo.torch_quantile_linear(np.arange(101.0), 0.05) == 5.0 and lo.torch_quantile_linear(np.arange(101.0), 0.955) == 95.5
    ap = lo.lipschitz_allpairs(X, a * X)
    assert ap["count"] == 300 * 299 // 2
    np.testing.assert_allclose([ap["max"], ap["min"], ap["sum"] / ap["count"]], [a, a, a], rtol=1e-9)
