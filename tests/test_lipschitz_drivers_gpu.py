"""The batched evaluation drivers (vae_song_b200/lipschitz.py) against a literal per-cell loop with the reference's
structure (lipschitz.py:48-222: one encode / reparameterize / estimator call per cell, same RNG call order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(kind="lidvae"):
    from vae_song_b200 import model
    torch.manual_seed(0)
    if kind == "lidvae":
        m = model.LIDVAE(inverse_lipschitz=0.2, beta=0.5, dataset="pinwheel", icnn_channels=[64, 128], hidden_channels=[16, 8, 2])
        rng = np.random.default_rng(3)
        with torch.no_grad():
            for ic in (m.decoder[0], m.decoder[1]):
                H = ic.hidden_channel
                ic.W[0].param.copy_(torch.tensor(rng.normal(np.log(1.0 / H), 1.0, (H, H)), dtype=torch.float32))
                ic.W[1].param.copy_(torch.tensor(rng.normal(np.log(2.0 / H), 1.0, (1, H)), dtype=torch.float32))
    else:
        m = model.LRVAE(alpha=0.1, dataset="pinwheel", hidden_channels=[16, 16, 2])
    m = m.cuda().train()
    with torch.no_grad():          # give BatchNorm non-trivial running statistics
        for _ in range(3):
            m.encode(torch.randn(64, 2, device="cuda") * 2 + 1)
    return m.eval()


def _kl(mu, lv):
    return -0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=1)


@pytest.mark.parametrize("kind", ["lidvae", "lrvae"])
def test_x_cells_match_per_cell_loop(kind):
    from vae_song_b200 import lipschitz as L
    from vae_song_b200.utils import estimate_local_lipschitz, reparameterize
    m = _model(kind)
    K = 4
    ds = L.GaussianMixture2D(6, 400, center_range=K, stds=0.3, pattern="corner_heavy", seed=1)
    ds.y[:3] = 9                                     # a 3-sample cell; cell 10 gets a single sample; cells 11.. stay empty
    ds.y[3] = 10
    torch.manual_seed(11)
    got = L._get_kl_and_lipschitz_for_x_cells(m, ds, K, "cuda", nsamples_z=5, num_pairs_lips=300)
    torch.manual_seed(11)
    want = [np.full(K * K, L.DEFAULT_EMPTY_CELL_FILL_VALUE, dtype=np.float32) for _ in range(4)]
    with torch.no_grad():
        for c in range(K * K):
            Xc = ds.X[ds.y == c].cuda()
            if Xc.size(0) == 0:
                continue
            mu, lv = m.encode(Xc)
            want[0][c] = _kl(mu, lv).mean().item()
            if Xc.size(0) < 2:
                continue
            z = reparameterize(mu, lv, nsamples=5).reshape(-1, mu.size(-1))
            inv_l, l, bi = estimate_local_lipschitz(m.decode, z, num_pairs=300)
            want[1][c], want[2][c], want[3][c] = l, inv_l, bi
    for g, w, name in zip(got, want, ("kl", "lips", "inv_lips", "bi_lips")):
        np.testing.assert_allclose(g, w, rtol=2e-5, atol=1e-6, err_msg=name)
    assert got[0][10] != L.DEFAULT_EMPTY_CELL_FILL_VALUE and got[1][10] == L.DEFAULT_EMPTY_CELL_FILL_VALUE
    assert got[0][12] == L.DEFAULT_EMPTY_CELL_FILL_VALUE


def test_z_cells_and_data_based_match_loop():
    from vae_song_b200 import lipschitz as L
    from vae_song_b200.utils import estimate_local_lipschitz, reparameterize
    m = _model("lidvae")
    Kz, ns = 3, 40
    torch.manual_seed(5)
    got = L._get_kl_and_lipschitz_for_z_cells(m, Kz, -2.0, 2.0, 2, "cuda", nsamples_z_per_cell=ns, num_pairs_lips=200)
    torch.manual_seed(5)
    cx = np.linspace(-2.0, 2.0, Kz)
    want = [np.zeros(Kz * Kz, dtype=np.float32) for _ in range(4)]
    zs = []
    for yi in range(Kz):                             # all randn draws first, in cell order (the only RNG use)
        for xi in range(Kz):
            c = torch.tensor([cx[xi], cx[yi]], dtype=torch.float32, device="cuda")
            zs.append(c.repeat(ns, 1) + torch.randn(ns, 2, device="cuda") * 0.1)
    with torch.no_grad():
        for i, z in enumerate(zs):
            mu, lv = m.encode(m.decode(z))
            want[0][i] = _kl(mu, lv).mean().item()
            inv_l, l, bi = estimate_local_lipschitz(m.decode, z, num_pairs=200)
            want[1][i], want[2][i], want[3][i] = l, inv_l, bi
    for g, w, name in zip(got, want, ("kl", "lips", "inv_lips", "bi_lips")):
        np.testing.assert_allclose(g, w, rtol=2e-5, atol=1e-6, err_msg=name)
    with pytest.raises(ValueError):
        L._get_kl_and_lipschitz_for_z_cells(m, Kz, -2.0, 2.0, 3, "cuda")
    # data-based estimates: both branches (data set larger / smaller than num_samples; the latter is reference defect D9)
    ds = L.GaussianMixture2D(4, 300, center_range=4, stds=0.3, seed=2)
    for num in (100, 1000):
        torch.manual_seed(7)
        inv_l, l, bi = L._get_data_based_lipschitz(m, ds, "cuda", num_samples=num, num_pairs_lips=500)
        assert np.isfinite([inv_l, l, bi]).all() and bi == max(inv_l, l)
        kl = L._get_data_based_kl(m, ds, "cuda", num_samples=num)
        assert np.isfinite(kl) and kl > 0


def test_cli_runs_end_to_end(tmp_path):
    from vae_song_b200 import lipschitz as L
    res = L.main(["--model", "lidvae", "--IL", "0.2", "--beta", "0.001", "--K", "4", "--K_z", "3", "--std", "0.3", "--epochs", "1",
                  "--hidden_channels", "8", "4", "2", "--train_total_samples", "512", "--batch_size", "128", "--device", "cuda",
                  "--output_dir", str(tmp_path), "--seed", "3"])
    assert (tmp_path / "metrics.npz").exists() and (tmp_path / "summary.json").exists()
    assert res["kl_x"].shape == (16,) and res["kl_z"].shape == (9,) and np.isfinite(res["data_kl"])
