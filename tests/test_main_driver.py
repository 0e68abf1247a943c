"""vae_song_b200/main.py: the reference's config-driven driver (main.py:174-580) over the B200 modules."""
import glob
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECTED = {   # parameter counts of the reference's models for these configs
    "config_pinwheel.yaml": ("LRVAE", 6958, 2),
    "config_mnist.yaml": ("LRVAE", 1896848, 1),
    "config_shapenet_setlrvae.yaml": ("SetLRVAE", 3260675, 1),
    "config_lidvae_chessboard.yaml": ("LIDVAE", 1321558, 2),
}


def test_configs_build_the_reference_models():
    from vae_song_b200 import main as M
    files = sorted(glob.glob(os.path.join(ROOT, "configs", "*.yaml")))
    assert {os.path.basename(f) for f in files} == set(EXPECTED)
    for f in files:
        cls, nparam, nmodels = EXPECTED[os.path.basename(f)]
        built = list(M.iter_models(M.load_config(f)))
        assert len(built) == nmodels
        for tag, model, kw in built:
            assert type(model).__name__ == cls
            assert sum(p.numel() for p in model.parameters()) == nparam
            assert {"epochs", "batch_size", "dataset_name", "num_mc_samples", "grad_clip"} <= set(kw)


def test_repeats_do_not_collide_and_runs_are_seeded():
    """niter > 1 (main.py:423: the outer repeat loop): every repeat gets its own results key and run directory suffix; the
    run directory carries the reference's " %m%d%H%M" timestamp; run_experiment seeds random / numpy / torch with 42
    (main.py:31-36)."""
    import random
    from vae_song_b200 import main as M
    cfg = M.load_config(os.path.join(ROOT, "configs", "config_pinwheel.yaml"))
    cfg["common_params"]["niter"] = 3
    built = list(M.iter_models(cfg))
    tags = [t for t, _, _ in built]
    assert len(tags) == len(set(tags)) and all(t.rsplit("_r", 1)[1] in "012" for t in tags)
    assert sorted({kw["run_tag"] for _, _, kw in built}) == [0, 1, 2]
    cfg["common_params"]["niter"] = 1
    assert all("run_tag" not in kw for _, _, kw in M.iter_models(cfg))
    M.seed_everything()
    a = (random.random(), float(np.random.rand()), float(torch.rand(())))
    M.seed_everything(42)
    assert a == (random.random(), float(np.random.rand()), float(torch.rand(())))


def test_synthetic_datasets_have_the_reference_shapes():
    from vae_song_b200 import main as M
    for name, shape in (("pinwheel", (2,)), ("chessboard", (2,)), ("mnist", (1, 28, 28)), ("cifar10", (3, 32, 32)),
                        ("shapenet", (64, 3))):
        tr, te = M.synthetic_dataset(name, 32, 8, num_points=64)
        assert tr.tensors[0].shape == (32, *shape) and te.tensors[0].shape == (8, *shape)
        assert tr.tensors[0].dtype == torch.float32 and tr.tensors[1].dtype == torch.int64
    with pytest.raises(ValueError):
        M.synthetic_dataset("nope")
    x = M.synthetic_dataset("chessboard", 500, 8)[0].tensors[0].numpy()
    assert ((np.floor(x[:, 0]) + np.floor(x[:, 1])) % 2 == 0).all()


@pytest.mark.gpu
def test_staged_backward_matches_its_definition():
    """main.py:262-284 on an LRVAE: grads == d recon + d reg + d lr, with the encoder's share of d lr scaled by 1e-4."""
    from vae_song_b200 import main as M, model
    torch.manual_seed(0)
    m = model.LRVAE(alpha=0.5, beta=0.3, dataset="pinwheel", hidden_channels=[16, 16], encoder_type="mlp", decoder_type="mlp").cuda().train()
    m.wu_alpha = 1.0
    x = torch.randn(64, 2, device="cuda")
    eps = torch.randn(2, 64, m.latent_channel, device="cuda")
    params = list(m.parameters())
    enc_ids = {id(p) for p in m.encoder.parameters()}

    def parts():
        torch.manual_seed(1)
        out = m(x, L=2, eps=eps)
        return m.loss(x, *out)
    loss, rec, reg, lr = parts()
    assert lr.requires_grad and reg.requires_grad and rec.requires_grad
    for p in params:
        p.grad = None
    M.staged_backward(m, loss, rec, reg, lr)
    got = [None if p.grad is None else p.grad.clone() for p in params]
    loss, rec, reg, lr = parts()
    g_rec = torch.autograd.grad(rec, params, retain_graph=True, allow_unused=True)
    g_reg = torch.autograd.grad(reg, params, retain_graph=True, allow_unused=True)
    g_lr = torch.autograd.grad(lr, params, allow_unused=True)
    for p, g, a, b, c in zip(params, got, g_rec, g_reg, g_lr):
        want = torch.zeros_like(p)
        for t, scale in ((a, 1.0), (b, 1.0), (c, M.ENCODER_LR_WEIGHT if id(p) in enc_ids else 1.0)):
            if t is not None:
                want = want + scale * t
        if g is None:
            assert float(want.abs().max()) == 0.0
        else:
            torch.testing.assert_close(g, want, rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["config_pinwheel.yaml", "config_lidvae_chessboard.yaml"])
def test_run_experiment_trains(cfg, tmp_path):
    from vae_song_b200 import main as M
    config = M.load_config(os.path.join(ROOT, "configs", cfg))
    config["model_params"]["beta_list"] = config["model_params"]["beta_list"][:1]
    if "il_list" in config["model_params"]:
        config["model_params"]["il_list"] = [0.2]
    config["common_params"]["batch_size"] = 256
    res = M.run_experiment(config, device="cuda", epochs=3, result_root=str(tmp_path),
                           dataset_params={"n_train": 2048, "n_test": 512})
    assert len(res) == 1
    h = next(iter(res.values()))
    assert len(h["train"]) == 3 and len(h["test"]) == 3
    if "lidvae" not in cfg:     # the ICNN's default exp(W) ~ 1 init gives ~1e20 losses: finiteness is not promised
        assert np.isfinite(h["train"]).all() and np.isfinite(h["test"]).all()
        # the TOTAL carries the latent-recon term whose weight is warmed up from epoch to epoch (model.warmup), so it may
        # rise; what must fall within three epochs is the reconstruction part
        assert h["train"][-1][1] < h["train"][0][1], h["train"]
    run_dir = os.path.join(str(tmp_path), os.listdir(str(tmp_path))[0])
    sub = os.path.join(run_dir, os.listdir(run_dir)[0])
    # main.py:307-310: the last epoch's state_dict under its 0-based index (utils.py:365 / test.py look for model_<E-1>.pt)
    assert os.path.exists(os.path.join(sub, "params", "model_2.pt")) and os.path.exists(os.path.join(sub, "history.json"))
    assert h["checkpoint"].endswith(os.path.join("params", "model_2.pt"))


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_trainer_staged_backward_and_graph_match_main_train_step(fused):
    """DataParallelTrainer(staged_backward=True): the same gradients as main.staged_backward, eagerly and as one CUDA graph."""
    import copy
    from vae_song_b200 import main as M, model, train
    torch.manual_seed(0)
    m = model.LRVAE(alpha=0.5, beta=0.3, dataset="pinwheel", hidden_channels=[16, 16], encoder_type="mlp", decoder_type="mlp").cuda().train()
    m.wu_alpha = 1.0
    m.fused_mlp = fused
    ref = copy.deepcopy(m)
    x = torch.randn(256, 2, device="cuda")
    eps = torch.randn(1, 256, m.latent_channel, device="cuda")
    out = ref(x, L=1, eps=eps)
    parts = ref.loss(x, *out)
    M.staged_backward(ref, *parts)
    want = {k: p.grad.clone() for k, p in ref.named_parameters()}
    tr = train.DataParallelTrainer(m, lr=1e-3, staged_backward=True)
    flat0 = tr.fp.flat.clone()
    tr.step(x, eps)
    for (k, p), view in zip(m.named_parameters(), tr.fp.views):
        torch.testing.assert_close(view, want[k], rtol=2e-4, atol=1e-6, msg=k)
    # graph replay == eager step from the same state
    tr2 = train.DataParallelTrainer(copy.deepcopy(ref), lr=1e-3, staged_backward=True)
    tr2.model.zero_grad(set_to_none=True)
    tr2.capture(x, eps)
    tr2.step_graphed(x, eps)
    torch.testing.assert_close(tr2.fp.flat, tr.fp.flat, rtol=1e-5, atol=1e-6)
    assert not torch.equal(tr.fp.flat, flat0)


@pytest.mark.gpu
def test_graph_mode_follows_the_eager_loop(tmp_path):
    """train_and_test(graph=True): warm-up factor and cosine learning rate live on the device -- the replayed graph must
    track the eager reference loop (torch Adam + CosineAnnealingLR + staged backward) over several epochs."""
    from vae_song_b200 import main as M, model
    X = M.synthetic_dataset("pinwheel", 1024, 256)[0].tensors[0]
    loader = [(X[i * 256:(i + 1) * 256], torch.zeros(256, dtype=torch.int64)) for i in range(4)]
    hist = {}
    for graph in (False, True):
        torch.manual_seed(0)
        m = model.LRVAE(alpha=0.3, beta=0.05, dataset="pinwheel", hidden_channels=[16, 16], encoder_type="mlp", decoder_type="mlp")
        torch.manual_seed(1)          # same eps stream (4 epochs x 4 steps of torch.randn on the device)
        hist[graph] = M.train_and_test(m, epochs=4, batch_size=256, device="cuda", dataset_name="pinwheel", num_mc_samples=2,
                                       grad_clip={"enabled": True, "clip_type": "norm", "max_norm": 1.0, "norm_type": 2.0},
                                       wu_strat="linear", loader_train=loader, loader_test=None, result_root=str(tmp_path),
                                       resultname="g%d" % graph, graph=graph)
    a, b = np.array(hist[False]["train"]), np.array(hist[True]["train"])
    assert a.shape == b.shape == (4, 4) and np.isfinite(b).all()
    # the eps draws differ between the two runs (capture consumes warm-up draws), so the per-epoch means agree up to
    # sampling noise, not bitwise: reconstruction and KL parts of every epoch within 20 %
    np.testing.assert_allclose(b[:, 1], a[:, 1], rtol=0.2, err_msg=f"recon per epoch: eager {a[:, 1]} graph {b[:, 1]}")
    np.testing.assert_allclose(b[:, 2], a[:, 2], rtol=0.2, err_msg=f"KL per epoch: eager {a[:, 2]} graph {b[:, 2]}")
    assert a[-1, 1] < a[0, 1] and b[-1, 1] < b[0, 1], (a[:, 1], b[:, 1])          # both learn to reconstruct
    # the latent-recon weight follows the linear warm-up INSIDE the replayed graph: weighted term / weight ~ comparable,
    # and the weighted term itself grows with the epochs like in the eager loop
    np.testing.assert_allclose(b[:, 3], a[:, 3], rtol=0.3, err_msg=f"alpha*wu*lr per epoch: eager {a[:, 3]} graph {b[:, 3]}")
