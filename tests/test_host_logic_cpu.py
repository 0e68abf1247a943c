"""Host-side schedules against outputs of the unmodified reference (tests/golden/host_cases.npz, written by
`python oracle/make_golden.py host`): the alpha warm-up strategies of model.py:37-63 and utils.apply_grad_clip."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_cases.npz"))
WARMUP_CASES = [
    ("linear", dict(wu_strat="linear")),
    ("linear_up", dict(wu_strat="linear", up_amount=0.07, start_epoch=3)),
    ("exponential", dict(wu_strat="exponential", start_epoch=2)),
    ("exponential_up", dict(wu_strat="exponential", up_amount=0.05)),
    ("repeat_linear", dict(wu_strat="repeat_linear", repeat_interval=4, start_epoch=1)),
    ("kl_adaptive", dict(wu_strat="kl_adaptive")),
]


@pytest.mark.parametrize("name,kw", WARMUP_CASES, ids=[c[0] for c in WARMUP_CASES])
def test_warmup_schedules_follow_the_reference(name, kw):
    from vae_song_b200 import model
    m = model.LRVAE(beta=0.01, alpha=0.1, dataset="pinwheel", hidden_channels=[4], encoder_type="mlp", decoder_type="mlp")
    assert m.wu_alpha == 0.0
    seq = []
    for epoch in range(24):
        m.last_kl_loss = float(G["warmup/kl_seq"][epoch])
        assert m.warmup(epoch=epoch, max_epoch=20, **kw) is True
        seq.append(m.wu_alpha)
    np.testing.assert_allclose(seq, G["warmup/" + name], rtol=1e-12, atol=0)


def test_warmup_factor_mirrors_into_the_graph_scalar():
    """After graph_scalars() the losses multiply with a device copy of wu_alpha that follows every assignment (on CPU here)."""
    from vae_song_b200 import model
    m = model.LRVAE(beta=0.01, alpha=0.1, dataset="pinwheel", hidden_channels=[4], encoder_type="mlp", decoder_type="mlp")
    m.wu_alpha = 0.25
    assert m._lr_weight() == pytest.approx(0.025)
    m.graph_scalars("cpu")
    m.warmup(epoch=0, max_epoch=3)                               # linear: += 1/4
    w = m._lr_weight()
    assert torch.is_tensor(w) and float(w) == pytest.approx(0.1 * 0.5) and m.wu_alpha == pytest.approx(0.5)


@pytest.mark.parametrize("tag,cfg", [("norm", {"enabled": True, "clip_type": "norm", "max_norm": 0.7, "norm_type": 2.0}),
                                      ("value", {"enabled": True, "clip_type": "value", "clip_value": 0.5}),
                                      ("off", {"enabled": False, "clip_type": "norm", "max_norm": 0.1})])
def test_apply_grad_clip_follows_the_reference(tag, cfg):
    from vae_song_b200.utils import apply_grad_clip
    lin = torch.nn.Linear(4, 3)
    lin.weight.grad = torch.tensor(G["clip/g_w"]); lin.bias.grad = torch.tensor(G["clip/g_b"])
    apply_grad_clip(lin, cfg)
    np.testing.assert_allclose(lin.weight.grad.numpy(), G[f"clip/{tag}/w"], rtol=1e-6)
    np.testing.assert_allclose(lin.bias.grad.numpy(), G[f"clip/{tag}/b"], rtol=1e-6)
    apply_grad_clip(lin, None)                                   # None is a no-op, like the reference
