"""CPU-side tests: the C-ABI library loads and exports every symbol include/b200vae.h declares, the
drop-in modules keep the reference's state_dict keys / shapes, host-side sharding logic, and the
data-parallel trainer (world_size 2, gloo) equals single-process full-batch training."""
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def test_abi_exports_match_header():
    from vae_song_b200 import _C
    if not os.path.exists(_C.LIB_PATH):
        _C.build()
    lib = _C.load()
    hdr = open(os.path.join(ROOT, "include", "b200vae.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|long long|const char\*)\s+(b200vae_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared == set(_C.EXPORTS), declared ^ set(_C.EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.b200vae_version()
    assert lib.b200vae_lipschitz_num_tiles(5000) == 79 * 80 // 2
    assert lib.b200vae_icnn_workspace_bytes(65536, 2, 1024, 0, 0) > 2 * 1024 * 1024 * 4


def test_no_cpu_fallback():
    from vae_song_b200 import _C, model
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[32, 32])
    with pytest.raises(_C.B200VaeError):
        m.decode(torch.zeros(4, 2))
    # every ICNN entry point, narrow and wide inputs, with and without autograd: CPU tensors are refused, never computed
    from vae_song_b200 import module
    for d in (2, 32):
        ic = module.ICNN(d, 16)
        for grad in (True, False):
            with torch.set_grad_enabled(grad):
                with pytest.raises(_C.B200VaeError):
                    ic(torch.zeros(4, d, requires_grad=grad))
                with pytest.raises(_C.B200VaeError):
                    ic.brenier(torch.zeros(4, d), 0.1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vae_song_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn


def test_state_dict_keys_match_reference_golden():
    from vae_song_b200 import model
    G = np.load(os.path.join(GOLDEN, "lidvae_cases.npz"))
    m = model.LIDVAE(dataset="pinwheel", icnn_channels=[64, 128], hidden_channels=[16, 8], inverse_lipschitz=0.3, beta=0.7)
    ref = {k[len("pin_small/sd/"):]: G[k].shape for k in G.files if k.startswith("pin_small/sd/")}
    ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref
    assert m.il_factor == 0.15 and not hasattr(m, "wu_alpha") and m.warmup(0, 10) is False


def test_constructor_surface():
    from vae_song_b200 import model
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(model.LIDVAE(dataset="pinwheel", hidden_channels=[128, 64, 64, 32, 16, 8, 4, 2])) == 1337744
    assert n(model.LRVAE(alpha=1e-4, beta=0.01, dataset="pinwheel", hidden_channels=[16] * 12)) == 6958
    assert n(model.LRVAE(beta=0.001, alpha=0.1, dataset="mnist", encoder_type="conv", decoder_type="mlp")) == 1896848
    m = model.LIDVAE(dataset="mnist")            # reference defect D1 fixed
    assert n(m) == 3776658 and tuple(m.B.shape) == (784, 32)
    with pytest.raises(ValueError):
        model.LIDVAE(dataset="nope")
    with pytest.raises(ValueError):
        model.LIDVAE(dataset="pinwheel", icnn_channels=[1, 2, 3])
    v = model.VanillaVAE(dataset="pinwheel", hidden_channels=[4])
    v.warmup(0, 10)
    assert 0 < v.wu_alpha <= 1


def test_same_seed_same_init_as_reference_order():
    """ICNN constructs its layers in the reference order, so one seed gives the golden's shapes/keys and
    PositiveLinear uses kaiming_uniform(a=sqrt(5))."""
    from vae_song_b200 import module
    torch.manual_seed(0)
    ic = module.ICNN(2, 16)
    assert list(dict(ic.named_parameters())) == ["W.0.param", "W.1.param", "A.0.weight", "A.0.bias", "A.1.weight",
                                                  "A.1.bias", "A0.weight", "A0.bias"]
    assert float(ic.W[0].param.abs().max()) <= 1 / 4 + 1e-6


def test_tile_and_row_sharding():
    from vae_song_b200 import train, utils
    for nt, w in ((10, 3), (3160, 8), (1, 4), (0, 2)):
        rs = [utils.tile_range(nt, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == nt and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    assert train.shard_rows(64, 3, 4) == (48, 64)
    with pytest.raises(ValueError):
        train.shard_rows(10, 0, 4)


class _ToyVAE(torch.nn.Module):
    """Stock-torch stand-in with the LIDVAE forward/loss contract (the CUDA kernels cannot run here)."""

    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Sequential(torch.nn.Linear(2, 6, bias=False), torch.nn.BatchNorm1d(6), torch.nn.LeakyReLU(),
                                       torch.nn.Linear(6, 4))
        self.dec = torch.nn.Linear(2, 2)
        self.beta = 0.3

    def forward(self, x, eps=None):
        mu, lv = self.enc(x).split(2, 1)
        z = mu + eps * torch.exp(0.5 * lv)
        return self.dec(z), mu, lv, z, None

    def loss(self, x, recon, mu, lv, z, zr):
        rec = ((x - recon) ** 2).mean(0).sum()
        reg = (-0.5 * (1 + lv - mu ** 2 - lv.exp())).mean(0).sum()
        return rec + self.beta * reg, rec.detach(), reg.detach(), 0.0


def _torch_adam(flat, grad, m, v, t, scale, hp):
    g = grad * scale
    b1, b2 = hp["betas"]
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(hp["eps"])
    flat.addcdiv_(m, denom, value=-hp["lr"] / (1 - b1 ** t))


def _dp_worker(rank, world, port, x, eps, out):
    import torch.distributed as dist
    from vae_song_b200 import train
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(123 + rank)                       # replicas start different; rank 0 is broadcast
    tr = train.DataParallelTrainer(_ToyVAE(), lr=1e-2, optimizer_step=_torch_adam)
    lo, hi = train.shard_rows(x.shape[0], rank, world)
    for _ in range(3):
        tr.step(x[lo:hi], eps[lo:hi])
    assert any(type(mod).__name__ == "SyncBatchNorm1d" for mod in tr.model.modules())
    if rank == 0:
        out.put((tr.fp.flat.clone().numpy(), tr.model.enc[1].running_var.clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_trainer_gloo_world2_equals_single():
    from vae_song_b200 import train
    torch.manual_seed(0)
    x, eps = torch.randn(16, 2), torch.randn(16, 2)
    torch.manual_seed(123)
    single = train.DataParallelTrainer(_ToyVAE(), lr=1e-2, optimizer_step=_torch_adam)
    for _ in range(3):
        single.step(x, eps)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, x, eps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, rvar = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, single.fp.flat.numpy(), rtol=2e-4, atol=2e-6)      # cross-rank BatchNorm statistics
    np.testing.assert_allclose(rvar, single.model.enc[1].running_var.numpy(), rtol=1e-5)


def test_flat_params_views():
    from vae_song_b200 import train
    m = _ToyVAE()
    fp = train.FlatParams(m)
    # every parameter starts on a 16-byte boundary inside the flat buffer (padding stays zero)
    assert fp.numel == sum((p.numel() + 3) // 4 * 4 for p in m.parameters()) and fp.numel % 4 == 0
    assert all(v.data_ptr() % 16 == fp.grad.data_ptr() % 16 for v in fp.views)
    assert all(p.data_ptr() % 16 == fp.flat.data_ptr() % 16 for p in m.parameters())
    m.enc[0].weight.data.fill_(2.0)
    assert float(fp.flat[:12].sum()) == 24.0
    x = torch.randn(5, 2)
    out = m(x, eps=torch.zeros(5, 2))
    m.loss(x, *out)[0].backward()
    assert float(fp.grad.abs().sum()) > 0 and m.enc[0].weight.grad.data_ptr() == fp.grad.data_ptr()


class _ToyLRVAE(torch.nn.Module):
    """Stock-torch stand-in with the LR-VAE contract of model.py:418-447 / 587-616: decode attached and detached, re-encode,
    ATTACHED loss parts, and a latent-recon term that is a mean over L but a SUM over the batch (Appendix B.2) -- the one
    term whose global value is the sum, not the mean, of the per-rank values."""

    def __init__(self, set_style=False):
        super().__init__()
        self.encoder = torch.nn.Sequential(torch.nn.Linear(2, 5), torch.nn.LeakyReLU(), torch.nn.Linear(5, 4))
        self.decoder = torch.nn.Linear(2, 2)
        self.beta, self.alpha = 0.3, 0.5
        # SetLRVAE (model.py:1093-1114): latents stay [B,D], so the same `.mean(dim=0)` IS a batch mean there, z_input is
        # attached, and the loss parts come back detached -- nothing to compensate when sharding
        self.set_style = set_style

    def forward(self, x, eps=None, L=1):
        mu, lv = self.encoder(x).split(2, 1)
        z = mu + eps * torch.exp(0.5 * lv)
        z_rec = self.encoder(self.decoder(z.detach())).split(2, 1)[0]
        if self.set_style:
            return self.decoder(z.detach()), mu, lv, z, z_rec
        return self.decoder(z), mu, lv, z.detach()[None], z_rec[None]

    def loss(self, x, recon, mu, lv, z_in, z_rec):
        rec = ((x - recon) ** 2).mean(0).sum()
        reg = self.beta * (-0.5 * (1 + lv - mu ** 2 - lv.exp())).mean(0).sum()
        lr = self.alpha * ((z_in - z_rec) ** 2).mean(0).sum()
        if self.set_style:
            return rec + reg + lr, rec.detach(), reg.detach(), lr.detach()
        return rec + reg + lr, rec, reg, lr


def _train_lr(x, eps, lo, hi, staged, clip, set_style=False):
    from vae_song_b200 import train
    torch.manual_seed(321)
    tr = train.DataParallelTrainer(_ToyLRVAE(set_style), lr=1e-2, optimizer_step=_torch_adam, staged_backward=staged, grad_clip=clip,
                                   forward_kwargs={"L": 1})
    for _ in range(3):
        tr.step(x[lo:hi], eps[lo:hi])
    return tr.fp.flat.clone().numpy()


def _lr_worker(rank, world, port, x, eps, staged, clip, out, set_style=False):
    import torch.distributed as dist
    from vae_song_b200 import train
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = train.shard_rows(x.shape[0], rank, world)
    flat = _train_lr(x, eps, lo, hi, staged, clip, set_style)
    if rank == 0:
        out.put(flat)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("staged", [False, True], ids=["one_backward", "staged_backward"])
@pytest.mark.parametrize("clip", [None, {"enabled": True, "clip_type": "norm", "max_norm": 0.05, "norm_type": 2.0}],
                         ids=["noclip", "clipnorm"])
def test_sharded_lr_vae_step_equals_single(staged, clip):
    """Batch-sharded LR-VAE training (world 2, gloo) == single-process training on the whole batch: the batch-summed
    latent-recon term is compensated before the 1/W gradient averaging, in the one-backward and in the staged
    (main.py:262-284) step, with and without gradient clipping."""
    _check_sharded_lr(staged, clip, False)


def test_sharded_set_lr_vae_step_equals_single():
    """SetLRVAE-style losses (batch-MEAN latent-recon term, detached parts): sharding needs no compensation."""
    _check_sharded_lr(True, None, True)


def _check_sharded_lr(staged, clip, set_style):
    torch.manual_seed(1)
    x, eps = torch.randn(12, 2), torch.randn(12, 2)
    single = _train_lr(x, eps, 0, 12, staged, clip, set_style)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lr_worker, args=(r, 2, port, x, eps, staged, clip, q, set_style)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, single, rtol=1e-5, atol=1e-7)


def test_flat_params_begin_step_and_gather():
    """begin_step() detaches every .grad so autograd hands gradients over as produced; gather() moves them into the flat
    buffer (one multi-tensor copy), re-attaches the views, and zeroes the segment of a parameter that stops receiving one."""
    from vae_song_b200 import train
    m = _ToyVAE()
    fp = train.FlatParams(m)
    x, eps = torch.randn(6, 2), torch.randn(6, 2)
    ref = _ToyVAE(); ref.load_state_dict(m.state_dict())

    def run(model, use_dec=True):
        out = model(x, eps=eps)
        loss = model.loss(x, *out)[0] if use_dec else out[1].pow(2).sum()      # second form: decoder gets no gradient
        loss.backward()

    fp.begin_step()
    assert all(p.grad is None for p in m.parameters())
    run(m)
    fp.gather()
    run(ref)
    for p, q, view in zip(m.parameters(), ref.parameters(), fp.views):
        assert p.grad.data_ptr() == view.data_ptr()
        np.testing.assert_allclose(p.grad.numpy(), q.grad.numpy(), rtol=1e-6, atol=1e-8)
    # padding between parameters stays zero, and the flat buffer is exactly the concatenation of the (padded) gradients
    assert float(fp.grad.abs().sum()) == pytest.approx(sum(float(q.grad.abs().sum()) for q in ref.parameters()), rel=1e-5)
    # next step: the decoder receives no gradient -> its segment must read zero, not the stale values
    fp.begin_step()
    run(m, use_dec=False)
    fp.gather()
    assert float(m.dec.weight.grad.abs().sum()) == 0.0 and float(m.dec.bias.grad.abs().sum()) == 0.0
    assert float(m.enc[0].weight.grad.abs().sum()) > 0.0


def _allpairs_worker(rank, world, port, X, Y, out):
    import torch.distributed as dist
    from vae_song_b200 import utils
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # this rank's share of the unordered pairs (any partition combines the same way as the kernel's tile ranges)
    i, j = np.triu_indices(X.shape[0], 1)
    mine = (np.arange(i.size) % world) == rank
    dx = np.maximum(np.sqrt(((X[i[mine]] - X[j[mine]]) ** 2).sum(1)), 1e-3)
    dy = np.maximum(np.sqrt(((Y[i[mine]] - Y[j[mine]]) ** 2).sum(1)), 1e-3)
    r = dy / dx
    stats = torch.tensor([r.max(), r.min(), r.sum(), float(r.size)], dtype=torch.float64)
    hist = torch.tensor(np.histogram(np.log2(r), bins=8, range=(-4, 4))[0], dtype=torch.float64)
    stats, hist = utils.combine_allpairs(stats, hist)
    if rank == 0:
        out.put((stats.numpy(), hist.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_allpairs_partials_combine_across_ranks():
    """utils.combine_allpairs: MAX / MIN / SUM all-reduce of per-rank [max, min, sum, count] (+ histogram) == the global
    all-pairs statistics of the oracle (world 3, gloo)."""
    from oracle import loss_oracle as lo
    rng = np.random.default_rng(8)
    X = rng.normal(0, 1, (90, 2)); Y = np.tanh(X @ rng.normal(0, 1, (2, 3))) * 2
    ref = lo.lipschitz_allpairs(X, Y)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allpairs_worker, args=(r, 3, port, X, Y, q)) for r in range(3)]
    for p in procs:
        p.start()
    stats, hist = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(stats, [ref["max"], ref["min"], ref["sum"], ref["count"]], rtol=1e-10)
    assert hist.sum() <= ref["count"] and hist.sum() > 0
