"""Config-driven experiment driver with the reference's main.py contract (main.py:174-393 train_and_test, :395-580
run_experiment, :582-590 CLI): the same YAML schema (experiment_type / common_params / model_params), the same mapping of
config keys to constructor kwargs, the same optimiser (Adam 1e-2 + cosine annealing over all steps) and -- verbatim in
behaviour -- the STAGED backward of main.py:255-292 (latent-recon term first, encoder gradients scaled by 1e-4, then the
KL term, then the reconstruction term; one `loss.backward()` when the model returns detached parts).  Every model class
comes from vae_song_b200.model, so the hot path runs on the B200 kernels.

Data: the reference's loaders (dataset.py: torchvision downloads, ShapeNet files) are host-side and out of scope, and
there is no network here.  `train_and_test` therefore accepts any `(loader_train, loader_test)`; without them it builds
SYNTHETIC data of the config's shape (`synthetic_dataset`).  Plots / TensorBoard are not produced; per-epoch metrics are
returned and written to `<result_dir>/history.json`.

    python -m vae_song_b200.main --config configs/config_pinwheel.yaml [--epochs 5] [--device cuda]
"""
from __future__ import annotations

import argparse
import json
import os
import random
from datetime import datetime

import numpy as np
import torch
import yaml
from torch.utils.data import DataLoader, TensorDataset

from . import model as Model
from .utils import apply_grad_clip

ENCODER_LR_WEIGHT = 0.0001     # `lam` of main.py:270
SEED = 42                      # main.py:31


def seed_everything(seed=SEED):
    """main.py:31-36 seeds random, numpy and torch (CPU + every CUDA device) at import time; the drivers here call this
    explicitly at the start of a run instead (importing a package should not reseed the caller's generators)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def load_config(config_path):
    with open(config_path, "r") as f:
        return yaml.safe_load(f)


# ----------------------------------------------------------------------------------------------------- synthetic data
def _pinwheel(n, rng, num_classes=5, radial_std=0.3, tangential_std=0.05, rate=0.25):
    """Five-arm pinwheel (same family as dataset.py:119-161; host numpy, not a parity target)."""
    lab = rng.integers(0, num_classes, n)
    feats = rng.normal(0, 1, (n, 2)) * np.array([radial_std, tangential_std]) + np.array([1.0, 0.0])
    ang = 2 * np.pi * lab / num_classes + rate * np.exp(feats[:, 0])
    rot = np.stack([np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)], 1).reshape(n, 2, 2)
    return np.einsum("ni,nij->nj", feats, rot).astype(np.float32), lab


def _chessboard(n, rng):
    """4x4 chessboard on [-2,2]^2 (dataset.py:72-102 family)."""
    out = np.empty((0, 2), np.float32)
    while out.shape[0] < n:
        p = rng.uniform(-2, 2, (2 * n, 2))
        keep = ((np.floor(p[:, 0]) + np.floor(p[:, 1])) % 2 == 0)
        out = np.concatenate([out, p[keep].astype(np.float32)])
    return out[:n], np.zeros(n, np.int64)


def synthetic_dataset(dataset_name, n_train=10000, n_test=2000, seed=42, num_points=2048):
    """(train, test) TensorDatasets of the shape the named reference data set has."""
    rng = np.random.default_rng(seed)

    def make(n):
        if dataset_name == "pinwheel":
            x, y = _pinwheel(n, rng)
        elif dataset_name == "chessboard":
            x, y = _chessboard(n, rng)
        elif dataset_name in ("mnist", "fashionmnist", "omniglot"):
            x, y = rng.uniform(0, 1, (n, 1, 28, 28)).astype(np.float32), rng.integers(0, 10, n)
        elif dataset_name in ("cifar10",):
            x, y = rng.uniform(0, 1, (n, 3, 32, 32)).astype(np.float32), rng.integers(0, 10, n)
        elif dataset_name == "celeba":
            x, y = rng.uniform(0, 1, (n, 3, 64, 64)).astype(np.float32), rng.integers(0, 2, n)
        elif dataset_name == "shapenet":
            x, y = rng.normal(0, 1, (n, num_points, 3)).astype(np.float32), np.zeros(n, np.int64)
        else:
            raise ValueError(f"Unsupported dataset: {dataset_name}")
        return TensorDataset(torch.from_numpy(x), torch.from_numpy(np.asarray(y, dtype=np.int64)))
    return make(n_train), make(n_test)


# ----------------------------------------------------------------------------------------------------- training
def staged_backward(model, loss, loss_recon, loss_reg, loss_lr):
    """main.py:262-284: backward the attached loss parts one by one, scaling the encoder's latent-recon gradients."""
    did_backward = False
    if hasattr(loss_lr, "requires_grad") and loss_lr.requires_grad:
        loss_lr.backward(retain_graph=True)
        did_backward = True
        for param in model.encoder.parameters():
            if param.grad is not None:
                param.grad *= ENCODER_LR_WEIGHT
    if hasattr(loss_reg, "requires_grad") and loss_reg.requires_grad:
        loss_reg.backward(retain_graph=True)
        did_backward = True
    if hasattr(loss_recon, "requires_grad") and loss_recon.requires_grad:
        loss_recon.backward()
        did_backward = True
    if not did_backward:
        loss.backward()


def train_step(model, x, optimizer, scheduler=None, num_mc_samples=1, grad_clip=None):
    """One iteration of main.py:255-292.  Returns the four loss parts as device scalars (no host sync here)."""
    result = model(x, L=num_mc_samples)
    loss, loss_recon, loss_reg, loss_lr = model.loss(x, *result)
    optimizer.zero_grad()
    staged_backward(model, loss, loss_recon, loss_reg, loss_lr)
    apply_grad_clip(model, grad_clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    as_t = lambda v: v.detach().float().reshape(()) if torch.is_tensor(v) else torch.tensor(float(v), device=x.device)
    return torch.stack([as_t(loss), as_t(loss_recon), as_t(loss_reg), as_t(loss_lr)])


def evaluate(model, loader, device):
    """main.py:91-172 reduced to its numbers: mean loss parts over the test loader, `model(x)` with the DEFAULT number of
    Monte-Carlo samples like main.py:103 (autograd stays enabled, like the reference, because its LIDVAE.decode needs
    it; ours does not care)."""
    model.eval()
    tot, n = torch.zeros(4, device=device), 0
    for x, _ in loader:
        x = x.to(device)
        result = model(x)
        parts = model.loss(x, *result)
        tot += torch.stack([p.detach().float().reshape(()) if torch.is_tensor(p) else torch.tensor(float(p), device=device)
                            for p in parts])
        n += 1
    return (tot / max(n, 1)).tolist()


def train_and_test(model, epochs=100, batch_size=128, device="cuda", dataset_name="mnist", logfilename="log.csv",
                   resultname="res", pt_param=None, num_mc_samples=1, grad_clip=None, wu_strat="linear",
                   loader_train=None, loader_test=None, result_root="./results", dataset_params=None, num_workers=0,
                   graph=False, run_tag=None):
    """main.py:174-393.  Returns {'train': [[loss, recon, reg, lr] per epoch], 'test': [...], 'name': run name}.

    graph=True: the same step (forward, staged backward, clipping, Adam, cosine schedule) captured ONCE into a CUDA graph
    and replayed (train.DataParallelTrainer): the warm-up factor and the learning rate live on the device, so no host
    work remains per step -- these small models are otherwise bound by ~1500 eager launches per step."""
    if loader_train is None:
        dp = dataset_params or {}
        tr, te = synthetic_dataset(dataset_name, dp.get("n_train", 10000), dp.get("n_test", 2000), dp.get("seed", 42),
                                   dp.get("num_points", 2048))
        test_shuffle = dataset_name in ("pinwheel", "chessboard")                       # main.py:180
        loader_train = DataLoader(tr, batch_size=batch_size, shuffle=True, num_workers=num_workers, drop_last=True,
                                  pin_memory=True)
        loader_test = DataLoader(te, batch_size=batch_size, shuffle=test_shuffle, num_workers=num_workers, drop_last=True,
                                 pin_memory=True)
    model = model.to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2)                          # main.py:200
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, max(epochs * len(loader_train), 1))
    if pt_param is not None:
        if not os.path.exists(pt_param):
            raise FileNotFoundError(f"No such file: {pt_param}")                        # the reference calls exit()
        model.load_state_dict(torch.load(pt_param, map_location=device))
    # run directory name of main.py:211-218: class name + " %m%d%H%M" timestamp + hyper-parameters.  `run_tag` (the repeat
    # index when a config asks for niter > 1) keeps repeats that start within the same minute from overwriting each other.
    name = type(model).__name__ + datetime.now().strftime(" %m%d%H%M") + (f"_r{run_tag}" if run_tag is not None else "")
    if not type(model).__name__.startswith("NaiveAE"):
        name += "_b=" + str(float(model.beta))
    if type(model).__name__.startswith("LR") or type(model).__name__.startswith("SetLR"):
        name += "_a=" + str(model.alpha)
    if getattr(model, "is_log_mse", False):
        name += "_logmse"
    if type(model).__name__ == "LIDVAE":
        name += "_il=" + str(float(model.il_factor))
    out_dir = os.path.join(result_root, resultname, name)
    os.makedirs(os.path.join(out_dir, "params"), exist_ok=True)
    history = {"name": name, "train": [], "test": []}
    trainer = None
    if graph:
        from .train import DataParallelTrainer
        model.train()
        model.graph_scalars(device)
        trainer = DataParallelTrainer(model, lr=1e-2, grad_clip=grad_clip, staged_backward=True,
                                      forward_kwargs={"L": num_mc_samples},
                                      lr_schedule=("cosine", max(epochs * len(loader_train), 1)))
        trainer.capture(next(iter(loader_train))[0].to(device))
    for epoch in range(epochs):
        model.train()
        model.warmup(epoch=epoch, max_epoch=epochs, wu_strat=wu_strat)
        tot, nb = torch.zeros(4, device=device), 0
        for x, _ in loader_train:
            x = x.to(device, non_blocking=True)
            if trainer is not None:
                trainer.step_graphed(x)
                tot += torch.stack([p.reshape(()).float() if torch.is_tensor(p) else torch.full((), p, device=device)
                                    for p in trainer.last_parts])
            else:
                tot += train_step(model, x, optimizer, scheduler, num_mc_samples, grad_clip)
            nb += 1
        history["train"].append((tot / max(nb, 1)).tolist())                             # one sync per epoch
        if loader_test is not None:
            history["test"].append(evaluate(model, loader_test, device))
    # main.py:307-310: the state_dict of the LAST epoch, named after its 0-based index (utils.py:365 and test.py look
    # for params/model_<epochs-1>.pt)
    ckpt = os.path.join(out_dir, "params", f"model_{max(epochs - 1, 0)}.pt")
    torch.save(model.state_dict(), ckpt)
    history["checkpoint"], history["out_dir"] = ckpt, out_dir
    with open(os.path.join(out_dir, "history.json"), "w") as f:
        json.dump(history, f)
    return history


# ----------------------------------------------------------------------------------------------------- config -> models
def iter_models(config):
    """Yield (tag, model, train kwargs) for every (alpha, beta, IL, repeat) of the config: main.py:423-578.  With
    niter > 1 the tag and the kwargs carry the repeat index so that repeats neither share a results key nor a run
    directory."""
    exp_type, cp, mp = config["experiment_type"], config["common_params"], config["model_params"]
    data = cp.get("exp_data", "shapenet")
    kw = dict(epochs=cp["exp_epochs"], batch_size=cp["batch_size"], dataset_name=data, pt_param=cp.get("pt_param", None),
              num_mc_samples=mp.get("num_mc_samples", 1), grad_clip=cp.get("grad_clip", None))
    flex = dict(dataset=data, hidden_channels=mp.get("hchans", None), encoder_type=mp.get("encoder_type", "conv"),
                decoder_type=mp.get("decoder_type", "mlp"))
    setkw = dict(latent_channel=mp.get("latent_channel", 128), num_points=mp.get("num_points", 2048),
                 encoder_hidden=mp.get("encoder_hidden", [128, 256, 512]), decoder_hidden=mp.get("decoder_hidden", [512, 256, 128]),
                 dataset="shapenet", pool_type=mp.get("pool_type", "max"), use_attention=mp.get("use_attention", True),
                 d_model=mp.get("d_model", 256), num_heads=mp.get("num_heads", 4),
                 num_encoder_layers=mp.get("num_encoder_layers", 2), num_decoder_layers=mp.get("num_decoder_layers", 2),
                 ff_dim=mp.get("ff_dim", 512), attn_dropout=mp.get("attn_dropout", 0.0))
    wu = dict(wu_strat=cp.get("wu_strat", "linear"))
    niter = cp["niter"]
    for rep, item in ((r, it) for r in range(niter) for it in _iter_once(exp_type, cp, mp, data, kw, flex, setkw, wu)):
        tag, model, k = item
        if niter > 1:
            tag, k = f"{tag}_r{rep}", {**k, "run_tag": rep}
        yield tag, model, k


def _iter_once(exp_type, cp, mp, data, kw, flex, setkw, wu):
    if True:
        if exp_type == "lidvae":
            for beta in mp["beta_list"]:
                for il in mp["il_list"]:
                    yield (f"lidvae_b{beta}_il{il}", Model.LIDVAE(is_log_mse=mp.get("log_mse", False), inverse_lipschitz=il,
                                                                  beta=beta, dataset=data, hidden_channels=mp.get("hchans", None),
                                                                  **({"precision": mp["precision"]} if "precision" in mp else {})), kw)
        elif exp_type == "vae":
            for beta in mp["beta_list"]:
                yield (f"vae_b{beta}", Model.VanillaVAE(beta=beta, fixed_var=mp.get("fixed_var", False),
                                                        residual_connection=mp.get("residual_connection", False), **flex), kw)
        elif exp_type == "nae":
            yield ("nae", Model.NaiveAE(**flex), kw)
        elif exp_type == "lrvae":
            for alpha in mp["alpha_list"]:
                for beta in mp["beta_list"]:
                    yield (f"lrvae_a{alpha}_b{beta}", Model.LRVAE(beta=beta, alpha=alpha, z_source=mp.get("z_source", "Ex"),
                                                                  pwise_reg=mp.get("pwise_reg", False),
                                                                  residual_connection=mp.get("residual_connection", False), **flex),
                           {**kw, **wu})
        elif exp_type == "setvae":
            for beta in mp.get("beta_list", [1.0]):
                yield (f"setvae_b{beta}", Model.SetVAE(beta=beta, **setkw), kw)
        elif exp_type == "setlrvae":
            for alpha in mp.get("alpha_list", [0.01]):
                for beta in mp.get("beta_list", [1.0]):
                    yield (f"setlrvae_a{alpha}_b{beta}", Model.SetLRVAE(alpha=alpha, beta=beta, **setkw), {**kw, **wu})
        else:
            raise ValueError(f"unknown experiment_type {exp_type!r}")


def run_experiment(config_path, device="cuda", epochs=None, result_root="./results", dataset_params=None, graph=False,
                   seed=SEED):
    if seed is not None:
        seed_everything(seed)                                                            # main.py:31-36
    config = load_config(config_path) if isinstance(config_path, (str, os.PathLike)) else config_path
    cp, mp = config["common_params"], config["model_params"]
    res_tag = "_res" if mp.get("residual_connection", False) else ""
    exp_str = f"{cp.get('exp_data', 'shapenet')}_{config['experiment_type']}{res_tag}_depth{len(mp.get('hchans') or [])}_mc{mp.get('num_mc_samples', 1)}"
    resultname = cp.get("resultname") or f"result_{exp_str}"
    dp = {**cp.get("dataset_params", {}), **(dataset_params or {})}
    out = {}
    for tag, model, kw in iter_models(config):
        if epochs is not None:
            kw = {**kw, "epochs": epochs}
        out[tag] = train_and_test(model, device=device, resultname=resultname, result_root=result_root, dataset_params=dp,
                                  graph=graph, **kw)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="configs/config_pinwheel.yaml")       # README.md:34-37 of the reference (defect D6)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--epochs", type=int, default=None, help="override common_params.exp_epochs")
    ap.add_argument("--result_root", default="./results")
    ap.add_argument("--graph", action="store_true", help="replay the whole train step as one CUDA graph")
    args = ap.parse_args(argv)
    res = run_experiment(args.config, device=args.device, epochs=args.epochs, result_root=args.result_root, graph=args.graph)
    for tag, h in res.items():
        print(tag, "final train [loss, recon, reg, lr] =", h["train"][-1] if h["train"] else None)
    return res


if __name__ == "__main__":
    main()
