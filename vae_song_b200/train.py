"""Training loops for the hot path.

train_model           the reference's LID-VAE / LR-VAE loop, lipschitz.py:23-44, same signature.
DataParallelTrainer   one process per GPU: batch sharded by rank, parameters and gradients live in ONE
                      flat fp32 buffer each, a single NCCL all-reduce per step over NVLink, fused Adam
                      (b200vae_adam_step) over the flat buffers.  The path has exactly one exchange
                      step (the gradient all-reduce), so nothing else communicates
                      except SyncBatchNorm statistics in the stock encoder.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .utils import apply_grad_clip


def train_model(model, loader, epochs, lr, device, grad_clip=None, wu_strat="linear", wu_start_epoch=0,
                wu_up_amount=None, wu_repeat_interval=10, experiment_logger=None):
    """lipschitz.py:23-44: Adam(lr); per batch forward -> loss -> backward -> clip -> step."""
    model.to(device).train()
    optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    for epoch in range(epochs):
        model.warmup(epoch=epoch, max_epoch=epochs, wu_strat=wu_strat, up_amount=wu_up_amount,
                     start_epoch=wu_start_epoch, repeat_interval=wu_repeat_interval)
        if experiment_logger and hasattr(model, "wu_alpha"):
            experiment_logger.log_alpha_value(epoch, model.wu_alpha)
        for X, _ in loader:
            X = X.to(device)
            optimizer.zero_grad()
            recon, mu, log_var, z_in, z_rec = model(X)
            total_loss, _, _, _ = model.loss(X, recon, mu, log_var, z_in, z_rec)
            total_loss.backward()
            apply_grad_clip(model, grad_clip)
            optimizer.step()
    return model


class FlatParams:
    """Re-homes every parameter (and its .grad) of `model` as a view into one flat fp32 buffer."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        dev, n = params[0].device, sum(p.numel() for p in params)
        self.params = params
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            off += k
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:   # autograd may have replaced .grad; re-attach the views
            k = p.numel()
            view = self.grad[off:off + k].view_as(p.data)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += k


def shard_rows(n_rows, rank, world):
    """Contiguous equal shards (rank r gets rows [r*n/W, (r+1)*n/W)); n must divide evenly so that the
    mean-over-batch losses average exactly across ranks."""
    if n_rows % world:
        raise ValueError(f"global batch {n_rows} is not divisible by world size {world}")
    per = n_rows // world
    return rank * per, (rank + 1) * per


class DataParallelTrainer:
    """Batch-sharded trainer.  `step(x_local, eps_local=None)` runs forward/loss/backward on this rank's
    shard, all-reduces the flat gradient (SUM) once, and applies Adam with grad_scale = 1/world.
    `optimizer_step` can be replaced (tests on CPU/gloo inject a torch implementation; on CUDA the fused
    kernel is used and there is no fallback)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 grad_clip=None, sync_bn=True, optimizer_step=None):
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.pg = process_group
        if self.world > 1 and sync_bn and next(model.parameters()).is_cuda:
            model = nn.SyncBatchNorm.convert_sync_batchnorm(model, process_group)
        self.model = model
        self.fp = FlatParams(model)
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat)
        self.t = 0
        self.hp = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.grad_clip = grad_clip
        self._opt = optimizer_step or self._fused_adam
        if self.world > 1:   # identical replicas to start from
            dist.broadcast(self.fp.flat, src=0, group=self.pg)

    def _fused_adam(self, flat, grad, m, v, t, scale, hp):
        ops.adam_step_(flat, grad, m, v, t, hp["lr"], hp["betas"], hp["eps"], hp["weight_decay"], scale)

    def step(self, x_local, eps_local=None):
        model, W = self.model, self.world
        self.fp.zero_grad()
        kw = {} if eps_local is None else {"eps": eps_local}
        out = model(x_local, **kw)
        total, rec, reg, lr_term = model.loss(x_local, *out)
        if W > 1 and torch.is_tensor(lr_term) and lr_term.requires_grad:
            # latent-recon term sums over the batch (Appendix B.2): global value = SUM over ranks, while
            # every other term is a batch mean -> compensate before the 1/W gradient averaging
            total = total + (W - 1) * lr_term
        total.backward()
        if W > 1:
            dist.all_reduce(self.fp.grad, op=dist.ReduceOp.SUM, group=self.pg)
        scale = 1.0 / W
        if self.grad_clip and self.grad_clip.get("enabled", False):
            self.fp.grad.mul_(scale)
            scale = 1.0
            apply_grad_clip(model, self.grad_clip)
        self.t += 1
        self._opt(self.fp.flat, self.fp.grad, self.m, self.v, self.t, scale, self.hp)
        return total.detach(), rec, reg

    @torch.no_grad()
    def global_losses(self, *local_scalars):
        """Mean over ranks of per-rank batch-mean scalars (one tiny all-reduce)."""
        t = torch.stack([torch.as_tensor(s, dtype=torch.float32, device=self.fp.flat.device).reshape(()) for s in local_scalars])
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
            t /= self.world
        return t
