"""vae_song_b200 -- B200-native (sm_100a) decoder-and-loss hot path of vae-song.

Drop-in surface (same names, signatures and state_dict keys as the reference):
    module.PositiveLinear, module.ICNN, model.LIDVAE, model.FlexibleVAE / VanillaVAE / LRVAE / NaiveAE,
    utils.estimate_local_lipschitz / reparameterize / kld / apply_grad_clip, train.train_model.
Every starred op runs in hand-written CUDA kernels behind the C ABI in include/b200vae.h.
"""
from . import _C  # noqa: F401

__all__ = ["_C", "ops", "module", "model", "utils", "train"]
__version__ = "0.1.0"
