import numpy as np
import torch

from oracle import icnn_oracle as io

KEYS = io.PARAM_KEYS   # == vae_song_b200._C.PARAM_FIELDS order


def params_to_torch(p, device="cuda"):
    return [torch.tensor(np.asarray(p[k], dtype=np.float32), device=device) for k in KEYS]


def params_f32_as_f64(p):
    """What the GPU actually sees (fp32-rounded values), evaluated by the oracle in fp64."""
    return {k: np.asarray(p[k], dtype=np.float32).astype(np.float64) for k in KEYS}


def f32_as_f64(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def close_report(ours, ref, rtol, name, bad_frac=0.0, floor=0.0):
    """|ours-ref| <= rtol*|ref| + rtol*max|ref| elementwise, except a fraction `bad_frac` of elements
    (kink flips: a LeakyReLU unit whose pre-activation is within rounding of 0 changes xhat by O(1/H),
    SURVEY.md section 7).  Returns the worst normalised error for logging."""
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert ours.shape == ref.shape, (name, ours.shape, ref.shape)
    scale = max(np.abs(ref).max(), floor) + 1e-300   # floor: tensors that are mathematically zero (rounding noise)
    err = np.abs(ours - ref)
    tol = rtol * np.abs(ref) + rtol * scale
    bad = (err > tol)
    frac = bad.mean()
    assert np.isfinite(ours).all(), f"{name}: non-finite values"
    assert frac <= bad_frac, f"{name}: {bad.sum()}/{bad.size} elements out of tolerance, worst {(err / scale).max():.3e} (rtol {rtol})"
    return float((err / scale).max())
