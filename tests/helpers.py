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


# ---------------------------------------------------------------------------------------------------- LeakyReLU kinks
# xhat = grad psi is piecewise linear in z: a hidden unit whose pre-activation h1[b,n] (or h2[b]) lies within the forward's
# rounding error of 0 may land on the other side of the kink and moves xhat[b] by O(1/H).  Instead of waving a FRACTION
# of elements through, the tests below pin those cases down with the oracle:
#   * `kink_rows`: the rows that have such a unit, computed from the oracle's own pre-activations and the DECLARED
#     accuracy `h_rtol` of the mode's forward (|h| < h_rtol * max|h|).  Every other row must meet the strict bound with
#     no exception; the kink rows must still meet `loose`.
#   * `check_decode_with_masks` (kernels that return their masks): every mask bit that differs from the oracle's must
#     belong to such a unit, and xhat must equal the oracle evaluated WITH THE KERNEL'S MASKS on every row, strictly.
H_RTOL = {0: 4e-6, 3: 1e-5, 4: 1e-5, 1: 3e-4}      # precision -> declared relative accuracy of the hidden pre-activations


def kink_rows(aux, h_rtol, with_h0=False):
    h1, h2 = np.asarray(aux["h1"]), np.asarray(aux["h2"])
    k = (np.abs(h1) < h_rtol * np.abs(h1).max()).any(1) | (np.abs(h2) < h_rtol * max(np.abs(h2).max(), 1e-300))
    if with_h0:      # the double-backward is also discontinuous where the FIRST layer crosses its kink (t0 carries s0^2)
        h0 = np.asarray(aux["h0"])
        k = k | (np.abs(h0) < h_rtol * np.abs(h0).max()).any(1)
    return k


def close_rows(ours, ref, rtol, name, kink, loose=5e-3):
    """Strict bound (no exceptions) on the rows outside `kink`, `loose` on the kink rows.  Returns the worst strict error."""
    ours, ref = np.asarray(ours, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert ours.shape == ref.shape, (name, ours.shape, ref.shape)
    assert np.isfinite(ours).all(), f"{name}: non-finite values"
    scale = np.abs(ref).max() + 1e-300
    err = np.abs(ours - ref).reshape(ours.shape[0], -1)
    tol = (rtol * np.abs(ref) + rtol * scale).reshape(ours.shape[0], -1)
    ok_rows = ~kink
    bad = (err > tol) & ok_rows[:, None]
    assert not bad.any(), (f"{name}: {bad.sum()} elements of {ok_rows.sum()} kink-free rows out of tolerance, worst "
                           f"{(err[ok_rows] / scale).max():.3e} (rtol {rtol}); {kink.sum()} kink rows set aside")
    if kink.any():
        worst = (err[kink] / scale).max()
        assert worst <= loose, f"{name}: kink row off by {worst:.3e} > {loose}"
    return float((err[ok_rows] / scale).max()) if ok_rows.any() else 0.0


def unpack_mask1(m1, H):
    """[B, Hp/32] int32 words (bit n of row b = h1[b,n] > 0) -> bool [B,H]."""
    w = np.asarray(m1.cpu().numpy() if hasattr(m1, "cpu") else m1).astype(np.uint32)
    bits = (w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1
    return bits.reshape(w.shape[0], -1)[:, :H].astype(bool)


def check_decode_with_masks(psi, xhat, m1, m2, z, p64, mode, kappa, psi_rtol, x_rtol, h_rtol, name=""):
    """Rigorous decode check for kernels that return their LeakyReLU masks (see the block comment above)."""
    from oracle import icnn_oracle as io
    z64 = np.asarray(z, dtype=np.float64)
    H = p64["A0w"].shape[0]
    rpsi, rx, aux = io.icnn_brenier(z64, p64, mode, kappa)
    mask1 = unpack_mask1(m1, H)
    mask2 = np.asarray(m2.cpu().numpy() if hasattr(m2, "cpu") else m2).astype(bool)
    d1 = mask1 != aux["mask1"]
    d2 = mask2 != aux["mask2"]
    lim1, lim2 = h_rtol * np.abs(aux["h1"]).max(), h_rtol * max(np.abs(aux["h2"]).max(), 1e-300)
    assert (np.abs(aux["h1"][d1]) < lim1).all(), (f"{name}: {d1.sum()} mask1 flips, largest |h1| "
                                                  f"{np.abs(aux['h1'][d1]).max():.3e} vs window {lim1:.3e}")
    assert (np.abs(aux["h2"][d2]) < lim2).all(), f"{name}: mask2 flip outside the rounding window"
    _, rx_m, _ = io.icnn_brenier(z64, p64, mode, kappa, masks=(mask1, mask2))
    e_psi = close_report(psi.cpu().numpy() if hasattr(psi, "cpu") else psi, rpsi, psi_rtol, name + " psi")
    e_x = close_report(xhat.cpu().numpy() if hasattr(xhat, "cpu") else xhat, rx_m, x_rtol, name + " xhat (kernel's masks)")
    return e_psi, e_x, int(d1.sum()) + int(d2.sum())
