"""tcgen05 (TF32 / 3xTF32 / 3xFP16) variants of the fused ICNN forward and backward.

Bounds (vs the fp64 oracle on fp32-rounded inputs, |a-b| <= rtol*|b| + rtol*max|b|, NO fraction of elements exempted):
    tf32x3 : psi 1e-5, xhat 1e-5, parameter gradients / dz 1e-4  -- north_star's FP32 bounds.  fp32-grade operand split
             (3 MMAs per product) + K accumulated in chunks of 512 with round-to-nearest adds between chunks
             (icnn_tc3.cu; the tensor core's own accumulation truncates).
    f16x3  : the same bounds.  Operands split into FP16 hi/lo pairs (22 mantissa bits like a tf32 pair), scaled by exact
             powers of two into the fp16 range (per tensor / per sample row), three kind::f16 MMAs per K step.
    tf32   : psi 2e-4, xhat 5e-3, gradients 5e-3 (one MMA, operands rounded to 11 bits): the stated looser bound.
Kink handling (helpers.py): where the kernels return their LeakyReLU masks, every mask bit that differs from the oracle's
must belong to a unit whose pre-activation is inside the mode's rounding window (|h| < H_RTOL * max|h|) and xhat is then
compared STRICTLY with the oracle evaluated with the kernel's masks; where only (psi, xhat) are visible, the rows the
oracle flags as kink-adjacent are held to the loose bound and every other row to the strict one."""
import numpy as np
import pytest
import torch

from oracle import icnn_oracle as io

from helpers import (H_RTOL, KEYS, check_decode_with_masks, close_report, close_rows, f32_as_f64, kink_rows, params_f32_as_f64,
                     params_to_torch)

pytestmark = pytest.mark.gpu
BOUNDS = {4: (1e-5, 1e-5), 3: (1e-5, 1e-5), 1: (2e-4, 5e-3)}          # precision -> (psi, xhat)
GRAD_BOUND = {4: 1e-4, 3: 1e-4, 1: 5e-3}
PRECS, PREC_IDS = [4, 3, 1], ["f16x3", "tf32x3", "tf32"]
# (d, H, B, regime, weight mode): ragged sizes, both weight reparameterisations (exp / clamp), and the two ICNN shapes of
# BASELINE configs[1] (ICNN(2,512), ICNN(2,1024)) in the trained-like AND the default-init regime
CASES = [(2, 256, 256, "mixed", 0), (2, 96, 77, "mixed", 0), (1, 64, 300, "mixed", 0), (3, 512, 1000, "mixed", 0),
         (2, 1024, 2048, "mixed", 0), (2, 512, 513, "default", 0), (2, 1024, 512, "default", 0),
         (2, 256, 300, "clampy", 1), (3, 512, 200, "clampy", 1), (2, 1024, 256, "default", 1)]


@pytest.mark.parametrize("prec", PRECS, ids=PREC_IDS)
@pytest.mark.parametrize("case", CASES, ids=[f"d{c[0]}_h{c[1]}_b{c[2]}_{c[3]}_mode{c[4]}" for c in CASES])
def test_tc_forward_vs_oracle(case, prec):
    from vae_song_b200 import ops
    d, H, B, regime, mode = case
    rng = np.random.default_rng(H + B)
    p = io.random_params(rng, d, H, np.float64, regime)
    z = rng.normal(0, 1, (B, d))
    zt = torch.tensor(z, dtype=torch.float32, device="cuda")
    P = params_to_torch(p)
    ws = ops.icnn_prepare(P, d, H, mode, prec, B, False)
    psi, xhat, m1, m2 = ops.icnn_decode_fwd(zt, ws, d, H, mode, 0.15, prec, True, True, True)
    rp, rx = BOUNDS[prec]
    e_psi, e_x, flips = check_decode_with_masks(psi, xhat, m1, m2, f32_as_f64(z), params_f32_as_f64(p), mode, 0.15, rp, rx,
                                                H_RTOL[prec], f"{case}")
    print(f"prec {prec} {case}: psi {e_psi:.2e} xhat {e_x:.2e} mask flips {flips}/{B * H}")
    # the autograd.Function (what the modules call) returns the same rows
    psi2, xhat2 = ops.IcnnBrenierFn.apply(zt, 0.15, mode, prec, *P)
    assert torch.equal(psi2, psi) and torch.equal(xhat2, xhat)


@pytest.mark.parametrize("prec", PRECS, ids=PREC_IDS)
def test_tc_matches_fp32_path_and_backward_runs(prec):
    """Same module, precision switched: outputs agree with the FP32 kernels within the stated bound, and the
    training backward (masks from the tensor-core forward) matches the oracle evaluated with those masks'
    semantics (FP32 bound on the gradients)."""
    from vae_song_b200 import ops
    rng = np.random.default_rng(77)
    d, H, B = 2, 512, 700
    p = io.random_params(rng, d, H, np.float64, "mixed")
    z, v = rng.normal(0, 1, (B, d)), rng.normal(0, 1, (B, d))
    outs = {}
    for pr in (0, prec):
        ps = [t.requires_grad_(True) for t in params_to_torch(p)]
        zt = torch.tensor(z, dtype=torch.float32, device="cuda", requires_grad=True)
        psi, xhat = ops.IcnnBrenierFn.apply(zt, 0.1, 0, pr, *ps)
        (xhat * torch.tensor(v, dtype=torch.float32, device="cuda")).sum().backward()
        outs[pr] = (psi.detach().cpu().numpy(), xhat.detach().cpu().numpy(), zt.grad.cpu().numpy(),
                    {k: t.grad.cpu().numpy() for k, t in zip(KEYS, ps)})
    rp, rx = BOUNDS[prec]
    _, _, aux = io.icnn_brenier(f32_as_f64(z), params_f32_as_f64(p), 0, 0.1, keep=True)
    close_report(outs[prec][0], outs[0][0], rp, "psi vs fp32 path")
    close_rows(outs[prec][1], outs[0][1], rx, "xhat vs fp32 path", kink_rows(aux, H_RTOL[prec]))
    gtol = GRAD_BOUND[prec]                # tensor-core backward (rows + dP0 kernels) in the same precision
    close_rows(outs[prec][2], outs[0][2], gtol, "dz vs fp32 path", kink_rows(aux, H_RTOL[prec], with_h0=True), loose=5e-2)
    for k in ("A0w", "A0b", "A1w", "A2w", "W0", "W1"):
        close_report(outs[prec][3][k], outs[0][3][k], gtol, "grad " + k)
    assert float(np.abs(outs[prec][3]["A1b"]).max()) == 0.0 and float(np.abs(outs[prec][3]["A2b"]).max()) == 0.0


def test_tc_full_size_tiling_invariance_and_sampled_rows_vs_oracle():
    """BASELINE decode size B = 65536, both ICNN widths: (a) a row decoded alone equals the row decoded inside the big
    batch, bit for bit (persistent scheduling, chunked accumulation and the ordered partial sums are tile independent);
    (b) every 256th row (256 rows) against the fp64 oracle at the mode's bound."""
    from vae_song_b200 import ops
    rng = np.random.default_rng(3)
    for H in (512, 1024):
        p = io.random_params(rng, 2, H, np.float64, "mixed")
        params = params_to_torch(p)
        z = torch.tensor(rng.normal(0, 1, (65536, 2)), dtype=torch.float32, device="cuda")
        for prec in (1, 3, 4):
            psi, xhat = ops.IcnnBrenierFn.apply(z, 0.1, 0, prec, *params)
            sel = torch.arange(0, 65536, 256, device="cuda")
            psi_s, xhat_s = ops.IcnnBrenierFn.apply(z[sel].contiguous(), 0.1, 0, prec, *params)
            assert torch.equal(psi[sel], psi_s) and torch.equal(xhat[sel], xhat_s)     # row results are tile independent
            assert torch.isfinite(xhat).all()
            rpsi, rx, aux = io.icnn_brenier(z[sel].double().cpu().numpy(), params_f32_as_f64(p), 0, 0.1)
            rp, rxb = BOUNDS[prec]
            close_report(psi_s.cpu().numpy(), rpsi, rp, f"psi H={H} prec={prec}")
            close_rows(xhat_s.cpu().numpy(), rx, rxb, f"xhat H={H} prec={prec}", kink_rows(aux, H_RTOL[prec]))


def test_tc_unsupported_is_loud():
    from vae_song_b200 import _C, ops
    rng = np.random.default_rng(1)
    p = io.random_params(rng, 4, 64, np.float64, "mixed")
    with pytest.raises(_C.B200VaeError):
        ops.IcnnBrenierFn.apply(torch.zeros(8, 4, device="cuda"), 0.0, 0, 1, *params_to_torch(p))   # d=4 on the TC path
    with pytest.raises(_C.B200VaeError):
        ops.IcnnBrenierFn.apply(torch.zeros(8, 2, device="cuda"), 0.0, 0, 2,
                                *params_to_torch(io.random_params(rng, 2, 64, np.float64, "mixed")))  # precision id 2 is reserved (never built)


BWD_CASES = [(1, 64, 300, 0), (2, 96, 77, 1), (3, 512, 1000, 0), (2, 1024, 600, 0), (2, 256, 256, 1)]


@pytest.mark.parametrize("prec", PRECS, ids=PREC_IDS)
@pytest.mark.parametrize("case", BWD_CASES, ids=[f"d{c[0]}_h{c[1]}_b{c[2]}_mode{c[3]}" for c in BWD_CASES])
def test_tc_backward_vs_fp32_kernels_same_masks_and_vs_oracle(case, prec):
    """Tensor-core double-backward (persistent pair rows kernel + dP0 kernel) against the FP32 SIMT kernels fed the SAME
    saved masks, and against the fp64 oracle: ragged batch, H not a multiple of 256 (padded units), d = 1..3, exp and
    clamp weights, tf32x3 at north_star's FP32 gradient bound (1e-4), tf32 at its stated 5e-3."""
    from vae_song_b200 import ops
    d, H, B, mode = case
    rng = np.random.default_rng(11 * H + B)
    p = io.random_params(rng, d, H, np.float64, "mixed")
    P = params_to_torch(p)
    z = torch.tensor(rng.normal(0, 1, (B, d)), dtype=torch.float32, device="cuda")
    v = torch.tensor(rng.normal(0, 1, (B, d)), dtype=torch.float32, device="cuda")
    ws = ops.icnn_prepare(P, d, H, mode, prec, B, True)
    _, _, m1, m2 = ops.icnn_decode_fwd(z, ws, d, H, mode, 0.1, prec, True, True, True)
    dz, g = ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws, d, H, mode, 0.1, prec)
    ws0 = ops.icnn_prepare(P, d, H, mode, 0, B, True)
    dz0, g0 = ops.icnn_decode_bwd(z, v, None, m1, m2, P, ws0, d, H, mode, 0.1, 0)
    gtol = GRAD_BOUND[prec]
    close_report(dz.cpu().numpy(), dz0.cpu().numpy(), gtol, "dz")
    for k, a, b in zip(KEYS, g, g0):
        if float(b.abs().max()) == 0.0:
            assert float(a.abs().max()) == 0.0, k
        else:
            close_report(a.cpu().numpy(), b.cpu().numpy(), gtol, "grad " + k)
    # ... and against the fp64 oracle evaluated with the masks this forward saved (given the masks the backward is linear
    # in everything else, so NO element is exempted); the masks themselves are checked by test_tc_forward_vs_oracle
    from helpers import unpack_mask1
    p64, z64, v64 = params_f32_as_f64(p), z.double().cpu().numpy(), v.double().cpu().numpy()
    km = (unpack_mask1(m1, H), m2.cpu().numpy().astype(bool))
    _, _, aux = io.icnn_brenier(z64, p64, mode, 0.1, keep=True)
    rdz, rg = io.icnn_brenier_backward(z64, v64, p64, mode, 0.1, None, masks=km)
    h0_kink = (np.abs(aux["h0"]) < H_RTOL[0] * np.abs(aux["h0"]).max()).any(1)      # first-layer kinks (h0 is FP32 FMA work)
    close_rows(dz.cpu().numpy(), rdz, gtol, "dz vs oracle", h0_kink, loose=5e-2)
    for k, a in zip(KEYS, g):
        if np.abs(rg[k]).max() > 0:
            close_report(a.cpu().numpy(), rg[k], gtol, "grad vs oracle " + k)


def test_tc_forward_workspace_reuse_and_mask_paths():
    """One prepared workspace, several launches: the per-tile unit counters reset themselves; psi-only launches
    (one GEMM), launches with caller-owned masks and launches with the internal mask scratch give identical rows."""
    from vae_song_b200 import ops
    rng = np.random.default_rng(9)
    d, H, B = 2, 512, 1111
    P = params_to_torch(io.random_params(rng, d, H, np.float64, "mixed"))
    z = torch.tensor(rng.normal(0, 1, (B, d)), dtype=torch.float32, device="cuda")
    for prec in (1, 3, 4):
        ws = ops.icnn_prepare(P, d, H, 0, prec, B, True)
        psi_a, xhat_a, m1, m2 = ops.icnn_decode_fwd(z, ws, d, H, 0, 0.1, prec, True, True, True)
        psi_b, xhat_b, _, _ = ops.icnn_decode_fwd(z, ws, d, H, 0, 0.1, prec, True, True, False)
        psi_c, _, _, _ = ops.icnn_decode_fwd(z, ws, d, H, 0, 0.1, prec, True, False, False)
        psi_d, xhat_d, m1d, m2d = ops.icnn_decode_fwd(z, ws, d, H, 0, 0.1, prec, True, True, True)
        assert torch.equal(psi_a, psi_b) and torch.equal(psi_a, psi_c) and torch.equal(psi_a, psi_d)
        assert torch.equal(xhat_a, xhat_b) and torch.equal(xhat_a, xhat_d)
        assert torch.equal(m1, m1d) and torch.equal(m2, m2d)
        assert torch.equal(m2.bool(), psi_a > 0)


@pytest.mark.parametrize("prec", [3, 4], ids=["tf32x3", "f16x3"])
@pytest.mark.parametrize("shape", [(2, 512, 700), (3, 1024, 300), (1, 256, 1000), (2, 1024, 65536)],
                         ids=["d2_h512", "d3_h1024", "d1_h256", "d2_h1024_b65536"])
def test_saved_accumulator_backward_equals_recompute(shape, prec):
    """3xTF32 / 3xFP16 training: the pair forward keeps its GEMM2 accumulators in the for_backward workspace and the backward reads
    them back (icnn_tc3.cu SV kernels) instead of redoing that GEMM.  Re-running prepare on the workspace invalidates the
    save (host bookkeeping, api.cu), so the same backward then RECOMPUTES.  The saved accumulators come from the forward's
    K-chunked accumulation (chunks of 256 added with round-to-nearest FP32 adds, icnn_tc3.cu), the recomputed ones from one
    plain accumulator: the two backward results agree to the accumulation error of the plain path (<= 2e-5 of the largest
    entry; they were bit-identical before the forward accumulated in chunks) and the decode itself is unchanged."""
    from vae_song_b200 import _C, ops
    import ctypes as C
    d, H, B = shape
    rng = np.random.default_rng(5 + H)
    p = io.random_params(rng, d, H, np.float64, "mixed")
    params = params_to_torch(p)
    z = torch.tensor(rng.normal(0, 1, (B, d)), dtype=torch.float32, device="cuda")
    v = torch.tensor(rng.normal(0, 1, (B, d)), dtype=torch.float32, device="cuda")
    res = []
    for invalidate in (False, True):
        ws = ops.icnn_prepare(params, d, H, 0, prec, B, True)
        psi, xhat, m1, m2 = ops.icnn_decode_fwd(z, ws, d, H, 0, 0.1, prec, True, True, True)
        if invalidate:      # same parameters, so nothing changes numerically -- but the saved accumulators are forgotten
            ps = ops._params_struct(params)
            _C.check(_C.load().b200vae_icnn_prepare(C.byref(ps), d, H, 0, prec, ops._ptr(ws), ws.numel(), ops._stream()), "prepare")
        dz, grads = ops.icnn_decode_bwd(z, v, None, m1, m2, params, ws, d, H, 0, 0.1, prec)
        torch.cuda.synchronize()
        res.append((xhat.clone(), dz.clone(), [g.clone() for g in grads]))
    assert torch.equal(res[0][0], res[1][0])
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
    assert rel(res[0][1], res[1][1]) <= 2e-5, rel(res[0][1], res[1][1])
    for k, a, b in zip(KEYS, res[0][2], res[1][2]):
        if float(b.abs().max()) > 0:
            assert rel(a, b) <= 2e-5, (k, rel(a, b))
        else:
            assert torch.equal(a, b), k
    if B > 2000:        # BASELINE batch: the agreement above is the size-independent check
        return
    # and against the fp64 oracle (3xTF32 bounds)
    z64, p64 = f32_as_f64(z.cpu().numpy()), params_f32_as_f64(p)
    _, _, aux = io.icnn_brenier(z64, p64, 0, 0.1, keep=True)
    rdz, rg = io.icnn_brenier_backward(z64, f32_as_f64(v.cpu().numpy()), p64, 0, 0.1, None)
    close_rows(res[0][1].cpu().numpy(), rdz, 1e-4, "dz", kink_rows(aux, H_RTOL[3], with_h0=True), loose=5e-2)
    for k, g in zip(KEYS, res[0][2]):
        if np.abs(rg[k]).max() > 0:
            close_report(g.cpu().numpy(), rg[k], 1e-4, "grad " + k)


@pytest.mark.parametrize("zscale,vscale", [(1e-3, 1e-6), (1.0, 1.5e-5), (300.0, 1e4), (1e-2, 1.0)],
                         ids=["tiny_z_tiny_v", "unit_z_batchmean_v", "huge_z_huge_v", "small_z_unit_v"])
def test_f16x3_is_scale_robust(zscale, vscale):
    """FP16 has a 5-bit exponent: the 3xFP16 mode scales every operand by an exact power of two (per sample row for the
    generated operands, per tensor for the prepared weights).  Inputs and cotangents far from 1 -- |z| ~ 1e-3 .. 3e2,
    v ~ 1e-6 (a batch-mean loss at B = 65536) .. 1e4 -- and weights spread over e^{+-4} must meet the same FP32 bounds
    relative to the size of the result."""
    from vae_song_b200 import ops
    from helpers import unpack_mask1
    d, H, B, mode, prec = 2, 512, 600, 0, 4
    rng = np.random.default_rng(int(zscale * 1e3) + 17)
    p = io.random_params(rng, d, H, np.float64, "mixed")
    p["W0"] = p["W0"] + rng.normal(0, 2.0, p["W0"].shape)          # exp(W) over ~4 decades
    P = params_to_torch(p)
    z = rng.normal(0, zscale, (B, d))
    v = rng.normal(0, vscale, (B, d))
    zt = torch.tensor(z, dtype=torch.float32, device="cuda")
    vt = torch.tensor(v, dtype=torch.float32, device="cuda")
    ws = ops.icnn_prepare(P, d, H, mode, prec, B, True)
    psi, xhat, m1, m2 = ops.icnn_decode_fwd(zt, ws, d, H, mode, 0.15, prec, True, True, True)
    assert torch.isfinite(psi).all() and torch.isfinite(xhat).all()
    p64 = params_f32_as_f64(p)
    check_decode_with_masks(psi, xhat, m1, m2, f32_as_f64(z), p64, mode, 0.15, 1e-5, 1e-5, H_RTOL[prec], f"scale {zscale}")
    dz, g = ops.icnn_decode_bwd(zt, vt, None, m1, m2, P, ws, d, H, mode, 0.15, prec)
    z64, v64 = zt.double().cpu().numpy(), vt.double().cpu().numpy()
    km = (unpack_mask1(m1, H), m2.cpu().numpy().astype(bool))
    _, _, aux = io.icnn_brenier(z64, p64, mode, 0.15, keep=True)
    rdz, rg = io.icnn_brenier_backward(z64, v64, p64, mode, 0.15, None, masks=km)
    h0_kink = (np.abs(aux["h0"]) < H_RTOL[0] * np.abs(aux["h0"]).max()).any(1)
    close_rows(dz.cpu().numpy(), rdz, 1e-4, "dz vs oracle", h0_kink, loose=5e-2)
    for k, a in zip(KEYS, g):
        assert torch.isfinite(a).all(), k
        if np.abs(rg[k]).max() > 0:
            close_report(a.cpu().numpy(), rg[k], 1e-4, "grad vs oracle " + k)


@pytest.mark.parametrize("prec", PRECS, ids=PREC_IDS)
def test_tc_hidden_width_beyond_the_tensor_core_kernels_is_loud(prec):
    """H = 1280 > 1024: the CTA-pair kernels do not hold such an ICNN (a row of mask bits must fit their shared-memory buffer,
    the operand tables too), so every tensor-core precision refuses it -- loudly, never with garbage (the pair kernels used
    to accept Hq = 1280 on the shared-memory test alone) -- while the FP32 kernels take it."""
    from vae_song_b200 import _C, ops
    d, H, B = 2, 1280, 300
    rng = np.random.default_rng(H)
    p = io.random_params(rng, d, H, np.float64, "mixed")
    P = params_to_torch(p)
    z = rng.normal(0, 1, (B, d))
    zt = torch.tensor(z, dtype=torch.float32, device="cuda")
    with pytest.raises(_C.B200VaeError):
        ws = ops.icnn_prepare(P, d, H, 0, prec, B, True)
        ops.icnn_decode_fwd(zt, ws, d, H, 0, 0.1, prec, True, True, True)
    torch.cuda.synchronize()
    ws = ops.icnn_prepare(P, d, H, 0, 0, B, True)
    psi, xhat, m1, m2 = ops.icnn_decode_fwd(zt, ws, d, H, 0, 0.1, 0, True, True, True)
    check_decode_with_masks(psi, xhat, m1, m2, f32_as_f64(z), params_f32_as_f64(p), 0, 0.1, 1e-5, 1e-5, H_RTOL[0], "H1280 fp32")
