"""Arithmetic of the 3xFP16 mode, restated in numpy (no GPU): the operand split hi = rn16(x), lo = rn16(x - hi) of
csrc/icnn_tc3.cu (split_f16), the power-of-two row / tensor scales, and the three-product contraction
a_lo.b_hi + a_hi.b_lo + a_hi.b_hi.  numpy's float16 conversion is round-to-nearest-even like cvt.rn.f16.f32, an fp16 x fp16
product is exact in fp32 -- so with the accumulation done in float64 what remains is exactly the OPERAND error budget of the
mode (the tensor core's own fp32 accumulation is measured on the GPU, tests/test_icnn_tc_gpu.py)."""
import numpy as np


def split16(x):
    x = np.asarray(x, np.float32)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def pow2_scale_to(bound, top_exp):
    """2^(top_exp - floor(log2 bound)): bound * scale in [2^top_exp, 2^(top_exp+1))."""
    e = np.floor(np.log2(np.asarray(bound, np.float64)))
    return np.exp2(top_exp - e)


def test_split_carries_22_bits_in_range_and_an_absolute_floor_below():
    rng = np.random.default_rng(0)
    x = (rng.uniform(1.0, 2.0, 200000) * np.exp2(rng.integers(-24, 15, 200000))).astype(np.float32)   # up to 2^15: no overflow
    hi, lo = split16(x)
    assert np.isfinite(hi.astype(np.float32)).all()
    err = np.abs(x.astype(np.float64) - (hi.astype(np.float64) + lo.astype(np.float64)))
    # two 11-bit pieces: relative 2^-22 where lo is a normal fp16 (|x| >= 2^-3); fp16's subnormal spacing 2^-24 below it
    assert (err <= np.maximum(np.exp2(-22.0) * np.abs(x), np.exp2(-25.0)) * (1 + 1e-6)).all()
    big = np.abs(x) >= 0.125
    assert (err[big] <= np.exp2(-22.0) * np.abs(x[big])).all()


def test_three_product_contraction_is_fp32_grade_after_power_of_two_scaling():
    """h1 = sum_k x1[m,k] P[n,k] the way GEMM1 forms it: x1 = leaky(A0 z + b0)^2 scaled per ROW through h0 * t (bound from
    max|A0|, |z|), P scaled per TENSOR from its maximum; products of fp16 pairs, unscaled at the end.  Inputs over many
    decades (|z| 1e-3 .. 3e2, exp(W) over e^{+-4})."""
    rng = np.random.default_rng(1)
    M, N, K, d = 64, 48, 512, 2
    A0 = rng.normal(0, 0.7, (K, d)).astype(np.float32)
    b0 = rng.normal(0, 0.5, K).astype(np.float32)
    P = np.exp(rng.normal(np.log(1.0 / K), 2.0, (N, K))).astype(np.float32)
    s1 = pow2_scale_to(P.max(), 14)
    Phi, Plo = split16((P.astype(np.float64) * s1).astype(np.float32))
    assert np.isfinite(Phi.astype(np.float32)).all()
    worst = 0.0
    for zscale in (1e-3, 1.0, 3e2):
        z = rng.normal(0, zscale, (M, d)).astype(np.float32)
        h0 = (z.astype(np.float64) @ A0.T.astype(np.float64) + b0).astype(np.float32)
        bound = np.abs(z) @ np.abs(A0).max(0) + np.abs(b0).max()                       # h0_bound of the kernel
        assert (np.abs(h0).max(1) <= bound * (1 + 1e-6)).all()
        t = pow2_scale_to(bound, 6)[:, None]                                          # bound * t in [2^6, 2^7): x1 t^2 < 2^14
        a0 = np.maximum(h0 * t, 0.2 * (h0 * t)).astype(np.float32)
        x1s = (a0 * a0).astype(np.float32)
        assert x1s.max() < 2.0 ** 14
        ahi, alo = split16(x1s)
        f = lambda t_: t_.astype(np.float64)
        acc = f(alo) @ f(Phi).T + f(ahi) @ f(Plo).T + f(ahi) @ f(Phi).T
        got = acc / (t.astype(np.float64) ** 2 * s1)
        x1 = np.maximum(h0, 0.2 * h0).astype(np.float32)
        want = f((x1 * x1).astype(np.float32)) @ f(P).T
        worst = max(worst, float(np.abs(got - want).max() / np.abs(want).max()))
    assert worst < 1e-6, worst                                  # operand budget: well inside the 1e-5 bound of psi / xhat


def test_slope_pattern_and_its_sample_scale_are_exact_in_fp16():
    for v in (1.0, 5.0):
        assert float(np.float16(v)) == v
