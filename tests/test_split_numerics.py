"""CPU checks of the arithmetic the tensor-core kernels rely on (numpy emulation; no GPU, no oracle import).

* the scaled 3xFP16 split of attn_f16.cu / gemm_f16.cu: x*s = hi + lo with hi = the top 11 significant bits (exact in
  fp16), lo = x*s - hi rounded to fp16, s the power of two that puts the operand's maximum into [2^14, 2^15);
  D = (A_lo B_hi + A_hi B_lo + A_hi B_hi) / (sa sb) must be as accurate as the 3xTF32 split and as plain fp32;
* the branch-free erf of the GELU epilogues (gelu_fit in tc_common.cuh, used by gemm_tc.cu / gemm_f16.cu): coefficients are read from the
  CUDA sources, so an accidental edit of a constant fails here.
"""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "easydgl_b200", "csrc")


def pow2_scale(m):
    """attn_f16.cu / gemm_f16.cu pow2_scale: exponent field clamped to [15, 239]."""
    e = int((np.float32(m).view(np.uint32) >> 23) & 0xFF)
    e = min(max(e, 15), 239)
    return np.float32(2.0) ** (14 - (e - 127))


def split16(x, scale):
    x = (x * scale).astype(np.float32)
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (x - hi).astype(np.float32)
    return hi.astype(np.float16).astype(np.float64), lo.astype(np.float16).astype(np.float64)


def split_tf32(x):
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (x - hi).astype(np.float32)
    lo = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi.astype(np.float64), lo.astype(np.float64)


def products(A, B, per_row):
    exact = A.astype(np.float64) @ B.astype(np.float64)
    den = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
    if per_row:
        sa = np.array([pow2_scale(np.abs(r).max()) for r in A], dtype=np.float32)[:, None]
    else:
        sa = np.full((A.shape[0], 1), pow2_scale(np.abs(A).max()), dtype=np.float32)
    sb = pow2_scale(np.abs(B).max())
    ah, al = split16(A, sa)
    bh, bl = split16(B, sb)
    assert np.isfinite(ah).all() and np.isfinite(bh).all() and np.abs(ah).max() < 65504
    f16 = (al @ bh + ah @ bl + ah @ bh) / (sa.astype(np.float64) * float(sb))
    th, tl = split_tf32(A)
    uh, ul = split_tf32(B)
    tf = tl @ uh + th @ ul + th @ uh
    f32 = (A @ B).astype(np.float64)
    err = lambda c: float(np.max(np.abs(c - exact) / den))
    return err(f16), err(tf), err(f32)


def test_scaled_fp16_split_matches_tf32_split_and_fp32():
    rng = np.random.default_rng(3)
    cases = {
        "normal": (rng.standard_normal((128, 104)), rng.standard_normal((104, 16))),
        "probabilities": (None, rng.standard_normal((104, 16)) * 0.05),
        "wide range": (rng.standard_normal((128, 104)) * np.exp(rng.uniform(-12, 0, (128, 104))),
                       rng.standard_normal((104, 16)) * np.exp(rng.uniform(-12, 0, (104, 16)))),
        "tiny": (rng.standard_normal((64, 144)) * 1e-20, rng.standard_normal((144, 32)) * 1e-12),
        "huge": (rng.standard_normal((64, 144)) * 1e18, rng.standard_normal((144, 32)) * 1e15),
    }
    for name, (A, B) in cases.items():
        if A is None:
            s = rng.standard_normal((128, 104)) * 3
            p = np.exp(s - s.max(1, keepdims=True))
            A = p / p.sum(1, keepdims=True)
        A, B = A.astype(np.float32), B.astype(np.float32)
        for per_row in (True, False):
            e16, etf, e32 = products(A, B, per_row)
            # relative to sum |a||b|: fp32-level, and no worse than 2x the 3xTF32 split
            assert e16 < 2e-6, (name, per_row, e16)
            assert e16 < 2.0 * max(etf, e32) + 1e-7, (name, per_row, e16, etf, e32)


def test_per_tensor_scale_keeps_small_rows_accurate_to_the_tensor_maximum():
    """gemm_f16.cu scales A per TENSOR: a row 2^-16 below the maximum keeps full relative precision; the floor is
    absolute (2^-25 after scaling), i.e. relative to max|A| * |B| the error stays at the fp32 level."""
    rng = np.random.default_rng(4)
    A = rng.standard_normal((32, 128)).astype(np.float32)
    A[1] *= np.float32(2.0 ** -16)
    A[2] *= np.float32(2.0 ** -30)
    B = rng.standard_normal((128, 64)).astype(np.float32)
    sa, sb = pow2_scale(np.abs(A).max()), pow2_scale(np.abs(B).max())
    ah, al = split16(A, sa)
    bh, bl = split16(B, sb)
    got = (al @ bh + ah @ bl + ah @ bh) / (float(sa) * float(sb))
    exact = A.astype(np.float64) @ B.astype(np.float64)
    row_rel = np.abs(got - exact).max(1) / np.abs(exact).max(1)
    assert row_rel[0] < 2e-6 and row_rel[1] < 2e-6        # within 2^18 of the maximum: full relative precision
    tensor_rel = np.abs(got - exact).max() / np.abs(exact).max()
    assert tensor_rel < 2e-6                               # every row: fp32-level relative to the tensor's scale
    assert np.abs(got[2] - exact[2]).max() < 1e-9 * np.abs(exact).max()


def _gelu_fit_coefficients(path):
    src = open(path).read()
    body = src[src.index("float gelu_fit(float x)"):]
    body = body[:body.index("return")]
    first = re.search(r"float p = ([-0-9.e+]+)f;", body).group(1)
    rest = re.findall(r"p = fmaf\(p, t, ([-0-9.e+]+)f\);", body)
    return [float(first)] + [float(c) for c in rest]  # highest degree first


def test_gelu_fit_constants_and_accuracy():
    from scipy.special import erf
    f32 = np.float32
    coefs = _gelu_fit_coefficients(os.path.join(CSRC, "tc_common.cuh"))  # shared by gemm_tc.cu and gemm_f16.cu
    assert len(coefs) == 8, coefs
    x = np.linspace(-9, 9, 1000001).astype(f32)
    t = np.minimum((np.abs(x) * f32(0.70710678118654752440)).astype(f32), f32(4.0))
    p = np.full_like(t, f32(coefs[0]))
    for c in coefs[1:]:
        p = (p * t + f32(c)).astype(f32)
    e = np.exp2((-(t * p)).astype(f32).astype(np.float64)).astype(f32)
    erf_fit = (f32(1.0) - e).astype(f32)
    assert np.max(np.abs(erf_fit - erf(t.astype(np.float64)))) < 1.3e-7
    gelu = (np.abs(x) * (e * f32(-0.5) + f32(0.5)).astype(f32) + (f32(0.5) * x).astype(f32)).astype(f32)
    ref = 0.5 * x.astype(np.float64) * (1.0 + erf(x.astype(np.float64) / math.sqrt(2.0)))
    assert np.max(np.abs(gelu - ref) / np.maximum(np.abs(x), 1e-3)) < 1.5e-7
