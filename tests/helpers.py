"""Shared helpers for the parity tests (the oracle is the checker, never the product)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import easydgl_oracle as O  # noqa: E402
from easydgl_b200 import synth  # noqa: E402

# small configurations covering both models, several head dims, E, blocks and ragged L
SMALL = {
    "easy_a": dict(model="EasyDGL", num_units=32, seqslen=12, num_items=200, num_heads=4, num_blocks=2, num_events=8),
    "easy_b": dict(model="EasyDGL", num_units=64, seqslen=30, num_items=500, num_heads=2, num_blocks=1, num_events=16),
    "easy_c": dict(model="EasyDGL", num_units=64, seqslen=36, num_items=333, num_heads=4, num_blocks=1, num_events=4),
    # dh = 16, E = 16 at a short L: the scaled 3xFP16 attention (NT = 4) and QKVT GEMM (K = 80) - the smoke() config
    "easy_d": dict(model="EasyDGL", num_units=64, seqslen=30, num_items=500, num_heads=4, num_blocks=1, num_events=16),
    "ctsma_a": dict(model="CTSMA", num_units=32, seqslen=13, num_items=200, num_heads=4, num_blocks=2, num_events=8),
    "ctsma_b": dict(model="CTSMA", num_units=64, seqslen=30, num_items=500, num_heads=4, num_blocks=2, num_events=16),
}


def case(name, batch=6, mode="parity", seed=synth.SEED, onehot=False, edge=True, **over):
    kw = dict(SMALL[name]) if name in SMALL else dict(synth.CONFIGS[name])
    kw.pop("batch", None)
    kw.update(over)
    cfg = synth.make_config(**kw)
    inp = synth.make_inputs(cfg, batch, seed=seed, edge_cases=edge)
    W = synth.make_weights(cfg, seed=seed, mode=mode, onehot_marks=onehot)
    return cfg, inp, W


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (b = reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, tol=1e-3, what=""):
    """SURVEY 8c tolerance: max|a-b| <= tol*max|b|  and  allclose(rtol=tol, atol=tol*rms(b))."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a).all(), what + ": non-finite values"
    e = float((a - b).abs().max())
    m = float(b.abs().max())
    rms = float(b.pow(2).mean().sqrt())
    assert e <= tol * max(m, 1e-30), "%s: max abs err %.3e > %.0e * max|ref| %.3e" % (what, e, tol, m)
    assert torch.allclose(a, b, rtol=tol, atol=tol * rms), "%s: allclose(rtol=%g, atol=%g*rms) failed" % (what, tol, tol)
    return e / max(m, 1e-30)
