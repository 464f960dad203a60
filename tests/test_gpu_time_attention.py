"""Ti / Tf / TgMultiHeadAttention (temporal.py:15-264) through the facade classes and the C ABI
(edgl_time_attention, edgl_row_nonzero, edgl_layernorm_last) against the literal fp64 restatement in the oracle.
Tolerance: max|a-b| <= 1e-5 max|ref| (exact-fp32 kernels; the oracle materialises the pairwise code tensors)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import O, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _inputs(N, T, C, seed, pad=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, T, C, generator=g)
    if pad:  # left padding: all-zero key rows (seqs_outs = seqs_units * seqs_masks, TiSASREC.py:66-68)
        for b in range(N):
            x[b, :int(torch.randint(0, T // 2, (1,), generator=g))] = 0.0
        x[0] = 0.0  # a sequence with every key masked: uniform attention
    ts = torch.cumsum(torch.rand(N, T + 1, generator=g) * 3.0, dim=1)
    spans = (ts[:, 1:].unsqueeze(2) - ts[:, :-1].unsqueeze(1)).clamp_min(0.0)   # TGAT.py:51-54
    return x, spans, g


def _dense_w(cq, ck, C, g, extra=False):
    def glorot(i, o):
        lim = float(np.sqrt(6.0 / (i + o)))
        return (torch.rand(i, o, generator=g) * 2 - 1) * lim
    w = {"q_w": glorot(cq, C), "q_b": 0.1 * torch.randn(C, generator=g), "k_w": glorot(ck, C),
         "k_b": 0.1 * torch.randn(C, generator=g), "v_w": glorot(ck, C), "v_b": 0.1 * torch.randn(C, generator=g)}
    if extra:
        w.update({"o_w": glorot(C, 2 * C), "o_b": 0.1 * torch.randn(2 * C, generator=g),
                  "ln_g": 1.0 + 0.1 * torch.randn(2 * C, generator=g), "ln_b": 0.1 * torch.randn(2 * C, generator=g)})
    return w


def _f64(d):
    return {k: v.double() for k, v in d.items()}


@pytest.mark.parametrize("N,T,C,h,causal", [(3, 31, 64, 4, True), (2, 100, 128, 8, True), (2, 50, 64, 2, False)])
def test_ti_multi_head_attention(N, T, C, h, causal):
    from easydgl_b200.module import coding as Cm, temporal as Tm
    keys, spans, g = _inputs(N, T, C, 5)
    queries = torch.randn(N, T, C, generator=g)
    queries[1, 3] = 0.0  # an all-zero query row: query masking (temporal.py:87-90)
    timelen = 40
    iv = spans.round().clamp(0, timelen).to(torch.int64)                          # TiSASREC.py:59
    tabs = [0.3 * torch.randn(timelen + 1 if i >= 2 else T, C, generator=g) for i in range(4)]
    w = _dense_w(C, C, C, g)
    pk, pv = (Cm.PositionCoding(T, C, initializer=t, device=DEV) for t in tabs[:2])
    tk, tv = (Cm.TimeIntervalCoding(timelen + 1, C, initializer=t, device=DEV) for t in tabs[2:])
    layer = Tm.TiMultiHeadAttention(C, h, 0.0, 0.0, pk, pv, tk, tv, weights=w, device=DEV)
    out = layer(queries.to(DEV), keys.to(DEV), iv.to(DEV), False, causal)
    ref = O.ti_attention(queries.double(), keys.double(), iv, _f64(w), tabs[0].double(), tabs[1].double(),
                         tabs[2].double(), tabs[3].double(), h, causal)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < TOL, rel_err(out, ref)


@pytest.mark.parametrize("N,T,C,h,causal", [(3, 31, 64, 4, True), (2, 100, 128, 8, True), (2, 50, 32, 1, False)])
def test_tf_multi_head_attention(N, T, C, h, causal):
    from easydgl_b200.module import coding as Cm, temporal as Tm
    keys, spans, g = _inputs(N, T, C, 7)
    queries = torch.randn(N, T, C, generator=g)
    ptab = 0.3 * torch.randn(T, C, generator=g)
    freq = torch.from_numpy(np.linspace(0, 9, C).astype(np.float32))              # coding.py:108
    phase = 0.5 * torch.randn(C, generator=g)
    w = _dense_w(C, C, C, g)
    pk = Cm.PositionCoding(T, C, initializer=ptab, device=DEV)
    tc = Cm.TimeFunctionCoding(C, device=DEV, basis_freq=freq, phase=phase)
    layer = Tm.TfMultiHeadAttention(C, h, 0.0, 0.0, pk, tc, weights=w, device=DEV)
    out = layer(queries.to(DEV), keys.to(DEV), spans.to(DEV), False, causal)
    ref = O.tf_attention(queries.double(), keys.double(), spans, _f64(w), ptab.double(), freq, phase, h, causal)
    assert rel_err(out, ref) < TOL, rel_err(out, ref)


@pytest.mark.parametrize("N,T,C,h,causal", [(3, 31, 64, 4, True), (2, 60, 32, 2, True), (2, 40, 64, 1, False)])
def test_tg_multi_head_attention(N, T, C, h, causal):
    from easydgl_b200.module import coding as Cm, temporal as Tm
    keys, _, g = _inputs(N, T, C, 9, pad=False)
    ts = torch.cumsum(torch.rand(N, T, generator=g) * 3.0, dim=1)
    spans = (ts.unsqueeze(2) - ts.unsqueeze(1)).clamp_min(0.0)                     # TGREC.py:44-47
    ids_ok = torch.ones(N, T)
    ids_ok[1, :7] = 0.0
    masks = ids_ok.unsqueeze(1).repeat(1, T, 1)                                    # TGREC.py:51-52
    freq = torch.from_numpy(np.linspace(0, 9, C).astype(np.float32))
    phase = 0.5 * torch.randn(C, generator=g)
    w = _dense_w(2 * C, 2 * C, C, g, extra=True)
    tc = Cm.TimeFunctionCoding(C, device=DEV, basis_freq=freq, phase=phase)
    layer = Tm.TgMultiHeadAttention(C, h, 0.0, 0.0, tc, weights=w, device=DEV)
    out = layer(keys.to(DEV), keys.to(DEV), masks.to(DEV), spans.to(DEV), False, causal)
    ref = O.tg_attention(keys.double(), keys.double(), masks.double(), spans, _f64(w), freq, phase, h, causal)
    assert out.shape == (N, T, 2 * C)
    assert rel_err(out, ref) < TOL, rel_err(out, ref)


def test_row_nonzero_and_layernorm_last():
    from easydgl_b200 import engine as E
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 96, generator=g)
    x[5] = 0.0
    x[11, 1:] = 0.0
    m = E.row_nonzero(x.to(DEV)).cpu()
    assert torch.equal(m, (x.abs().sum(-1) != 0).to(torch.uint8))
    gam, bet = 1.0 + 0.1 * torch.randn(96, generator=g), 0.1 * torch.randn(96, generator=g)
    y = E.layernorm_last(x.to(DEV), gam.to(DEV), bet.to(DEV), 1e-8)
    ref = O.layernorm_last(x.double(), gam.double(), bet.double(), 1e-8)
    assert rel_err(y, ref) < 1e-5
