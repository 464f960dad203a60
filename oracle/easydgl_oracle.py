"""CPU oracle for the EasyDGL / CTSMA eval forward pass.  TEST INFRASTRUCTURE ONLY.

This file is a torch-CPU *restatement* of the reference's TensorFlow graph for the
hot path.  It is the checker for the CUDA path; it is never
shipped, never imported by ``easydgl_b200`` and never the thing measured, except as
the ``cpu_baseline`` / ``--impl reference`` arm of ``bench.py``.

PARITY UNPINNED: the reference (cchao0116/EasyDGL @ 1489428) has no tests, no
golden vectors and no fixtures for this path, and its arithmetic lives in
TensorFlow 2.3.4 (``requirements.txt:4``), which is not installed here and cannot be
installed (no network, Python 3.12).  The only pin is the source text of the
reference files cited on every function below, plus an independent plain-C
restatement (``oracle/easydgl_ref.c``) that the tests cross-check against this one.

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  One code base serves three purposes through switches:

* ``dtype=torch.float64``  -> ground truth (all contractions in fp64; the time
  arguments ``ts/time_scale`` and ``ts/scale_j`` are still rounded through fp32
  exactly as the fp32 reference graph does, because ``sin`` of a ~1e4 rad argument
  amplifies that rounding to ~5e-4 -- it is part of the reference's semantics).
* ``dtype=torch.float32``  -> the "reference CPU path" (what TF-CPU would compute).
* ``literal=True``         -> materialises the two ``[hB,L,L,E]`` tensors exactly
  like ``src/module/temporal.py:309-313``; ``literal=False`` contracts them.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

MASK_FILL = float(-2 ** 32 + 1)  # src/module/temporal.py:358,425


# ----------------------------------------------------------------------------
# src/model/Base.py
# ----------------------------------------------------------------------------
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """Base.py:12-67.  begin_norm_axis=1 -> moments over ALL axes but the batch axis
    (Base.py:51-52), population variance (tf.nn.moments), eps 1e-12 (Base.py:56),
    gamma/beta on the last axis (Base.py:29-49), tf.nn.batch_normalization form
    ``x*inv + (beta - mean*inv)`` with ``inv = rsqrt(var+eps)*gamma`` (Base.py:57-63)."""
    axes = tuple(range(1, x.dim()))
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    inv = torch.rsqrt(var + 1e-12) * gamma
    return x * inv + (beta - mean * inv)


def feedforward(x: torch.Tensor, w0, b0, w1, b1) -> torch.Tensor:
    """Base.FeedForward.__call__ (Base.py:77-87): conv1x1+relu, conv1x1, residual."""
    y = torch.relu(x @ w0 + b0)
    y = y @ w1 + b1
    return y + x


def output_bias(bias: torch.Tensor) -> torch.Tensor:
    """Sequential.output_bias(inf_pad=True) (Base.py:106-110): concat([-1000], bias)."""
    return torch.cat([torch.full((1,), -1000.0, dtype=bias.dtype), bias], dim=0)


def mask_seen_logits(logits: torch.Tensor, seqs_i: torch.Tensor) -> torch.Tensor:
    """Sequential.eval (Base.py:156-163): logits[b, seqs_i[b, l]] += -inf for all l."""
    out = logits.clone()
    rows = torch.arange(logits.shape[0]).unsqueeze(1).expand_as(seqs_i)
    out[rows.reshape(-1), seqs_i.reshape(-1).long()] = float("-inf")
    return out


def topk_lower_index_first(scores: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """tf.nn.top_k semantics (Base.py:181): sorted descending, ties -> lower index.
    torch.topk does not promise the tie order, so sort stably instead."""
    order = torch.sort(-scores, dim=-1, stable=True).indices[:, :k]
    return torch.gather(scores, 1, order), order


def eval_topk(logits: torch.Tensor, seqs_i: torch.Tensor, mask_seen: bool = True, k: int = 100,
              rank_on: str = "probs") -> Tuple[torch.Tensor, torch.Tensor]:
    """Sequential.eval ranking (Base.py:150-181).  ``rank_on='probs'`` is literal
    (softmax then top_k, Base.py:164,181); ``'logits'`` ranks the masked logits
    (identical wherever softmax is injective in floating point)."""
    if mask_seen:
        logits = mask_seen_logits(logits, seqs_i)
    if rank_on == "probs":
        scores = torch.softmax(logits, dim=-1)
    else:
        scores = logits
    return topk_lower_index_first(scores, k)


def ranking_metrics(topk_idx: torch.Tensor, labels_last: torch.Tensor) -> Dict[str, float]:
    """HR@{10,50,100}, NDCG@{10,50,100} (Base.py:167-201) for one batch (the
    reference then streams the batch means through tf.metrics.mean)."""
    tp = (topk_idx == labels_last.reshape(-1, 1)).to(torch.float64)
    gain = torch.tensor(1.0 / np.log2(np.arange(2, 100 + 2)), dtype=torch.float64)
    out = {}
    for kk in (100, 50, 10):
        t = tp[:, :kk]
        out["H%d" % kk] = float(torch.sign(t.sum(-1)).mean())
        out["N%d" % kk] = float((t * gain[:kk]).sum(-1).mean())
    return out


# ----------------------------------------------------------------------------
# src/module/coding.py
# ----------------------------------------------------------------------------
def zero_pad_table(table: torch.Tensor) -> torch.Tensor:
    """Embedding.__init__ zero_pad=True (coding.py:56-57): row 0 replaced by zeros."""
    return torch.cat([torch.zeros_like(table[:1]), table[1:]], dim=0)


def embedding(table: torch.Tensor, ids: torch.Tensor, scale: bool, num_units: int) -> torch.Tensor:
    """Embedding.__call__ (coding.py:60-64): lookup, optionally times sqrt(num_units)."""
    out = table[ids.long()]
    if scale:
        out = out * (num_units ** 0.5)
    return out


def sinusoid_scale(num_units: int) -> np.ndarray:
    """TimeSinusoidCoding.__init__ (coding.py:134-135): float64 power, stored as fp32."""
    return np.power(10000, np.arange(0, num_units, 2) * 1. / num_units).astype(np.float32)


def time_sinusoid_code(ts32: torch.Tensor, num_units: int, dtype) -> torch.Tensor:
    """TimeSinusoidCoding.code (coding.py:137-149).  ``ts32`` is the fp32 scaled
    timestamp tensor [B,L].  x = ts/scale is an fp32 divide (coding.py:142) in every
    mode; sin/cos are evaluated in ``dtype``; [sin, cos] interleaved (coding.py:147-148)."""
    assert ts32.dim() == 2, "the tensor rank should be 2."  # coding.py:139
    scale = torch.from_numpy(sinusoid_scale(num_units))
    x = (ts32.to(torch.float32).unsqueeze(-1) / scale).to(dtype)
    code = torch.stack([torch.sin(x), torch.cos(x)], dim=-1)
    return code.reshape(ts32.shape[0], ts32.shape[1], num_units)


def time_function_code(inputs: torch.Tensor, basis_freq: torch.Tensor, phase: torch.Tensor, dtype) -> torch.Tensor:
    """TimeFunctionCoding.code (coding.py:112-122): reshape to [B,L,-1], tile over num_units, x*freq
    (fp32 multiply) + phase (fp32 bias_add), cos evaluated in ``dtype``."""
    B, L = inputs.shape[0], inputs.shape[1]
    x = inputs.to(torch.float32).reshape(B, L, -1).unsqueeze(-1)
    arg = (x * basis_freq.to(torch.float32) + phase.to(torch.float32)).to(dtype)
    return torch.cos(arg)


# ----------------------------------------------------------------------------
# src/module/temporal.py
# ----------------------------------------------------------------------------
def fold_heads(x: torch.Tensor, h: int) -> torch.Tensor:
    """tf.concat(tf.split(x, h, axis=2), axis=0) (temporal.py:346-349,413-416):
    [B,L,d] -> [h*B,L,d/h], head-major (index = head*B + b)."""
    return torch.cat(torch.split(x, x.shape[2] // h, dim=2), dim=0)


def unfold_heads(x: torch.Tensor, h: int) -> torch.Tensor:
    """tf.concat(tf.split(x, h, axis=0), axis=2) (temporal.py:382,444)."""
    return torch.cat(torch.split(x, x.shape[0] // h, dim=0), dim=2)


def intensity(H: torch.Tensor, intervals: torch.Tensor, marks: torch.Tensor, w: dict, num_heads: int,
              num_events: int, literal: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """MAU.intensity (temporal.py:281-315).
    H [hB,L,dh]; intervals [B,L]; marks [B,L,E] (integer multi-hot) -> G [hB,L,L], lam [hB,L,E]."""
    dh = H.shape[-1]
    E = num_events
    iv = intervals.unsqueeze(-1).repeat(num_heads, 1, 1)                      # :283
    x = torch.cat([H, iv], dim=-1)                                            # :287
    Z = torch.sigmoid(x @ w["int_w"] + w["int_b"])                            # :289-290
    Z = Z.reshape(Z.shape[0], Z.shape[1], E, dh)                              # :291 split by event
    s = torch.exp(w["int_scaling"])                                           # :302
    mi = torch.einsum("nqed,ed->nqe", Z, w["int_weight"]) / s                 # :293-305
    lam = s * torch.log(1. + torch.exp(mi))                                   # :306-307 (naive softplus, Q6)
    mk = marks.to(H.dtype)                                                    # :311 tf.to_float
    if literal:
        L = H.shape[1]
        lam4 = lam.unsqueeze(2).repeat(1, 1, L, 1)                            # :309-310
        mk4 = mk.unsqueeze(1).repeat(num_heads, L, 1, 1)                      # :311-312
        G = (lam4 * mk4).sum(-1)                                              # :313
    else:
        G = torch.einsum("nqe,nke->nqk", lam, mk.repeat(num_heads, 1, 1))
    return G, lam


def _attention_core(Q, K, V, T, kmask, intervals, marks, w, num_heads, num_events, causal: bool,
                    diag_one: bool, literal: bool):
    """Shared body of MAU.__call__ (temporal.py:345-382) and BiMAU.__call__ (:412-444)."""
    h = num_heads
    Q_, K_, V_, T_ = (fold_heads(t, h) for t in (Q, K, V, T))
    dh = K_.shape[-1]
    S = Q_ @ K_.transpose(1, 2)                                               # :352 / :419
    S = S / (dh ** 0.5)                                                       # :355 / :422
    L = S.shape[1]
    km = kmask.unsqueeze(1).repeat(h, L, 1)                                   # EasyDGL.py:94-95
    S = torch.where(km == 0, torch.full_like(S, MASK_FILL), S)                # :358-359 / :425-426
    if causal:                                                                # :362-367
        tril = torch.tril(torch.ones(L, L, dtype=S.dtype))
        S = torch.where(tril.unsqueeze(0) == 0, torch.full_like(S, MASK_FILL), S)
    P = torch.softmax(S, dim=-1)                                              # :370 / :429
    Hs = P @ T_                                                               # :375 / :434
    G, lam = intensity(Hs, intervals, marks, w, h, num_events, literal)       # :376 / :435
    if diag_one:                                                              # :438-439 (BiMAU only)
        eye = torch.eye(L, dtype=torch.bool).unsqueeze(0)
        G = torch.where(eye, torch.ones_like(G), G)
    O = (G * P) @ V_                                                          # :379-381 / :441-443
    return unfold_heads(O, h), lam


def bimau(queries, kmask, intervals, marks, w, num_units, num_heads, num_events, literal=False):
    """BiMAU.__call__ (temporal.py:404-452).  queries [B,L,Cin]; `keys`, `causality`
    ignored by the reference (Q15).  Returns (outputs [B,L,d], lam [hB,L,E])."""
    d = num_units
    QKVT = queries @ w["qkvt_w"] + w["qkvt_b"]                                # :409
    Q, K, V, T = torch.split(QKVT, d, dim=-1)                                 # :410
    O, lam = _attention_core(Q, K, V, T, kmask, intervals, marks, w, num_heads, num_events,
                             causal=False, diag_one=True, literal=literal)
    O = O + queries[:, :, :d]                                                 # :447
    return O, lam


def mgau(queries, kmask, intervals, marks, w, num_units, num_heads, num_events, literal=False):
    """MGAU.__call__ (temporal.py:463-508): BiMAU without set_diag."""
    d = num_units
    QKVT = queries @ w["qkvt_w"] + w["qkvt_b"]                                # :468
    Q, K, V, T = torch.split(QKVT, d, dim=-1)
    O, lam = _attention_core(Q, K, V, T, kmask, intervals, marks, w, num_heads, num_events,
                             causal=False, diag_one=False, literal=literal)
    O = O + queries[:, :, :d]                                                 # :503
    return O, lam


def mau(queries, keys, kmask, intervals, marks, w, num_units, num_heads, num_events, causality=True,
        literal=False):
    """MAU.__call__ (temporal.py:335-390).  Four separate dense layers (:340-343)."""
    d = num_units
    Q = queries @ w["q_w"] + w["q_b"]
    K = keys @ w["k_w"] + w["k_b"]
    V = keys @ w["v_w"] + w["v_b"]
    T = keys @ w["t_w"] + w["t_b"]
    O, lam = _attention_core(Q, K, V, T, kmask, intervals, marks, w, num_heads, num_events,
                             causal=bool(causality), diag_one=False, literal=literal)
    O = O + queries[:, :, :d]                                                 # :385
    return O, lam


# ----------------------------------------------------------------------------
# src/model/EasyDGL.py
# ----------------------------------------------------------------------------
def _time_attn_softmax(scores, keys, causality, num_heads, masks=None):
    """Scale is applied by the caller.  Key masking from ``keys`` (temporal.py:64-70,149-155) or from a given
    ``masks`` [N,T_q,T_k] (temporal.py:232-233), causality (temporal.py:72-79), softmax (temporal.py:82)."""
    hN, Tq, Tk = scores.shape
    if masks is None:
        km = torch.sign(keys.abs().sum(-1))                                   # [N,T_k]
        km = km.repeat(num_heads, 1).unsqueeze(1).expand(hN, Tq, Tk)
    else:
        km = masks.repeat(num_heads, 1, 1) if masks.shape[0] * num_heads == hN else masks
    fill = torch.full_like(scores, MASK_FILL)
    scores = torch.where(km == 0, fill, scores)
    if causality:
        tril = torch.tril(torch.ones(Tq, Tk, dtype=scores.dtype)).unsqueeze(0).expand(hN, Tq, Tk)
        scores = torch.where(tril == 0, fill, scores)
    return torch.softmax(scores, dim=-1)


def ti_attention(queries, keys, intervals, w, pos_k, pos_v, tab_k, tab_v, num_heads, causality=True):
    """TiMultiHeadAttention.__call__ (temporal.py:37-109), eval mode, literal: the [h*N,T_q,T_k,C/h] interval-code
    tensors are materialised like the reference.  pos_* [T,C] = PositionCoding tables (rows 0..T-1, coding.py:76-79),
    tab_* [vocab,C] = TimeIntervalCoding tables (coding.py:93-94), intervals int64 [N,T_q,T_k]."""
    h = num_heads
    N, T, _ = queries.shape
    Q = queries @ w["q_w"] + w["q_b"]
    K = keys @ w["k_w"] + w["k_b"]
    V = keys @ w["v_w"] + w["v_b"]
    Q_, K_, V_ = fold_heads(Q, h), fold_heads(K, h), fold_heads(V, h)
    Kp = fold_heads(pos_k[:T].unsqueeze(0).expand(N, T, -1), h)
    Vp = fold_heads(pos_v[:T].unsqueeze(0).expand(N, T, -1), h)
    Kt = torch.cat(torch.split(tab_k[intervals], tab_k.shape[1] // h, dim=3), dim=0)   # [hN,T_q,T_k,dh]
    Vt = torch.cat(torch.split(tab_v[intervals], tab_v.shape[1] // h, dim=3), dim=0)
    out = Q_ @ K_.transpose(1, 2) + Q_ @ Kp.transpose(1, 2) + (Kt @ Q_.unsqueeze(3)).squeeze(3)   # temporal.py:55-59
    out = out / (K_.shape[-1] ** 0.5)
    P = _time_attn_softmax(out, keys, causality, h)
    qm = torch.sign(queries.abs().sum(-1)).repeat(h, 1).unsqueeze(-1)                 # temporal.py:87-90
    P = P * qm
    o = P @ V_ + P @ Vp + (P.unsqueeze(2) @ Vt).squeeze(2)                            # temporal.py:96-100
    return unfold_heads(o, h) + queries                                                # temporal.py:103-106


def tf_attention(queries, keys, intervals, w, pos_k, basis_freq, phase, num_heads, causality=True):
    """TfMultiHeadAttention.__call__ (temporal.py:126-185), eval mode, literal.  intervals fp32 [N,T_q,T_k]."""
    h = num_heads
    N, T, _ = queries.shape
    dtype = queries.dtype
    Q = queries @ w["q_w"] + w["q_b"]
    K = keys @ w["k_w"] + w["k_b"]
    V = keys @ w["v_w"] + w["v_b"]
    Q_, K_, V_ = fold_heads(Q, h), fold_heads(K, h), fold_heads(V, h)
    Kp = fold_heads(pos_k[:T].unsqueeze(0).expand(N, T, -1), h)
    code = time_function_code(intervals, basis_freq, phase, dtype)                     # [N,T_q,T_k,C]
    Kt = torch.cat(torch.split(code, code.shape[3] // h, dim=3), dim=0)
    out = Q_ @ K_.transpose(1, 2) + Q_ @ Kp.transpose(1, 2) + (Kt @ Q_.unsqueeze(3)).squeeze(3)   # temporal.py:144-147
    out = out / (K_.shape[-1] ** 0.5)
    P = _time_attn_softmax(out, keys, causality, h)
    return unfold_heads(P @ V_, h) + queries                                           # temporal.py:176-183


def layernorm_last(x, gamma, beta, eps=1e-8):
    """module.normalize.layernorm (normalize.py:9-19)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return gamma * ((x - mean) / torch.sqrt(var + eps)) + beta


def tg_attention(queries, keys, masks, intervals, w, basis_freq, phase, num_heads, causality=True):
    """TgMultiHeadAttention.__call__ (temporal.py:204-264), eval mode, literal: [N,T_q,T_k,2C] keys are built and
    projected like the reference.  masks [N,T_q,T_k] (or [h*N,...]); output [N,T_q,2C]."""
    h = num_heads
    N, Tq, C = queries.shape
    dtype = queries.dtype
    q_t = time_function_code(torch.zeros(N, Tq, 1), basis_freq, phase, dtype)          # [N,T_q,1,C]
    q4 = torch.cat([queries.unsqueeze(2), q_t], dim=-1)                                # [N,T_q,1,2C]
    k_t = time_function_code(intervals, basis_freq, phase, dtype)                      # [N,T_q,T_k,C]
    k4 = torch.cat([keys.unsqueeze(1).expand(N, Tq, keys.shape[1], keys.shape[2]), k_t], dim=-1)
    Q = q4 @ w["q_w"] + w["q_b"]
    K = k4 @ w["k_w"] + w["k_b"]
    V = k4 @ w["v_w"] + w["v_b"]
    split = lambda x: torch.cat(torch.split(x, x.shape[3] // h, dim=3), dim=0)         # noqa: E731
    Q_, K_, V_ = split(Q), split(K), split(V)
    out = (Q_ @ K_.transpose(2, 3)).squeeze(2)                                         # [hN,T_q,T_k]
    out = out / (K_.shape[-1] ** 0.5)
    P = _time_attn_softmax(out, None, causality, h, masks=masks)
    o = (P.unsqueeze(2) @ V_).squeeze(2)                                               # [hN,T_q,dh]
    o = unfold_heads(o, h)
    o = o @ w["o_w"] + w["o_b"] + q4.squeeze(2)                                        # temporal.py:260-261
    return layernorm_last(o, w["ln_g"], w["ln_b"])


def gelu(x: torch.Tensor) -> torch.Tensor:
    """EasyDGL.gelu (EasyDGL.py:19-32): exact erf form (Q18)."""
    cdf = 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    return x * cdf


def _cast(wd: dict, dtype) -> dict:
    out = {}
    for k, v in wd.items():
        if isinstance(v, list):
            out[k] = [_cast(b, dtype) for b in v]
        elif torch.is_tensor(v) and v.is_floating_point():
            out[k] = v.to(dtype)
        else:
            out[k] = v
    return out


def scaled_time(seqs_t: torch.Tensor, time_scale: float) -> torch.Tensor:
    """features['seqs_t'] / self.time_scale (EasyDGL.py:71, CTSMA.py:47): an fp32 divide."""
    return seqs_t.to(torch.float32) / torch.tensor(time_scale, dtype=torch.float32)


def easydgl_inputs(seqs_i, seqs_t, W, cfg, dtype):
    """EasyDGL.__call__ input assembly (EasyDGL.py:70-95).  Returns X0 [B,L,3d], kmask [B,L],
    spans [B,L], marks [B,L,E] int64."""
    d = cfg.num_units
    ids = seqs_i.long()
    ts32 = scaled_time(seqs_t, cfg.time_scale)                                # :71
    spans = torch.clamp(ts32[:, 1:] - ts32[:, :-1], 0., 100.)                 # :73 (fp32 subtract)
    spans = torch.cat([spans[:, :1], spans], dim=-1).to(dtype)                # :74
    mark_ids = torch.where(ids == cfg.mask_id, torch.zeros_like(ids), ids)    # :76
    marks = W["mark_table"][mark_ids]                                         # :77  [B,L,E] int
    tcode = time_sinusoid_code(ts32, d, dtype)                                # :80
    x = embedding(zero_pad_table(W["item_embs"]), ids, True, d) + tcode       # :83
    B, L = ids.shape
    pos = W["pos_embs"][torch.arange(L)].unsqueeze(0).expand(B, L, d)         # :84 coding.py:76-79
    mcode = zero_pad_table(W["mark_embs"])[marks].sum(dim=2)                  # :87-88 (Q1: values as indices)
    X0 = torch.cat([x, pos, mcode], dim=-1)                                   # :89
    kmask = (ids != 0).to(dtype)                                              # :94
    return X0, kmask, spans, marks


def easydgl_forward(seqs_i, seqs_t, W, cfg, dtype=torch.float64, literal=False, return_all=False):
    """EasyDGL.__call__(features, is_training=False) (EasyDGL.py:69-151) -> logits [B,N1]."""
    W = _cast(W, dtype)
    d, h, E = cfg.num_units, cfg.num_heads, cfg.num_events
    X0, kmask, spans, marks = easydgl_inputs(seqs_i, seqs_t, W, cfg, dtype)
    prev = X0
    lams = []
    for blk in W["blocks"]:                                                   # :100
        att, lam = bimau(prev, kmask, spans, marks, blk, d, h, E, literal)    # :108-109
        lams.append(lam)
        att = att @ blk["ao_w"] + blk["ao_b"]                                 # :113
        att = layernorm(att + prev[:, :, :d], blk["ao_ln_g"], blk["ao_ln_b"])  # :116
        inter = gelu(att @ blk["ff1_w"] + blk["ff1_b"])                       # :120-121
        out = inter @ blk["ff2_w"] + blk["ff2_b"]                             # :125
        prev = layernorm(out + att, blk["ff_ln_g"], blk["ff_ln_b"])           # :128
    Y = gelu(prev @ W["tr_w"] + W["tr_b"])                                    # :138
    Y = layernorm(Y, W["tr_ln_g"], W["tr_ln_b"])                              # :139
    y = Y[:, -1]                                                              # :146
    logits = y @ zero_pad_table(W["item_embs"]).t()                           # :149 (unscaled table, Q11)
    logits = logits + output_bias(W["output_bias"])                           # :150
    if return_all:
        return SimpleNamespace(logits=logits, y=y, X0=X0, spans=spans, marks=marks, kmask=kmask,
                               lams=lams, last_hidden=prev)
    return logits


# ----------------------------------------------------------------------------
# src/model/CTSMA.py
# ----------------------------------------------------------------------------
def ctsma_inputs(seqs_i, seqs_t, W, cfg, dtype):
    """CTSMA.__call__ input assembly (CTSMA.py:47-60).  seqs_i [B,S]; seqs_t [B,S+1]."""
    d = cfg.num_units
    ids = seqs_i.long()
    ts32 = scaled_time(seqs_t, cfg.time_scale)                                # :48
    spans = (ts32[:, 1:] - ts32[:, :-1]).to(dtype)                            # :49 unclipped (Q16)
    marks = W["mark_table"][ids]                                              # :52
    x = embedding(zero_pad_table(W["item_embs"]), ids, True, d)               # :53
    B, S = ids.shape
    pos = W["pos_embs"][torch.arange(S)].unsqueeze(0).expand(B, S, d)
    X = torch.cat([x, pos], dim=-1)                                           # :54 coding.py:72-74
    kmask = (ids != 0).to(dtype)                                              # :59
    return X, kmask, spans, marks


def ctsma_forward(seqs_i, seqs_t, W, cfg, dtype=torch.float64, literal=False, return_all=False):
    """CTSMA.__call__(features, is_training=False) (CTSMA.py:46-91) -> logits [B,N]."""
    W = _cast(W, dtype)
    d, h, E = cfg.num_units, cfg.num_heads, cfg.num_events
    X, kmask, spans, marks = ctsma_inputs(seqs_i, seqs_t, W, cfg, dtype)
    out = X
    lams = []
    for blk in W["blocks"]:                                                   # :64
        qin = layernorm(out, blk["ln1_g"], blk["ln1_b"])                      # :68
        out, lam = mau(qin, out, kmask, spans, marks, blk, d, h, E, True, literal)
        lams.append(lam)
        f_in = layernorm(out, blk["ln2_g"], blk["ln2_b"])                     # :73
        out = feedforward(f_in, blk["ff1_w"], blk["ff1_b"], blk["ff2_w"], blk["ff2_b"])
    Y = layernorm(out, W["out_ln_g"], W["out_ln_b"])                          # :80
    y = Y[:, -1]                                                              # :87
    logits = y @ zero_pad_table(W["item_embs"]).t()                           # :89
    logits = logits + output_bias(W["output_bias"])                           # :90
    if return_all:
        return SimpleNamespace(logits=logits, y=y, X0=X, spans=spans, marks=marks, kmask=kmask,
                               lams=lams, last_hidden=out)
    return logits


def forward(seqs_i, seqs_t, W, cfg, **kw):
    if cfg.model == "EasyDGL":
        return easydgl_forward(seqs_i, seqs_t, W, cfg, **kw)
    if cfg.model == "CTSMA":
        return ctsma_forward(seqs_i, seqs_t, W, cfg, **kw)
    raise NotImplementedError("The ranking model: {0} not implemented".format(cfg.model))  # util.py:96


# ----------------------------------------------------------------------------
# training-mode forward: Model.train() up to the loss (dropout rates 0; the backward pass / Adam are not restated)
# ----------------------------------------------------------------------------
def biased_likelihood(mark_intensity: torch.Tensor, next_mark_onehot: torch.Tensor, intervals: torch.Tensor):
    """MAU.biased_likelihood (temporal.py:317-333).  mark_intensity [hN,Tq,E]; next_mark_onehot [hN,Tq,E];
    intervals [hN,Tq]."""
    mi = mark_intensity * torch.sign(next_mark_onehot.sum(dim=2, keepdim=True))          # :321
    event_intensity = (mi * next_mark_onehot).sum(dim=2)                                 # :322
    event_ll = torch.log(torch.where(event_intensity == 0, torch.ones_like(event_intensity), event_intensity))  # :324
    event_ll = event_ll.sum()                                                            # :325
    entire_intensity = mi.sum(dim=2)                                                     # :327
    non_event_ll = (entire_intensity * intervals * .5).sum()                             # :328-329
    num_events = next_mark_onehot.sum()                                                  # :331
    return -(event_ll - non_event_ll) / num_events                                       # :332


def l2_regularization(W: dict, cfg, l2_reg: float, dtype) -> torch.Tensor:
    """tf.losses.get_regularization_loss(): l2_reg * tf.nn.l2_loss(table) = l2_reg * sum(table**2) / 2 for every
    C.Embedding table built with l2_reg (coding.py:13-44,48,55): item_embs, spatial_embs (and mark_embs for
    EasyDGL, EasyDGL.py:50-54; CTSMA.py:33-35).  The regulariser sees the raw variable (row 0 included)."""
    if l2_reg == 0.:
        return torch.zeros((), dtype=dtype)
    names = ["item_embs", "pos_embs"] + (["mark_embs"] if cfg.model == "EasyDGL" else [])
    tot = torch.zeros((), dtype=dtype)
    for n in names:
        tot = tot + l2_reg * (W[n].to(dtype) ** 2).sum() / 2
    return tot


def train_forward(seqs_i, seqs_t, labels, W, cfg, masked_positions=None, l2_reg=0., ct_reg=0.,
                  dtype=torch.float64, return_logits=False):
    """Model.train(features, labels) up to `loss` with both dropout rates 0 (EasyDGL.py:140-189, CTSMA.py:82-124).
    EasyDGL: features carry `masked_positions` [B,M] (dataloader.py:181-201), labels [B,M]; CTSMA: every position
    is predicted, labels [B,S] (dataloader.py:95-98).  Returns dict(loss, ce, l2, ct[, logits])."""
    out = forward(seqs_i, seqs_t, W, cfg, dtype=dtype, return_all=True)
    Wd = _cast(W, dtype)
    h = cfg.num_heads
    B = seqs_i.shape[0]
    if cfg.model == "EasyDGL":
        Y = gelu(out.last_hidden @ Wd["tr_w"] + Wd["tr_b"])                              # :138
        Y = layernorm(Y, Wd["tr_ln_g"], Wd["tr_ln_b"])                                   # :139
        pos = masked_positions.long()
        Yg = torch.gather(Y, 1, pos.unsqueeze(-1).expand(-1, -1, Y.shape[-1]))           # :141 tf.batch_gather
        Yg = Yg.reshape(B * pos.shape[1], -1)                                            # :142
    else:
        Y = layernorm(out.last_hidden, Wd["out_ln_g"], Wd["out_ln_b"])                   # CTSMA.py:80
        Yg = Y.reshape(B * Y.shape[1], -1)                                               # :83
    logits = Yg @ zero_pad_table(Wd["item_embs"]).t() + output_bias(Wd["output_bias"])   # :149-150 / CTSMA :89-90
    log_probs = torch.log(torch.softmax(logits, -1) + 1e-5)                              # :155 / CTSMA :95
    reg = l2_regularization(W, cfg, l2_reg, dtype)                                       # :158
    ct = torch.zeros((), dtype=dtype)
    if ct_reg != 0.:
        st = seqs_t.to(torch.float32)                                                    # UNSCALED timestamps (Q9)
        if cfg.model == "EasyDGL":
            spans = torch.clamp(st[:, 1:] - st[:, :-1], 0., 100.)                        # :161 clip_by_value
            spans = torch.cat([spans[:, :1], spans], dim=-1)                             # :162
            spans = torch.gather(spans, 1, masked_positions.long())                      # :163
        else:
            spans = st[:, 1:] - st[:, :-1]                                               # CTSMA.py:100
        spans = spans.to(dtype)
        nm = W["mark_table"][labels.long()].to(dtype)                                    # :164 / CTSMA :101
        posr = masked_positions.long() if cfg.model == "EasyDGL" else None
        if h != 1:                                                                       # :166-169 / CTSMA :103-105
            spans = spans.repeat(h, 1)
            nm = nm.repeat(h, 1, 1)
            if posr is not None:
                posr = posr.repeat(h, 1)
        for lam in out.lams:                                                             # :171 collection "LLE_PP"
            if posr is not None:
                lam = torch.gather(lam, 1, posr.unsqueeze(-1).expand(-1, -1, lam.shape[-1]))  # :172
            bl = biased_likelihood(lam, nm, spans)
            ct = ct + (ct_reg * bl / h if cfg.model == "EasyDGL" else ct_reg * bl)       # :175 / CTSMA :110
    lab = labels.reshape(-1).long()                                                      # :178
    weights = (lab != 0).to(dtype)                                                       # :180
    per_example = -log_probs.gather(1, lab.unsqueeze(1)).squeeze(1)                      # :179,182 one_hot . log_probs
    ce = (weights * per_example).sum() / (weights.sum() + 1e-5)                          # :183-185
    res = dict(loss=ce + reg + ct, ce=ce, l2=reg, ct=ct)
    if return_logits:
        res["logits"] = logits
    return res


# ----------------------------------------------------------------------------
# comparators
# ----------------------------------------------------------------------------
def topk_set_compare(idx_test: torch.Tensor, logits64_masked: torch.Tensor, k: int, tau: float):
    """Tie-aware top-K set comparison against the fp64 oracle's masked logits.

    A row is *exact* if the index sets are equal.  A differing row is *excused* iff
    every item in the symmetric difference has an fp64 logit within ``tau`` of the
    oracle's k-th largest value (i.e. only near-ties at the cut were swapped).
    Returns dict(exact=, excused=, bad=, rows=)."""
    B = idx_test.shape[0]
    ref_val, ref_idx = topk_lower_index_first(logits64_masked, k + 1)
    kth = ref_val[:, k - 1]
    exact = excused = bad = 0
    bad_rows = []
    for b in range(B):
        a = set(idx_test[b].tolist())
        r = set(ref_idx[b, :k].tolist())
        if a == r:
            exact += 1
            continue
        diff = list(a ^ r)
        v = logits64_masked[b, diff]
        if bool(((v - kth[b]).abs() <= tau).all()):
            excused += 1
        else:
            bad += 1
            bad_rows.append(b)
    return dict(exact=exact, excused=excused, bad=bad, rows=B, bad_rows=bad_rows[:8])
