"""Facade of ``src/module/temporal.py``: MAU (temporal.py:267-390), BiMAU (temporal.py:396-452), MGAU
(temporal.py:455-508) and the time-aware attention layers of the baselines - TiMultiHeadAttention (temporal.py:15-109),
TfMultiHeadAttention (temporal.py:112-185), TgMultiHeadAttention (temporal.py:188-264).

Constructor and call signatures follow the reference.  TF creates the layer's variables on the
first call; so does this facade (reference initialisers) unless ``weights`` is supplied.  The
attention runs in ``edgl_attention_layer`` / ``edgl_intensity`` of libeasydgl_b200.so.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .. import engine as _E
from ..engine import Engine


def _glorot(i, o):
    lim = float(np.sqrt(6.0 / (i + o)))
    return (torch.rand(i, o) * 2 - 1) * lim


class MAU(object):
    _model = "CTSMA"
    _no_diag = False

    def __init__(self, num_units, num_heads, num_events, dropout_rate, scope="modulating_attention", weights=None,
                 device="cuda:0"):
        self.num_units = num_units
        self.num_heads = num_heads
        self.num_events = num_events
        self.dropout_rate = dropout_rate
        self.scope = scope
        self.weights = weights
        self.device = device
        self._eng = None
        self._key = None

    # -- variables ------------------------------------------------------------------------------
    def _init_weights(self, cin):
        d, dh, E = self.num_units, self.num_units // self.num_heads, self.num_events
        w = {"int_w": _glorot(dh + 1, dh * E), "int_b": torch.zeros(dh * E), "int_weight": _glorot(E, dh),
             "int_scaling": torch.zeros(E)}
        if self._model == "EasyDGL":
            w.update({"qkvt_w": torch.randn(cin, 4 * d) * 0.02, "qkvt_b": torch.zeros(4 * d)})  # temporal.py:393
        else:
            for n in "qkvt":
                w[n + "_w"] = _glorot(cin, d)
                w[n + "_b"] = torch.zeros(d)
        return w

    def _engine(self, B, L, cin):
        key = (L, cin)
        if self._eng is not None and self._key == key and self._eng.max_batch >= B:
            return self._eng
        d, E = self.num_units, self.num_events
        if self.weights is None:
            self.weights = self._init_weights(cin)
        easy = self._model == "EasyDGL"
        if cin != (3 * d if easy else 2 * d) and cin != d:
            raise ValueError("layer input width %d unsupported (expected %d or %d)" % (cin, d, (3 if easy else 2) * d))
        nb = 1 if cin != d else 2  # block index 1 has a width-d input (Q17)
        cfg = SimpleNamespace(model=self._model, num_units=d, num_heads=self.num_heads, num_blocks=nb, num_events=E,
                              L=L, ts_len=L if easy else L + 1, num_rows=2, mask_id=-1, time_scale=1.0, topk=1)
        z = torch.zeros
        blk = dict(self.weights)
        if easy:
            filler = {"ao_w": z(d, d), "ao_b": z(d), "ao_ln_g": z(d), "ao_ln_b": z(d), "ff1_w": z(d, 2 * d),
                      "ff1_b": z(2 * d), "ff2_w": z(2 * d, d), "ff2_b": z(d), "ff_ln_g": z(d), "ff_ln_b": z(d)}
            model_w = {"mark_embs": z(E, d), "tr_w": z(d, d), "tr_b": z(d), "tr_ln_g": z(d), "tr_ln_b": z(d)}
        else:
            filler = {"ln1_g": z(cin), "ln1_b": z(cin), "ln2_g": z(d), "ln2_b": z(d), "ff1_w": z(d, d), "ff1_b": z(d),
                      "ff2_w": z(d, d), "ff2_b": z(d)}
            model_w = {"out_ln_g": z(d), "out_ln_b": z(d)}
        for k, v in filler.items():
            blk.setdefault(k, v)
        blocks = [blk]
        if nb == 2:  # block 0 is a placeholder with the wide input
            wide = 3 * d if easy else 2 * d
            b0 = dict(blk)
            if easy:
                b0["qkvt_w"] = z(wide, 4 * d)
            else:
                b0.update({n + "_w": z(wide, d) for n in "qkvt"})
                b0.update({"ln1_g": z(wide), "ln1_b": z(wide)})
            blocks = [b0, blk]
        W = dict(item_embs=z(2, d), pos_embs=z(L, d), output_bias=z(1), mark_table=torch.zeros(1, E, dtype=torch.int64),
                 blocks=blocks, **model_w)
        self._eng = Engine(cfg, W, max_batch=B, device=self.device, topk=1)
        self._key = key
        self._block = nb - 1
        return self._eng

    @staticmethod
    def _key_mask(masks, B):
        """The reference passes masks [h*B, T_q, T_k] tiled from the per-key padding mask
        (EasyDGL.py:94-95); recover kmask [B, T_k]."""
        if masks.dim() == 3:
            masks = masks[:B, 0, :]
        return (masks != 0).to(torch.uint8).contiguous()

    def intensity(self, H, intervals, mark_onehot):
        """temporal.py:281-315 -> (G [hB,L,L], lam [hB,L,E])."""
        hB, L, dh = H.shape
        eng = self._engine(hB // self.num_heads, L, self._key[1] if self._key else self.num_units)
        return eng.intensity(self._block, H, intervals.to(torch.float32), mark_onehot.to(torch.uint8))

    def __call__(self, queries, keys, masks, intervals, marks, is_training, causality):
        if is_training:
            raise NotImplementedError("training-mode forward (dropout, likelihood) is out of scope (SURVEY 8f)")
        B, L, cin = queries.shape
        eng = self._engine(B, L, cin)
        return eng.attention_layer(self._block, queries, keys, self._key_mask(masks, B), intervals.to(torch.float32),
                                   marks.to(torch.uint8), causality=bool(causality), no_diag=self._no_diag)


class BiMAU(MAU):
    _model = "EasyDGL"

    def __init__(self, num_units, num_heads, num_events, dropout_rate, scope="TMAU", weights=None, device="cuda:0"):
        super().__init__(num_units, num_heads, num_events, dropout_rate, scope, weights, device)

    def __call__(self, queries, keys, masks, intervals, marks, is_training, causality=None):
        # `keys` and `causality` are ignored exactly like the reference (temporal.py:404-429)
        return super().__call__(queries, None, masks, intervals, marks, is_training, False)


class MGAU(BiMAU):
    """temporal.py:455-508: Modulating Gated Attention Unit = BiMAU without ``tf.linalg.set_diag``
    (instantiated by no model of the reference; provided for completeness)."""
    _no_diag = True

    def __init__(self, num_units, num_heads, num_events, dropout_rate, scope="GAU_SMA", weights=None, device="cuda:0"):
        super().__init__(num_units, num_heads, num_events, dropout_rate, scope, weights, device)


class _TimeAttention(object):
    """Shared plumbing of the Ti / Tf / Tg layers: the dense kernels (tf.layers.dense creates them on the first call with
    glorot-uniform kernels and zero biases; pass ``weights`` to use given ones) and the eval-only guard."""
    _names = ("q", "k", "v")

    def __init__(self, num_units, num_heads, dropout_rate, l2_reg, scope, weights=None, device="cuda:0"):
        self.num_units = num_units
        self.num_heads = num_heads
        self.dropout_rate = dropout_rate
        self.l2_reg = l2_reg
        self.scope = scope
        self.weights = weights
        self.device = device

    def _w(self, cin_q, cin_k):
        if self.weights is None:
            C = self.num_units
            self.weights = {"q_w": _glorot(cin_q, C), "q_b": torch.zeros(C), "k_w": _glorot(cin_k, C), "k_b": torch.zeros(C),
                            "v_w": _glorot(cin_k, C), "v_b": torch.zeros(C)}
        return {k: v.to(device=self.device, dtype=torch.float32).contiguous() for k, v in self.weights.items()}

    @staticmethod
    def _eval_only(is_training):
        if is_training:
            raise NotImplementedError("training-mode forward (dropout) of the baseline attention layers is out of scope")


class TiMultiHeadAttention(_TimeAttention):
    """temporal.py:15-109 (TiSASRec).  ``pcoding_K/V``: PositionCoding, ``tcoding_K/V``: TimeIntervalCoding;
    ``intervals`` int64 [N, T_q, T_k] (already clipped, TiSASREC.py:56-59)."""

    def __init__(self, num_units, num_heads, dropout_rate, l2_reg, pcoding_K, pcoding_V, tcoding_K, tcoding_V,
                 scope="attention/timeinterval", weights=None, device="cuda:0"):
        super().__init__(num_units, num_heads, dropout_rate, l2_reg, scope, weights, device)
        self.pcoding_K, self.pcoding_V, self.tcoding_K, self.tcoding_V = pcoding_K, pcoding_V, tcoding_K, tcoding_V

    def __call__(self, queries, keys, intervals, is_training, causality):
        self._eval_only(is_training)
        w = self._w(queries.shape[-1], keys.shape[-1])
        Q = _E.dense(queries, w["q_w"], w["q_b"])   # temporal.py:41-43
        K = _E.dense(keys, w["k_w"], w["k_b"])
        V = _E.dense(keys, w["v_w"], w["v_b"])
        T = queries.shape[1]
        return _E.time_attention(Q, K, V, self.num_heads, key_mask=_E.row_nonzero(keys),
                                 query_mask=_E.row_nonzero(queries),
                                 pos_k=self.pcoding_K.pembs.lookup_table[:T], pos_v=self.pcoding_V.pembs.lookup_table[:T],
                                 time_mode=1, intervals=intervals.to(torch.int64),
                                 time_k=self.tcoding_K.pembs.lookup_table, time_v=self.tcoding_V.pembs.lookup_table,
                                 residual=queries, causality=causality)


class TfMultiHeadAttention(_TimeAttention):
    """temporal.py:112-185 (TGAT).  ``pcoding_K``: PositionCoding, ``tcoding_K``: TimeFunctionCoding (the Bochner / Mercer
    time kernel); ``intervals`` fp32 [N, T_q, T_k]."""

    def __init__(self, num_units, num_heads, dropout_rate, l2_reg, pcoding_K, tcoding_K, scope="attention/timeinterval",
                 weights=None, device="cuda:0"):
        super().__init__(num_units, num_heads, dropout_rate, l2_reg, scope, weights, device)
        self.pcoding_K, self.tcoding_K = pcoding_K, tcoding_K

    def __call__(self, queries, keys, intervals, is_training, causality):
        self._eval_only(is_training)
        w = self._w(queries.shape[-1], keys.shape[-1])
        Q = _E.dense(queries, w["q_w"], w["q_b"])   # temporal.py:131-133
        K = _E.dense(keys, w["k_w"], w["k_b"])
        V = _E.dense(keys, w["v_w"], w["v_b"])
        T = queries.shape[1]
        return _E.time_attention(Q, K, V, self.num_heads, key_mask=_E.row_nonzero(keys),
                                 pos_k=self.pcoding_K.pembs.lookup_table[:T], time_mode=2,
                                 intervals=intervals.to(torch.float32), basis_freq=self.tcoding_K.basis_freq,
                                 phase=self.tcoding_K.phase, residual=queries, causality=causality)


class TgMultiHeadAttention(_TimeAttention):
    """temporal.py:188-264 (TGSRec).  Keys and values are dense projections of ``[key_k | tcoding(interval(q, k))]``
    (width 2C); the [N, T_q, T_k, 2C] tensor of the reference is never built: the key half of the projection is taken
    per key, the time half is moved to the query side (scores) / applied after the probability-weighted sum of the time
    codes (values).  Output width 2C (dense to 2C + [queries | tcoding(0)] + last-axis layernorm)."""

    def __init__(self, num_units, num_heads, dropout_rate, l2_reg, tcoding, scope="attention/TgMultiHeadAttention",
                 weights=None, device="cuda:0"):
        super().__init__(num_units, num_heads, dropout_rate, l2_reg, scope, weights, device)
        self.tcoding = tcoding

    def _w(self, cin_q, cin_k):
        if self.weights is None:
            C = self.num_units
            self.weights = {"q_w": _glorot(2 * cin_q, C), "q_b": torch.zeros(C), "k_w": _glorot(2 * cin_k, C),
                            "k_b": torch.zeros(C), "v_w": _glorot(2 * cin_k, C), "v_b": torch.zeros(C),
                            "o_w": _glorot(C, 2 * C), "o_b": torch.zeros(2 * C), "ln_g": torch.ones(2 * C),
                            "ln_b": torch.zeros(2 * C)}
        return {k: v.to(device=self.device, dtype=torch.float32).contiguous() for k, v in self.weights.items()}

    def __call__(self, queries, keys, masks, intervals, is_training, causality):
        self._eval_only(is_training)
        B, Tq, Cq = queries.shape
        C, h = self.num_units, self.num_heads
        dh = C // h
        w = self._w(Cq, keys.shape[-1])
        Ck = keys.shape[-1]
        intervals = intervals.to(torch.float32)
        tc0 = self.tcoding.code(torch.zeros((B, Tq, 1), device=queries.device)).reshape(B, Tq, -1)   # temporal.py:209
        queries2 = torch.cat([queries, tc0], dim=-1).contiguous()                                     # temporal.py:211
        Q = _E.dense(queries2, w["q_w"], w["q_b"])                                                    # temporal.py:218
        A_k = _E.dense(keys, w["k_w"][:Ck].contiguous(), w["k_b"])     # key half of K = dense([keys | keys_t])
        A_v = _E.dense(keys, w["v_w"][:Ck].contiguous(), w["v_b"])
        # time half of K on the query side: U[b,q,head,:] = W_k2[:, head] Q[b,q,head]
        U = torch.empty((B, Tq, h, w["k_w"].shape[0] - Ck), dtype=torch.float32, device=queries.device)
        for hd in range(h):
            wk2 = w["k_w"][Ck:, hd * dh:(hd + 1) * dh].t().contiguous()          # [dh, C]
            U[:, :, hd, :] = _E.dense(Q[:, :, hd * dh:(hd + 1) * dh].contiguous(), wk2, None)
        kmask = MAU._key_mask(masks, B) if masks is not None else None
        out, TC = _E.time_attention(Q, A_k, A_v, h, key_mask=kmask, time_mode=3, intervals=intervals,
                                    basis_freq=self.tcoding.basis_freq, phase=self.tcoding.phase, U=U.contiguous(),
                                    causality=causality)
        for hd in range(h):  # time half of V: (sum_k P cos(.)) W_v2[:, head]
            wv2 = w["v_w"][Ck:, hd * dh:(hd + 1) * dh].contiguous()               # [C, dh]
            out[:, :, hd * dh:(hd + 1) * dh] += _E.dense(TC[:, :, hd, :].contiguous(), wv2, None)
        out = _E.dense(out, w["o_w"], w["o_b"]) + queries2                        # temporal.py:260-261
        return _E.layernorm_last(out, w["ln_g"], w["ln_b"], 1e-8)                 # temporal.py:262
