"""Facade of ``src/module/temporal.py``: MAU (temporal.py:267-390) and BiMAU (temporal.py:396-452).

Constructor and call signatures follow the reference.  TF creates the layer's variables on the
first call; so does this facade (reference initialisers) unless ``weights`` is supplied.  The
attention runs in ``edgl_attention_layer`` / ``edgl_intensity`` of libeasydgl_b200.so.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

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
