"""Facade of ``src/model/CTSMA.py``: ``CTSMA(num_items, FLAGS)`` and
``model(features, is_training=False) -> logits [B, num_items]`` (CTSMA.py:22-91)."""
from __future__ import annotations

from .Base import Sequential


class CTSMA(Sequential):
    _model = "CTSMA"

    def __init__(self, num_items, FLAGS, weights=None, mark_table=None, device="cuda:0", max_batch=512):
        super().__init__(num_items, FLAGS)
        self._raw_items = num_items
        self.ct_reg = getattr(FLAGS, "ct_reg", 0.)
        self.time_scale = getattr(FLAGS, "time_scale", 1.0)
        self._setup(FLAGS, weights, mark_table, device, max_batch)

    def __call__(self, features, is_training):
        return self._forward(features, is_training)
