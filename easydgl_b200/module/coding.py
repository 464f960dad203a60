"""Facade of the reference's ``src/module/coding.py`` layers that sit on the hot path.

Same class names, ctor arguments and call signatures (coding.py:45-79,125-149); torch CUDA
tensors stand in for TF tensors and the arithmetic runs in libeasydgl_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import engine as _E


def _glorot(shape, gen=None):
    lim = float(np.sqrt(6.0 / (shape[0] + shape[1])))
    return (torch.rand(shape, generator=gen) * 2 - 1) * lim


class Embedding(object):
    """coding.py:45-64.  ``lookup_table`` is the (zero-padded) table like the reference attribute."""

    def __init__(self, vocab_size, num_units, l2_reg=0., zero_pad=True, scale=True, initializer=None,
                 scope="embedding", device="cuda:0"):
        self._num_units = num_units
        self._scale = scale
        self._zero_pad = zero_pad
        table = initializer if torch.is_tensor(initializer) else _glorot((vocab_size, num_units))
        table = table.to(device=device, dtype=torch.float32).contiguous()
        if zero_pad:  # coding.py:56-57
            table = torch.cat((torch.zeros(1, num_units, device=table.device), table[1:, :]), 0)
        self.lookup_table = table

    def __call__(self, inputs):
        # zero_pad already applied to lookup_table; scale = x * num_units ** 0.5 (coding.py:61-63)
        return _E.embedding_lookup(self.lookup_table, inputs, False, self._scale)


class PositionCoding(object):
    """coding.py:67-79."""

    def __init__(self, vocab_size, num_units, l2_reg=0., initializer=None, scope="coding/pos", device="cuda:0"):
        self.pembs = Embedding(vocab_size, num_units, l2_reg, zero_pad=False, initializer=initializer, scale=False,
                               device=device)

    def __call__(self, inputs, **kwargs):
        return torch.cat([inputs, self.code(inputs)], dim=-1)

    def code(self, inputs):
        batch_size, seqs_len = inputs.shape[0], inputs.shape[1]
        pos = torch.arange(seqs_len, device=inputs.device, dtype=torch.int64).unsqueeze(0).repeat(batch_size, 1)
        return self.pembs(pos)


class TimeIntervalCoding(object):
    """coding.py:82-94 (Li et al., TiSASRec): a learned embedding per clipped integer time interval."""

    def __init__(self, vocab_size, num_units, l2_reg=0., scope="coding/tim", device="cuda:0", initializer=None):
        self.pembs = Embedding(vocab_size, num_units, l2_reg, zero_pad=False, scale=False, initializer=initializer,
                               device=device)

    def code(self, inputs):
        return self.pembs(inputs)


class TimeFunctionCoding(object):
    """coding.py:97-122 (Xu et al., TGAT): learnable harmonic (Bochner / Mercer) time kernel
    ``cos(t * basis_freq + phase)``; ``basis_freq`` initialised to linspace(0, 9, d), ``phase`` to zeros."""

    def __init__(self, num_units, scope="coding/tif", device="cuda:0", basis_freq=None, phase=None):
        self._num_units = num_units
        self.basis_freq = (torch.from_numpy(np.linspace(0, 9, num_units).astype(np.float32)) if basis_freq is None
                           else basis_freq).to(device=device, dtype=torch.float32)
        self.phase = (torch.zeros(num_units) if phase is None else phase).to(device=device, dtype=torch.float32)

    def code(self, inputs):
        batch_size, seqslen = inputs.shape[0], inputs.shape[1]
        x = inputs.to(torch.float32).reshape(batch_size, seqslen, -1)   # coding.py:114 (tf.to_float + reshape)
        return _E.time_function_code(x, self.basis_freq, self.phase)     # [B, L, M, d] like tile + cos


class TimeSinusoidCoding(object):
    """coding.py:125-149."""

    def __init__(self, num_units):
        self.num_units = num_units
        self.scale = np.power(10000, np.arange(0, num_units, 2) * 1. / num_units).astype(np.float32)

    def code(self, inputs):
        assert inputs.dim() == 2, "the tensor rank should be 2."  # coding.py:139
        return _E.time_sinusoid_code(inputs.to(torch.float32), self.num_units)
