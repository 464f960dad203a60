"""Facade of ``src/model/Base.py``: ``layernorm`` (Base.py:12-67) and ``Sequential``
(Base.py:90-207) - ctor, ``__call__`` and ``eval`` keep the reference signatures; ``train`` is out
of scope and raises."""
from __future__ import annotations

import abc

import numpy as np
import torch

from .. import checkpoint
from .. import engine as _E
from .. import synth


def layernorm(inputs, gamma=None, beta=None, begin_norm_axis=1, begin_params_axis=-1, scope=None):
    """Base.layernorm: statistics over every axis but the batch axis; gamma/beta default to ones/zeros
    like freshly created TF variables (Base.py:36-49)."""
    if inputs.dim() != 3:
        raise ValueError("Inputs has unsupported rank %d (expected [B,L,C])" % inputs.dim())
    if begin_norm_axis != 1 or begin_params_axis not in (-1, 2):
        raise ValueError("begin_params_axis (%d) and begin_norm_axis (%d) must be -1 and 1" %
                         (begin_params_axis, begin_norm_axis))
    C = inputs.shape[-1]
    gamma = torch.ones(C, device=inputs.device) if gamma is None else gamma
    beta = torch.zeros(C, device=inputs.device) if beta is None else beta
    return _E.layernorm(inputs, gamma, beta)


class Sequential(object):
    _model = None

    def __init__(self, num_items, FLAGS):
        self.num_items = num_items
        self.num_units = FLAGS.num_units
        self.num_heads = FLAGS.num_heads
        self.hidden_dropout_rate = getattr(FLAGS, "hidden_dropout_rate", 0.)
        self.attention_probs_dropout_rate = getattr(FLAGS, "attention_probs_dropout_rate", 0.)
        self.seqslen = FLAGS.seqslen
        self.learning_rate = getattr(FLAGS, "learning_rate", None)
        self.l2_reg = getattr(FLAGS, "l2_reg", 0.)
        self.num_train_steps = getattr(FLAGS, "num_train_steps", None)
        self.num_warmup_steps = getattr(FLAGS, "num_warmup_steps", None)
        self._flags = FLAGS
        self._engine = None
        self._metric_sums = {}
        self._metric_n = 0

    # ---- weights / engine ----------------------------------------------------------------------
    def _load_mark_table(self, FLAGS, mark_table):
        if mark_table is not None:
            return torch.as_tensor(np.asarray(mark_table)).to(torch.int64)
        path = getattr(FLAGS, "mark", None)
        if path:
            return checkpoint.load_mark_table(path)  # EasyDGL.py:45
        return None

    def _setup(self, FLAGS, weights, mark_table, device, max_batch):
        mt = self._load_mark_table(FLAGS, mark_table)
        num_events = int(mt.shape[-1]) if mt is not None else int(getattr(FLAGS, "num_events", 16))
        self.cfg = synth.make_config(model=self._model, num_units=FLAGS.num_units, seqslen=FLAGS.seqslen,
                                     num_items=self._raw_items, num_heads=FLAGS.num_heads, num_blocks=FLAGS.num_blocks,
                                     num_events=num_events, time_scale=getattr(FLAGS, "time_scale", 1.0))
        self.num_events = num_events
        if weights is None:  # tf.get_variable default initialisers
            weights = synth.make_weights(self.cfg, mode="reference")
        weights = dict(weights)
        if mt is not None:
            weights["mark_table"] = mt
        self.mark_lookup_table = weights["mark_table"]
        self.weights = weights
        self.device = device
        self._max_batch = max_batch

    def _get_engine(self, B):
        if self._engine is None or self._engine.max_batch < B:
            if self._engine is not None:
                self._engine.close()
            self._engine = _E.Engine(self.cfg, self.weights, max_batch=max(B, self._max_batch), device=self.device)
        return self._engine

    def load_weights(self, weights):
        self.weights = dict(weights)
        if self._engine is not None:
            self._engine.load_weights(self.weights)

    def restore(self, ckpt, overrides=None):
        """``saver.restore(sess, FLAGS.ckpt)`` (analytics.py:83-88): load the model variables of a
        ``tf.train.Saver`` checkpoint (prefix, or a directory holding a ``checkpoint`` state file)."""
        W = checkpoint.load_checkpoint(self.cfg, ckpt, overrides)
        W["mark_table"] = self.mark_lookup_table
        self.load_weights(W)

    def save(self, ckpt=None):
        """``EarlyStopping.save_ckpt`` (util.py:53-55): write the weights as ``ckpt/{model}`` in Saver format."""
        ckpt = ckpt or "ckpt/%s" % self._model
        checkpoint.save_checkpoint(self.cfg, self.weights, ckpt)
        return ckpt

    # ---- reference protocol --------------------------------------------------------------------
    @abc.abstractmethod
    def __call__(self, features, is_training):
        raise NotImplementedError("the model is not implemented")  # Base.py:117

    def _check_no_dropout(self):
        if self.hidden_dropout_rate or self.attention_probs_dropout_rate:
            raise NotImplementedError(
                "training-mode forward with dropout > 0 is not implemented: TensorFlow's random stream "
                "(EasyDGL.py:92,114,126; temporal.py:442) cannot be reproduced; set both dropout rates to 0")

    def train(self, features, labels):
        """The forward half of ``Model.train`` (EasyDGL.py:153-189 / CTSMA.py:93-124 / Base.py:119-131): returns the
        scalar ``loss`` tensor (masked softmax cross entropy + l2 regulariser + ct_reg * TPP likelihood) and a dict of
        its parts.  The reference returns ``(train_op, loss_op, loss_init_op)``; the backward pass and the Adam step
        (Base.py:142-144) are outside this library's scope."""
        self._check_no_dropout()
        ids, ts = features['seqs_i'], features['seqs_t']
        eng = self._get_engine(ids.shape[0])
        out = eng.forward_train_loss(ids, ts, labels, features.get('masked_positions'), float(self.l2_reg or 0.),
                                     float(getattr(self, 'ct_reg', 0.) or 0.))
        return out[0], {"ce": out[1], "l2": out[2], "ct": out[3]}

    def _forward(self, features, is_training):
        ids, ts = features['seqs_i'], features['seqs_t']
        if is_training:
            # model(features, is_training=True): logits at the predicted positions, [B * masklen, N]
            # (EasyDGL.py:140-151) or [B * seqslen, N] (CTSMA.py:82-91); dropout rates must be 0
            self._check_no_dropout()
            return self._get_engine(ids.shape[0]).forward_train_logits(ids, ts, features.get('masked_positions'))
        return self._get_engine(ids.shape[0]).forward_logits(ids, ts)

    def eval(self, features, labels, mask_seen=True):
        """Base.py:150-207.  Returns (metrics, topk_idx): the running means of H@{10,50,100} and
        N@{10,50,100} over every batch since ``reset_metrics()`` (tf.metrics.mean semantics), and this
        batch's top-100 indices.

        Deviation from the reference: the ranking is taken on the seen-masked LOGITS, not on
        ``softmax(logits)`` (Base.py:164,181).  Softmax is monotone, so the order is the same wherever fp32
        softmax is injective; where probabilities underflow to 0 (> ~88 nats below the row maximum) the
        reference's ``top_k`` returns the lowest indices among the zeros, which is not reproduced.  The
        cut-offs 100/50/10 are the reference's constants, so the engine must rank at least 100 items."""
        ids, ts = features['seqs_i'], features['seqs_t']
        eng = self._get_engine(ids.shape[0])
        if eng.K < 100:
            raise ValueError("eval() reports H/N@100 (Base.py:181): the engine must be built with topk >= 100 "
                             "(got %d)" % eng.K)
        idx, _ = eng.forward_topk(ids, ts, mask_seen)
        idx = idx[:, :100]
        real = labels[:, -1:].to(idx.device).to(torch.int32)           # Base.py:169
        tp = (idx == real).to(torch.float64)                           # one-hot gather, Base.py:182-185
        gain = torch.tensor(1. / np.log2(np.arange(2, 100 + 2)), dtype=torch.float64, device=idx.device)
        B = ids.shape[0]
        for k in (100, 50, 10):
            t = tp[:, :k]
            self._metric_sums['H%d' % k] = self._metric_sums.get('H%d' % k, 0.) + float(torch.sign(t.sum(-1)).sum())
            self._metric_sums['N%d' % k] = self._metric_sums.get('N%d' % k, 0.) + float((t * gain[:k]).sum())
        self._metric_n += B
        return {k: v / self._metric_n for k, v in self._metric_sums.items()}, idx

    def reset_metrics(self):
        """metric_init_op (Base.py:204-206)."""
        self._metric_sums, self._metric_n = {}, 0
