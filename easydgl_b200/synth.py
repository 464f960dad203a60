"""Synthetic Netflix-schema inputs, mark tables and weights.

The reference ships neither its dataset nor ``mark.pkl`` nor a checkpoint, so every
run here uses tensors generated with the conventions of the reference's own ETL:

* ``data/linkpred.py:142-157`` - sequences are right-aligned and left-padded with 0
  (ids and timestamps alike), truncated to the last ``seqslen+1`` events;
* ``src/dataloader.py:166-179`` - EasyDGL eval replaces the LAST token by ``[MASK]``
  (= ``FLAGS.num_items``); ``src/dataloader.py:95-99`` - CTSMA gets ``tokens[:-1]`` but
  keeps all ``seqslen+1`` timestamps;
* seed 9876 is the reference's (``src/main.py:157``).

Plain numpy + torch CPU tensors; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

SEED = 9876  # src/main.py:157

# BASELINE.json configs.  num_items is FLAGS.num_items.
CONFIGS = {
    "C1": dict(model="EasyDGL", num_units=64, seqslen=99, num_items=18000, batch=32, num_heads=8,
               num_blocks=1, num_events=16),
    "C2": dict(model="EasyDGL", num_units=128, seqslen=99, num_items=18000, batch=4096, num_heads=8,
               num_blocks=1, num_events=16),
    "C3": dict(model="CTSMA", num_units=64, seqslen=100, num_items=18000, batch=4096, num_heads=4,
               num_blocks=2, num_events=16),
    "C4": dict(model="EasyDGL", num_units=128, seqslen=199, num_items=100000, batch=8192, num_heads=8,
               num_blocks=1, num_events=16),
    "C5": dict(model="EasyDGL", num_units=256, seqslen=511, num_items=1000000, batch=16384, num_heads=8,
               num_blocks=1, num_events=16),
}


def make_config(model="EasyDGL", num_units=64, seqslen=30, num_items=1000, num_heads=1, num_blocks=1,
                num_events=16, time_scale=86400.0, batch=None, topk=100, **extra) -> SimpleNamespace:
    """A FLAGS-like namespace (src/main.py:22-75 names) plus the derived sizes the
    reference computes in its ctors (EasyDGL.py:39-41, CTSMA.py:23-32)."""
    cfg = SimpleNamespace(model=model, num_units=num_units, seqslen=seqslen, num_items=num_items,
                          num_heads=num_heads, num_blocks=num_blocks, num_events=num_events,
                          time_scale=float(time_scale), batch_size=batch, topk=topk,
                          hidden_dropout_rate=0.0, attention_probs_dropout_rate=0.0, masklen=6,
                          learning_rate=5e-4, l2_reg=0.0, ct_reg=0.0, num_train_steps=None,
                          num_warmup_steps=None, mask_seen=True, mark=None)
    for k, v in extra.items():
        setattr(cfg, k, v)
    if model == "EasyDGL":
        cfg.L = seqslen + 1           # EasyDGL.py:40
        cfg.ts_len = cfg.L
        cfg.num_rows = num_items + 1  # EasyDGL.py:41
        cfg.mask_id = num_items       # EasyDGL.py:39
    elif model == "CTSMA":
        cfg.L = seqslen
        cfg.ts_len = seqslen + 1      # dataloader.py:99
        cfg.num_rows = num_items
        cfg.mask_id = -1
    else:
        raise NotImplementedError("The ranking model: {0} not implemented".format(model))
    return cfg


def named_config(name: str, **over) -> SimpleNamespace:
    kw = dict(CONFIGS[name])
    kw.update(over)
    return make_config(**kw)


def _zipf_ids(rng, n, num_ids, alpha=1.05):
    """Zipf(alpha) over 1..num_ids via inverse CDF (bounded support)."""
    w = 1.0 / np.power(np.arange(1, num_ids + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    return (np.searchsorted(cdf, rng.random(n)) + 1).astype(np.int64)


def make_inputs(cfg, batch: int, seed: int = SEED, min_len: int = 5, edge_cases: bool = False):
    """Returns dict(seqs_i int64 [B,L], seqs_t float32 [B,ts_len], labels int64 [B]) on CPU."""
    rng = np.random.default_rng(seed)
    T = cfg.ts_len                       # tokens stored per example (= seqslen+1)
    B = batch
    min_len = min(min_len, T)
    lens = rng.integers(min_len, T + 1, size=B)
    if edge_cases and B >= 4:
        lens[0] = T                      # full history
        lens[1] = 1                      # only the target token
        lens[2] = 0                      # all padding (empty history)
        lens[3] = 2
    tokens = np.zeros((B, T), dtype=np.int64)
    times = np.zeros((B, T), dtype=np.float32)
    flat = _zipf_ids(rng, int(lens.sum()), cfg.num_items - 1)
    start = rng.uniform(9.4e8, 1.10e9, size=B)
    pos = 0
    for b in range(B):
        n = int(lens[b])
        if n == 0:
            continue
        tokens[b, T - n:] = flat[pos:pos + n]
        pos += n
        inc = rng.exponential(3 * 86400.0, size=n)
        inc[0] = 0.0
        times[b, T - n:] = (start[b] + np.cumsum(inc)).astype(np.float32)
    labels = tokens[:, -1].copy()
    if cfg.model == "EasyDGL":
        seqs_i = tokens.copy()
        # dataloader.py:166-169 mask_last: the last slot becomes [MASK] unconditionally
        seqs_i[:, -1] = cfg.mask_id
        seqs_t = times
    else:
        seqs_i = tokens[:, :-1].copy()   # dataloader.py:98
        seqs_t = times                   # dataloader.py:99 keeps all seqslen+1 timestamps
    return dict(seqs_i=torch.from_numpy(seqs_i), seqs_t=torch.from_numpy(seqs_t),
                labels=torch.from_numpy(labels))


def make_train_inputs(cfg, batch: int, seed: int = SEED, masklen: int = None):
    """Training-mode features / labels with the reference's post-processors.  EasyDGL (MAUPostProcessor.mask_random,
    dataloader.py:181-201): `masklen` distinct positions in 1..T-1 become [MASK], labels = the tokens there (0 on
    padding: weight 0 in the loss); CTSMA (RegressivePostProcessor, dataloader.py:95-98): seqs_i = tokens[:-1],
    labels = tokens[1:].  Returns dict(seqs_i, seqs_t, labels[, masked_positions])."""
    rng = np.random.default_rng(seed + 7)
    base = make_inputs(cfg, batch, seed=seed)
    if cfg.model != "EasyDGL":
        # make_inputs built tokens [B, S+1]; seqs_i = tokens[:, :-1], labels_last = tokens[:, -1]
        tokens = torch.cat([base["seqs_i"], base["labels"].unsqueeze(1)], dim=1)
        return dict(seqs_i=tokens[:, :-1].contiguous(), seqs_t=base["seqs_t"], labels=tokens[:, 1:].contiguous())
    T = cfg.ts_len
    M = int(masklen if masklen is not None else getattr(cfg, "masklen", 6))
    tokens = base["seqs_i"].clone()
    tokens[:, -1] = base["labels"]                        # undo mask_last: the raw sequence
    pos = np.stack([rng.choice(T - 1, M, replace=False) + 1 for _ in range(batch)]).astype(np.int64)  # dataloader.py:34-36
    pos = torch.from_numpy(pos)
    labels = torch.gather(tokens, 1, pos)
    masked = tokens.clone()
    masked.scatter_(1, pos, cfg.mask_id)
    return dict(seqs_i=masked, seqs_t=base["seqs_t"], labels=labels, masked_positions=pos)


def make_mark_table(cfg, seed: int = SEED, onehot: bool = False) -> torch.Tensor:
    """int64 [num_items, E] multi-hot item->event-mark table standing in for mark.pkl
    (EasyDGL.py:45): row 0 (padding) all-zero, every item 1-3 marks (exactly 1 if onehot)."""
    rng = np.random.default_rng(seed + 1)
    N, E = cfg.num_items, cfg.num_events
    tab = np.zeros((N, E), dtype=np.int64)
    nm = np.ones(N, dtype=np.int64) if onehot else rng.integers(1, min(3, E) + 1, size=N)
    for j in range(3):
        cols = rng.integers(0, E, size=N)
        sel = nm > j
        tab[np.nonzero(sel)[0], cols[sel]] = 1
    tab[0] = 0
    return torch.from_numpy(tab)


def _glorot(rng, *shape):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[-2], shape[-1])
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return torch.from_numpy(rng.uniform(-lim, lim, size=shape).astype(np.float32))


def make_weights(cfg, seed: int = SEED, mode: str = "reference", onehot_marks: bool = False) -> dict:
    """Weights as a nested dict of fp32 CPU tensors, names per SURVEY.md section 8(a-params).

    mode="reference": the reference's initialisers (glorot-uniform tables/kernels,
    N(0,0.02) BiMAU QKVT kernel temporal.py:393, zeros for biases/scaling/output_bias/beta,
    ones for gamma).  mode="parity": every bias / gamma / beta / scaling / output_bias is
    randomised and kernels are scaled up so that no term of the forward pass is vacuous
    (attention scores O(1), logits spread O(1))."""
    rng = np.random.default_rng(seed + 2)
    d, h, E, L, N1 = cfg.num_units, cfg.num_heads, cfg.num_events, cfg.L, cfg.num_rows
    dh = d // h
    par = mode == "parity"

    def vec(n, fill=0.0, std=0.1):
        if par:
            return torch.from_numpy((fill + std * rng.standard_normal(n)).astype(np.float32))
        return torch.full((n,), fill, dtype=torch.float32)

    def kern(i, o, gain=1.0):
        return _glorot(rng, i, o) * (gain if par else 1.0)

    def normal(i, o, std):
        return torch.from_numpy((std * rng.standard_normal((i, o))).astype(np.float32))

    W = {"mark_table": make_mark_table(cfg, seed, onehot_marks)}
    if par:
        W["item_embs"] = normal(N1, d, 1.0 / np.sqrt(d))
        W["pos_embs"] = normal(L, d, 0.5)
    else:
        W["item_embs"] = _glorot(rng, N1, d)
        W["pos_embs"] = _glorot(rng, L, d)
    W["output_bias"] = vec(N1 - 1, 0.0, 0.1)

    def intensity_w():
        return {"int_w": kern(dh + 1, dh * E, 2.0), "int_b": vec(dh * E), "int_weight": kern(E, dh, 2.0),
                "int_scaling": torch.from_numpy(rng.uniform(-0.5, 0.5, E).astype(np.float32)) if par
                else torch.zeros(E)}

    blocks = []
    if cfg.model == "EasyDGL":
        W["mark_embs"] = normal(E, d, 0.3) if par else _glorot(rng, E, d)
        for i in range(cfg.num_blocks):
            cin = 3 * d if i == 0 else d
            blk = {"qkvt_w": normal(cin, 4 * d, (1.0 / np.sqrt(cin)) if par else 0.02), "qkvt_b": vec(4 * d)}
            blk.update(intensity_w())
            blk.update({"ao_w": kern(d, d, 1.5), "ao_b": vec(d), "ao_ln_g": vec(d, 1.0), "ao_ln_b": vec(d),
                        "ff1_w": kern(d, 2 * d, 1.5), "ff1_b": vec(2 * d),
                        "ff2_w": kern(2 * d, d, 1.5), "ff2_b": vec(d),
                        "ff_ln_g": vec(d, 1.0), "ff_ln_b": vec(d)})
            blocks.append(blk)
        W.update({"tr_w": kern(d, d, 1.5), "tr_b": vec(d), "tr_ln_g": vec(d, 1.0), "tr_ln_b": vec(d)})
    else:
        for i in range(cfg.num_blocks):
            cin = 2 * d if i == 0 else d
            blk = {"ln1_g": vec(cin, 1.0), "ln1_b": vec(cin)}
            for nm in ("q", "k", "v", "t"):
                blk[nm + "_w"] = kern(cin, d, 2.0)
                blk[nm + "_b"] = vec(d)
            blk.update(intensity_w())
            blk.update({"ln2_g": vec(d, 1.0), "ln2_b": vec(d),
                        "ff1_w": kern(d, d, 1.5), "ff1_b": vec(d), "ff2_w": kern(d, d, 1.5), "ff2_b": vec(d)})
            blocks.append(blk)
        W.update({"out_ln_g": vec(d, 1.0), "out_ln_b": vec(d)})
    W["blocks"] = blocks
    return W
