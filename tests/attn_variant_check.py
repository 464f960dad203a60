"""Layer- and model-level parity of ONE attention-kernel variant (selected with EDGL_ATTN, read once per
process) against the fp64 oracle, including operands rescaled over a wide dynamic range.

Run as a script on a GPU box (``EDGL_ATTN=f16 python tests/attn_variant_check.py``); the pytest wrapper is
``test_gpu_model.py::test_attention_variant_layers``.  Prints one JSON line per case and ``VARIANT_OK`` when
every case is inside its tolerance.  The oracle is the checker only.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import O, case, rel_err  # noqa: E402
from easydgl_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"
TOL = float(os.environ.get("EDGL_VARIANT_TOL", "2e-5"))  # fp32-level; the north-star bar is 1e-3


def scaled_weights(W, cfg, q=1.0, k=1.0, v=1.0, t=1.0, mlp=1.0):
    """Rescale the Q/K/V/T projections (and the intensity MLP) of block 0: changes operand magnitudes by
    orders of magnitude while the oracle computes the same function of the new weights."""
    import copy
    W = copy.deepcopy(W)
    d = cfg.num_units
    blk = W["blocks"][0]
    if cfg.model == "EasyDGL":
        for j, s in enumerate((q, k, v, t)):
            blk["qkvt_w"][:, j * d:(j + 1) * d] *= s
            blk["qkvt_b"][j * d:(j + 1) * d] *= s
    else:
        for n, s in (("q", q), ("k", k), ("v", v), ("t", t)):
            blk[n + "_w"] *= s
            blk[n + "_b"] *= s
    blk["int_w"][:-1] *= mlp
    return W


def layer_case(name, batch, causal, scales, **over):
    cfg, inp, W = case(name, batch=batch, **over)
    W = scaled_weights(W, cfg, **scales)
    eng = Engine(cfg, W, max_batch=batch, device=DEV)
    W64 = O._cast(W, torch.float64)
    blk = W64["blocks"][0]
    if cfg.model == "EasyDGL":
        X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
        rO, rl = O.bimau(X0, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events)
        out, lam = eng.attention_layer(0, X0.float().to(DEV), None, kmask.to(torch.uint8).to(DEV),
                                       spans.float().to(DEV), marks.to(torch.uint8).to(DEV))
    else:
        X, kmask, spans, marks = O.ctsma_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
        qin = O.layernorm(X, blk["ln1_g"], blk["ln1_b"])
        rO, rl = O.mau(qin, X, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, causal)
        out, lam = eng.attention_layer(0, qin.float().to(DEV), X.float().to(DEV), kmask.to(torch.uint8).to(DEV),
                                       spans.float().to(DEV), marks.to(torch.uint8).to(DEV), causality=causal)
    fin = bool(torch.isfinite(out).all() and torch.isfinite(lam).all())
    return {"case": name, "causal": causal, "scales": scales, "out_err": rel_err(out, rO), "lam_err": rel_err(lam, rl),
            "finite": fin}


def model_case(name, batch):
    cfg, inp, W = case(name, batch=batch)
    eng = Engine(cfg, W, max_batch=batch, device=DEV)
    lg = eng.forward_logits(inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)).cpu()
    ref = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float64)
    err = rel_err(lg[:, 1:], ref[:, 1:])
    idx, _ = eng.forward_topk(inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV), mask_seen=True)
    masked = O.mask_seen_logits(ref, inp["seqs_i"])
    abs_err = float((lg[:, 1:].double() - ref[:, 1:]).abs().max())
    res = O.topk_set_compare(idx.cpu().long(), masked, idx.shape[1], tau=4 * abs_err)
    return {"case": name, "logits_err": err, "topk": {k: (int(v) if isinstance(v, (int, bool)) else v)
                                                       for k, v in res.items() if not torch.is_tensor(v)}}


def main():
    ok = True
    unit = dict(q=1.0, k=1.0, v=1.0, t=1.0, mlp=1.0)
    layer_runs = [
        ("ctsma_b", 6, True, unit, {}), ("ctsma_b", 6, False, unit, {}),
        ("C2", 4, False, unit, {}), ("C3", 4, True, unit, {}),
        ("C2", 3, False, dict(q=2.0 ** -20, k=2.0 ** 20, v=2.0 ** 12, t=2.0 ** -9, mlp=2.0 ** 9), {}),
        ("C2", 3, False, dict(q=2.0 ** 30, k=2.0 ** -30, v=2.0 ** -40, t=2.0 ** 6, mlp=2.0 ** -6), {}),
        ("C3", 3, True, dict(q=37.0, k=1.0 / 37.0, v=1.0e-6, t=3.0e3, mlp=1.0 / 3.0e3), {}),
        ("C2", 3, False, unit, dict(seqslen=119)), ("C2", 2, False, unit, dict(seqslen=199)),
        # head dim 32, L = 512 (C5's shape; attn_f16_long.cu by default), also over a wide operand range, and a
        # length that is neither a multiple of the 64-key chunk nor of the 256-row query block
        ("C5", 2, False, unit, dict(num_items=3000)),
        ("C5", 1, False, dict(q=2.0 ** -20, k=2.0 ** 20, v=2.0 ** 12, t=2.0 ** -9, mlp=2.0 ** 9), dict(num_items=3000)),
        ("C5", 2, False, unit, dict(num_items=3000, seqslen=300)),
    ]
    for name, batch, causal, scales, over in layer_runs:
        r = layer_case(name, batch, causal, scales, **over)
        r["over"] = over
        bad = (not r["finite"]) or r["out_err"] > TOL or r["lam_err"] > TOL
        r["ok"] = not bad
        ok &= not bad
        print(json.dumps(r), flush=True)
    for name in ("ctsma_b", "C2", "C3"):
        r = model_case(name, 8)
        bad = r["logits_err"] > 1e-4 or r["topk"].get("bad", 0) != 0
        r["ok"] = not bad
        ok &= not bad
        print(json.dumps(r), flush=True)
    print("VARIANT_OK" if ok else "VARIANT_FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
