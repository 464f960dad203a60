#!/usr/bin/env python
"""Precision ablation at the headline size: which of the three tensor-core products of each
contraction (A_lo B_hi, A_hi B_lo, A_hi B_hi) can be dropped?  For every setting the C2 forward (B = 4096 rows, parity
weights) runs on the GPU and is compared with the fp64 oracle on ALL rows:

  logits error (max |a - b| / max |ref|), and the number of rows whose top-100 SET differs from the fp64 oracle's.

The fp32 oracle (the "reference CPU path") goes through the same comparison: its count is the noise floor an fp32
implementation has against fp64.  Switches: EDGL_ABL_GEMM / EDGL_ABL_ATTN (easydgl_b200/csrc/common.cuh), read per
launch.  The oracle is the checker only.      python tools/ablation.py [rows] > profiles/r2_ablation.md"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import easydgl_oracle as O  # noqa: E402
from easydgl_b200 import synth  # noqa: E402
from easydgl_b200.engine import Engine  # noqa: E402

DEV = "cuda:0"
SETTINGS = [
    ("default: every contraction 3 products", {}),
    ("attention S = Q K^T: hi hi only", {"EDGL_ABL_ATTN": "3"}),
    ("attention P T, (G o P) V: P_lo dropped", {"EDGL_ABL_ATTN": "4"}),
    ("attention P T, (G o P) V: T_lo / V_lo dropped", {"EDGL_ABL_ATTN": "8"}),
    ("attention P T, (G o P) V: hi hi only", {"EDGL_ABL_ATTN": "12"}),
    ("intensity MLP H W1: hi hi only", {"EDGL_ABL_ATTN": "48"}),
    ("gate G = lam M^T: lam_hi only", {"EDGL_ABL_ATTN": "64"}),
    ("attention: every product hi hi only (plain fp16)", {"EDGL_ABL_ATTN": "127"}),
    ("QKVT dense: A_lo dropped", {"EDGL_ABL_GEMM": "100000"}),
    ("QKVT dense: W_lo dropped", {"EDGL_ABL_GEMM": "200000"}),
    ("QKVT dense: hi hi only (plain fp16)", {"EDGL_ABL_GEMM": "300000"}),
    ("attention-out dense: hi hi only (plain TF32)", {"EDGL_ABL_GEMM": "030000"}),
    ("FF1 dense: hi hi only", {"EDGL_ABL_GEMM": "003000"}),
    ("FF2 dense: hi hi only", {"EDGL_ABL_GEMM": "000300"}),
    ("transform dense: hi hi only", {"EDGL_ABL_GEMM": "000030"}),
    ("logits dense: A_lo dropped", {"EDGL_ABL_GEMM": "000001"}),
    ("logits dense: W_lo dropped", {"EDGL_ABL_GEMM": "000002"}),
    ("logits dense: hi hi only (plain TF32)", {"EDGL_ABL_GEMM": "000003"}),
    ("all six dense layers: hi hi only", {"EDGL_ABL_GEMM": "333333"}),
    ("everything hi hi only (1 product everywhere)", {"EDGL_ABL_GEMM": "333333", "EDGL_ABL_ATTN": "127"}),
]


def topk_sets(masked, k=100):
    return torch.topk(masked, k, dim=1).indices.sort(dim=1).values


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cfg = synth.named_config("C2")
    inp = synth.make_inputs(cfg, B, edge_cases=True)
    W = synth.make_weights(cfg, mode="parity")
    t0 = time.time()
    ref64, ref32 = [], []
    for r0 in range(0, B, 256):
        sl = slice(r0, min(B, r0 + 256))
        ref64.append(O.forward(inp["seqs_i"][sl], inp["seqs_t"][sl], W, cfg, dtype=torch.float64))
        ref32.append(O.forward(inp["seqs_i"][sl], inp["seqs_t"][sl], W, cfg, dtype=torch.float32))
    ref64, ref32 = torch.cat(ref64), torch.cat(ref32)
    scale = float(ref64[:, 1:].abs().max())
    want = topk_sets(O.mask_seen_logits(ref64, inp["seqs_i"]))
    # gap between the 100th and 101st unmasked fp64 logit: rows below ~4x the logit error are decided by rounding
    srt = torch.sort(O.mask_seen_logits(ref64, inp["seqs_i"]), dim=1, descending=True).values
    gap = (srt[:, 99] - srt[:, 100])

    def row(name, logits):
        err = float((logits.double() - ref64)[:, 1:].abs().max())
        got = topk_sets(O.mask_seen_logits(logits.double(), inp["seqs_i"]))
        diff = (got != want).any(dim=1)
        n = int(diff.sum())
        near = int((gap[diff] < 4 * err).sum()) if n else 0
        return {"setting": name, "logits_rel_err": err / scale, "rows_with_other_top100_set": n,
                "of_which_gap_below_4x_err": near, "rows": B}
    recs = [row("fp32 CPU oracle (reference CPU path) vs fp64", ref32)]
    sys.stderr.write("oracles: %.1f s\n" % (time.time() - t0))
    eng = Engine(cfg, W, max_batch=B, device=DEV)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    for name, env in SETTINGS:
        for k in ("EDGL_ABL_GEMM", "EDGL_ABL_ATTN"):
            os.environ.pop(k, None)
        os.environ.update(env)
        logits = eng.forward_logits(ids, ts).cpu()
        recs.append(row(name, logits))
        recs[-1]["env"] = env
        sys.stderr.write(json.dumps(recs[-1]) + "\n")
    for k in ("EDGL_ABL_GEMM", "EDGL_ABL_ATTN"):
        os.environ.pop(k, None)
    print("| setting | logits max err / max\\|ref\\| | rows (of %d) whose top-100 set differs from the fp64 oracle's | of which the fp64 gap at the cut is < 4x the logit error |" % B)
    print("|---|---|---|---|")
    for r in recs:
        print("| %s | %.2e | %d | %d |" % (r["setting"], r["logits_rel_err"], r["rows_with_other_top100_set"],
                                        r["of_which_gap_below_4x_err"]))
    with open(os.path.join(ROOT, "gpurun_out", "r2_ablation.jsonl"), "w") as fh:
        for r in recs:
            fh.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
