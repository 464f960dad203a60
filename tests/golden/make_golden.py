"""Generates the committed golden fixtures from the fp64 CPU oracle.

PARITY UNPINNED: the reference cannot run here (TensorFlow 2.3.4 absent) and
ships no golden vectors, so these fixtures pin the ORACLE (a regression pin for both the
oracle and the CUDA path), not the reference binary.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import easydgl_oracle as O  # noqa: E402
from easydgl_b200 import synth  # noqa: E402

CASES = {
    "easydgl_d32_L13_h4_b2": dict(model="EasyDGL", num_units=32, seqslen=12, num_items=150, num_heads=4,
                                  num_blocks=2, num_events=8),
    "easydgl_d64_L31_h4_b1": dict(model="EasyDGL", num_units=64, seqslen=30, num_items=300, num_heads=4,
                                  num_blocks=1, num_events=16),
    "ctsma_d32_S14_h2_b2": dict(model="CTSMA", num_units=32, seqslen=14, num_items=150, num_heads=2,
                                num_blocks=2, num_events=8),
}

if __name__ == "__main__":
    for name, kw in CASES.items():
        cfg = synth.make_config(**kw)
        inp = synth.make_inputs(cfg, 6, seed=4242, edge_cases=True)
        W = synth.make_weights(cfg, seed=4242, mode="parity")
        r = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float64, return_all=True)
        k = 20
        _, idx = O.eval_topk(r.logits, inp["seqs_i"], True, k, rank_on="probs")
        torch.save(dict(cfg=kw, seqs_i=inp["seqs_i"], seqs_t=inp["seqs_t"], weights=W, logits64=r.logits,
                        y64=r.y, X064=r.X0, lam64=r.lams[0], topk_idx=idx), os.path.join(HERE, name + ".pt"))
        print(name, os.path.getsize(os.path.join(HERE, name + ".pt")))
