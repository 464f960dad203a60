#!/usr/bin/env python
"""Time the per-row top-K kernels on many short rows (the column shards of the multi-GPU path):
    python tools/bench_topk_rows.py [rows] [cols] [K]      (CTA per row, first and second warp-per-row kernels)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easydgl_b200 import engine  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 2251
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(rows, cols, device="cuda", generator=g)
res = {}
for mode in ("0", "1", "default"):
    if mode == "default":
        os.environ.pop("EDGL_TOPK_WARP", None)
    else:
        os.environ["EDGL_TOPK_WARP"] = mode
    for _ in range(3):
        idx, val = engine.topk(x.clone(), K)
    torch.cuda.synchronize()
    xs = [x.clone() for _ in range(10)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in xs:
        idx, val = engine.topk(t, K)
    e1.record()
    torch.cuda.synchronize()
    res[mode] = (e0.elapsed_time(e1) / 10, idx.clone(), val.clone())
    print("EDGL_TOPK_WARP=%s: %.4f ms per launch (%d rows x %d columns, K=%d)" % (mode, res[mode][0], rows, cols, K))
print("identical:", all(bool(torch.equal(res["0"][1], res[m][1]) and torch.equal(res["0"][2], res[m][2])) for m in ("1", "default")))
