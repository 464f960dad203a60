"""Scaled 3xFP16 dense layers inside the model pipeline: parity per EDGL_F16_MASK (bit 0 QKVT, 1 attention-out, 2 FF1,
3 FF2, 4 transform, 5 logits; FF2 needs FF1's published maximum, so bit 3 is only tested together with bit 2)."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import O, case, rel_err
from easydgl_b200.engine import Engine
cfg, inp, W = case("C2", batch=8)
eng = Engine(cfg, W, max_batch=8, device="cuda:0")
lg = eng.forward_logits(inp["seqs_i"].cuda(), inp["seqs_t"].cuda()).cpu()
ref = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float64)
fin = bool(torch.isfinite(lg).all())
print("finite", fin, "err", rel_err(lg[:, 1:], ref[:, 1:]) if fin else float("nan"))
'''
for mask in [int(a) for a in sys.argv[1:]] or (0, 1, 2, 12, 63):
    env = dict(os.environ, EDGL_F16_MASK=str(mask))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    print("mask", mask, (r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip().splitlines()[-1:])
