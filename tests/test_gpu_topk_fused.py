"""Fused logits + seen-mask + top-K (candidate filter in the epilogue of the logits GEMM, api.cu logits_topk) against
the materialised path (logits -> mask_seen -> top-K): idx AND val must be identical bit for bit, for a single table, for
column shards, for rows that overflow the candidate list (ties / constant rows: predicated fallback) and for rows with
ids of every kind in the seen list.  Reference semantics: EasyDGL.py:149-150, Base.py:156-181."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import case  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _engine(cfg, W, B, **kw):
    from easydgl_b200.engine import Engine
    return Engine(cfg, W, max_batch=B, device=DEV, **kw)


def _both(monkeypatch, fn):
    monkeypatch.setenv("EDGL_TOPK_FUSE", "0")
    ref = fn()
    monkeypatch.setenv("EDGL_TOPK_FUSE", "1")
    got = fn()
    torch.cuda.synchronize()
    return ref, got


@pytest.mark.parametrize("name,B,items", [("C2", 300, 18000), ("C3", 129, 18000), ("C2", 64, 40000)])
def test_fused_equals_materialised_forward(name, B, items, monkeypatch):
    cfg, inp, W = case(name, batch=B, num_items=items)
    eng = _engine(cfg, W, B)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    for mask in (True, False):
        (ri, rv), (gi, gv) = _both(monkeypatch, lambda: tuple(t.clone() for t in eng.forward_topk(ids, ts, mask)))
        assert torch.equal(ri, gi) and torch.equal(rv.view(torch.int32), gv.view(torch.int32)), (name, mask)
    eng.close()


@pytest.mark.parametrize("sample", [0, 4096])
def test_fused_column_shards_and_merge(sample, monkeypatch):
    """Column shards (global ids, seen-mask restricted to the shard) and the K-way merge."""
    from easydgl_b200 import engine
    if sample:
        monkeypatch.setenv("EDGL_TOPK_SAMPLE", str(sample))
    B = 96
    cfg, inp, W = case("C2", batch=B, num_items=36000)
    full = _engine(cfg, W, B)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    y = full.encode(ids, ts).clone()
    monkeypatch.setenv("EDGL_TOPK_FUSE", "0")
    idx, val = (t.clone() for t in full.forward_topk(ids, ts, True))
    monkeypatch.setenv("EDGL_TOPK_FUSE", "1")
    ci, cv = [], []
    for r in range(3):
        sh = _engine(cfg, W, B, shard_rank=r, shard_world=3)
        i, v = sh.logits_topk(y, ids)
        ci.append(i.clone())
        cv.append(v.clone())
        sh.close()
    mi, mv = engine.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, idx) and torch.equal(mv, val)
    full.close()


def test_fused_overflow_rows_fall_back(monkeypatch):
    """Rows on which the sample threshold is useless - y = 0 makes every logit equal to its bias (zeros: one 18000-way
    tie), and a row pointing along one table row - must come out exactly as the materialised path gives them."""
    B = 40
    cfg, inp, W = case("C2", batch=B, num_items=18000, mode="reference")  # reference initialisers: output_bias = 0
    eng = _engine(cfg, W, B)
    ids = inp["seqs_i"].to(DEV)
    y = eng.encode(ids, inp["seqs_t"].to(DEV)).clone()
    y[3] = 0.0
    y[17] = 0.0
    y[21] = W["item_embs"][777].to(DEV) * 50.0
    (ri, rv), (gi, gv) = _both(monkeypatch, lambda: tuple(t.clone() for t in eng.logits_topk(y, ids)))
    assert torch.equal(ri, gi) and torch.equal(rv.view(torch.int32), gv.view(torch.int32))
    # ties -> lower index first, seen ids skipped (Base.py:156-163,181): check the all-tie row against that rule
    assert float(W["output_bias"].abs().max()) == 0.0
    seen = set(ids[3].tolist())
    want = [c for c in range(1, 400) if c not in seen][:eng.K]  # column 0 is -1000 (Base.py:110)
    assert gi[3].tolist() == want
    eng.close()


def test_fused_path_is_taken(monkeypatch):
    """The fused path must not write the [B, N1] logits: its launch count differs from the materialised path's, and
    the predicated fallback launches leave no trace in the result."""
    from easydgl_b200 import engine
    B = 64
    cfg, inp, W = case("C2", batch=B, num_items=18000)
    eng = _engine(cfg, W, B)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    counts = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("EDGL_TOPK_FUSE", mode)
        eng.forward_topk(ids, ts, True)
        n0 = engine.launch_count()
        eng.forward_topk(ids, ts, True)
        counts[mode] = engine.launch_count() - n0
    assert counts["1"] > counts["0"], counts
    eng.close()
