"""BASELINE.json's full sizes (C2: B=4096, L=100, d=128, 18K items; C3: CTSMA) through
size-independent properties, plus an oracle spot check on a random subset of rows."""
import pytest
import torch

from helpers import O, assert_close, case, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_full(name, B):
    from easydgl_b200 import engine
    from easydgl_b200.engine import Engine
    cfg = synth.named_config(name)
    inp = synth.make_inputs(cfg, B, edge_cases=True)
    W = synth.make_weights(cfg, mode="parity")
    eng = Engine(cfg, W, max_batch=B, device=DEV)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    idx, val = eng.forward_topk(ids, ts, True)
    idx2, val2 = eng.forward_topk(ids, ts, True)
    assert torch.equal(idx, idx2) and torch.equal(val, val2), "two runs must agree bit for bit"
    K = eng.K
    # sortedness: values non-increasing, ties broken by increasing index
    dv = val[:, 1:] - val[:, :-1]
    assert bool((dv <= 0).all())
    tie = dv == 0
    assert bool((idx[:, 1:][tie] > idx[:, :-1][tie]).all())
    # index sanity: in range, unique per row, never a seen id, never column 0
    assert int(idx.min()) >= 1 and int(idx.max()) < cfg.num_rows
    srt = torch.sort(idx, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate index in a top-K row"
    seen = torch.zeros((B, cfg.num_rows), dtype=torch.bool, device=DEV)
    seen.scatter_(1, ids, True)
    assert not bool(torch.gather(seen, 1, idx.long()).any()), "a seen item was ranked"
    # checksum of checksums: sharded (4 logical shards) merge == single
    y = eng.encode(ids, ts)
    ci, cv = [], []
    for r in range(4):
        sh = Engine(cfg, W, max_batch=B, device=DEV, shard_rank=r, shard_world=4)
        i, v = sh.logits_topk(y, ids)
        ci.append(i)
        cv.append(v)
        sh.close()
    mi, mv = engine.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, idx) and torch.equal(mv, val)
    # batch invariance on a slice
    sl = slice(1000, 1064)
    i3, v3 = eng.forward_topk(ids[sl].contiguous(), ts[sl].contiguous(), True)
    assert torch.equal(i3, idx[sl]) and torch.equal(v3, val[sl])
    # oracle spot check on 48 rows (fp64)
    g = torch.Generator().manual_seed(11)
    rows = torch.randperm(B, generator=g)[:48]
    rows[:4] = torch.arange(4)  # the edge-case rows
    ref = O.forward(inp["seqs_i"][rows], inp["seqs_t"][rows], W, cfg, dtype=torch.float64)
    ref32 = O.forward(inp["seqs_i"][rows], inp["seqs_t"][rows], W, cfg, dtype=torch.float32)
    logits = eng.forward_logits(ids[rows.to(DEV)].contiguous(), ts[rows.to(DEV)].contiguous()).cpu()
    well = (ref32.double() - ref)[:, 1:].abs().amax(1) <= 1e-4 * ref[:, 1:].abs().max()
    assert int(well.sum()) >= 44
    assert_close(logits[well, 1:], ref[well, 1:], 1e-3, name + " full-size logits (48 rows)")
    err = float((logits.double() - ref)[well].abs().max())
    res = O.topk_set_compare(idx[rows.to(DEV)].cpu().long()[well], O.mask_seen_logits(ref, inp["seqs_i"][rows])[well],
                             K, tau=4 * err)
    assert res["bad"] == 0, res
    return res


def test_c2_full_size_properties():
    _run_full("C2", 4096)


def test_c3_full_size_properties():
    _run_full("C3", 4096)
