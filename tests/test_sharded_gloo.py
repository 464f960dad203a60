"""world_size-2 gloo test (CPU) of the multi-GPU host logic in easydgl_b200/sharded.py.

The CUDA engine cannot run here, so a stand-in with the engine's four-method surface is built
from the oracle (test infrastructure standing in for the device, never a product fallback); what
is under test is the packing, the two all-gathers, the shard bounds and the merge addressing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShardEngine:
    def __init__(self, cfg, W, rank, world, O):
        from easydgl_b200.sharded import shard_bounds
        self.cfg, self.W, self.O = cfg, W, O
        self.d, self.K = cfg.num_units, 20
        self.c0, self.c1 = shard_bounds(cfg.num_rows, rank, world)
        self.table = O.zero_pad_table(W["item_embs"]).double()
        self.bias = O.output_bias(W["output_bias"]).double()

    def encode(self, ids, ts):
        return self.O.forward(ids, ts, self.W, self.cfg, dtype=torch.float64, return_all=True).y.float()

    def logits_topk(self, y_all, seen_all, out, out_stride=0):
        lg = y_all.double() @ self.table[self.c0:self.c1].t() + self.bias[self.c0:self.c1]
        if seen_all is not None:
            for b in range(lg.shape[0]):
                s = seen_all[b]
                s = s[(s >= self.c0) & (s < self.c1)] - self.c0
                lg[b, s] = float("-inf")
        v, i = self.O.topk_lower_index_first(lg, self.K)
        out[0].copy_((i + self.c0).to(torch.int32))
        out[1].copy_(v.float())


def _cpu_merge(buf, row0, B):
    G, rows, _, K = buf.shape
    idx = buf[:, row0:row0 + B, 0].permute(1, 0, 2).reshape(B, G * K).long()
    val = buf[:, row0:row0 + B, 1].contiguous().view(torch.float32).permute(1, 0, 2).reshape(B, G * K).double()
    # sort by (val desc, idx asc): stable sort on idx first, then on -val
    o1 = torch.sort(idx, dim=1, stable=True).indices
    idx, val = torch.gather(idx, 1, o1), torch.gather(val, 1, o1)
    o2 = torch.sort(-val, dim=1, stable=True).indices[:, :K]
    return torch.gather(idx, 1, o2).to(torch.int32), torch.gather(val, 1, o2).float()


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import easydgl_oracle as O
    from easydgl_b200 import synth
    from easydgl_b200.sharded import ShardedRanker
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.make_config(model="EasyDGL", num_units=16, seqslen=9, num_items=301, num_heads=2, num_blocks=1,
                                num_events=4)
        W = synth.make_weights(cfg, mode="parity")
        inp = synth.make_inputs(cfg, 5, seed=100 + rank, edge_cases=False)
        eng = OracleShardEngine(cfg, W, rank, world, O)
        ranker = ShardedRanker(eng, merge_fn=_cpu_merge, exchange="all_to_all")
        idx, val = ranker.forward_topk(inp["seqs_i"], inp["seqs_t"], mask_seen=True)
        ranker_ag = ShardedRanker(eng, merge_fn=_cpu_merge, exchange="all_gather")
        idx_ag, val_ag = ranker_ag.forward_topk(inp["seqs_i"], inp["seqs_t"], mask_seen=True)
        # single-device answer for this rank's rows
        full = OracleShardEngine(cfg, W, 0, 1, O)
        out = (torch.empty(5, 20, dtype=torch.int32), torch.empty(5, 20))
        full.logits_topk(full.encode(inp["seqs_i"], inp["seqs_t"]), inp["seqs_i"], out)
        ok = torch.equal(idx, out[0]) and torch.equal(val, out[1])
        ok = ok and torch.equal(idx_ag, out[0]) and torch.equal(val_ag, out[1])
        # and without the seen-mask
        idx2, _ = ranker.forward_topk(inp["seqs_i"], inp["seqs_t"], mask_seen=False)
        full.logits_topk(full.encode(inp["seqs_i"], inp["seqs_t"]), None, out)
        ok = ok and torch.equal(idx2, out[0])
        # ragged batches (InputReader's short last batch): rank 1 passes 3 rows while rank 0 passes 5; the ranker
        # pads to the batch agreed on the first call, so the collectives keep identical shapes on both ranks
        nb = 5 if rank == 0 else 3
        idx3, val3 = ranker.forward_topk(inp["seqs_i"][:nb], inp["seqs_t"][:nb], mask_seen=True)
        out3 = (torch.empty(nb, 20, dtype=torch.int32), torch.empty(nb, 20))
        full.logits_topk(full.encode(inp["seqs_i"][:nb], inp["seqs_t"][:nb]), inp["seqs_i"][:nb], out3)
        ok = ok and idx3.shape[0] == nb and torch.equal(idx3, out3[0]) and torch.equal(val3, out3[1])
        try:  # a batch larger than the agreed one must raise, not issue a mismatched collective
            big = torch.cat([inp["seqs_i"], inp["seqs_i"]]), torch.cat([inp["seqs_t"], inp["seqs_t"]])
            ranker.forward_topk(big[0], big[1])
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_ranker_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == [(0, True), (1, True)], res


def test_shard_bounds_partition_the_columns():
    from easydgl_b200.sharded import shard_bounds
    for n in (1, 7, 100, 18001, 1000001):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= a <= b <= n for a, b in spans)
