"""Column-sharded multi-GPU ranking: one process per GPU.

Each rank encodes its OWN batch of sequences (the encoder is per-sample, so the batch is
data-parallel with no exchange) and owns the item-table rows / logit columns
``[rank*ceil(N1/G), (rank+1)*ceil(N1/G))``.  Two exchange steps:

1. all-gather of the packed rows ``[y (d fp32) | seqs_i (L int64, for the seen-mask)]``
   so that every rank can score every sequence against its column shard;
2. exchange of the per-shard top-K candidates ``(global idx int32 | val fp32)`` (one interleaved
   buffer), followed by a local K-way merge (ties -> lower global id, Base.py:181) of this
   rank's own rows.  ``exchange="all_to_all"`` (default) sends every rank only the candidates of
   ITS rows (1/G of the bytes); ``exchange="all_gather"`` is the literal "all-gather of per-shard
   top-K" of the north star and leaves every rank holding every row's candidates.  Both produce
   the same result.

The reference has no multi-GPU path at all; parity is defined as: the merged
top-K equals the single-GPU top-K bit for bit (tests/test_gpu_model.py,
tests/test_sharded_gloo.py, tools/check_sharded_nccl.py).

``torch.distributed`` is plumbing only.  ``engine`` is an ``easydgl_b200.engine.Engine``
created with ``shard_rank=rank, shard_world=world``; the tests substitute a stand-in with
the same methods to exercise this file under gloo on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(num_rows: int, rank: int, world: int):
    """Same partition as edgl_create (api.cu): ceil-div blocks, last one short."""
    per = (num_rows + world - 1) // world
    c0 = min(per * rank, num_rows)
    return c0, min(c0 + per, num_rows)


class ShardedRanker:
    def __init__(self, engine, group=None, merge_fn=None, exchange="all_to_all"):
        if exchange not in ("all_to_all", "all_gather"):
            raise ValueError("exchange must be 'all_to_all' or 'all_gather'")
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.exchange = exchange
        self.merge_fn = merge_fn if merge_fn is not None else _merge_packed_cuda
        self._bufs = {}

    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = b
        return b

    def forward_topk(self, seqs_i: torch.Tensor, seqs_t: torch.Tensor, mask_seen: bool = True):
        """Top-K (global item ids) for THIS rank's sequences. seqs_i [B,L] int64, seqs_t [B,ts_len]."""
        eng, G = self.engine, self.world
        B, L = seqs_i.shape
        d, K = eng.d, eng.K
        dev = seqs_i.device
        # ---- exchange 1: packed [y | ids] rows (ids travel as raw bytes in fp32 lanes)
        y = eng.encode(seqs_i, seqs_t)
        W = d + 2 * L
        mine = self._buf("pack", (B, W), torch.float32, dev)
        mine[:, :d].copy_(y)
        mine[:, d:].copy_(seqs_i.contiguous().view(torch.float32))
        allp = self._buf("allpack", (G * B, W), torch.float32, dev)
        dist.all_gather_into_tensor(allp, mine, group=self.group)
        # strided views into the gathered rows (no copies): y = first d floats, ids = the remaining 2L floats
        y_all = allp[:, :d]
        seen_all = allp.view(torch.int64)[:, d // 2:] if (mask_seen and d % 2 == 0) else (
            allp[:, d:].contiguous().view(torch.int64) if mask_seen else None)
        # ---- local: logits of ALL G*B rows against this rank's column shard + local top-K, written
        #      interleaved: row r = [idx(K) | val bits(K)], rows of destination rank j are contiguous
        cand = self._buf("cand", (G * B, 2, K), torch.int32, dev)
        eng.logits_topk(y_all, seen_all, out=(cand[:, 0], cand[:, 1].view(torch.float32)), out_stride=2 * K)
        # ---- exchange 2 + merge of this rank's rows
        if self.exchange == "all_to_all":
            recv = self._buf("recv", (G, B, 2, K), torch.int32, dev)       # recv[g] = shard g's candidates, my rows
            dist.all_to_all_single(recv.view(G * B, 2, K), cand, group=self.group)
            return self.merge_fn(recv, 0, B)
        allc = self._buf("allcand", (G, G * B, 2, K), torch.int32, dev)    # allc[g] = all rows of shard g
        dist.all_gather_into_tensor(allc.view(G * G * B, 2, K), cand, group=self.group)
        return self.merge_fn(allc, self.rank * B, B)


def _merge_packed_cuda(buf: torch.Tensor, row0: int, B: int):
    """buf int32 [G, rows, 2, K] (per row: idx | val bits) -> merged (idx, val) of rows row0..row0+B of every
    shard.  Shard stride = rows*2K elements, row stride = 2K: edgl_topk_merge's strides."""
    from .engine import topk_merge_raw
    G, rows, _, K = buf.shape
    idx_base = buf[0, row0, 0]
    val_base = buf[0, row0, 1]
    return topk_merge_raw(val_base.data_ptr(), idx_base.data_ptr(), G, B, K, rows * 2 * K, 2 * K, buf.device)
