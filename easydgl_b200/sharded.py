"""Column-sharded multi-GPU ranking: one process per GPU.

Each rank encodes its OWN batch of sequences (the encoder is per-sample, so the batch is
data-parallel with no exchange) and owns the item-table rows / logit columns
``[rank*ceil(N1/G), (rank+1)*ceil(N1/G))``.  Two exchange steps:

1. all-gather of the packed rows ``[y (d fp32) | seqs_i (L int64, for the seen-mask)]``
   so that every rank can score every sequence against its column shard;
2. all-gather of the per-shard top-K candidates ``(global idx int32, val fp32)`` - the
   "single NCCL all-gather of per-shard top-K" of the north star - followed by a local
   K-way merge (ties -> lower global id, Base.py:181) of this rank's own rows.

The reference has no multi-GPU path at all; parity is defined as: the merged
top-K equals the single-GPU top-K bit for bit (tests/test_gpu_model.py,
tests/test_sharded_gloo.py).

``torch.distributed`` is plumbing only.  ``engine`` is an ``easydgl_b200.engine.Engine``
created with ``shard_rank=rank, shard_world=world``; the tests substitute a stand-in with
the same four methods to exercise this file under gloo on CPU.
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
    def __init__(self, engine, group=None, merge_fn=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
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
        y_all = allp[:, :d].contiguous()
        seen_all = allp[:, d:].contiguous().view(torch.int64) if mask_seen else None
        # ---- local: logits of ALL G*B rows against this rank's column shard + local top-K
        cand = self._buf("cand", (2, G * B, K), torch.int32, dev)      # plane 0 = idx, plane 1 = val bits
        eng.logits_topk(y_all, seen_all, out=(cand[0], cand[1].view(torch.float32)))
        # ---- exchange 2: all-gather of the per-shard top-K
        allc = self._buf("allcand", (G, 2, G * B, K), torch.int32, dev)
        dist.all_gather_into_tensor(allc.view(G * 2, G * B, K), cand, group=self.group)  # concat along dim 0
        # ---- merge this rank's rows [rank*B, (rank+1)*B) out of every shard's candidate block
        return self.merge_fn(allc, self.rank * B, B)


def _merge_packed_cuda(allc: torch.Tensor, row0: int, B: int):
    """allc int32 [G, 2, Bt, K] (plane 0 idx, plane 1 val bits) -> merged (idx, val) of rows row0..row0+B.
    Shard g's block starts g * (2*Bt*K) elements after shard 0's: edgl_topk_merge's shard_stride."""
    from .engine import topk_merge_raw
    G, _, Bt, K = allc.shape
    idx_base = allc[0, 0, row0]
    val_base = allc[0, 1, row0]
    return topk_merge_raw(val_base.data_ptr(), idx_base.data_ptr(), G, B, K, 2 * Bt * K, allc.device)
