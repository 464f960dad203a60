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
   top-K" of the north star and leaves every rank holding every row's candidates.
   ``exchange="p2p"`` fuses BOTH exchanges into the kernels: the packed rows and the top-K
   candidates are written straight into the peers' memory over NVLink (CUDA-IPC-mapped buffers,
   system-scope release/acquire flags) by ``edgl_xchg_put_rows`` and by the top-K kernel itself
   (``edgl_logits_topk_p2p``); no NCCL call remains on the data path.  All three give the same result.

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
    """``batch`` is the per-rank batch size every collective is sized for.  It must be the same on every rank,
    so it is agreed ONCE: either passed to the constructor or taken from the first call, and in both cases
    all-reduced (MAX) over the group - a collective every rank executes at the same point.  Later calls may pass
    any ``B <= batch`` (e.g. the short last batch of ``InputReader``, drop_remainder=False): the rows are padded
    with all-padding sequences (ids 0, ts 0) up to ``batch`` so that the exchanged buffers have identical
    shapes on every rank, and the result is sliced back to ``B``.  A larger ``B`` raises instead of issuing a
    size-mismatched collective (NCCL hang / wrong rows merged)."""

    def __init__(self, engine, group=None, merge_fn=None, exchange="all_to_all", batch=None):
        if exchange not in ("all_to_all", "all_gather", "p2p"):
            raise ValueError("exchange must be 'all_to_all', 'all_gather' or 'p2p'")
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.exchange = exchange
        self.merge_fn = merge_fn if merge_fn is not None else _merge_packed_cuda
        self._bufs = {}
        self._peer = None
        self.batch = None
        if batch is not None:
            self._agree_batch(int(batch), engine.device)

    def _agree_batch(self, B, device):
        t = torch.tensor([B], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.batch = int(t.item())
        mb = getattr(self.engine, "max_batch", None)
        if mb is not None and self.batch > mb:
            raise ValueError("agreed per-rank batch %d exceeds the engine's max_batch %d" % (self.batch, mb))

    def close(self):
        if self._peer is not None:
            self._peer.close()
            self._peer = None
        self._bufs = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buf(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = b
        return b

    def forward_topk(self, seqs_i: torch.Tensor, seqs_t: torch.Tensor, mask_seen: bool = True):
        """Top-K (global item ids) for THIS rank's sequences. seqs_i [B,L] int64, seqs_t [B,ts_len]."""
        Bin = int(seqs_i.shape[0])
        if self.batch is None:
            self._agree_batch(Bin, seqs_i.device)
        if Bin > self.batch:
            raise ValueError("batch %d exceeds the per-rank batch %d agreed for the exchange; build the "
                             "ShardedRanker with batch=<largest per-rank batch>" % (Bin, self.batch))
        if Bin < self.batch:  # pad with all-padding sequences; their candidates are dropped below
            pad = self.batch - Bin
            seqs_i = torch.cat([seqs_i, seqs_i.new_zeros((pad, seqs_i.shape[1]))])
            seqs_t = torch.cat([seqs_t, seqs_t.new_zeros((pad, seqs_t.shape[1]))])
        idx, val = self._forward_full(seqs_i, seqs_t, mask_seen)
        return (idx, val) if Bin == self.batch else (idx[:Bin], val[:Bin])

    def _forward_full(self, seqs_i, seqs_t, mask_seen):
        eng, G = self.engine, self.world
        B, L = seqs_i.shape
        dev = seqs_i.device
        if self.exchange == "p2p":
            if self._peer is None:
                self._peer = PeerExchange(eng, B, self.group)  # sized once for the agreed batch
            return self._peer.forward_topk(seqs_i, seqs_t, mask_seen)
        if dev.type != "cuda" or not hasattr(eng, "encode_packed"):
            return self._exchange_rows(seqs_i, seqs_t, mask_seen, 0, B, None)
        # Two micro-batches: the collectives and the shard-local ranking of micro-batch 0 run on a side stream while
        # the main stream already encodes micro-batch 1, so the NVLink transfers hide behind kernels.
        K = eng.K
        M = self.micro_batches if B >= 2 * 256 else 1
        bounds = [(m * B // M, (m + 1) * B // M) for m in range(M)]
        idx = torch.empty((B, K), dtype=torch.int32, device=dev)
        val = torch.empty((B, K), dtype=torch.float32, device=dev)
        if M == 1:
            self._exchange_rows(seqs_i, seqs_t, mask_seen, 0, B, (idx, val))
            return idx, val
        main = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        side.wait_stream(main)                     # buffers of the previous call are free again
        idx.record_stream(side)
        val.record_stream(side)
        for m, (r0, r1) in enumerate(bounds):
            mine = self._encode_rows(seqs_i[r0:r1], seqs_t[r0:r1], m)          # main stream
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                self._rank_rows(mine, mask_seen, r1 - r0, m, (idx[r0:r1], val[r0:r1]))
        main.wait_stream(side)
        return idx, val

    # micro-batches per call (EDGL_MICRO overrides; 1 = everything on the caller's stream)
    # Measured at 8 GPUs (C2, B=4096/GPU): 2 micro-batches 4.32 ms vs 4.22 ms for 1 - the two collectives are
    # latency-bound (43 MB and 26 MB over NVSwitch) and the kernels of the two streams slow each other down more
    # than the transfers cost, so the default is 1.
    micro_batches = int(__import__("os").environ.get("EDGL_MICRO", "1"))
    _side = None

    def _encode_rows(self, seqs_i, seqs_t, tag):
        """Encoder of this rank's rows into the packed message rows [y | ids]."""
        eng = self.engine
        B, L = seqs_i.shape
        W = eng.d + 2 * L
        mine = self._buf("pack%d" % tag, (B, W), torch.float32, seqs_i.device)
        if hasattr(eng, "encode_packed"):
            eng.encode_packed(seqs_i, seqs_t, mine)
        else:  # stand-in engines of the CPU tests
            mine[:, :eng.d].copy_(eng.encode(seqs_i, seqs_t))
            mine[:, eng.d:].copy_(seqs_i.contiguous().view(torch.float32))
        return mine

    def _rank_rows(self, mine, mask_seen, B, tag, out):
        """Exchange 1 (all-gather of the packed rows), shard-local logits + top-K of all G*B rows, exchange 2 (the
        candidates) and the merge of this rank's rows into ``out`` (or new tensors)."""
        eng, G = self.engine, self.world
        d, K = eng.d, eng.K
        dev = mine.device
        W = mine.shape[1]
        allp = self._buf("allpack%d" % tag, (G * B, W), torch.float32, dev)
        dist.all_gather_into_tensor(allp, mine, group=self.group)
        # strided views into the gathered rows (no copies): y = first d floats, ids = the remaining 2L floats
        y_all = allp[:, :d]
        seen_all = allp.view(torch.int64)[:, d // 2:] if (mask_seen and d % 2 == 0) else (
            allp[:, d:].contiguous().view(torch.int64) if mask_seen else None)
        # local: logits of ALL G*B rows against this rank's column shard + local top-K, written interleaved:
        # row r = [idx(K) | val bits(K)], rows of destination rank j are contiguous
        cand = self._buf("cand%d" % tag, (G * B, 2, K), torch.int32, dev)
        eng.logits_topk(y_all, seen_all, out=(cand[:, 0], cand[:, 1].view(torch.float32)), out_stride=2 * K)
        if self.exchange == "all_to_all":
            recv = self._buf("recv%d" % tag, (G, B, 2, K), torch.int32, dev)   # recv[g] = shard g's candidates, my rows
            dist.all_to_all_single(recv.view(G * B, 2, K), cand, group=self.group)
            return self._merge(recv, 0, B, out)
        allc = self._buf("allcand%d" % tag, (G, G * B, 2, K), torch.int32, dev)  # allc[g] = all rows of shard g
        dist.all_gather_into_tensor(allc.view(G * G * B, 2, K), cand, group=self.group)
        return self._merge(allc, self.rank * B, B, out)

    def _merge(self, buf, row0, B, out):
        if out is None:
            return self.merge_fn(buf, row0, B)
        try:
            return self.merge_fn(buf, row0, B, out)
        except TypeError:  # merge functions of the CPU tests take no output argument
            i, v = self.merge_fn(buf, row0, B)
            out[0].copy_(i)
            out[1].copy_(v)
            return out

    def _exchange_rows(self, seqs_i, seqs_t, mask_seen, r0, r1, out):
        mine = self._encode_rows(seqs_i[r0:r1], seqs_t[r0:r1], 0)
        return self._rank_rows(mine, mask_seen, r1 - r0, 0, out)


def _merge_packed_cuda(buf: torch.Tensor, row0: int, B: int, out=None):
    """buf int32 [G, rows, 2, K] (per row: idx | val bits) -> merged (idx, val) of rows row0..row0+B of every
    shard.  Shard stride = rows*2K elements, row stride = 2K: edgl_topk_merge's strides."""
    from .engine import topk_merge_raw
    G, rows, _, K = buf.shape
    idx_base = buf[0, row0, 0]
    val_base = buf[0, row0, 1]
    return topk_merge_raw(val_base.data_ptr(), idx_base.data_ptr(), G, B, K, rows * 2 * K, 2 * K, buf.device, out)


class PeerExchange:
    """Peer-memory exchange region of one rank + the mapped regions of its peers (see include/easydgl_b200.h,
    "fused exchange over peer memory").  NCCL is used once, at construction, to ship the 64-byte IPC handles."""

    def __init__(self, engine, B: int, group=None):
        import ctypes as C
        from ._lib import check
        self.eng, self.lib, self.B, self.group = engine, engine.lib, int(B), group
        self.G, self.rank = dist.get_world_size(group), dist.get_rank(group)
        G, d, L, K = self.G, engine.d, engine.L, engine.K
        self.W = d + 2 * L
        if self.W % 4 or d % 4:
            raise ValueError("p2p exchange needs d and d + 2L to be multiples of 4")
        self.off_rows = 0
        self.off_cand = self.off_rows + G * B * self.W * 4
        self.off_frows = (self.off_cand + G * B * 2 * K * 4 + 255) // 256 * 256
        self.off_fcand = self.off_frows + 256
        total = self.off_fcand + 256
        dev = engine.device
        with torch.cuda.device(dev):
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            check(self.lib.edgl_xchg_alloc(total, C.byref(ptr), handle))
            self.base = ptr.value
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
            allh = torch.empty((G, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
            allh = allh.cpu()
            self.peers, self._opened = [], []
            for g in range(G):
                if g == self.rank:
                    self.peers.append(self.base)
                    continue
                hb = (C.c_ubyte * 64).from_buffer_copy(bytes(allh[g].tolist()))
                pp = C.c_void_p()
                check(self.lib.edgl_xchg_open(hb, C.byref(pp)))
                self.peers.append(pp.value)
                self._opened.append(pp.value)

            def table(off):
                return torch.tensor([p + off for p in self.peers], dtype=torch.int64, device=dev)
            self.t_rows, self.t_cand = table(self.off_rows), table(self.off_cand)
            self.t_frows, self.t_fcand = table(self.off_frows), table(self.off_fcand)
        self.epoch = 0
        dist.barrier(group=group)

    def forward_topk(self, seqs_i, seqs_t, mask_seen=True):
        from ._lib import check
        from .engine import _stream, topk_merge_raw
        eng, lib, G, B, K, W, d, L = self.eng, self.lib, self.G, self.B, self.eng.K, self.W, self.eng.d, self.eng.L
        self.epoch += 1
        e = self.epoch
        seqs_i = seqs_i.contiguous()
        y = eng.encode(seqs_i, seqs_t)
        h = eng._handle
        with torch.cuda.device(eng.device):
            st = _stream()
            check(lib.edgl_xchg_put_rows(h, y.data_ptr(), 0, seqs_i.data_ptr(), B, self.t_rows.data_ptr(),
                                         self.t_frows.data_ptr(), G, self.rank, e, st))
            check(lib.edgl_xchg_wait(self.base + self.off_frows, G, e, st))
            rows = self.base + self.off_rows
            seen = rows + d * 4 if mask_seen else None
            check(lib.edgl_logits_topk_p2p(h, rows, W, seen, L, W // 2, G * B, B, self.t_cand.data_ptr(),
                                           self.t_fcand.data_ptr(), G, self.rank, e, st))
            check(lib.edgl_xchg_wait(self.base + self.off_fcand, G, e, st))
            cand = self.base + self.off_cand
            return topk_merge_raw(cand + K * 4, cand, G, B, K, B * 2 * K, 2 * K, eng.device)

    def close(self):
        for p in self._opened:
            self.lib.edgl_xchg_close(p)
        self._opened = []
        if self.base:
            self.lib.edgl_xchg_free(self.base)
            self.base = 0
