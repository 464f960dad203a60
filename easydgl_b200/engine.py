"""Thin host-side owner of one ``edgl_handle`` (include/easydgl_b200.h).

PyTorch tensors are containers only: this class allocates them, passes
``data_ptr()`` + the current CUDA stream to the C ABI, and keeps the weight tensors
alive (the library borrows them).  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import EdglConfig, check

_MODEL_ID = {"EasyDGL": _lib.EDGL_MODEL_EASYDGL, "CTSMA": _lib.EDGL_MODEL_CTSMA}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str, device=None) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise ValueError("%s must be a torch tensor" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if device is not None and t.device != device:
        raise ValueError("%s must live on %s (got %s)" % (name, device, t.device))
    return t.contiguous()


class Engine:
    """One model instance on one GPU.  ``cfg`` is a FLAGS-like namespace produced by
    ``easydgl_b200.synth.make_config`` (or the facade models); ``weights`` the nested
    dict of SURVEY.md 8(a-params) names."""

    def __init__(self, cfg, weights: Dict, max_batch: int, device="cuda:0", shard_rank: int = 0,
                 shard_world: int = 1, topk: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("easydgl_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.cfg = cfg
        self.K = int(topk if topk is not None else getattr(cfg, "topk", 100))
        self.L, self.d, self.h, self.E = cfg.L, cfg.num_units, cfg.num_heads, cfg.num_events
        self.ts_len = cfg.ts_len
        self.N1 = cfg.num_rows
        self.max_batch = int(max_batch)
        self.shard_rank, self.shard_world = shard_rank, shard_world
        per = (self.N1 + shard_world - 1) // shard_world
        self.c0 = min(per * shard_rank, self.N1)
        self.c1 = min(self.c0 + per, self.N1)
        self._handle = C.c_void_p()
        self._w = {}
        mark_rows = int(weights["mark_table"].shape[0])
        c = EdglConfig(model=_MODEL_ID[cfg.model], max_batch=self.max_batch, seq_len=self.L, num_units=self.d,
                       num_heads=self.h, num_blocks=cfg.num_blocks, num_events=self.E, num_rows=self.N1,
                       mark_rows=mark_rows, topk=self.K, time_scale=float(cfg.time_scale),
                       mask_id=int(cfg.mask_id), shard_rank=shard_rank, shard_world=shard_world)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_create(C.byref(c), C.byref(self._handle)))
        self.load_weights(weights)

    # ------------------------------------------------------------------ weights
    def load_weights(self, weights: Dict):
        with torch.cuda.device(self.device):
            for name, v in weights.items():
                if name == "blocks":
                    for i, blk in enumerate(v):
                        for n2, t in blk.items():
                            self._bind(n2, i, t)
                else:
                    self._bind(name, -1, v)
            check(self.lib.edgl_commit(self._handle, _stream()))

    def _bind(self, name: str, block: int, t: torch.Tensor):
        dt = torch.int64 if name == "mark_table" else torch.float32
        t = t.detach().to(device=self.device, dtype=dt).contiguous()
        self._w[(name, block)] = t  # borrowed by the library: keep alive
        check(self.lib.edgl_set_tensor(self._handle, name.encode(), block, t.data_ptr(), t.numel()))

    def close(self):
        if self._handle:
            self.lib.edgl_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ profiling
    def profile(self, enable: bool):
        check(self.lib.edgl_profile(self._handle, int(bool(enable))))

    def profile_read(self) -> Dict[str, Tuple[float, int]]:
        """{stage: (summed device ms, launches)} since the last read (CUDA events on the launch stream)."""
        n = self.lib.edgl_num_stages()
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        check(self.lib.edgl_profile_read(self._handle, ms, cnt, n))
        return {self.lib.edgl_stage_name(i).decode(): (ms[i], cnt[i]) for i in range(n) if cnt[i] > 0}

    # ------------------------------------------------------------------ model-level
    def _inputs(self, seqs_i, seqs_t) -> Tuple[torch.Tensor, torch.Tensor, int]:
        seqs_i = _req(seqs_i, torch.int64, "seqs_i", self.device)
        seqs_t = _req(seqs_t, torch.float32, "seqs_t", self.device)
        if seqs_i.dim() != 2 or seqs_i.shape[1] != self.L:
            raise ValueError("seqs_i must be [B,%d] (got %s)" % (self.L, tuple(seqs_i.shape)))
        if seqs_t.dim() != 2:
            raise AssertionError("the tensor rank should be 2.")  # coding.py:139
        if seqs_t.shape != (seqs_i.shape[0], self.ts_len):
            raise ValueError("seqs_t must be [B,%d] (got %s)" % (self.ts_len, tuple(seqs_t.shape)))
        return seqs_i, seqs_t, int(seqs_i.shape[0])

    def forward_logits(self, seqs_i, seqs_t) -> torch.Tensor:
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        out = torch.empty((B, self.c1 - self.c0), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_forward_logits(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B,
                                               out.data_ptr(), _stream()))
        return out

    def forward_topk(self, seqs_i, seqs_t, mask_seen: bool = True, out=None):
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        if out is None:
            idx = torch.empty((B, self.K), dtype=torch.int32, device=self.device)
            val = torch.empty((B, self.K), dtype=torch.float32, device=self.device)
        else:
            idx, val = out
        with torch.cuda.device(self.device):
            check(self.lib.edgl_forward_topk(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B,
                                             int(bool(mask_seen)), idx.data_ptr(), val.data_ptr(), _stream()))
        return idx, val

    def forward_topk_host(self, seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu, mask_seen: bool = True):
        """End-to-end call on HOST buffers (ideally pinned): H2D, forward, top-K, D2H, sync."""
        self._check_host(seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu)
        B = int(seqs_i_cpu.shape[0])
        with torch.cuda.device(self.device):
            check(self.lib.edgl_forward_topk_host(self._handle, seqs_i_cpu.data_ptr(), seqs_t_cpu.data_ptr(), B,
                                                  int(bool(mask_seen)), idx_cpu.data_ptr(), val_cpu.data_ptr(),
                                                  _stream()))
        return idx_cpu, val_cpu

    def forward_topk_host_submit(self, seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu, mask_seen: bool = True) -> int:
        """Asynchronous half of ``forward_topk_host``: enqueue upload, kernels and download of one batch and return
        its slot (0/1).  Up to two batches may be in flight; the host buffers must stay alive (and, to overlap, be
        pinned) until ``forward_topk_host_wait(slot)`` returns."""
        self._check_host(seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu)
        B = int(seqs_i_cpu.shape[0])
        with torch.cuda.device(self.device):
            slot = self.lib.edgl_forward_topk_host_submit(self._handle, seqs_i_cpu.data_ptr(), seqs_t_cpu.data_ptr(), B,
                                                          int(bool(mask_seen)), idx_cpu.data_ptr(), val_cpu.data_ptr(),
                                                          _stream())
        if slot < 0:
            check(slot)
        return slot

    def forward_topk_host_wait(self, slot: int):
        check(self.lib.edgl_forward_topk_host_wait(self._handle, int(slot)))

    def _check_host(self, seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu):
        B = int(seqs_i_cpu.shape[0])
        for t in (seqs_i_cpu, seqs_t_cpu, idx_cpu, val_cpu):
            if t.device.type != "cpu" or not t.is_contiguous():
                raise ValueError("forward_topk_host takes contiguous CPU tensors")
        if seqs_i_cpu.dtype != torch.int64 or seqs_t_cpu.dtype != torch.float32:
            raise ValueError("seqs_i must be int64 and seqs_t float32")
        if idx_cpu.dtype != torch.int32 or val_cpu.dtype != torch.float32:
            raise ValueError("idx must be int32 and val float32")
        if tuple(seqs_i_cpu.shape) != (B, self.L) or tuple(seqs_t_cpu.shape) != (B, self.ts_len):
            raise ValueError("bad input shapes")
        if idx_cpu.numel() < B * self.K or val_cpu.numel() < B * self.K:
            raise ValueError("output buffers too small")

    # ------------------------------------------------------------------ training-mode forward
    def _train_args(self, seqs_i, seqs_t, masked_positions):
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        if self.cfg.model == "EasyDGL":
            if masked_positions is None:
                raise ValueError("EasyDGL training needs features['masked_positions'] (dataloader.py:181-201)")
            pos = _req(masked_positions, torch.int64, "masked_positions", self.device)
            if pos.dim() != 2 or pos.shape[0] != B:
                raise ValueError("masked_positions must be [B, masklen]")
            return seqs_i, seqs_t, B, pos, int(pos.shape[1])
        if masked_positions is not None:
            raise ValueError("CTSMA predicts every position in training (CTSMA.py:82-83): no masked_positions")
        return seqs_i, seqs_t, B, None, self.L

    def forward_train_logits(self, seqs_i, seqs_t, masked_positions=None) -> torch.Tensor:
        """model(features, is_training=True) with dropout 0: logits [B*M, N] at the predicted positions."""
        seqs_i, seqs_t, B, pos, M = self._train_args(seqs_i, seqs_t, masked_positions)
        out = torch.empty((B * M, self.c1 - self.c0), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_forward_train_logits(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B, _ptr(pos), M,
                                                     out.data_ptr(), _stream()))
        return out

    def forward_train_loss(self, seqs_i, seqs_t, labels, masked_positions=None, l2_reg: float = 0.,
                           ct_reg: float = 0.) -> torch.Tensor:
        """The loss of model.train (dropout 0): tensor [4] = (loss, cross entropy, l2 term, continuous-time term)."""
        seqs_i, seqs_t, B, pos, M = self._train_args(seqs_i, seqs_t, masked_positions)
        labels = _req(labels, torch.int64, "labels", self.device)
        if tuple(labels.shape) != (B, M):
            raise ValueError("labels must be [B, %d] (got %s)" % (M, tuple(labels.shape)))
        out = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_forward_train_loss(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B, _ptr(pos),
                                                   labels.data_ptr(), M, float(l2_reg), float(ct_reg), out.data_ptr(),
                                                   _stream()))
        return out

    def encode(self, seqs_i, seqs_t) -> torch.Tensor:
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        y = torch.empty((B, self.d), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_encode(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B, y.data_ptr(), _stream()))
        return y

    def encode_packed(self, seqs_i, seqs_t, rows: torch.Tensor) -> torch.Tensor:
        """Encoder output written straight into the packed exchange rows ``rows`` [B, >= d + 2L] fp32:
        ``[y | seqs_i as raw bytes]`` per row (the all-gather message of the multi-GPU path)."""
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        if rows.dtype != torch.float32 or rows.device != self.device or rows.dim() != 2 or rows.stride(1) != 1 \
                or rows.shape[0] < B:
            raise ValueError("rows must be a 2-D fp32 tensor on %s with at least B rows" % self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_encode_packed(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B, rows.data_ptr(),
                                              int(rows.stride(0)), _stream()))
        return rows

    def logits_topk(self, y, seen_ids=None, out=None, out_stride=0):
        """Local top-K of this handle's item shard for rows ``y`` [Bt,d]; global column ids.
        ``y`` / ``seen_ids`` may be row-strided views (e.g. columns of the packed exchange buffer);
        ``out=(idx, val)`` may be views into one interleaved buffer with ``out_stride`` elements per row."""
        def rows_view(t, dtype, name):
            if not torch.is_tensor(t) or t.dtype != dtype or t.device != self.device or t.dim() != 2:
                raise ValueError("%s must be a 2-D %s tensor on %s" % (name, dtype, self.device))
            if t.stride(1) != 1:
                t = t.contiguous()
            return t, int(t.stride(0))
        y, ys = rows_view(y, torch.float32, "y")
        if ys % 4 != 0 or y.data_ptr() % 16 != 0:
            y = y.contiguous()
            ys = int(y.stride(0))
        Bt = int(y.shape[0])
        seen_len, ss = 0, 0
        if seen_ids is not None:
            seen_ids, ss = rows_view(seen_ids, torch.int64, "seen_ids")
            seen_len = int(seen_ids.shape[1])
        if out is None:
            idx = torch.empty((Bt, self.K), dtype=torch.int32, device=self.device)
            val = torch.empty((Bt, self.K), dtype=torch.float32, device=self.device)
        else:
            idx, val = out
        with torch.cuda.device(self.device):
            check(self.lib.edgl_logits_topk(self._handle, y.data_ptr(), ys, _ptr(seen_ids), seen_len, ss, Bt,
                                            out_stride, idx.data_ptr(), val.data_ptr(), _stream()))
        return idx, val

    # ------------------------------------------------------------------ layer-level
    def embed(self, seqs_i, seqs_t):
        seqs_i, seqs_t, B = self._inputs(seqs_i, seqs_t)
        width = (3 if self.cfg.model == "EasyDGL" else 2) * self.d
        X0 = torch.empty((B, self.L, width), dtype=torch.float32, device=self.device)
        spans = torch.empty((B, self.L), dtype=torch.float32, device=self.device)
        marks = torch.empty((B, self.L, self.E), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_embed(self._handle, seqs_i.data_ptr(), seqs_t.data_ptr(), B, X0.data_ptr(),
                                      spans.data_ptr(), marks.data_ptr(), _stream()))
        return X0, spans, marks

    def attention_layer(self, block, queries, keys, kmask, intervals, marks, causality=False, want_lam=True,
                        no_diag=False):
        queries = _req(queries, torch.float32, "queries", self.device)
        B, L, Cq = queries.shape
        if L != self.L:
            raise ValueError("queries must be [B,%d,C]" % self.L)
        keys_c = None if keys is None else _req(keys, torch.float32, "keys", self.device)
        kmask = _req(kmask, torch.uint8, "kmask", self.device)
        intervals = _req(intervals, torch.float32, "intervals", self.device)
        marks = _req(marks, torch.uint8, "marks", self.device)
        out = torch.empty((B, L, self.d), dtype=torch.float32, device=self.device)
        lam = torch.empty((self.h * B, L, self.E), dtype=torch.float32, device=self.device) if want_lam else None
        with torch.cuda.device(self.device):
            check(self.lib.edgl_attention_layer(self._handle, block, queries.data_ptr(), Cq, _ptr(keys_c),
                                                0 if keys_c is None else keys_c.shape[2], kmask.data_ptr(),
                                                intervals.data_ptr(), marks.data_ptr(), B,
                                                int(bool(causality)) | (2 if no_diag else 0),
                                                out.data_ptr(), _ptr(lam), _stream()))
        return out, lam

    def intensity(self, block, H, intervals, marks):
        H = _req(H, torch.float32, "H", self.device)
        intervals = _req(intervals, torch.float32, "intervals", self.device)
        marks = _req(marks, torch.uint8, "marks", self.device)
        hB, L, dh = H.shape
        B = hB // self.h
        G = torch.empty((hB, L, L), dtype=torch.float32, device=self.device)
        lam = torch.empty((hB, L, self.E), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.edgl_intensity(self._handle, block, H.data_ptr(), intervals.data_ptr(), marks.data_ptr(),
                                          B, G.data_ptr(), lam.data_ptr(), _stream()))
        return G, lam


# ---------------------------------------------------------------------- handle-free ops
def topk_merge(cand_val: torch.Tensor, cand_idx: torch.Tensor):
    """[G,Bt,K] per-shard candidates -> merged (idx [Bt,K] int32, val [Bt,K]); ties -> lower id."""
    cand_val = _req(cand_val, torch.float32, "cand_val")
    cand_idx = _req(cand_idx, torch.int32, "cand_idx")
    G, Bt, K = cand_val.shape
    return topk_merge_raw(cand_val.data_ptr(), cand_idx.data_ptr(), G, Bt, K, 0, 0, cand_val.device)


def topk_merge_raw(val_ptr: int, idx_ptr: int, G: int, Bt: int, K: int, shard_stride: int, row_stride: int, device,
                   out=None):
    """Merge with explicit base pointers / strides (used on the packed exchange buffer).  ``out=(idx, val)``:
    contiguous [Bt, K] int32 / fp32 tensors to write into."""
    lib = _lib.load()
    if out is None:
        idx = torch.empty((Bt, K), dtype=torch.int32, device=device)
        val = torch.empty((Bt, K), dtype=torch.float32, device=device)
    else:
        idx, val = out
        if not (idx.is_contiguous() and val.is_contiguous()):
            raise ValueError("merge outputs must be contiguous")
    with torch.cuda.device(device):
        check(lib.edgl_topk_merge(val_ptr, idx_ptr, G, Bt, K, shard_stride, row_stride, idx.data_ptr(), val.data_ptr(),
                                  _stream()))
    return idx, val


def time_sinusoid_code(ts: torch.Tensor, num_units: int) -> torch.Tensor:
    lib = _lib.load()
    if ts.dim() != 2:
        raise AssertionError("the tensor rank should be 2.")  # coding.py:139
    ts = _req(ts, torch.float32, "ts")
    B, L = ts.shape
    out = torch.empty((B, L, num_units), dtype=torch.float32, device=ts.device)
    with torch.cuda.device(ts.device):
        check(lib.edgl_time_sinusoid_code(ts.data_ptr(), B, L, num_units, out.data_ptr(), _stream()))
    return out


def time_function_code(x: torch.Tensor, basis_freq: torch.Tensor, phase: torch.Tensor) -> torch.Tensor:
    """cos(x[..., None] * basis_freq + phase) -> x.shape + (d,)  (C.TimeFunctionCoding, coding.py:112-122)."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    basis_freq = _req(basis_freq, torch.float32, "basis_freq", x.device)
    phase = _req(phase, torch.float32, "phase", x.device)
    d = int(basis_freq.numel())
    out = torch.empty(tuple(x.shape) + (d,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.edgl_time_function_code(x.data_ptr(), basis_freq.data_ptr(), phase.data_ptr(), x.numel(), d,
                                          out.data_ptr(), _stream()))
    return out


def row_nonzero(x: torch.Tensor) -> torch.Tensor:
    """tf.sign(tf.reduce_sum(tf.abs(x), -1)) as uint8 (key / query masks, temporal.py:65, 87)."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    out = torch.empty(tuple(x.shape[:-1]), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.edgl_row_nonzero(x.data_ptr(), x.numel() // x.shape[-1], int(x.shape[-1]), out.data_ptr(), _stream()))
    return out


def layernorm_last(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """module.normalize.layernorm (normalize.py:9-19): last-axis moments."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    gamma = _req(gamma, torch.float32, "gamma", x.device)
    beta = _req(beta, torch.float32, "beta", x.device)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.edgl_layernorm_last(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), x.numel() // x.shape[-1],
                                      int(x.shape[-1]), float(eps), out.data_ptr(), _stream()))
    return out


def time_attention(Q, K, V, num_heads, key_mask=None, query_mask=None, pos_k=None, pos_v=None, time_mode=0,
                   intervals=None, time_k=None, time_v=None, basis_freq=None, phase=None, U=None, residual=None,
                   causality=False):
    """Attention core of the Ti / Tf / Tg layers (edgl_time_attention).  Returns out [B,Tq,C] (and TC for mode 3)."""
    lib = _lib.load()
    Q = _req(Q, torch.float32, "Q")
    dev = Q.device
    K = _req(K, torch.float32, "K", dev)
    V = _req(V, torch.float32, "V", dev)
    B, Tq, C = Q.shape
    Tk = int(K.shape[1])

    def opt(t, dtype, name):
        return None if t is None else _req(t, dtype, name, dev)
    key_mask = opt(key_mask, torch.uint8, "key_mask")
    query_mask = opt(query_mask, torch.uint8, "query_mask")
    pos_k, pos_v = opt(pos_k, torch.float32, "pos_k"), opt(pos_v, torch.float32, "pos_v")
    time_k, time_v = opt(time_k, torch.float32, "time_k"), opt(time_v, torch.float32, "time_v")
    basis_freq, phase = opt(basis_freq, torch.float32, "basis_freq"), opt(phase, torch.float32, "phase")
    U, residual = opt(U, torch.float32, "U"), opt(residual, torch.float32, "residual")
    if time_mode != 0:
        intervals = _req(intervals, torch.int64 if time_mode == 1 else torch.float32, "intervals", dev)
        if tuple(intervals.shape) != (B, Tq, Tk):
            raise ValueError("intervals must be [B,Tq,Tk]")
    out = torch.empty_like(Q)
    TC = torch.empty((B, Tq, num_heads, C), dtype=torch.float32, device=dev) if time_mode == 3 else None
    vocab = int(time_k.shape[0]) if time_k is not None else 0
    with torch.cuda.device(dev):
        check(lib.edgl_time_attention(Q.data_ptr(), K.data_ptr(), V.data_ptr(), _ptr(key_mask), _ptr(query_mask),
                                      _ptr(pos_k), _ptr(pos_v), int(time_mode), _ptr(intervals), _ptr(time_k),
                                      _ptr(time_v), vocab, _ptr(basis_freq), _ptr(phase), _ptr(U), _ptr(TC),
                                      _ptr(residual), B, Tq, Tk, C, int(num_heads), int(bool(causality)),
                                      out.data_ptr(), _stream()))
    return (out, TC) if time_mode == 3 else out


def embedding_lookup(table: torch.Tensor, ids: torch.Tensor, zero_pad: bool, scale: bool) -> torch.Tensor:
    lib = _lib.load()
    table = _req(table, torch.float32, "table")
    ids = _req(ids, torch.int64, "ids", table.device)
    vocab, d = table.shape
    out = torch.empty(tuple(ids.shape) + (d,), dtype=torch.float32, device=table.device)
    with torch.cuda.device(table.device):
        check(lib.edgl_embedding_lookup(table.data_ptr(), vocab, d, int(zero_pad), int(scale), ids.data_ptr(),
                                        ids.numel(), out.data_ptr(), _stream()))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    if x.dim() != 3:
        raise ValueError("layernorm expects [B,L,C]")
    gamma = _req(gamma, torch.float32, "gamma", x.device)
    beta = _req(beta, torch.float32, "beta", x.device)
    B, L, Cc = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.edgl_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, L, Cc, out.data_ptr(), _stream()))
    return out


def dense(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act: int = 0) -> torch.Tensor:
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    w = _req(w, torch.float32, "w", x.device)
    K, N = w.shape
    M = x.numel() // K
    out = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=torch.float32, device=x.device)
    bb = None if b is None else _req(b, torch.float32, "b", x.device)
    with torch.cuda.device(x.device):
        check(lib.edgl_dense(x.data_ptr(), w.data_ptr(), _ptr(bb), M, K, N, act, out.data_ptr(), _stream()))
    return out


def dense_nk(x: torch.Tensor, wt: torch.Tensor, b: Optional[torch.Tensor], act: int = 0, f16: bool = False) -> torch.Tensor:
    """dense layer with the kernel stored K-major (wt [N,K]): the tcgen05 tensor-core GEMM (3xTF32, or the scaled
    3xFP16 split with f16=True)."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x")
    wt = _req(wt, torch.float32, "wt", x.device)
    N, K = wt.shape
    M = x.numel() // K
    out = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=torch.float32, device=x.device)
    bb = None if b is None else _req(b, torch.float32, "b", x.device)
    with torch.cuda.device(x.device):
        fn = lib.edgl_dense_nk_f16 if f16 else lib.edgl_dense_nk
        check(fn(x.data_ptr(), wt.data_ptr(), _ptr(bb), M, K, N, act, out.data_ptr(), _stream()))
    return out


def topk(logits: torch.Tensor, k: int, seen_ids: Optional[torch.Tensor] = None):
    """Sequential.eval ranking on given logits (modified in place when seen_ids is given)."""
    lib = _lib.load()
    logits = _req(logits, torch.float32, "logits")
    B, N = logits.shape
    seen_len = 0
    if seen_ids is not None:
        seen_ids = _req(seen_ids, torch.int64, "seen_ids", logits.device)
        seen_len = int(seen_ids.shape[1])
    idx = torch.empty((B, k), dtype=torch.int32, device=logits.device)
    val = torch.empty((B, k), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        check(lib.edgl_topk(logits.data_ptr(), B, N, _ptr(seen_ids), seen_len, k, idx.data_ptr(), val.data_ptr(),
                            _stream()))
    return idx, val


def launch_count() -> int:
    return int(_lib.load().edgl_launch_count())
