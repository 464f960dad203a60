"""Weights in and out of the reference's on-disk formats, without TensorFlow.

The reference persists a trained model with ``tf.train.Saver(max_to_keep=1).save(sess, "ckpt/{model}")``
(util.py:26,53-55) and reloads it with ``saver.restore(sess, FLAGS.ckpt)`` (analytics.py:83-88); the
item -> event-mark table comes from ``pickle.load(open(FLAGS.mark,'rb')).toarray()`` (EasyDGL.py:45), a
scipy CSR matrix.  This module provides

* ``read_tensor_bundle(prefix)`` / ``write_tensor_bundle(prefix, tensors)`` - the Saver-V2 "tensor bundle":
  ``prefix.index`` is a leveldb-format table (prefix-compressed key/value blocks, 5-byte block trailers with
  a masked CRC32C, 48-byte footer ending in magic 0xdb4775248b80fb57) whose key "" holds a
  ``BundleHeaderProto`` and whose other keys are variable names holding ``BundleEntryProto``
  {dtype, shape, shard_id, offset, size, crc32c}; ``prefix.data-00000-of-0000N`` are the raw little-endian
  tensor bytes.
* ``tf_variable_names(cfg)`` - the name / shape every parameter of SURVEY 8(a-params) has in the
  reference's graph (read off the ``variable_scope`` nesting, see the table in the function), and
  ``import_weights`` / ``export_weights`` between that naming and this package's weight dict.
* ``load_checkpoint`` / ``save_checkpoint`` - the two composed; ``latest_checkpoint(dir)`` reads the
  ``checkpoint`` state file; ``load_mark_table`` / ``save_mark_table`` for ``mark.pkl``.

PARITY UNPINNED: TensorFlow is not installable in this image and the reference ships no checkpoint, so the
container format is restated from TensorFlow's published tensor_bundle / table format and the variable
names from reading the reference's scopes; neither has been checked against a file written by TensorFlow.
``import_weights`` therefore matches by scope *suffix* and shape, reports every candidate when a name is
missing or ambiguous, and accepts explicit ``overrides``.
"""
from __future__ import annotations

import os
import pickle
import re
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from .dataloader import _enc_varint, _fields, _ld, _varint, crc32c

_MAGIC = 0xdb4775248b80fb57
_FOOTER_LEN = 48
_RESTART_INTERVAL = 16
_BLOCK_SIZE = 262144

# tensorflow/core/framework/types.proto
_DT_TO_NP = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
             9: np.int64, 10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_NP_TO_DT = {np.dtype(v): k for k, v in _DT_TO_NP.items()}


def _mask(c: int) -> int:
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _unmask(m: int) -> int:
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- snappy (block decode only)
def _snappy_uncompress(buf: bytes) -> bytes:
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):  # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch (%d != %d)" % (len(out), n))
    return bytes(out)


# ----------------------------------------------------------------------------- leveldb-format table
def _read_block(data: bytes, offset: int, size: int, verify: bool) -> bytes:
    if offset + size + 5 > len(data):
        raise ValueError("table block handle (%d,%d) runs past the end of the file" % (offset, size))
    body, kind = data[offset:offset + size], data[offset + size]
    if verify:
        want = _unmask(struct.unpack_from("<I", data, offset + size + 1)[0])
        if crc32c(data[offset:offset + size + 1]) != want:
            raise ValueError("table block at %d: crc32c mismatch" % offset)
    if kind == 0:
        return body
    if kind == 1:
        return _snappy_uncompress(body)
    raise ValueError("table block at %d: unknown compression type %d" % (offset, kind))


def _block_entries(block: bytes) -> Iterable[Tuple[bytes, bytes]]:
    if len(block) < 4:
        raise ValueError("table block too short")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise ValueError("table block: bad restart count")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise ValueError("table block: corrupt entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _handle(buf: bytes, pos: int = 0) -> Tuple[int, int, int]:
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def read_table(path: str, verify_crc: bool = True) -> List[Tuple[bytes, bytes]]:
    """All (key, value) pairs of a leveldb-format table file, in key order."""
    data = open(path, "rb").read()
    if len(data) < _FOOTER_LEN:
        raise ValueError("%s: too short to be a table file" % path)
    footer = data[-_FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
        raise ValueError("%s: bad table magic (not a tensor-bundle index)" % path)
    _, _, pos = _handle(footer, 0)          # metaindex (unused: no filter block in bundles)
    ioff, isize, _ = _handle(footer, pos)
    out = []
    for _, hv in _block_entries(_read_block(data, ioff, isize, verify_crc)):
        boff, bsize, _ = _handle(hv)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify_crc)))
    return out


class _BlockBuilder:
    def __init__(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % _RESTART_INTERVAL == 0:
            if self.count:
                self.restarts.append(len(self.buf))
        else:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_size: int = _BLOCK_SIZE):
    """Write sorted (key, value) pairs as an uncompressed leveldb-format table (what BundleWriter emits)."""
    keys = [k for k, _ in items]
    if keys != sorted(keys) or len(set(keys)) != len(keys):
        raise ValueError("table keys must be unique and sorted")
    out = bytearray()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block)
        out.append(0)  # kNoCompression
        out.extend(struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        return _enc_varint(off) + _enc_varint(len(block))

    index = _BlockBuilder()
    cur = _BlockBuilder()
    for k, v in items:
        cur.add(k, v)
        if cur.size() >= block_size:
            index.add(cur.last, emit(cur.finish()))
            cur = _BlockBuilder()
    if cur.count or not items:
        index.add(cur.last, emit(cur.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out.extend(footer)
    with open(path, "wb") as fh:
        fh.write(out)


# ----------------------------------------------------------------------------- tensor bundle
def _shape_proto(shape) -> bytes:
    return b"".join(_ld(2, _enc_varint((1 << 3) | 0) + _enc_varint(int(s))) for s in shape)


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for fno, _, v in _fields(buf):
        if fno == 2:
            size = 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    size = v2 - (1 << 64) if v2 >> 63 else v2
            dims.append(size)
        elif fno == 3 and v:
            raise ValueError("tensor of unknown rank in bundle")
    return tuple(dims)


def _data_path(prefix: str, shard: int, num_shards: int) -> str:
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def read_bundle_index(prefix: str, verify_crc: bool = True):
    """-> (num_shards, {name: dict(dtype, shape, shard, offset, size, crc)})."""
    items = read_table(prefix + ".index", verify_crc)
    if not items or items[0][0] != b"":
        raise ValueError("%s.index: missing bundle header entry" % prefix)
    num_shards, endian = 1, 0
    for fno, _, v in _fields(items[0][1]):
        if fno == 1:
            num_shards = v
        elif fno == 2:
            endian = v
    if endian != 0:
        raise ValueError("big-endian tensor bundles are not supported")
    entries = {}
    for k, val in items[1:]:
        e = dict(dtype=0, shape=(), shard=0, offset=0, size=0, crc=None, sliced=False)
        for fno, wt, v in _fields(val):
            if fno == 1:
                e["dtype"] = v
            elif fno == 2:
                e["shape"] = _parse_shape(v)
            elif fno == 3:
                e["shard"] = v
            elif fno == 4:
                e["offset"] = v
            elif fno == 5:
                e["size"] = v
            elif fno == 6:
                e["crc"] = struct.unpack("<I", v)[0]
            elif fno == 7:
                e["sliced"] = True
        entries[k.decode("utf-8")] = e
    return num_shards, entries


def read_tensor_bundle(prefix: str, names: Optional[Iterable[str]] = None, verify_crc: bool = True
                       ) -> Dict[str, np.ndarray]:
    """Every (or the named) tensor of a Saver-V2 checkpoint ``prefix`` as numpy arrays."""
    num_shards, entries = read_bundle_index(prefix, verify_crc)
    want = list(entries) if names is None else list(names)
    files = {}
    out = {}
    try:
        for name in want:
            if name not in entries:
                raise KeyError("%s: no tensor named %r" % (prefix, name))
            e = entries[name]
            if e["sliced"]:
                raise ValueError("%r is a partitioned variable; not supported" % name)
            if e["dtype"] not in _DT_TO_NP:
                raise ValueError("%r: unsupported DataType %d" % (name, e["dtype"]))
            dt = np.dtype(_DT_TO_NP[e["dtype"]])
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise ValueError("%r: %d bytes on disk, shape %s needs %d" % (name, e["size"], e["shape"],
                                                                               count * dt.itemsize))
            fh = files.get(e["shard"])
            if fh is None:
                fh = files[e["shard"]] = open(_data_path(prefix, e["shard"], num_shards), "rb")
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("%r: data shard truncated" % name)
            if verify_crc and e["crc"] is not None and _unmask(e["crc"]) != crc32c(raw):
                raise ValueError("%r: crc32c mismatch in data shard" % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for fh in files.values():
            fh.close()
    return out


def write_tensor_bundle(prefix: str, tensors: Dict[str, np.ndarray]):
    """Write ``tensors`` as a one-shard Saver-V2 checkpoint (``prefix.index`` + ``prefix.data-00000-of-00001``)
    plus the ``checkpoint`` state file ``tf.train.latest_checkpoint`` reads."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    header = _enc_varint((1 << 3) | 0) + _enc_varint(1)                    # num_shards = 1
    header += _ld(3, _enc_varint((1 << 3) | 0) + _enc_varint(1))           # version { producer: 1 }
    items = [(b"", header)]
    off = 0
    with open(_data_path(prefix, 0, 1), "wb") as fh:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])   # tobytes() below is C order whatever the strides
            if a.dtype not in _NP_TO_DT:
                raise ValueError("%r: dtype %s has no TensorFlow DataType here" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            fh.write(raw)
            e = _enc_varint((1 << 3) | 0) + _enc_varint(_NP_TO_DT[a.dtype])
            e += _ld(2, _shape_proto(a.shape))
            if off:
                e += _enc_varint((4 << 3) | 0) + _enc_varint(off)
            e += _enc_varint((5 << 3) | 0) + _enc_varint(len(raw))
            e += _enc_varint((6 << 3) | 5) + struct.pack("<I", _mask(crc32c(raw)))
            items.append((name.encode("utf-8"), e))
            off += len(raw)
    write_table(prefix + ".index", items)
    base = os.path.basename(prefix)
    with open(os.path.join(d or ".", "checkpoint"), "w") as fh:
        fh.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the prefix named by ``checkpoint_dir/checkpoint``, or None."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    p = p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
    return p if os.path.exists(p + ".index") else None


# ----------------------------------------------------------------------------- variable names
def _intensity_names(scope: str, d: int, h: int, E: int):
    dh = d // h
    s = scope + "/sequential_temporal_combined/"      # temporal.py:288-303
    return [("int_w", s + "dense/kernel", (dh + 1, dh * E)), ("int_b", s + "dense/bias", (dh * E,)),
            ("int_weight", s + "weight", (E, dh)), ("int_scaling", s + "scaling", (E,))]


def tf_variable_names(cfg) -> List[Tuple[str, int, str, Tuple[int, ...]]]:
    """[(our name, block or -1, TF variable name, TF shape)] for ``cfg`` (synth.make_config).

    Names follow the reference's scopes under ``tf.variable_scope("main")`` (main.py:91).  Variables are
    created where the *call* happens, so the per-block tensors live under ``main/layer_i/...`` (EasyDGL.py:
    101-125) or ``main/num_blocks_i/...`` (CTSMA.py:65-73), not under the ``CSTMA/num_blocks_i`` scope the
    layer objects are constructed in; un-named ``tf.layers.dense`` calls are numbered ``dense, dense_1, ...``
    within their scope (temporal.py:340-343: Q, K, V, T); ``layernorm`` opens ``LayerNorm`` (Base.py:15-16).
    """
    d, h, E, L, N1 = cfg.num_units, cfg.num_heads, cfg.num_events, cfg.L, cfg.num_rows
    out = [("item_embs", -1, "main/CSTMA/item_embs/lookup_table", (N1, d)),                  # coding.py:52-55
           ("pos_embs", -1, "main/CSTMA/spatial_embs/embedding/lookup_table", (L, d)),       # coding.py:69-70
           ("output_bias", -1, "main/CSTMA/output_bias", (N1 - 1,))]                         # Base.py:106-110
    if cfg.model == "EasyDGL":
        out.append(("mark_embs", -1, "main/CSTMA/mark_embs/lookup_table", (E, d)))           # EasyDGL.py:52-53
        for i in range(cfg.num_blocks):
            p = "main/layer_%d/" % i
            cin = 3 * d if i == 0 else d
            blk = [("qkvt_w", p + "attention/self/TMAU/dense/kernel", (cin, 4 * d)),         # temporal.py:407-409
                   ("qkvt_b", p + "attention/self/TMAU/dense/bias", (4 * d,))]
            blk += _intensity_names(p + "attention/self/TMAU", d, h, E)
            blk += [("ao_w", p + "attention/output/dense/kernel", (d, d)),                   # EasyDGL.py:112-116
                    ("ao_b", p + "attention/output/dense/bias", (d,)),
                    ("ao_ln_g", p + "attention/output/LayerNorm/gamma", (d,)),
                    ("ao_ln_b", p + "attention/output/LayerNorm/beta", (d,)),
                    ("ff1_w", p + "intermediate/dense/kernel", (d, 2 * d)),                  # EasyDGL.py:119-121
                    ("ff1_b", p + "intermediate/dense/bias", (2 * d,)),
                    ("ff2_w", p + "output/dense/kernel", (2 * d, d)),                        # EasyDGL.py:124-128
                    ("ff2_b", p + "output/dense/bias", (d,)),
                    ("ff_ln_g", p + "output/LayerNorm/gamma", (d,)),
                    ("ff_ln_b", p + "output/LayerNorm/beta", (d,))]
            out += [(n, i, t, s) for n, t, s in blk]
        p = "main/cls/predictions/transform/"                                                # EasyDGL.py:136-139
        out += [("tr_w", -1, p + "dense/kernel", (d, d)), ("tr_b", -1, p + "dense/bias", (d,)),
                ("tr_ln_g", -1, p + "LayerNorm/gamma", (d,)), ("tr_ln_b", -1, p + "LayerNorm/beta", (d,))]
    else:
        for i in range(cfg.num_blocks):
            p = "main/num_blocks_%d/" % i
            cin = 2 * d if i == 0 else d
            a = p + "attention/modulating_attention/"                                        # temporal.py:338-343
            blk = [("ln1_g", p + "attention/LayerNorm/gamma", (cin,)),                       # CTSMA.py:68
                   ("ln1_b", p + "attention/LayerNorm/beta", (cin,))]
            for nm, dn in (("q", "dense"), ("k", "dense_1"), ("v", "dense_2"), ("t", "dense_3")):
                blk += [(nm + "_w", a + dn + "/kernel", (cin, d)), (nm + "_b", a + dn + "/bias", (d,))]
            blk += _intensity_names(p + "attention/modulating_attention", d, h, E)
            blk += [("ln2_g", p + "feed-forward/LayerNorm/gamma", (d,)),                     # CTSMA.py:73
                    ("ln2_b", p + "feed-forward/LayerNorm/beta", (d,)),
                    ("ff1_w", p + "feed-forward/Inner/kernel", (1, d, d)),                   # Base.py:73 Conv1D(k=1)
                    ("ff1_b", p + "feed-forward/Inner/bias", (d,)),
                    ("ff2_w", p + "feed-forward/Readout/kernel", (1, d, d)),                 # Base.py:74
                    ("ff2_b", p + "feed-forward/Readout/bias", (d,))]
            out += [(n, i, t, s) for n, t, s in blk]
        out += [("out_ln_g", -1, "main/outln/LayerNorm/gamma", (d,)),                        # CTSMA.py:79-80
                ("out_ln_b", -1, "main/outln/LayerNorm/beta", (d,))]
    return out


_SLOT = re.compile(r"(/Adam(_\d+)?$)|(^|/)(beta[12]_power|global_step)$|(^|/)Sequential/")


def model_variables(variables: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Drop optimizer slots (``.../Adam``, ``.../Adam_1``, ``beta?_power``), ``global_step`` and the
    ``Sequential/{TRAIN,EVAL}`` metric accumulators (Base.py:133,171) a Saver over all globals also writes."""
    return {k: v for k, v in variables.items() if not _SLOT.search(k)}


def _suffixes(name: str) -> List[str]:
    parts = name.split("/")
    return ["/".join(parts[i:]) for i in range(len(parts))]


def import_weights(cfg, variables: Dict[str, np.ndarray], overrides: Optional[Dict[str, str]] = None) -> dict:
    """TF-named arrays -> this package's weight dict (``synth.make_weights`` layout, without ``mark_table``).

    Each parameter is looked up by its expected name; failing that by the longest scope suffix that singles
    out one variable of the right shape (so a different outer scope than ``main/`` still loads).
    ``overrides`` maps ``"name"`` or ``"name@block"`` to an exact checkpoint variable name."""
    overrides = overrides or {}
    pool = model_variables(variables)
    used = set()
    W = {"blocks": [dict() for _ in range(cfg.num_blocks)]}
    for name, blk, tfname, shape in tf_variable_names(cfg):
        key = name if blk < 0 else "%s@%d" % (name, blk)
        src = overrides.get(key)
        if src is None and tfname in pool:
            src = tfname
        if src is None:
            for suf in _suffixes(tfname)[1:]:
                if "/" not in suf:
                    break  # a bare "kernel"/"gamma" identifies nothing
                cand = [k for k in pool if (k == suf or k.endswith("/" + suf)) and tuple(pool[k].shape) == shape
                        and k not in used]
                if len(cand) == 1:
                    src = cand[0]
                    break
                if len(cand) > 1:
                    raise KeyError("%s: %d variables end in %r with shape %s: %s - pass overrides={%r: ...}" %
                                   (key, len(cand), suf, shape, sorted(cand), key))
        if src is None:
            near = sorted(k for k in pool if tuple(pool[k].shape) == shape and k not in used)
            raise KeyError("%s: expected variable %r %s not found; same-shape variables left: %s" %
                           (key, tfname, shape, near))
        if src not in variables:
            raise KeyError("%s: override names %r, which is not in the checkpoint" % (key, src))
        a = np.asarray(variables[src])
        if tuple(a.shape) != shape:
            raise ValueError("%s: %r has shape %s, expected %s" % (key, src, tuple(a.shape), shape))
        used.add(src)
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        if len(shape) == 3:
            t = t.reshape(shape[1], shape[2])   # Conv1D kernel_size 1 == dense [in,out]
        (W if blk < 0 else W["blocks"][blk])[name] = t
    W["_unused"] = sorted(set(pool) - used)
    return W


def export_weights(cfg, weights: dict) -> Dict[str, np.ndarray]:
    """The inverse of ``import_weights``: {TF variable name: array} that ``saver.restore`` of the reference
    graph for ``cfg`` expects."""
    out = {}
    for name, blk, tfname, shape in tf_variable_names(cfg):
        t = weights[name] if blk < 0 else weights["blocks"][blk][name]
        a = t.detach().cpu().numpy().astype(np.float32)
        if int(np.prod(a.shape)) != int(np.prod(shape)):
            raise ValueError("%s: has shape %s, the reference variable is %s" % (name, tuple(a.shape), shape))
        out[tfname] = a.reshape(shape)
    return out


def load_checkpoint(cfg, prefix: str, overrides: Optional[Dict[str, str]] = None, verify_crc: bool = True) -> dict:
    """``saver.restore(sess, prefix)`` (analytics.py:88): read only the tensors that are model variables."""
    if os.path.isdir(prefix):
        p = latest_checkpoint(prefix)
        if p is None:
            raise FileNotFoundError("no checkpoint state in %s" % prefix)
        prefix = p
    if not os.path.exists(prefix + ".index"):
        raise FileNotFoundError("%s.index not found" % prefix)
    _, entries = read_bundle_index(prefix, verify_crc)
    keep = list(model_variables({k: None for k in entries}))
    W = import_weights(cfg, read_tensor_bundle(prefix, keep, verify_crc), overrides)
    W.pop("_unused")
    return W


def save_checkpoint(cfg, weights: dict, prefix: str):
    """``saver.save(sess, "ckpt/{model}")`` (util.py:53-55) for the model variables."""
    write_tensor_bundle(prefix, export_weights(cfg, weights))


# ----------------------------------------------------------------------------- mark.pkl
def load_mark_table(path: str) -> torch.Tensor:
    """``pickle.load(open(FLAGS.mark,'rb')).toarray()`` (EasyDGL.py:45): scipy sparse [num_items+1, E] ->
    int64 tensor."""
    obj = pickle.load(open(path, "rb"))
    a = obj.toarray() if hasattr(obj, "toarray") else np.asarray(obj)
    if a.ndim != 2:
        raise ValueError("%s: mark table must be 2-D, got shape %s" % (path, a.shape))
    if not np.array_equal(a, np.round(a)):
        raise ValueError("%s: mark table holds non-integer values (they index mark_embs, EasyDGL.py:87)" % path)
    return torch.from_numpy(a.astype(np.int64))


def save_mark_table(path: str, table) -> None:
    import scipy.sparse as sp
    a = table.cpu().numpy() if isinstance(table, torch.Tensor) else np.asarray(table)
    with open(path, "wb") as fh:
        pickle.dump(sp.csr_matrix(a), fh)
