"""Eval-side data pipeline of the reference, without TensorFlow.

Mirrors ``src/dataloader.py`` for the two models on the hot path:

* ``TfExampleDecoder(seqslen, ...)``            dataloader.py:11-31   (FixedLenFeature schema: seqs_i int64, seqs_t float32)
* ``MAUPostProcessor.mask_last``                dataloader.py:166-179 (EasyDGL eval: last token -> [MASK])
* ``RegressivePostProcessor``                   dataloader.py:88-108  (CTSMA: tokens[:-1], all timestamps)
* ``InputReader(pattern, is_training, decoder, processor)(batch_size)``  dataloader.py:209-246
* ``TFRecordWriter`` / ``serialize_example``    data/linkpred.py:26-39 (the writer side, for synthetic shards)

The TFRecord container (``u64 length | u32 masked_crc32c(length) | payload | u32 masked_crc32c(payload)``) and
the ``tf.train.Example`` protobuf wire format are decoded by hand; batches come out as torch tensors ready
for ``model(features, False)`` / ``model.eval(features, labels, mask_seen)``.
"""
from __future__ import annotations

import glob
import struct
from typing import Dict, Iterator, List, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------- CRC32C (Castagnoli)
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = tab
    return _CRC_TABLE


def _crc32c_py(data: bytes) -> int:
    tab = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data) -> int:
    """CRC32C of a bytes-like object.  Anything longer than a frame header goes through the library's
    host routine ``edgl_crc32c`` (include/easydgl_b200.h) when the library is built; the byte loop above is
    the independent restatement the tests check it against, and the fallback on a CPU-only preprocessing host
    where libeasydgl_b200.so was never compiled (file framing is not the GPU hot path)."""
    mv = memoryview(data).cast("B")
    if len(mv) <= 16:
        return _crc32c_py(bytes(mv))
    lib = _native_lib()
    if lib is None:
        return _crc32c_py(bytes(mv))
    buf = np.frombuffer(mv, dtype=np.uint8)
    return int(lib.edgl_crc32c(buf.ctypes.data, buf.size, 0))


_NATIVE = []


def _native_lib():
    """The shared library if it can be loaded, else None (cached)."""
    if not _NATIVE:
        try:
            from . import _lib
            _NATIVE.append(_lib.load())
        except Exception:
            _NATIVE.append(None)
    return _NATIVE[0]


def masked_crc32c(data: bytes) -> int:
    """TFRecord's masked CRC: rotate right by 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- protobuf wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) of one message; value is int (varint), bytes (len-delimited)
    or raw 4/8 bytes (fixed)."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


def _decode_feature(buf: bytes):
    """tf.train.Feature: oneof bytes_list=1 / float_list=2 / int64_list=3, each `repeated value = 1`."""
    for fno, _, v in _fields(buf):
        if fno == 2:  # FloatList
            vals: List[float] = []
            for f2, wt, x in _fields(v):
                if f2 != 1:
                    continue
                if wt == 2:
                    vals.extend(np.frombuffer(x, dtype="<f4").tolist())
                else:
                    vals.append(struct.unpack("<f", x)[0])
            return np.asarray(vals, dtype=np.float32)
        if fno == 3:  # Int64List
            ints: List[int] = []
            for f2, wt, x in _fields(v):
                if f2 != 1:
                    continue
                if wt == 2:
                    p = 0
                    while p < len(x):
                        val, p = _varint(x, p)
                        ints.append(val - (1 << 64) if val >= (1 << 63) else val)
                else:
                    ints.append(x - (1 << 64) if x >= (1 << 63) else x)
            return np.asarray(ints, dtype=np.int64)
        if fno == 1:  # BytesList
            return [x for f2, _, x in _fields(v) if f2 == 1]
    return np.zeros(0, dtype=np.float32)


def parse_example(serialized: bytes) -> Dict[str, np.ndarray]:
    """tf.train.Example -> {name: array}.  Example{features=1} / Features{map feature=1} / entry{key=1,value=2}."""
    out = {}
    for fno, _, feats in _fields(serialized):
        if fno != 1:
            continue
        for f2, _, entry in _fields(feats):
            if f2 != 1:
                continue
            key, val = None, b""
            for f3, _, x in _fields(entry):
                if f3 == 1:
                    key = x.decode("utf-8")
                elif f3 == 2:
                    val = x
            if key is not None:
                out[key] = _decode_feature(val)
    return out


def _enc_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fno: int, payload: bytes) -> bytes:
    return _enc_varint((fno << 3) | 2) + _enc_varint(len(payload)) + payload


def serialize_example(features: Dict[str, np.ndarray]) -> bytes:
    """The writer side of data/linkpred.py:26-39 (packed Int64List / FloatList)."""
    feats = b""
    for name in sorted(features):
        arr = np.asarray(features[name])
        if arr.dtype.kind == "f":
            feature = _ld(2, _ld(1, arr.astype("<f4").tobytes()))
        else:
            feature = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in arr.reshape(-1))))
        feats += _ld(1, _ld(1, name.encode()) + _ld(2, feature))
    return _ld(1, feats)


# ----------------------------------------------------------------------------- TFRecord container
class TFRecordWriter:
    def __init__(self, path: str):
        self._fh = open(path, "wb")

    def write(self, record: bytes):
        hdr = struct.pack("<Q", len(record))
        self._fh.write(hdr + struct.pack("<I", masked_crc32c(hdr)) + record + struct.pack("<I", masked_crc32c(record)))

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def read_tfrecords(path: str, verify_crc: bool = True) -> Iterator[bytes]:
    with open(path, "rb") as fh:
        while True:
            hdr = fh.read(8)
            if not hdr:
                return
            if len(hdr) < 8:
                raise IOError("truncated TFRecord header in %s" % path)
            (ln,) = struct.unpack("<Q", hdr)
            (hcrc,) = struct.unpack("<I", fh.read(4))
            data = fh.read(ln)
            tail = fh.read(4)
            if len(data) < ln or len(tail) < 4:
                raise IOError("truncated TFRecord in %s" % path)
            if verify_crc:
                if hcrc != masked_crc32c(hdr) or struct.unpack("<I", tail)[0] != masked_crc32c(data):
                    raise IOError("corrupted TFRecord (crc mismatch) in %s" % path)
            yield data


# ----------------------------------------------------------------------------- reference classes
class TfExampleDecoder(object):
    """dataloader.py:11-31.  FixedLenFeature semantics: a feature of the wrong length is an error."""

    def __init__(self, seqslen, has_labels=False, has_datetime=False):
        self._keys = {"seqs_i": np.int64, "seqs_t": np.float32}
        if has_labels:
            self._keys["labels"] = np.int64
        if has_datetime:
            for k in ("seqs_month", "seqs_day", "seqs_weekday", "seqs_hour"):
                self._keys[k] = np.int64
        self._seqslen = seqslen

    def decode(self, serialized_example: bytes) -> Dict[str, np.ndarray]:
        ex = parse_example(serialized_example)
        out = {}
        for k, dt in self._keys.items():
            if k not in ex:
                raise ValueError("Feature: %s (data type: %s) is required but could not be found." % (k, dt.__name__))
            v = np.asarray(ex[k])
            if v.shape != (self._seqslen,):
                raise ValueError("Key: %s.  Can't parse serialized Example: expected shape [%d], got %s"
                                 % (k, self._seqslen, list(v.shape)))
            out[k] = v.astype(dt)
        return out


class MAUPostProcessor(object):
    """dataloader.py:159-206 (EasyDGL).  Eval = mask_last; the random masking of training is out of scope."""

    def __init__(self, seqslen: int, maskslen: int, mask: int, is_training):
        self.seqslen, self.maskslen, self.mask, self.is_training = seqslen, maskslen, mask, is_training

    def mask_last(self, decoded: dict):
        tokens = decoded["seqs_i"]
        masked = tokens.copy()
        masked[self.seqslen - 1] = self.mask  # one_hot(seqslen-1) * (mask - tokens) + tokens
        return {"seqs_i": masked, "seqs_t": decoded["seqs_t"]}, tokens

    def __call__(self, decoded: dict):
        if not self.is_training:
            return self.mask_last(decoded)
        raise NotImplementedError("random masking for training is out of scope (SURVEY.md 8f rank 3)")


class RegressivePostProcessor(object):
    """dataloader.py:88-108 (CTSMA): features = tokens[:-1] with ALL timestamps; labels = tokens (eval)."""

    def __init__(self, is_training, has_datetime=False, keep_entire=False):
        self.is_training, self.has_datetime, self.keep_entire = is_training, has_datetime, keep_entire

    def __call__(self, decoded: dict):
        tokens, timestamps = decoded["seqs_i"], decoded["seqs_t"]
        features = {"seqs_i": tokens[:-1], "seqs_t": timestamps}
        labels = tokens[1:] if self.is_training else tokens
        return features, labels


class InputReader(object):
    """dataloader.py:209-246: file pattern -> batches of (features, labels) as torch tensors.
    Eval order is the sorted file order then record order (tf.data list_files(shuffle=False) + interleave over
    a single eval file behaves the same); the last batch may be short like map_and_batch(drop_remainder=False)."""

    def __init__(self, file_pattern, is_training, decoder, processor=None):
        assert decoder is not None, "decoder is not specified"
        assert processor is not None, "postprocessor is not specified"
        if is_training:
            raise NotImplementedError("the shuffling training pipeline is out of scope (SURVEY.md 8f rank 3)")
        self._file_pattern, self.decoder, self.processor = file_pattern, decoder, processor

    def __call__(self, batch_size, device=None, verify_crc=True):
        files = sorted(glob.glob(self._file_pattern))
        if not files:
            raise FileNotFoundError("no files match %s" % self._file_pattern)

        def flush(fs, ls):
            feats = {k: torch.from_numpy(np.stack([f[k] for f in fs])) for k in fs[0]}
            labels = torch.from_numpy(np.stack(ls))
            if device is not None:
                feats = {k: v.to(device) for k, v in feats.items()}
                labels = labels.to(device)
            return feats, labels

        fs, ls = [], []
        for path in files:
            for rec in read_tfrecords(path, verify_crc):
                f, lab = self.processor(self.decoder.decode(rec))
                fs.append(f)
                ls.append(lab)
                if len(fs) == batch_size:
                    yield flush(fs, ls)
                    fs, ls = [], []
        if fs:
            yield flush(fs, ls)


def reader(FLAGS, file_pattern, is_training: bool):
    """util.reader (src/util.py:99-129) for the two models on this path."""
    if FLAGS.model == "CTSMA":
        return InputReader(file_pattern, is_training=is_training,
                           decoder=TfExampleDecoder(FLAGS.seqslen + 1, has_datetime=False),
                           processor=RegressivePostProcessor(is_training, has_datetime=False, keep_entire=True))
    if FLAGS.model == "EasyDGL":
        seqslen = FLAGS.seqslen + 1
        return InputReader(file_pattern, is_training=is_training,
                           decoder=TfExampleDecoder(seqslen, has_datetime=False),
                           processor=MAUPostProcessor(seqslen, FLAGS.masklen, FLAGS.num_items, is_training))
    raise NotImplementedError("The ranking model: {0} not implemented".format(FLAGS.model))


def write_synthetic_shard(path: str, cfg, n: int, seed: int = 9876):
    """Synthetic Netflix-schema shard in the layout of data/linkpred.py:126-191 (right-aligned, left-padded
    seqslen+1 tokens and timestamps per user, plus the four calendar features the reference also stores)."""
    from . import synth
    raw = synth.make_inputs(cfg, n, seed=seed)
    tokens = raw["seqs_i"].numpy().copy()
    if cfg.model == "EasyDGL":
        tokens[:, -1] = raw["labels"].numpy()  # undo mask_last: the file holds the true last item
    else:
        tokens = np.concatenate([tokens, raw["labels"].numpy()[:, None]], axis=1)
    times = raw["seqs_t"].numpy()
    with TFRecordWriter(path) as w:
        for i in range(n):
            t = times[i].astype(np.int64)
            w.write(serialize_example({
                "seqs_i": tokens[i], "seqs_t": times[i],
                "seqs_month": (t // 2592000) % 12, "seqs_day": (t // 86400) % 31,
                "seqs_weekday": (t // 86400) % 7, "seqs_hour": (t // 3600) % 24}))
    return tokens, times
