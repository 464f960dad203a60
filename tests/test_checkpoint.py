"""tf.train.Saver tensor bundles, the reference's variable names, mark.pkl - CPU only.
(No TensorFlow here: the container format is checked through its own invariants - magic, block CRCs,
prefix-compressed keys, known-answer CRC32C - and a full write -> read -> import round trip.)"""
import os
import struct

import numpy as np
import pytest
import torch

from easydgl_b200 import checkpoint as CK
from easydgl_b200 import dataloader as D
from easydgl_b200 import synth
from helpers import SMALL


def _flat(W):
    out = {k: v for k, v in W.items() if k not in ("blocks", "mark_table")}
    for i, b in enumerate(W["blocks"]):
        out.update({"%s@%d" % (k, i): v for k, v in b.items()})
    return out


def test_crc32c_native_matches_restatement():
    assert D._crc32c_py(b"123456789") == 0xE3069283          # RFC 3720 B.4 check value
    assert D.crc32c(b"\x00" * 32) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros
    assert D.crc32c(b"\xff" * 32) == 0x62A8AB43               # 32 bytes of ones
    assert D.crc32c(bytes(range(32))) == 0x46DD794E           # 32 incrementing bytes
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 16, 17, 23, 24, 25, 1000, 4099):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert D.crc32c(b) == D._crc32c_py(b), n
    assert CK._unmask(CK._mask(0xDEADBEEF)) == 0xDEADBEEF


def test_table_roundtrip_many_blocks(tmp_path):
    rng = np.random.default_rng(1)
    keys = sorted({("main/layer_%d/scope_%d/kernel" % (i % 7, i)).encode() for i in range(500)} | {b""})
    items = [(k, rng.integers(0, 256, int(rng.integers(0, 90)), dtype=np.uint8).tobytes()) for k in keys]
    p = str(tmp_path / "t.index")
    CK.write_table(p, items, block_size=512)      # force several data blocks + restart points
    assert CK.read_table(p) == items
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) > 48
    bad = bytearray(raw)
    bad[10] ^= 0x40
    open(p, "wb").write(bad)
    with pytest.raises(ValueError, match="crc32c"):
        CK.read_table(p)
    assert len(CK.read_table(p, verify_crc=False)) == len(items)
    with pytest.raises(ValueError, match="sorted"):
        CK.write_table(p, [(b"b", b""), (b"a", b"")])


def test_snappy_block_decode():
    # literal "abcd" + copy(offset 4, len 8) + long literal
    lit = bytes(range(70))
    buf = bytes([4 + 8 + 70]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + \
        bytes([60 << 2, 69]) + lit
    assert CK._snappy_uncompress(buf) == b"abcd" + b"abcdabcd" + lit
    with pytest.raises(ValueError):
        CK._snappy_uncompress(bytes([5, (3 << 2) | 2, 9, 0]))


def test_bundle_roundtrip_dtypes_and_crc(tmp_path):
    rng = np.random.default_rng(2)
    T = {"main/a/kernel": rng.standard_normal((3, 5)).astype(np.float32),
         "main/a/bias": np.zeros((0,), np.float32),
         "global_step": np.array(7, dtype=np.int64),
         "main/b/table": rng.integers(-5, 5, (4, 2, 3)).astype(np.int32),
         "main/b/f64": rng.standard_normal(6)}
    prefix = str(tmp_path / "ckpt" / "EasyDGL")
    CK.write_tensor_bundle(prefix, T)
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["EasyDGL.data-00000-of-00001", "EasyDGL.index", "checkpoint"]
    assert CK.latest_checkpoint(str(tmp_path / "ckpt")) == prefix
    R = CK.read_tensor_bundle(prefix)
    assert set(R) == set(T)
    for k in T:
        assert R[k].dtype == T[k].dtype and R[k].shape == T[k].shape and np.array_equal(R[k], T[k]), k
    assert list(CK.read_tensor_bundle(prefix, ["main/a/kernel"])) == ["main/a/kernel"]
    with pytest.raises(KeyError):
        CK.read_tensor_bundle(prefix, ["nope"])
    path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(path, "rb").read())
    raw[-1] ^= 1
    open(path, "wb").write(raw)
    with pytest.raises(ValueError, match="crc32c mismatch in data shard"):
        CK.read_tensor_bundle(prefix)


@pytest.mark.parametrize("name", ["easy_b", "ctsma_b"])
def test_variable_names_roundtrip(tmp_path, name):
    cfg = synth.make_config(**SMALL[name])
    W = synth.make_weights(cfg, mode="parity")
    names = CK.tf_variable_names(cfg)
    assert len({t for _, _, t, _ in names}) == len(names)                 # no two parameters share a variable
    assert len(names) == len(_flat(W))                                    # every parameter has a name
    tfv = CK.export_weights(cfg, W)
    # what a Saver over all globals adds: Adam slots, power accumulators, metric locals are not restored
    tfv["main/CSTMA/item_embs/lookup_table/Adam"] = np.ones_like(tfv["main/CSTMA/item_embs/lookup_table"])
    tfv["main/CSTMA/item_embs/lookup_table/Adam_1"] = np.ones_like(tfv["main/CSTMA/item_embs/lookup_table"])
    tfv["main/beta1_power"] = np.float32(0.5)
    tfv["main/global_step"] = np.int64(3)
    prefix = str(tmp_path / cfg.model)
    CK.write_tensor_bundle(prefix, tfv)
    W2 = CK.load_checkpoint(cfg, prefix)
    a, b = _flat(W), _flat(W2)
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    W3 = CK.load_checkpoint(cfg, str(tmp_path))                           # directory + `checkpoint` state file
    assert torch.equal(W3["item_embs"], W["item_embs"])


def test_import_is_robust_to_outer_scope_and_reports_misses():
    cfg = synth.make_config(**SMALL["ctsma_a"])
    W = synth.make_weights(cfg, mode="parity")
    tfv = CK.export_weights(cfg, W)
    moved = {("tower0/" + k[len("main/"):]): v for k, v in tfv.items()}  # a different outer scope
    W2 = CK.import_weights(cfg, moved)
    assert W2["_unused"] == []
    assert torch.equal(W2["blocks"][0]["k_w"], W["blocks"][0]["k_w"])
    assert torch.equal(W2["blocks"][0]["ff1_w"], W["blocks"][0]["ff1_w"])   # Conv1D [1,d,d] -> [d,d]
    assert tfv["main/num_blocks_0/feed-forward/Inner/kernel"].shape == (1, cfg.num_units, cfg.num_units)
    broken = dict(tfv)
    k = "main/num_blocks_0/attention/modulating_attention/dense_2/kernel"
    broken["main/num_blocks_0/attention/modulating_attention/dense_9/kernel"] = broken.pop(k)
    with pytest.raises(KeyError, match="dense_9"):                         # the message lists the candidates
        CK.import_weights(cfg, broken)
    W3 = CK.import_weights(cfg, broken, overrides={"v_w@0": k.replace("dense_2", "dense_9")})
    assert torch.equal(W3["blocks"][0]["v_w"], W["blocks"][0]["v_w"])
    wrong = dict(tfv)
    wrong["main/CSTMA/output_bias"] = np.zeros(3, np.float32)
    with pytest.raises((KeyError, ValueError)):
        CK.import_weights(cfg, wrong)


def test_mark_pkl_roundtrip(tmp_path):
    cfg = synth.make_config(**SMALL["easy_a"])
    tab = synth.make_mark_table(cfg)
    p = str(tmp_path / "mark.pkl")
    CK.save_mark_table(p, tab)
    import pickle
    import scipy.sparse as sp
    assert sp.issparse(pickle.load(open(p, "rb")))                         # what EasyDGL.py:45 unpickles
    got = CK.load_mark_table(p)
    assert got.dtype == torch.int64 and torch.equal(got, tab)
    pickle.dump(sp.csr_matrix(np.array([[0.5, 0.0]])), open(p, "wb"))
    with pytest.raises(ValueError, match="non-integer"):
        CK.load_mark_table(p)
