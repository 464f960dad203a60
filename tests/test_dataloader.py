"""TFRecord reader + eval post-processing (no TensorFlow).  The wire format is pinned
against the official protobuf runtime (a dynamically built tf.train.Example schema) and the CRC32C
known-answer vector; the post-processors against the reference's conventions (dataloader.py)."""
import os
import struct
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from easydgl_b200 import dataloader as D
from easydgl_b200 import synth


def test_crc32c_known_answers():
    assert D.crc32c(b"123456789") == 0xE3069283          # CRC-32C check value (RFC 3720 appendix)
    assert D.crc32c(b"") == 0
    assert D.crc32c(bytes(32)) == 0x8A9136AA              # 32 zero bytes (iSCSI test vector)
    c = D.crc32c(b"abc")
    assert D.masked_crc32c(b"abc") == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def _official_example_class():
    """tf.train.Example built with google.protobuf from its published .proto definition."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="example_test.proto", package="tensorflow_t", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname, packed in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
            if packed:
                f.options.packed = True
        return m
    msg("BytesList", [("value", 1, T.TYPE_BYTES, T.LABEL_REPEATED, None, False)])
    msg("FloatList", [("value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, None, True)])
    msg("Int64List", [("value", 1, T.TYPE_INT64, T.LABEL_REPEATED, None, True)])
    ft = msg("Feature", [("bytes_list", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".tensorflow_t.BytesList", False),
                         ("float_list", 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".tensorflow_t.FloatList", False),
                         ("int64_list", 3, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".tensorflow_t.Int64List", False)])
    ft.oneof_decl.add(name="kind")
    for f in ft.field:
        f.oneof_index = 0
    feats = msg("Features", [("feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow_t.Features.FeatureEntry", False)])
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL, type_name=".tensorflow_t.Feature")
    msg("Example", [("features", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".tensorflow_t.Features", False)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow_t.Example"))


def test_example_wire_format_matches_official_protobuf():
    Example = _official_example_class()
    ids = np.array([0, 0, 5, 17770, -3, 2 ** 40], dtype=np.int64)
    ts = np.array([0.0, 0.0, 9.4e8, 1.1e9, 3.5, -1.25], dtype=np.float32)
    ex = Example()
    ex.features.feature["seqs_i"].int64_list.value.extend(ids.tolist())
    ex.features.feature["seqs_t"].float_list.value.extend(ts.tolist())
    got = D.parse_example(ex.SerializeToString())            # official writer -> our parser
    assert np.array_equal(got["seqs_i"], ids) and np.array_equal(got["seqs_t"], ts)
    ex2 = Example()
    ex2.ParseFromString(D.serialize_example({"seqs_i": ids, "seqs_t": ts}))   # our writer -> official parser
    assert list(ex2.features.feature["seqs_i"].int64_list.value) == ids.tolist()
    assert np.array_equal(np.array(ex2.features.feature["seqs_t"].float_list.value, dtype=np.float32), ts)


def test_tfrecord_roundtrip_crc_and_corruption(tmp_path):
    p = str(tmp_path / "a.tfrec")
    recs = [b"hello", b"", bytes(range(256)) * 5]
    with D.TFRecordWriter(p) as w:
        for r in recs:
            w.write(r)
    assert list(D.read_tfrecords(p)) == recs
    raw = bytearray(open(p, "rb").read())
    assert struct.unpack("<Q", raw[:8])[0] == 5
    raw[14] ^= 0x01                                          # flip a payload bit
    open(p, "wb").write(raw)
    with pytest.raises(IOError):
        list(D.read_tfrecords(p))
    assert len(list(D.read_tfrecords(p, verify_crc=False))) == 3


@pytest.mark.parametrize("model", ["EasyDGL", "CTSMA"])
def test_input_reader_follows_reference_postprocessing(model, tmp_path):
    cfg = synth.make_config(model=model, num_units=16, seqslen=9, num_items=120, num_heads=2, num_events=4)
    FLAGS = SimpleNamespace(model=model, seqslen=9, num_items=120, masklen=6)
    path = str(tmp_path / "test.tfrec")
    tokens, times = D.write_synthetic_shard(path, cfg, 11, seed=5)
    assert tokens.shape == (11, 10) and times.shape == (11, 10)          # seqslen + 1 stored per user
    batches = list(D.reader(FLAGS, str(tmp_path / "*.tfrec"), is_training=False)(4))
    assert [b[1].shape[0] for b in batches] == [4, 4, 3]                 # last batch is short
    feats = {k: torch.cat([b[0][k] for b in batches]) for k in batches[0][0]}
    labels = torch.cat([b[1] for b in batches])
    assert feats["seqs_i"].dtype == torch.int64 and feats["seqs_t"].dtype == torch.float32
    assert torch.equal(labels, torch.from_numpy(tokens))                 # labels = the unmasked tokens
    if model == "EasyDGL":                                               # dataloader.py:166-179
        want = tokens.copy()
        want[:, -1] = 120
        assert torch.equal(feats["seqs_i"], torch.from_numpy(want))
        assert torch.equal(feats["seqs_t"], torch.from_numpy(times))
    else:                                                                # dataloader.py:95-99
        assert torch.equal(feats["seqs_i"], torch.from_numpy(tokens[:, :-1]))
        assert feats["seqs_t"].shape == (11, 10)
    # the synthetic shard is exactly what synth.make_inputs hands to the kernels
    direct = synth.make_inputs(cfg, 11, seed=5)
    assert torch.equal(feats["seqs_i"], direct["seqs_i"]) and torch.equal(feats["seqs_t"], direct["seqs_t"])
    assert torch.equal(labels[:, -1], direct["labels"])


def test_decoder_and_reader_errors(tmp_path):
    dec = D.TfExampleDecoder(5)
    with pytest.raises(ValueError):
        dec.decode(D.serialize_example({"seqs_i": np.arange(4), "seqs_t": np.zeros(5, np.float32)}))
    with pytest.raises(ValueError):
        dec.decode(D.serialize_example({"seqs_i": np.arange(5)}))
    with pytest.raises(NotImplementedError):
        D.reader(SimpleNamespace(model="SASREC", seqslen=5, num_items=9, masklen=1), "x", False)
    with pytest.raises(NotImplementedError):
        D.reader(SimpleNamespace(model="EasyDGL", seqslen=5, num_items=9, masklen=1), "x", True)
    with pytest.raises(FileNotFoundError):
        next(D.reader(SimpleNamespace(model="EasyDGL", seqslen=5, num_items=9, masklen=1),
                      str(tmp_path / "none*.tfrec"), False)(2))
