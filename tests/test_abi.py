"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/easydgl_b200.h declares, fails loudly without a GPU, and never falls back to the CPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            txt = open(os.path.join(inc, f)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names.update(re.findall(r"\b(edgl_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    from easydgl_b200 import _lib
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libeasydgl_b200.so does not export " + name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.edgl_version() >= 100
    assert lib.edgl_num_stages() >= 10 and lib.edgl_stage_name(0) == b"embed"


def test_config_struct_matches_header_layout():
    from easydgl_b200 import _lib
    # 10 int32 + float + (4-byte pad) + int64 + 2 int32  -> 64 bytes with natural alignment
    assert ctypes.sizeof(_lib.EdglConfig) == 64
    assert _lib.EdglConfig.mask_id.offset == 48 and _lib.EdglConfig.shard_rank.offset == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from easydgl_b200 import _lib, synth
    from easydgl_b200.engine import Engine
    lib = _lib.load()
    cfg = _lib.EdglConfig(model=0, max_batch=4, seq_len=8, num_units=16, num_heads=2, num_blocks=1, num_events=4,
                          num_rows=50, mark_rows=49, topk=10, time_scale=1.0, mask_id=49, shard_rank=0, shard_world=1)
    h = ctypes.c_void_p()
    rc = lib.edgl_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"cuda" in lib.edgl_last_error().lower()
    c = synth.make_config(model="EasyDGL", num_units=16, seqslen=7, num_items=49, num_heads=2, num_events=4)
    with pytest.raises(RuntimeError):
        Engine(c, synth.make_weights(c), max_batch=4)


def test_invalid_config_is_rejected_before_touching_the_device():
    from easydgl_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.EdglConfig(model=7, max_batch=4, seq_len=8, num_units=16, num_heads=2, num_blocks=1, num_events=4,
                          num_rows=50, mark_rows=49, topk=10, time_scale=1.0, mask_id=49, shard_rank=0, shard_world=1)
    assert lib.edgl_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert b"not implemented" in lib.edgl_last_error()  # util.py:96 wording
    bad.model, bad.num_heads = 0, 3
    assert lib.edgl_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert lib.edgl_create(None, ctypes.byref(h)) == -1


def test_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under easydgl_b200/ may reference it."""
    pkg = os.path.join(ROOT, "easydgl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "easydgl_oracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
