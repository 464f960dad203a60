"""Cross-check of the two independent oracle restatements: oracle/easydgl_oracle.py (torch) and
oracle/easydgl_ref.c (plain C, double).  They were written separately from the reference's source
text; agreement to 1e-9 on every logit is the pin that stands in for the reference's missing tests."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from helpers import O, case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")


def _f32(t):
    return t.detach().to(torch.float32).contiguous().numpy().tobytes()


def _write_blob(path, cfg, inp, W):
    easy = cfg.model == "EasyDGL"
    with open(path, "wb") as fh:
        fh.write(struct.pack("<10i", 0 if easy else 1, inp["seqs_i"].shape[0], cfg.L, cfg.num_units, cfg.num_heads,
                             cfg.num_blocks, cfg.num_events, cfg.num_rows, W["mark_table"].shape[0], cfg.ts_len))
        fh.write(struct.pack("<d", cfg.time_scale))
        fh.write(struct.pack("<q", cfg.mask_id))
        fh.write(inp["seqs_i"].to(torch.int64).contiguous().numpy().tobytes())
        fh.write(_f32(inp["seqs_t"]))
        fh.write(_f32(W["item_embs"]))
        fh.write(_f32(W["pos_embs"]))
        fh.write(_f32(W["output_bias"]))
        fh.write(W["mark_table"].to(torch.int64).contiguous().numpy().tobytes())
        if easy:
            fh.write(_f32(W["mark_embs"]))
        for blk in W["blocks"]:
            if easy:
                names = ["qkvt_w", "qkvt_b", "int_w", "int_b", "int_weight", "int_scaling", "ao_w", "ao_b", "ao_ln_g",
                         "ao_ln_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b", "ff_ln_g", "ff_ln_b"]
            else:
                names = ["ln1_g", "ln1_b", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "t_w", "t_b", "int_w", "int_b",
                         "int_weight", "int_scaling", "ln2_g", "ln2_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b"]
            for n in names:
                fh.write(_f32(blk[n]))
        if easy:
            for n in ("tr_w", "tr_b", "tr_ln_g", "tr_ln_b"):
                fh.write(_f32(W[n]))
        else:
            for n in ("out_ln_g", "out_ln_b"):
                fh.write(_f32(W[n]))


@pytest.fixture(scope="module")
def ref_bin():
    subprocess.run(["make", "-s", "-C", ORACLE], check=True)
    return os.path.join(ORACLE, "easydgl_ref")


@pytest.mark.parametrize("name", ["easy_a", "easy_c", "ctsma_a", "ctsma_b"])
def test_c_and_torch_restatements_agree(name, ref_bin, tmp_path):
    cfg, inp, W = case(name, batch=5)
    blob, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_blob(blob, cfg, inp, W)
    subprocess.run([ref_bin, blob, out], check=True, timeout=120)
    got = torch.from_numpy(np.fromfile(out, dtype=np.float64).reshape(5, cfg.num_rows))
    ref = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float64, literal=True)
    ref32 = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float32)
    well = (ref32.double() - ref)[:, 1:].abs().amax(1) <= 1e-4 * ref[:, 1:].abs().max()  # well-posed rows
    assert int(well.sum()) >= 3
    assert torch.equal(got[:, 0], torch.full((5,), -1000.0, dtype=torch.float64))
    err = float((got - ref)[well].abs().max() / ref[:, 1:].abs().max())
    assert err < 1e-9, err
