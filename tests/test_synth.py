"""Synthetic Netflix-schema generator follows the reference's ETL conventions."""
import torch

from easydgl_b200 import synth


def test_easydgl_inputs_follow_linkpred_and_mask_last():
    cfg = synth.named_config("C1")
    inp = synth.make_inputs(cfg, 64, edge_cases=True)
    ids, ts = inp["seqs_i"], inp["seqs_t"]
    assert ids.dtype == torch.int64 and ts.dtype == torch.float32
    assert ids.shape == (64, 100) and ts.shape == (64, 100)
    assert (ids[:, -1] == cfg.mask_id).all()                      # dataloader.py:166-169
    body = ids[:, :-1]
    # right-aligned, left-padded with zeros (data/linkpred.py:142-157): no zero after a non-zero
    nz = body != 0
    assert bool((nz[:, 1:] >= nz[:, :-1]).all())
    assert bool((ts[:, :-1][~nz] == 0).all())
    assert int(body.max()) < cfg.num_items and int(body.min()) >= 0
    # timestamps non-decreasing over the real part, Netflix-era epoch seconds
    real = ts[0][ids[0] != 0]
    assert bool((real[1:] >= real[:-1]).all()) and 9.3e8 < float(real[0]) < 1.3e9
    # determinism (seed 9876, main.py:157)
    again = synth.make_inputs(cfg, 64, edge_cases=True)
    assert torch.equal(again["seqs_i"], ids) and torch.equal(again["seqs_t"], ts)


def test_ctsma_inputs_keep_extra_timestamp():
    cfg = synth.named_config("C3")
    inp = synth.make_inputs(cfg, 8)
    assert inp["seqs_i"].shape == (8, 100) and inp["seqs_t"].shape == (8, 101)   # dataloader.py:98-99
    assert cfg.num_rows == 18000 and cfg.mask_id == -1


def test_mark_table_and_weights_shapes():
    cfg = synth.named_config("C2")
    W = synth.make_weights(cfg)
    mt = W["mark_table"]
    assert mt.shape == (18000, 16) and mt.dtype == torch.int64
    assert int(mt[0].sum()) == 0 and int(mt[1:].sum(1).min()) >= 1 and int(mt.sum(1).max()) <= 3
    assert W["item_embs"].shape == (18001, 128) and W["output_bias"].shape == (18000,)
    assert W["blocks"][0]["qkvt_w"].shape == (384, 512)
    assert W["blocks"][0]["int_w"].shape == (17, 256) and W["blocks"][0]["int_weight"].shape == (16, 16)
    # reference initialisers: zeros / ones
    assert float(W["blocks"][0]["qkvt_b"].abs().max()) == 0.0 and float(W["tr_ln_g"].min()) == 1.0
    oh = synth.make_mark_table(cfg, onehot=True)
    assert int(oh[1:].sum(1).max()) == 1
