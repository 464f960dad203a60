"""Training-mode forward: model(features, is_training=True) and the loss of model.train() with
dropout 0, against the fp64 oracle's restatement of EasyDGL.py:140-189 / CTSMA.py:82-124 / temporal.py:317-333."""
from types import SimpleNamespace

import pytest
import torch

from helpers import O, assert_close, case, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", ["easy_a", "easy_b", "easy_d", "ctsma_a", "ctsma_b"])
def test_train_logits_and_loss_match_the_oracle(name):
    from easydgl_b200.engine import Engine
    cfg, _, W = case(name, batch=6)
    tr = synth.make_train_inputs(cfg, 6, masklen=5)
    eng = Engine(cfg, W, max_batch=6, device=DEV)
    pos = tr.get("masked_positions")
    posd = None if pos is None else pos.to(DEV)
    l2, ct = 1e-3, 1e-2 if cfg.model == "EasyDGL" else 1e-7
    ref = O.train_forward(tr["seqs_i"], tr["seqs_t"], tr["labels"], W, cfg, pos, l2_reg=l2, ct_reg=ct,
                          dtype=torch.float64, return_logits=True)
    lg = eng.forward_train_logits(tr["seqs_i"].to(DEV), tr["seqs_t"].to(DEV), posd).cpu()
    assert lg.shape == ref["logits"].shape
    well = (ref["logits"][:, 1:].abs().amax(1) < 1e6)
    assert_close(lg[well][:, 1:], ref["logits"][well][:, 1:], 1e-3, name + " training-mode logits")
    out = eng.forward_train_loss(tr["seqs_i"].to(DEV), tr["seqs_t"].to(DEV), tr["labels"].to(DEV), posd, l2, ct).cpu()
    out2 = eng.forward_train_loss(tr["seqs_i"].to(DEV), tr["seqs_t"].to(DEV), tr["labels"].to(DEV), posd, l2, ct).cpu()
    assert torch.equal(out, out2), "the loss must be reproducible bit for bit"
    for i, k in enumerate(("loss", "ce", "l2", "ct")):
        r = float(ref[k])
        assert abs(float(out[i]) - r) <= 1e-4 * max(abs(r), 1e-3), (name, k, float(out[i]), r)
    # without the regularisers the loss is the cross entropy
    out0 = eng.forward_train_loss(tr["seqs_i"].to(DEV), tr["seqs_t"].to(DEV), tr["labels"].to(DEV), posd, 0., 0.).cpu()
    assert float(out0[2]) == 0. and float(out0[3]) == 0. and abs(float(out0[0]) - float(ref["ce"])) <= 1e-4 * float(ref["ce"])
    eng.close()


def test_train_forward_rejects_bad_arguments():
    from easydgl_b200.engine import Engine
    cfg, _, W = case("easy_d", batch=4)
    tr = synth.make_train_inputs(cfg, 4)
    eng = Engine(cfg, W, max_batch=4, device=DEV)
    ids, ts = tr["seqs_i"].to(DEV), tr["seqs_t"].to(DEV)
    with pytest.raises(ValueError):
        eng.forward_train_logits(ids, ts, None)                      # EasyDGL needs masked_positions
    bad = tr["masked_positions"].clone()
    bad[0, 0] = cfg.L + 3
    with pytest.raises(ValueError):
        eng.forward_train_logits(ids, ts, bad.to(DEV))               # tf.batch_gather would raise
    lab = tr["labels"].clone()
    lab[1, 1] = cfg.num_rows + 5
    with pytest.raises(ValueError):
        eng.forward_train_loss(ids, ts, lab.to(DEV), tr["masked_positions"].to(DEV))
    eng.close()


def test_facade_train_returns_the_loss():
    from easydgl_b200.model.EasyDGL import EasyDGL
    cfg, _, W = case("easy_d", batch=5)
    tr = synth.make_train_inputs(cfg, 5)
    flags = SimpleNamespace(num_units=cfg.num_units, num_heads=cfg.num_heads, num_blocks=cfg.num_blocks, seqslen=cfg.seqslen,
                            time_scale=cfg.time_scale, l2_reg=1e-4, ct_reg=1e-3, hidden_dropout_rate=0.,
                            attention_probs_dropout_rate=0., masklen=6)
    m = EasyDGL(cfg.num_items, flags, weights=W, mark_table=W["mark_table"], device=DEV, max_batch=8)
    feats = {k: tr[k].to(DEV) for k in ("seqs_i", "seqs_t", "masked_positions")}
    loss, parts = m.train(feats, tr["labels"].to(DEV))
    ref = O.train_forward(tr["seqs_i"], tr["seqs_t"], tr["labels"], W, cfg, tr["masked_positions"], l2_reg=1e-4, ct_reg=1e-3)
    assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * float(ref["loss"])
    lg = m(feats, True)
    assert tuple(lg.shape) == (5 * 6, cfg.num_rows)
    flags.hidden_dropout_rate = 0.1
    m2 = EasyDGL(cfg.num_items, flags, weights=W, mark_table=W["mark_table"], device=DEV, max_batch=8)
    with pytest.raises(NotImplementedError):
        m2.train(feats, tr["labels"].to(DEV))
