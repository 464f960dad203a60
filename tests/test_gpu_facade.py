"""The reference-facing facade: same class names / ctor args / call signatures as the reference's
model and layer classes, checked against the oracle the way a reference test would."""
from types import SimpleNamespace

import pytest
import torch

from helpers import O, assert_close, case, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _flags(cfg, **kw):
    return SimpleNamespace(model=cfg.model, num_items=cfg.num_items, num_units=cfg.num_units, num_heads=cfg.num_heads,
                           num_blocks=cfg.num_blocks, seqslen=cfg.seqslen, time_scale=cfg.time_scale, masklen=6,
                           hidden_dropout_rate=0.1, attention_probs_dropout_rate=0.1, learning_rate=5e-4, l2_reg=1e-4,
                           ct_reg=1e-7, num_train_steps=None, num_warmup_steps=None, mark=None, mask_seen=True, **kw)


@pytest.mark.parametrize("name", ["easy_b", "ctsma_b"])
def test_util_ranking_model_call_and_eval(name):
    from easydgl_b200.util import ranking
    cfg, inp, W = case(name, batch=8)
    model = ranking(_flags(cfg), weights=W, mark_table=W["mark_table"].numpy(), device=DEV)
    assert model.num_items == cfg.num_rows and model.num_events == cfg.num_events
    features = {"seqs_i": inp["seqs_i"].to(DEV), "seqs_t": inp["seqs_t"].to(DEV)}
    logits = model(features, is_training=False)
    ref = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float64)
    assert_close(logits.cpu()[:, 1:], ref[:, 1:], 1e-3, "facade logits")
    labels = torch.stack([inp["labels"]] * 2, dim=1)  # labels[:, -1] is the target (Base.py:169)
    metrics, idx = model.eval(features, labels, mask_seen=True)
    _, ridx = O.eval_topk(ref, inp["seqs_i"], True, 100, rank_on="probs")
    want = O.ranking_metrics(ridx, inp["labels"])
    res = O.topk_set_compare(idx.cpu().long(), O.mask_seen_logits(ref, inp["seqs_i"]), 100,
                             tau=4 * float((logits.cpu().double() - ref).abs().max()))
    assert res["bad"] == 0
    if res["exact"] == res["rows"]:
        for k, v in want.items():
            assert abs(metrics[k] - v) < 1e-9, (k, metrics[k], v)
    # streaming means (tf.metrics.mean) and reset (metric_init_op)
    m2, _ = model.eval(features, labels, mask_seen=True)
    assert all(abs(m2[k] - metrics[k]) < 1e-12 for k in metrics)
    model.reset_metrics()
    with pytest.raises(NotImplementedError):
        model.train(features, labels)
    with pytest.raises(NotImplementedError):
        ranking(SimpleNamespace(model="SASREC", num_items=10))


def test_bimau_layer_facade_signature():
    """T.BiMAU(num_units, num_heads, num_events, dropout)(queries, keys, masks, intervals, marks, is_training)."""
    from easydgl_b200.module import temporal as T
    cfg, inp, W = case("easy_b", batch=4)
    W64 = O._cast(W, torch.float64)
    X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
    blk = W["blocks"][0]
    keep = {k: blk[k] for k in ("qkvt_w", "qkvt_b", "int_w", "int_b", "int_weight", "int_scaling")}
    layer = T.BiMAU(cfg.num_units, cfg.num_heads, cfg.num_events, 0.1, weights=keep, device=DEV)
    masks = kmask.float().unsqueeze(1).repeat(cfg.num_heads, cfg.L, 1).to(DEV)   # EasyDGL.py:94-95
    out, lam = layer(X0.float().to(DEV), X0.float().to(DEV), masks, spans.float().to(DEV), marks.to(DEV), False)
    rO, rl = O.bimau(X0, kmask, spans, marks, W64["blocks"][0], cfg.num_units, cfg.num_heads, cfg.num_events)
    assert_close(out.cpu(), rO, 1e-4, "BiMAU facade out")
    assert_close(lam.cpu(), rl, 1e-4, "BiMAU facade lam")
    assert lam.shape == (cfg.num_heads * 4, cfg.L, cfg.num_events)
    G, lam2 = layer.intensity(torch.randn(cfg.num_heads * 4, cfg.L, cfg.num_units // cfg.num_heads, device=DEV),
                              spans.float().to(DEV), marks.to(DEV))
    assert G.shape == (cfg.num_heads * 4, cfg.L, cfg.L)


def test_mgau_layer_facade():
    """T.MGAU (temporal.py:455-508) = BiMAU without set_diag."""
    from easydgl_b200.module import temporal as T
    cfg, inp, W = case("easy_c", batch=4, num_events=8)
    W64 = O._cast(W, torch.float64)
    X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
    blk = W["blocks"][0]
    keep = {k: blk[k] for k in ("qkvt_w", "qkvt_b", "int_w", "int_b", "int_weight", "int_scaling")}
    layer = T.MGAU(cfg.num_units, cfg.num_heads, cfg.num_events, 0.1, weights=keep, device=DEV)
    masks = kmask.float().unsqueeze(1).repeat(cfg.num_heads, cfg.L, 1).to(DEV)
    out, lam = layer(X0.float().to(DEV), X0.float().to(DEV), masks, spans.float().to(DEV), marks.to(DEV), False)
    rO, rl = O.mgau(X0, kmask, spans, marks, W64["blocks"][0], cfg.num_units, cfg.num_heads, cfg.num_events)
    assert_close(out.cpu(), rO, 1e-4, "MGAU out")
    bO, _ = O.bimau(X0, kmask, spans, marks, W64["blocks"][0], cfg.num_units, cfg.num_heads, cfg.num_events)
    assert (rO - bO).abs().max() > 1e-6                       # and it differs from BiMAU


def test_coding_layer_facades():
    from easydgl_b200.module import coding as C
    g = torch.Generator().manual_seed(3)
    table = torch.randn(40, 16, generator=g)
    emb = C.Embedding(40, 16, 0.0, zero_pad=True, scale=True, initializer=table, scope="item_embs", device=DEV)
    ids = torch.randint(0, 40, (3, 7), generator=g)
    assert torch.equal(emb(ids.to(DEV)).cpu(), O.embedding(O.zero_pad_table(table), ids, True, 16))
    assert torch.equal(emb.lookup_table[0].cpu(), torch.zeros(16))
    pc = C.PositionCoding(7, 16, 0.0, initializer=table[:7], device=DEV)
    x = emb(ids.to(DEV))
    assert pc(x).shape == (3, 7, 32) and torch.equal(pc.code(x)[1].cpu(), table[:7])
    tc = C.TimeSinusoidCoding(16)
    ts = torch.rand(3, 7, generator=g) * 12000
    assert_close(tc.code(ts.to(DEV)).cpu(), O.time_sinusoid_code(ts, 16, torch.float64), 1e-5, "tcoding")
    with pytest.raises(AssertionError):
        tc.code(torch.zeros(2, 3, 4, device=DEV))


def test_c1_from_tfrecords_end_to_end(tmp_path):
    """BASELINE.json configs[0] literally: EasyDGL d=64 L=100, 18K-item synthetic Netflix-schema TFRecords,
    B=32 - file -> reader (dataloader.py conventions) -> model.eval -> HR/NDCG, against the oracle."""
    from easydgl_b200 import dataloader as D
    from easydgl_b200.util import ranking
    cfg = synth.named_config("C1")
    W = synth.make_weights(cfg, mode="parity")
    D.write_synthetic_shard(str(tmp_path / "test.tfrec"), cfg, 64, seed=321)
    flags = _flags(cfg)
    model = ranking(flags, weights=W, mark_table=W["mark_table"].numpy(), device=DEV)
    ref_tp = []
    for features, labels in D.reader(flags, str(tmp_path / "*.tfrec"), is_training=False)(32, device=DEV):
        assert features["seqs_i"].shape == (32, 100)
        metrics, idx = model.eval(features, labels, mask_seen=True)
        ids_cpu = features["seqs_i"].cpu()
        ref = O.forward(ids_cpu, features["seqs_t"].cpu(), W, cfg, dtype=torch.float64)
        lg = model(features, is_training=False).cpu().double()
        res = O.topk_set_compare(idx.cpu().long(), O.mask_seen_logits(ref, ids_cpu), 100,
                                 tau=4 * float((lg - ref).abs().max()))
        assert res["bad"] == 0, res                       # identical sets up to near-ties at the cut
        ref_tp.append((idx.cpu().long(), labels[:, -1].cpu()))
    # HR/NDCG formulas (Base.py:181-201) on the returned rankings, streamed over both batches
    want = O.ranking_metrics(torch.cat([a for a, _ in ref_tp]), torch.cat([b for _, b in ref_tp]))
    for k, v in want.items():
        assert abs(metrics[k] - v) < 1e-9, (k, metrics[k], v)


@pytest.mark.parametrize("name", ["easy_b", "ctsma_a"])
def test_restore_from_saver_checkpoint_and_mark_pkl(tmp_path, name):
    """analytics.py:83-90 flow: FLAGS.mark -> mark.pkl, model built with fresh variables, then
    saver.restore(sess, FLAGS.ckpt); logits must equal those of a model given the weights directly."""
    from easydgl_b200 import checkpoint as CK
    from easydgl_b200.util import ranking
    cfg, inp, W = case(name, batch=6)
    CK.save_mark_table(str(tmp_path / "mark.pkl"), W["mark_table"])
    trained = ranking(_flags(cfg), weights=W, mark_table=W["mark_table"].numpy(), device=DEV)
    prefix = trained.save(str(tmp_path / "ckpt" / cfg.model))                 # util.py:53-55
    features = {"seqs_i": inp["seqs_i"].to(DEV), "seqs_t": inp["seqs_t"].to(DEV)}
    want = trained(features, False).cpu()
    flags = _flags(cfg)
    flags.mark = str(tmp_path / "mark.pkl")
    fresh = ranking(flags, device=DEV)                                       # tf default initialisers
    assert torch.equal(fresh.mark_lookup_table, W["mark_table"])
    before = fresh(features, False).cpu()
    assert not torch.equal(before, want)
    fresh.restore(prefix)                                                    # engine already built: weights rebind
    assert torch.equal(fresh(features, False).cpu(), want)
    again = ranking(flags, device=DEV)
    again.restore(str(tmp_path / "ckpt"))                                    # directory with a `checkpoint` file
    assert torch.equal(again(features, False).cpu(), want)


def test_host_entry_point_submit_wait_pipeline():
    """edgl_forward_topk_host_submit / _wait: four batches through the two staging slots give the results of the
    synchronous device-resident call, in order, and a third outstanding submit is refused."""
    from easydgl_b200 import synth
    from easydgl_b200.engine import Engine
    cfg = synth.make_config(model="EasyDGL", num_units=64, seqslen=30, num_items=500, num_heads=4, num_blocks=1,
                            num_events=16)
    W = synth.make_weights(cfg, mode="parity")
    B = 16
    eng = Engine(cfg, W, max_batch=B, device=DEV)
    batches = [synth.make_inputs(cfg, B, seed=77 + i) for i in range(4)]
    ref = [eng.forward_topk(b["seqs_i"].to(DEV), b["seqs_t"].to(DEV), True) for b in batches]
    ins = [(b["seqs_i"].pin_memory(), b["seqs_t"].pin_memory()) for b in batches]
    outs = [(torch.empty((B, eng.K), dtype=torch.int32).pin_memory(), torch.empty((B, eng.K)).pin_memory())
            for _ in batches]
    prev = None
    for i in range(4):
        slot = eng.forward_topk_host_submit(ins[i][0], ins[i][1], outs[i][0], outs[i][1], True)
        assert slot == i % 2
        if prev is not None:
            eng.forward_topk_host_wait(prev)
        prev = slot
    eng.forward_topk_host_wait(prev)
    for i in range(4):
        assert torch.equal(outs[i][0], ref[i][0].cpu()) and torch.equal(outs[i][1], ref[i][1].cpu()), i
    # two outstanding submits are the limit
    s0 = eng.forward_topk_host_submit(ins[0][0], ins[0][1], outs[0][0], outs[0][1], True)
    s1 = eng.forward_topk_host_submit(ins[1][0], ins[1][1], outs[1][0], outs[1][1], True)
    with pytest.raises(ValueError):
        eng.forward_topk_host_submit(ins[2][0], ins[2][1], outs[2][0], outs[2][1], True)
    eng.forward_topk_host_wait(s0)
    eng.forward_topk_host_wait(s1)
    # the synchronous call still works
    i2, v2 = eng.forward_topk_host(ins[2][0], ins[2][1], outs[2][0], outs[2][1], True)
    assert torch.equal(i2, ref[2][0].cpu())
    eng.close()
