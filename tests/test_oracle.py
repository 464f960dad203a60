"""CPU tests of the oracle itself: the reference has no tests or golden vectors for this path
(PARITY UNPINNED, SURVEY.md sections 4 and 8c), so each quirk Q1-Q18 that an "obviously right"
re-implementation would get wrong is pinned here against the reference's source text, together
with structural invariants and the committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import O, case, synth

D64 = torch.float64


def _w64(W):
    return O._cast(W, D64)


def test_q1_mark_values_index_mark_embs():
    """EasyDGL.py:87-88: multi-hot VALUES are used as indices -> mcode = nnz * mark_embs[1] (row 0 zero)."""
    cfg, inp, W = case("easy_a", batch=5)
    W = _w64(W)
    X0, _, _, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W, cfg, D64)
    d = cfg.num_units
    nnz = marks.sum(-1, keepdim=True).to(D64)
    assert torch.allclose(X0[:, :, 2 * d:], nnz * W["mark_embs"][1], atol=1e-12)


def test_q2_layernorm_joint_axes_population_variance():
    """Base.py:13,51-56: statistics over (L,d) jointly, population variance, eps 1e-12."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 8, generator=g, dtype=D64)
    gam, bet = torch.rand(8, generator=g, dtype=D64) + 0.5, torch.randn(8, generator=g, dtype=D64)
    y = O.layernorm(x, gam, bet)
    z = (x - x.mean(dim=(1, 2), keepdim=True)) / torch.sqrt(x.var(dim=(1, 2), unbiased=False, keepdim=True) + 1e-12)
    assert torch.allclose(y, z * gam + bet, atol=1e-12)
    # padded rows participate: changing one position changes every position's output
    x2 = x.clone()
    x2[:, 0] += 1.0
    assert (O.layernorm(x2, gam, bet)[:, 1:] - y[:, 1:]).abs().min() > 0


def test_q3_q15_bimau_residual_and_bidirectional():
    """temporal.py:447 residual adds queries[:,:,:d]; BiMAU has no causal mask (Q15)."""
    cfg, inp, W = case("easy_a", batch=4)
    W = _w64(W)
    X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W, cfg, D64)
    blk = dict(W["blocks"][0])
    d = cfg.num_units
    zero_v = dict(blk)
    zero_v["qkvt_w"] = blk["qkvt_w"].clone()
    zero_v["qkvt_w"][:, 2 * d:3 * d] = 0
    zero_v["qkvt_b"] = blk["qkvt_b"].clone()
    zero_v["qkvt_b"][2 * d:3 * d] = 0
    out, _ = O.bimau(X0, kmask, spans, marks, zero_v, d, cfg.num_heads, cfg.num_events)
    assert torch.allclose(out, X0[:, :, :d], atol=1e-12)
    # bidirectional: perturbing the LAST position changes the output at the first real position of row 0
    X1 = X0.clone()
    X1[0, -1] += 0.5
    a, _ = O.bimau(X0, kmask, spans, marks, blk, d, cfg.num_heads, cfg.num_events)
    b, _ = O.bimau(X1, kmask, spans, marks, blk, d, cfg.num_heads, cfg.num_events)
    assert (a[0, 0] - b[0, 0]).abs().max() > 1e-9


def test_q4_gate_not_renormalised_and_diag_one():
    """temporal.py:438-441: G multiplies P after softmax with no renormalisation; BiMAU sets diag(G)=1."""
    g = torch.Generator().manual_seed(1)
    B, L, d, h, E = 2, 6, 8, 2, 4
    w = {"int_w": torch.randn(d // h + 1, d // h * E, generator=g, dtype=D64),
         "int_b": torch.randn(d // h * E, generator=g, dtype=D64),
         "int_weight": torch.randn(E, d // h, generator=g, dtype=D64),
         "int_scaling": torch.randn(E, generator=g, dtype=D64) * 0.3}
    Q, K, V, T = (torch.randn(B, L, d, generator=g, dtype=D64) for _ in range(4))
    km = torch.ones(B, L, dtype=D64)
    iv = torch.rand(B, L, generator=g, dtype=D64)
    marks = torch.randint(0, 2, (B, L, E), generator=g)
    V1 = torch.ones_like(V)  # then O[q] = sum_k G[q,k] P[q,k]
    O_d, lam = O._attention_core(Q, K, V1, T, km, iv, marks, w, h, E, causal=False, diag_one=True, literal=True)
    O_n, _ = O._attention_core(Q, K, V1, T, km, iv, marks, w, h, E, causal=False, diag_one=False, literal=True)
    assert (O_d - 1.0).abs().max() > 1e-3, "rows of G o P must not sum to one"
    # difference between diag_one and not = (1 - G[q,q]) * P[q,q]
    Q_, K_ = O.fold_heads(Q, h), O.fold_heads(K, h)
    P = torch.softmax(Q_ @ K_.transpose(1, 2) / math.sqrt(d // h), -1)
    G = torch.einsum("nqe,nke->nqk", lam, marks.to(D64).repeat(h, 1, 1))
    diag = ((1 - torch.diagonal(G, dim1=1, dim2=2)) * torch.diagonal(P, dim1=1, dim2=2)).unsqueeze(-1)
    want = O.unfold_heads(diag.expand(-1, -1, d // h), h)
    assert torch.allclose(O_d - O_n, want, atol=1e-12)


def test_q5_head_major_fold_and_shared_intensity_weights():
    """temporal.py:413-416: index = head*B + b."""
    B, L, d, h = 3, 4, 6, 2
    x = torch.arange(B * L * d, dtype=D64).reshape(B, L, d)
    f = O.fold_heads(x, h)
    for head in range(h):
        for b in range(B):
            assert torch.equal(f[head * B + b], x[b, :, head * (d // h):(head + 1) * (d // h)])
    assert torch.equal(O.unfold_heads(f, h), x)


def test_q6_naive_softplus_overflows_like_tf():
    """temporal.py:305-306: s*log(1+exp(x/s)) overflows to inf for x/s > 88.7 in fp32."""
    E, dh = 4, 2
    w = {"int_w": torch.zeros(dh + 1, dh * E), "int_b": torch.full((dh * E,), 50.0),  # sigmoid -> 1
         "int_weight": torch.full((E, dh), 60.0), "int_scaling": torch.zeros(E)}
    H = torch.zeros(1, 3, dh)
    _, lam = O.intensity(H, torch.zeros(1, 3), torch.ones(1, 3, E, dtype=torch.int64), w, 1, E)
    assert torch.isinf(lam).all()  # 120 > 88.7
    w["int_weight"] = torch.full((E, dh), 10.0)
    _, lam = O.intensity(H, torch.zeros(1, 3), torch.ones(1, 3, E, dtype=torch.int64), w, 1, E)
    assert torch.allclose(lam, torch.full_like(lam, 20.0), atol=1e-5)


def test_q7_mask_token_semantics():
    """EasyDGL.py:76: [MASK] -> mark row 0, but it is a valid key with its own embedding and logit column."""
    cfg, inp, W = case("easy_a", batch=4)
    W64 = _w64(W)
    X0, kmask, _, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, D64)
    assert (inp["seqs_i"][:, -1] == cfg.mask_id).all()
    assert (marks[:, -1] == 0).all() and (kmask[:, -1] == 1).all()
    logits = O.easydgl_forward(inp["seqs_i"], inp["seqs_t"], W, cfg)
    assert logits.shape[1] == cfg.num_items + 1 and torch.isfinite(logits[:, cfg.mask_id]).all()


def test_q8_all_padding_row_gives_uniform_attention():
    """temporal.py:425-426: fill is -2^32+1 (finite), so an all-masked row is uniform 1/L, not NaN."""
    g = torch.Generator().manual_seed(2)
    B, L, d, h, E = 1, 5, 4, 1, 4
    w = {"int_w": torch.randn(d + 1, d * E, generator=g, dtype=D64), "int_b": torch.zeros(d * E, dtype=D64),
         "int_weight": torch.randn(E, d, generator=g, dtype=D64), "int_scaling": torch.zeros(E, dtype=D64)}
    Q, K, V, T = (torch.randn(B, L, d, generator=g, dtype=D64) for _ in range(4))
    marks = torch.zeros(B, L, E, dtype=torch.int64)  # G = 0 off-diagonal -> O[q] = P[q,q] V[q] = V[q]/L
    out, _ = O._attention_core(Q, K, V, T, torch.zeros(B, L, dtype=D64), torch.zeros(B, L, dtype=D64), marks, w, h, E,
                               causal=False, diag_one=True, literal=False)
    assert torch.allclose(out, V / L, atol=1e-12)
    assert O.MASK_FILL == -4294967295.0 and np.float32(O.MASK_FILL) == np.float32(-4294967296.0)


def test_q9_q10_spans_and_padded_time_code():
    """EasyDGL.py:73-74 spans clipped to [0,100] with spans[0]=spans[1]; Q10: ts=0 -> code [0,1,0,1,...]."""
    cfg = synth.make_config(model="EasyDGL", num_units=8, seqslen=4, num_items=50, num_events=4, time_scale=10.0)
    W = _w64(synth.make_weights(cfg, mode="parity"))
    ids = torch.tensor([[0, 0, 3, 4, cfg.mask_id]])
    ts = torch.tensor([[0.0, 0.0, 5000.0, 5020.0, 5010.0]])
    X0, _, spans, _ = O.easydgl_inputs(ids, ts, W, cfg, D64)
    assert spans.tolist() == [[0.0, 0.0, 100.0, 2.0, 0.0]]  # first real event sees a huge gap -> 100; negative -> 0
    code = X0[0, 0, :8] - 0.0  # id 0 -> zero item row
    assert torch.allclose(code, torch.tensor([0., 1.] * 4, dtype=D64))


def test_q11_q13_logits_layout():
    """Base.py:110, coding.py:57: column 0 = exactly -1000 (zero row, unscaled tied table); EasyDGL sizes (Q13)."""
    cfg, inp, W = case("easy_b", batch=3)
    logits = O.easydgl_forward(inp["seqs_i"], inp["seqs_t"], W, cfg)
    assert torch.equal(logits[:, 0], torch.full((3,), -1000.0, dtype=D64))
    assert cfg.L == cfg.seqslen + 1 and W["pos_embs"].shape[0] == cfg.L and W["item_embs"].shape[0] == cfg.num_items + 1
    assert W["output_bias"].shape[0] == cfg.num_items
    r = O.easydgl_forward(inp["seqs_i"], inp["seqs_t"], W, cfg, return_all=True)
    want = r.y @ O.zero_pad_table(W["item_embs"].double()).t() + O.output_bias(W["output_bias"].double())
    assert torch.allclose(r.logits, want)


def test_q12_ranking_semantics():
    """Base.py:156-164,181: -inf at every id in seqs_i (incl. 0 and [MASK]); top_k ties -> lower index."""
    logits = torch.tensor([[0.0, 3.0, 3.0, 1.0, 3.0, 2.0]], dtype=D64)
    ids = torch.tensor([[0, 4, 4]])
    v, i = O.eval_topk(logits, ids, True, 4, rank_on="probs")
    assert i.tolist() == [[1, 2, 5, 3]]
    v2, i2 = O.eval_topk(logits, ids, True, 4, rank_on="logits")
    assert torch.equal(i, i2)
    m = O.ranking_metrics(torch.tensor([[1, 2, 5, 3] + list(range(6, 102))]), torch.tensor([5]))
    assert m["H10"] == 1.0 and abs(m["N10"] - 1 / math.log2(4)) < 1e-12


def test_q16_ctsma_inputs():
    """CTSMA.py:48-49: S+1 timestamps, spans unclipped and forward-looking; X = [item*sqrt(d) | pos]."""
    cfg = synth.make_config(model="CTSMA", num_units=8, seqslen=3, num_items=50, num_events=4, time_scale=10.0)
    W = _w64(synth.make_weights(cfg, mode="parity"))
    ids = torch.tensor([[0, 3, 4]])
    ts = torch.tensor([[0.0, 5000.0, 5020.0, 4000.0]])
    X, kmask, spans, marks = O.ctsma_inputs(ids, ts, W, cfg, D64)
    assert spans.tolist() == [[500.0, 2.0, -102.0]]
    assert X.shape == (1, 3, 16) and kmask.tolist() == [[0.0, 1.0, 1.0]]
    assert torch.allclose(X[0, 1, :8], W["item_embs"][3] * math.sqrt(8))
    assert torch.equal(X[0, :, 8:], W["pos_embs"])


def test_q16_mau_is_causal_at_layer_level():
    cfg, _, W = case("ctsma_a", batch=3, edge=False)
    # full-length histories: a query whose visible keys are ALL padding attends uniformly to every key,
    # future ones included (Q8), so causality is only observable where a real key is visible
    inp = synth.make_inputs(cfg, 3, min_len=cfg.ts_len)
    W = _w64(W)
    X, kmask, spans, marks = O.ctsma_inputs(inp["seqs_i"], inp["seqs_t"], W, cfg, D64)
    blk = W["blocks"][0]
    a, _ = O.mau(X, X, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, True)
    X2 = X.clone()
    X2[:, -1] += 1.0  # perturb the last key/query
    b, _ = O.mau(X2, X2, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, True)
    assert torch.allclose(a[:, :-1], b[:, :-1], atol=1e-12)
    c, _ = O.mau(X2, X2, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, False)
    assert (c[:, 0] - a[:, 0]).abs().max() > 1e-9


def test_q17_second_block_is_width_d():
    cfg, inp, W = case("easy_a", batch=2)
    d = cfg.num_units
    assert W["blocks"][0]["qkvt_w"].shape == (3 * d, 4 * d) and W["blocks"][1]["qkvt_w"].shape == (d, 4 * d)


def test_q18_gelu_is_erf_form():
    x = torch.linspace(-3, 3, 61, dtype=D64)
    assert torch.allclose(O.gelu(x), torch.nn.functional.gelu(x), atol=1e-12)
    assert (O.gelu(x) - torch.nn.functional.gelu(x, approximate="tanh")).abs().max() > 1e-4


def test_time_sinusoid_code_layout():
    """coding.py:134-148: scale_j = 10000^(2j/d) (float64 -> fp32), [sin, cos] interleaved."""
    ts = torch.tensor([[3.0, 12000.5]])
    d = 6
    code = O.time_sinusoid_code(ts, d, D64)
    sc = np.power(10000, np.arange(0, d, 2) / d).astype(np.float32)
    for j in range(d // 2):
        x = np.float32(np.float32(12000.5) / sc[j])
        assert abs(code[0, 1, 2 * j].item() - math.sin(float(x))) < 1e-12
        assert abs(code[0, 1, 2 * j + 1].item() - math.cos(float(x))) < 1e-12
    with pytest.raises(AssertionError):
        O.time_sinusoid_code(torch.zeros(2, 3, 4), d, D64)


@pytest.mark.parametrize("name", ["easy_a", "easy_c", "ctsma_a"])
def test_literal_equals_contracted_and_fp32_close_to_fp64(name):
    cfg, inp, W = case(name, batch=5, edge=False)
    a = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=D64, literal=True)
    b = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=D64, literal=False)
    c = O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg, dtype=torch.float32)
    assert (a - b).abs().max() < 1e-11
    assert (a - c.double())[:, 1:].abs().max() <= 1e-5 * a[:, 1:].abs().max()


def test_head_permutation_equals_weight_column_permutation():
    cfg, inp, W = case("easy_a", batch=3)
    W64 = _w64(W)
    d, h = cfg.num_units, cfg.num_heads
    dh = d // h
    X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, D64)
    blk = W64["blocks"][0]
    perm = torch.tensor([2, 0, 3, 1])
    cols = torch.cat([torch.arange(p * dh, (p + 1) * dh) for p in perm])
    allcols = torch.cat([cols + i * d for i in range(4)])
    blk2 = dict(blk)
    blk2["qkvt_w"], blk2["qkvt_b"] = blk["qkvt_w"][:, allcols], blk["qkvt_b"][allcols]
    a, _ = O.bimau(X0, kmask, spans, marks, blk, d, h, cfg.num_events)
    b, _ = O.bimau(X0, kmask, spans, marks, blk2, d, h, cfg.num_events)
    assert torch.allclose((a - X0[:, :, :d])[:, :, cols], b - X0[:, :, :d], atol=1e-12)


def test_unknown_model_raises_like_util_ranking():
    cfg, inp, W = case("easy_a", batch=2)
    cfg.model = "SASREC"
    with pytest.raises(NotImplementedError):
        O.forward(inp["seqs_i"], inp["seqs_t"], W, cfg)


def test_golden_fixtures_reproduce():
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".pt"))
    assert len(files) >= 3
    for f in files:
        g = torch.load(os.path.join(gdir, f))
        cfg = synth.make_config(**g["cfg"])
        for literal in (False, True):
            r = O.forward(g["seqs_i"], g["seqs_t"], g["weights"], cfg, dtype=D64, literal=literal, return_all=True)
            assert (r.logits - g["logits64"]).abs().max() < 1e-9, f
            assert (r.lams[0] - g["lam64"]).abs().max() < 1e-9, f
        _, idx = O.eval_topk(g["logits64"], g["seqs_i"], True, g["topk_idx"].shape[1], rank_on="probs")
        assert torch.equal(idx, g["topk_idx"])


def test_topk_set_compare_excuses_only_near_ties():
    logits = torch.tensor([[0.0, 5.0, 4.0, 3.0, 2.9999999, 1.0]], dtype=D64)
    good = torch.tensor([[1, 2, 3]])
    swapped = torch.tensor([[1, 2, 4]])
    wrong = torch.tensor([[1, 2, 5]])
    assert O.topk_set_compare(good, logits, 3, 1e-6)["exact"] == 1
    assert O.topk_set_compare(swapped, logits, 3, 1e-6)["excused"] == 1
    assert O.topk_set_compare(wrong, logits, 3, 1e-6)["bad"] == 1


def test_train_forward_loss_restatement():
    """EasyDGL.py:153-189 / temporal.py:317-333: the training loss against an independent numpy computation of its
    three parts on a tiny case, and the weight-0 treatment of padded labels (EasyDGL.py:180)."""
    import numpy as np
    cfg, _, W = case("easy_a", batch=4)
    tr = synth.make_train_inputs(cfg, 4, masklen=3)
    r = O.train_forward(tr["seqs_i"], tr["seqs_t"], tr["labels"], W, cfg, tr["masked_positions"], l2_reg=1e-3, ct_reg=1e-2,
                        return_logits=True)
    lg = r["logits"].numpy()
    lab = tr["labels"].reshape(-1).numpy()
    p = np.exp(lg - lg.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    pe = -np.log(p[np.arange(len(lab)), lab] + 1e-5)
    w = (lab != 0).astype(np.float64)
    assert abs(float(r["ce"]) - (w * pe).sum() / (w.sum() + 1e-5)) < 1e-12
    l2 = 1e-3 * 0.5 * sum(float((W[n].double() ** 2).sum()) for n in ("item_embs", "pos_embs", "mark_embs"))
    assert abs(float(r["l2"]) - l2) < 1e-12
    assert float(r["ct"]) != 0. and abs(float(r["loss"]) - float(r["ce"] + r["l2"] + r["ct"])) < 1e-12
    # labels that are padding (0) carry weight 0: changing their logits row does not change the loss
    lab0 = tr["labels"].clone()
    lab0[:, 0] = 0
    r0 = O.train_forward(tr["seqs_i"], tr["seqs_t"], lab0, W, cfg, tr["masked_positions"])
    w0 = (lab0.reshape(-1).numpy() != 0)
    assert abs(float(r0["ce"]) - pe[w0].sum() / (w0.sum() + 1e-5)) < 1e-9
    # biased_likelihood: a row whose next item has no mark contributes nothing (sign(0) = 0, log(1) = 0)
    lam = torch.rand(2, 3, 4, dtype=torch.float64) + 0.1
    nm = torch.zeros(2, 3, 4, dtype=torch.float64)
    nm[0, 0, 1] = 1
    iv = torch.rand(2, 3, dtype=torch.float64)
    bl = O.biased_likelihood(lam.clone(), nm, iv)
    want = -(torch.log(lam[0, 0, 1]) - lam[0, 0].sum() * iv[0, 0] * .5) / 1.0
    assert abs(float(bl) - float(want)) < 1e-12


def test_tg_attention_decomposition_matches_literal():
    """TgMultiHeadAttention (temporal.py:204-264): the facade never builds the [N,T_q,T_k,2C] keys - it splits the K / V
    kernels into a per-key half and a time half (moved to the query side for the scores, applied after the weighted sum
    of the time codes for the values).  That algebra, restated in fp64 torch, must equal the literal restatement."""
    import numpy as np
    g = torch.Generator().manual_seed(4)
    N, T, C, h = 2, 9, 8, 2
    dh = C // h
    x = torch.randn(N, T, C, generator=g, dtype=torch.float64)
    ts = torch.cumsum(torch.rand(N, T, generator=g), 1)
    iv = (ts.unsqueeze(2) - ts.unsqueeze(1)).clamp_min(0.0).float()
    masks = torch.ones(N, T, T, dtype=torch.float64)
    masks[1, :, :3] = 0
    freq = torch.from_numpy(np.linspace(0, 9, C).astype(np.float32))
    phase = torch.randn(C, generator=g)
    w = {"q_w": torch.randn(2 * C, C, generator=g, dtype=torch.float64), "q_b": torch.randn(C, generator=g, dtype=torch.float64),
         "k_w": torch.randn(2 * C, C, generator=g, dtype=torch.float64), "k_b": torch.randn(C, generator=g, dtype=torch.float64),
         "v_w": torch.randn(2 * C, C, generator=g, dtype=torch.float64), "v_b": torch.randn(C, generator=g, dtype=torch.float64),
         "o_w": torch.randn(C, 2 * C, generator=g, dtype=torch.float64), "o_b": torch.randn(2 * C, generator=g, dtype=torch.float64),
         "ln_g": torch.ones(2 * C, dtype=torch.float64), "ln_b": torch.zeros(2 * C, dtype=torch.float64)}
    ref = O.tg_attention(x, x, masks, iv, w, freq, phase, h, True)
    code = O.time_function_code(iv, freq, phase, torch.float64)                    # [N,T,T,C]
    q2 = torch.cat([x, O.time_function_code(torch.zeros(N, T, 1), freq, phase, torch.float64).squeeze(2)], -1)
    Q = q2 @ w["q_w"] + w["q_b"]
    A_k, A_v = x @ w["k_w"][:C] + w["k_b"], x @ w["v_w"][:C] + w["v_b"]
    out = torch.zeros(N, T, C, dtype=torch.float64)
    for hd in range(h):
        sl = slice(hd * dh, (hd + 1) * dh)
        U = Q[:, :, sl] @ w["k_w"][C:, sl].t()                                      # [N,T,C]
        s = (Q[:, :, sl] @ A_k[:, :, sl].transpose(1, 2) + torch.einsum("nqc,nqkc->nqk", U, code)) / dh ** 0.5
        s = torch.where(masks == 0, torch.full_like(s, O.MASK_FILL), s)
        s = torch.where(torch.tril(torch.ones(T, T)) == 0, torch.full_like(s, O.MASK_FILL), s)
        P = torch.softmax(s, -1)
        TC = torch.einsum("nqk,nqkc->nqc", P, code)
        out[:, :, sl] = P @ A_v[:, :, sl] + TC @ w["v_w"][C:, sl]
    got = O.layernorm_last(out @ w["o_w"] + w["o_b"] + q2, w["ln_g"], w["ln_b"])
    assert float((got - ref).abs().max()) < 1e-10


def _ti_tf_inputs(N=2, T=7, C=8, seed=3):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(N, T, C, generator=g, dtype=torch.float64)
    k = torch.randn(N, T, C, generator=g, dtype=torch.float64)
    k[0, :2] = 0.0   # two padded keys in sequence 0
    w = {n: torch.randn(C, C, generator=g, dtype=torch.float64) for n in ("q_w", "k_w", "v_w")}
    w.update({n: torch.randn(C, generator=g, dtype=torch.float64) for n in ("q_b", "k_b", "v_b")})
    return g, q, k, w


def test_ti_attention_restatement_invariants():
    """TiMultiHeadAttention (temporal.py:37-109): with zero position / interval tables it is plain masked multi-head
    attention + residual; padded keys get no weight; causality blinds the future; an all-zero query row gets its
    attention output zeroed (query masking, temporal.py:87-90) and keeps only the residual."""
    g, q, k, w = _ti_tf_inputs()
    N, T, C = q.shape
    h = 2
    iv = torch.randint(0, 5, (N, T, T), generator=g)
    z_pos, z_tab = torch.zeros(T, C, dtype=torch.float64), torch.zeros(5, C, dtype=torch.float64)
    out = O.ti_attention(q, k, iv, w, z_pos, z_pos, z_tab, z_tab, h, causality=True)
    # plain attention by hand
    Q, K, V = q @ w["q_w"] + w["q_b"], k @ w["k_w"] + w["k_b"], k @ w["v_w"] + w["v_b"]
    ref = torch.zeros_like(out)
    dh = C // h
    for hd in range(h):
        sl = slice(hd * dh, (hd + 1) * dh)
        s = Q[:, :, sl] @ K[:, :, sl].transpose(1, 2) / dh ** 0.5
        km = (k.abs().sum(-1) != 0).unsqueeze(1).expand(N, T, T)
        s = torch.where(km, s, torch.full_like(s, O.MASK_FILL))
        s = torch.where(torch.tril(torch.ones(T, T)) == 0, torch.full_like(s, O.MASK_FILL), s)
        ref[:, :, sl] = torch.softmax(s, -1) @ V[:, :, sl]
    assert float((out - (ref + q)).abs().max()) < 1e-12
    # a non-trivial interval table changes the result, and only through intervals actually used
    tab = torch.randn(5, C, generator=g, dtype=torch.float64)
    out2 = O.ti_attention(q, k, iv, w, z_pos, z_pos, tab, tab, h, causality=True)
    assert float((out2 - out).abs().max()) > 1e-3
    # query masking: zero query row -> output row = residual = 0
    q2 = q.clone()
    q2[1, 3] = 0.0
    out3 = O.ti_attention(q2, k, iv, w, z_pos, z_pos, tab, tab, h, causality=True)
    assert float(out3[1, 3].abs().max()) == 0.0


def test_tf_attention_restatement_invariants():
    """TfMultiHeadAttention (temporal.py:126-185): the Bochner / Mercer code cos(dt w + phi) enters the scores only;
    with w = 0, phi = pi/2 the code vanishes (cos = 0 up to rounding) and the layer is plain attention."""
    g, q, k, w = _ti_tf_inputs(seed=8)
    N, T, C = q.shape
    iv = torch.rand(N, T, T, generator=g) * 10.0
    z_pos = torch.zeros(T, C, dtype=torch.float64)
    zero_code = O.tf_attention(q, k, iv, w, z_pos, torch.zeros(C), torch.full((C,), float(np.pi / 2)), 2, True)
    ti_plain = O.ti_attention(q, k, torch.zeros(N, T, T, dtype=torch.int64), w, z_pos, z_pos,
                              torch.zeros(1, C, dtype=torch.float64), torch.zeros(1, C, dtype=torch.float64), 2, True)
    # Tf has no query masking; these inputs have no all-zero query rows, so the two must agree
    assert float((zero_code - ti_plain).abs().max()) < 1e-6
    with_code = O.tf_attention(q, k, iv, w, z_pos, torch.linspace(0, 9, C), torch.zeros(C), 2, True)
    assert float((with_code - zero_code).abs().max()) > 1e-3
    # causality: rows 0..2 of the unpadded sequence see keys 0..2 only.  (Sequence 0 has its first two keys padded: its
    # rows 0, 1 have every visible key masked, all scores equal the fill value and the softmax is uniform over ALL keys,
    # future ones included - the reference's behaviour, Q8.)
    k2 = k.clone()
    k2[:, 3:] += 1.0
    a = O.tf_attention(q, k, iv, w, z_pos, torch.linspace(0, 9, C), torch.zeros(C), 2, True)
    b = O.tf_attention(q, k2, iv, w, z_pos, torch.linspace(0, 9, C), torch.zeros(C), 2, True)
    assert float((a[1, :3] - b[1, :3]).abs().max()) < 1e-12
    assert float((a[0, :2] - b[0, :2]).abs().max()) > 1e-3
