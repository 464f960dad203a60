"""GPU parity of every layer-level C-ABI entry point against the CPU oracle.

Each test reads like a test the reference could have had for the layer it names
(the reference has none).  Tolerance: 1e-3 relative (north star),
stated per assertion; integer outputs bit-exact.
"""
import math

import pytest
import numpy as np
import torch

from helpers import O, assert_close, case, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _engine(cfg, W, max_batch, **kw):
    from easydgl_b200.engine import Engine
    return Engine(cfg, W, max_batch=max_batch, device=DEV, **kw)


def test_library_loads_and_reports_version():
    from easydgl_b200 import _lib
    lib = _lib.load()
    assert lib.edgl_version() >= 100


@pytest.mark.parametrize("d", [8, 50, 64, 128])
def test_time_sinusoid_code(d):
    """C.TimeSinusoidCoding.code (coding.py:137-149) on Netflix-scale scaled timestamps."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(1)
    ts = (torch.rand(7, 33, generator=g) * 2000 + 10800).float()  # days since epoch / 1
    ts[0, :5] = 0.0  # padded slots (Q10)
    out = engine.time_sinusoid_code(ts.to(DEV), d).cpu()
    ref32 = O.time_sinusoid_code(ts, d, torch.float32)
    ref64 = O.time_sinusoid_code(ts, d, torch.float64)
    assert_close(out, ref64, 1e-5, "time code vs fp64")
    assert_close(out, ref32, 1e-5, "time code vs fp32")
    assert torch.equal(out[0, 0, 0::2], torch.zeros(d // 2)) and torch.equal(out[0, 0, 1::2], torch.ones(d // 2))


def test_time_function_code_bochner_mercer():
    """C.TimeFunctionCoding.code (coding.py:97-122): cos(t * basis_freq + phase), rank-2 and rank-3 inputs."""
    from easydgl_b200.module import coding as C
    g = torch.Generator().manual_seed(8)
    tc = C.TimeFunctionCoding(24, device=DEV)
    assert np.array_equal(tc.basis_freq.cpu().numpy(), np.linspace(0, 9, 24).astype(np.float32))  # coding.py:109
    ts = (torch.rand(5, 17, generator=g) * 2000 + 10800).float()
    out = tc.code(ts.to(DEV)).cpu()
    ref = O.time_function_code(ts, tc.basis_freq.cpu(), tc.phase.cpu(), torch.float64)
    assert out.shape == (5, 17, 1, 24)
    assert_close(out, ref, 1e-5, "time function code")
    iv = (ts[:, :, None] - ts[:, None, :])                       # [B,L,L] interval matrix (TGAT / TGSRec usage)
    tc2 = C.TimeFunctionCoding(8, device=DEV, phase=torch.randn(8, generator=g))
    out2 = tc2.code(iv.to(DEV)).cpu()
    assert out2.shape == (5, 17, 17, 8)
    assert_close(out2, O.time_function_code(iv, tc2.basis_freq.cpu(), tc2.phase.cpu(), torch.float64), 1e-5, "tif 3-D")


def test_time_sinusoid_code_rank_assert():
    from easydgl_b200 import engine
    with pytest.raises(AssertionError):
        engine.time_sinusoid_code(torch.zeros(3, 4, 5, device=DEV), 8)


@pytest.mark.parametrize("zero_pad,scale", [(True, True), (True, False), (False, False)])
def test_embedding_lookup(zero_pad, scale):
    """C.Embedding.__call__ (coding.py:45-64): bit-exact gather (x sqrt(d) is one fp32 multiply)."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(2)
    table = torch.randn(97, 24, generator=g)
    ids = torch.randint(0, 97, (5, 11), generator=g)
    ids[0, 0] = 0
    out = engine.embedding_lookup(table.to(DEV), ids.to(DEV), zero_pad, scale).cpu()
    tab = O.zero_pad_table(table) if zero_pad else table
    ref = O.embedding(tab, ids, scale, 24)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("name", ["easy_a", "easy_b", "easy_c", "ctsma_a", "ctsma_b"])
def test_embed_input_assembly(name):
    """EasyDGL.py:70-95 / CTSMA.py:47-60: X0, spans, marks."""
    cfg, inp, W = case(name, batch=9)
    eng = _engine(cfg, W, 9)
    X0, spans, marks = eng.embed(inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV))
    fn = O.easydgl_inputs if cfg.model == "EasyDGL" else O.ctsma_inputs
    rX, rk, rs, rm = fn(inp["seqs_i"], inp["seqs_t"], O._cast(W, torch.float64), cfg, torch.float64)
    assert torch.equal(marks.cpu().long(), rm), "marks must be bit-exact"
    assert_close(spans.cpu(), rs, 1e-6, "spans")
    assert_close(X0.cpu(), rX, 1e-5, "X0")


@pytest.mark.parametrize("B,L,C", [(3, 7, 8), (5, 31, 64), (2, 100, 128), (2, 300, 256)])
def test_layernorm_joint_axes(B, L, C):
    """Base.layernorm (Base.py:12-67): statistics over (L,C) jointly (Q2)."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, L, C, generator=g) * 2.0 + 0.7
    gam = torch.randn(C, generator=g) * 0.2 + 1.0
    bet = torch.randn(C, generator=g) * 0.3
    out = engine.layernorm(x.to(DEV), gam.to(DEV), bet.to(DEV)).cpu()
    ref = O.layernorm(x.double(), gam.double(), bet.double())
    assert_close(out, ref, 1e-5, "layernorm")
    # it is NOT the per-position layernorm
    per_pos = torch.nn.functional.layer_norm(x, (C,), gam, bet, 1e-12)
    assert (out - per_pos).abs().max() > 1e-3


@pytest.mark.parametrize("M,K,N,act", [(5, 17, 9, 0), (130, 64, 128, 1), (257, 144, 512, 0), (64, 130, 66, 2),
                                       (1000, 384, 512, 1)])
def test_dense(M, K, N, act):
    """tf.layers.dense with none / gelu-erf (EasyDGL.py:19-32) / relu."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(4)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    out = engine.dense(x.to(DEV), w.to(DEV), b.to(DEV), act).cpu()
    ref = x.double() @ w.double() + b.double()
    if act == 1:
        ref = O.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    assert_close(out, ref, 1e-5, "dense act=%d" % act)


@pytest.mark.parametrize("M,K,N,act", [(128, 32, 128, 0), (5, 32, 16, 0), (130, 64, 128, 1), (257, 144, 512, 0),
                                       (300, 128, 256, 2), (1000, 256, 128, 1), (77, 128, 18001, 0),
                                       (40000, 128, 128, 0), (20000, 144, 512, 1)])
def test_dense_tensor_core(M, K, N, act):
    """The tcgen05 (3xTF32) GEMM: fp32-level accuracy, M/N/K tails, both tile widths, many tiles per CTA."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(40 + M % 7)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    out = engine.dense_nk(x.to(DEV), w.t().contiguous().to(DEV), b.to(DEV), act).cpu()
    ref = x.double() @ w.double() + b.double()
    if act == 1:
        ref = O.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    assert_close(out, ref, 1e-5, "tensor-core dense act=%d" % act)


@pytest.mark.parametrize("M,K,N,act", [(128, 32, 128, 0), (5, 32, 16, 0), (130, 64, 128, 1), (257, 144, 512, 0),
                                       (300, 128, 256, 2), (1000, 256, 128, 1), (77, 128, 18001, 0),
                                       (40000, 128, 128, 0), (20000, 144, 512, 1)])
def test_dense_tensor_core_f16_split(M, K, N, act):
    """The tcgen05 kind::f16 GEMM with the scaled 3xFP16 split (gemm_f16.cu): same bar as the 3xTF32 kernel -
    fp32-level accuracy relative to max|ref| - with rows whose magnitudes differ by 2^12 and a K tail."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(60 + M % 7)
    x = torch.randn(M, K, generator=g) * torch.exp2(-torch.randint(0, 13, (M, 1), generator=g).float())
    w = torch.randn(K, N, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 1e-3
    out = engine.dense_nk(x.to(DEV), w.t().contiguous().to(DEV), b.to(DEV), act, f16=True).cpu()
    ref = x.double() @ w.double() + b.double()
    if act == 1:
        ref = O.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    assert_close(out, ref, 1e-5, "3xFP16 tensor-core dense act=%d" % act)


@pytest.mark.parametrize("name", ["easy_a", "easy_b", "easy_c", "ctsma_b"])
def test_intensity(name):
    """T.MAU.intensity (temporal.py:281-315): G [hB,L,L] and lam [hB,L,E], literal 4-D form as reference."""
    cfg, inp, W = case(name, batch=5)
    eng = _engine(cfg, W, 5)
    h, dh, E, L = cfg.num_heads, cfg.num_units // cfg.num_heads, cfg.num_events, cfg.L
    g = torch.Generator().manual_seed(5)
    H = torch.randn(h * 5, L, dh, generator=g)
    iv = torch.rand(5, L, generator=g) * 30
    fn = O.easydgl_inputs if cfg.model == "EasyDGL" else O.ctsma_inputs
    _, _, _, marks = fn(inp["seqs_i"], inp["seqs_t"], O._cast(W, torch.float64), cfg, torch.float64)
    blk = O._cast(W["blocks"][0], torch.float64)
    rG, rl = O.intensity(H.double(), iv.double(), marks, blk, h, E, literal=True)
    G, lam = eng.intensity(0, H.to(DEV), iv.to(DEV), marks.to(torch.uint8).to(DEV))
    assert_close(lam.cpu(), rl, 1e-4, "lam")
    assert_close(G.cpu(), rG, 1e-4, "G")


@pytest.mark.parametrize("proj", ["tc", "simt"])
@pytest.mark.parametrize("name", ["easy_a", "easy_b", "easy_c", "easy_d"])
def test_bimau_layer(name, proj, monkeypatch):
    """T.BiMAU.__call__ (temporal.py:404-452) incl. all-padding rows (Q8) and set_diag (Q4); QKVT projected by the
    tcgen05 GEMM the pipeline uses (proj="tc") and by the exact-fp32 CUDA-core GEMM (proj="simt")."""
    monkeypatch.setenv("EDGL_LAYER_GEMM", proj)
    cfg, inp, W = case(name, batch=6)
    eng = _engine(cfg, W, 6)
    W64 = O._cast(W, torch.float64)
    X0, kmask, spans, marks = O.easydgl_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
    blk = W64["blocks"][0]
    rO, rl = O.bimau(X0, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, literal=True)
    out, lam = eng.attention_layer(0, X0.float().to(DEV), None, kmask.to(torch.uint8).to(DEV),
                                   spans.float().to(DEV), marks.to(torch.uint8).to(DEV))
    assert_close(lam.cpu(), rl, 1e-4, "BiMAU lam")
    assert_close(out.cpu(), rO, 1e-4, "BiMAU out")


@pytest.mark.parametrize("proj", ["tc", "simt"])
@pytest.mark.parametrize("name,causal", [("ctsma_a", True), ("ctsma_b", True), ("ctsma_b", False)])
def test_mau_layer(name, causal, proj, monkeypatch):
    """T.MAU.__call__ (temporal.py:335-390): separate Q (from LN'd queries) and K/V/T (raw keys), causal mask."""
    monkeypatch.setenv("EDGL_LAYER_GEMM", proj)
    cfg, inp, W = case(name, batch=6)
    eng = _engine(cfg, W, 6)
    W64 = O._cast(W, torch.float64)
    X, kmask, spans, marks = O.ctsma_inputs(inp["seqs_i"], inp["seqs_t"], W64, cfg, torch.float64)
    blk = W64["blocks"][0]
    qin = O.layernorm(X, blk["ln1_g"], blk["ln1_b"])
    rO, rl = O.mau(qin, X, kmask, spans, marks, blk, cfg.num_units, cfg.num_heads, cfg.num_events, causal, literal=True)
    out, lam = eng.attention_layer(0, qin.float().to(DEV), X.float().to(DEV), kmask.to(torch.uint8).to(DEV),
                                   spans.float().to(DEV), marks.to(torch.uint8).to(DEV), causality=causal)
    assert_close(lam.cpu(), rl, 1e-4, "MAU lam")
    assert_close(out.cpu(), rO, 1e-4, "MAU out")
    if causal:
        # causal + all-padded history -> uniform attention over ALL keys (the -2^32+1 fill is finite, Q8)
        assert torch.isfinite(out).all()


def test_topk_ties_and_masking():
    """tf.nn.top_k (Base.py:181): sorted, ties -> lower index; seen ids -> -inf (Base.py:156-163)."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(6)
    B, N, K = 9, 1000, 100
    logits = torch.randn(B, N, generator=g)
    logits[0] = 0.5                       # all tied -> indices 0..K-1
    logits[1, :] = torch.round(logits[1] * 2) / 2  # heavy ties
    logits[2, 10:400] = float("-inf")     # many masked
    logits[3, :] = float("-inf")          # everything masked -> ties at -inf, lowest indices
    logits[4, 5] = float("inf")
    seen = torch.randint(0, N, (B, 17), generator=g)
    ref_v, ref_i = O.eval_topk(logits.double(), seen, True, K, rank_on="logits")
    lg = logits.clone().to(DEV)
    idx, val = engine.topk(lg, K, seen.to(DEV))
    assert torch.equal(idx.cpu().long(), ref_i), "top-K indices must be bit-identical"
    assert torch.equal(val.cpu().double(), ref_v)
    # in-place masking happened
    assert torch.isinf(lg.cpu()[5, seen[5]]).all()


@pytest.mark.parametrize("N,K", [(31, 100), (1000, 100), (2251, 100), (3072, 100), (2251, 7), (1500, 128)])
@pytest.mark.parametrize("warp", ["default", "1", "0"])
def test_topk_short_rows_warp_kernel(N, K, warp, monkeypatch):
    """Many short rows (the column shards of the multi-GPU path) through the two-pass one-warp-per-row kernel with its
    overflow list (the default for 256 <= N <= 4096), the register-resident first version (EDGL_TOPK_WARP=1) and the
    CTA-per-row kernel (EDGL_TOPK_WARP=0): bit-identical to tf.nn.top_k semantics, including ties at the cut, -inf and
    rows shorter than K."""
    from easydgl_b200 import engine
    if warp == "default":
        monkeypatch.delenv("EDGL_TOPK_WARP", raising=False)
    else:
        monkeypatch.setenv("EDGL_TOPK_WARP", warp)
    g = torch.Generator().manual_seed(60 + N + K)
    B = 300
    logits = torch.randn(B, N, generator=g)
    logits[0] = 0.5                                             # all tied -> indices 0..K-1
    logits[1, :] = torch.round(logits[1] * 2) / 2               # heavy ties, also at the cut
    logits[2, N // 10:N // 2] = float("-inf")                   # many masked
    logits[3, :] = float("-inf")                                # everything masked
    logits[4, min(5, N - 1)] = float("inf")
    logits[5, :] = torch.round(logits[5])                       # a handful of distinct values
    logits[6, ::2] = -0.0
    logits[6, 1::2] = 0.0                                       # -0.0 < +0.0 in the key order
    seen = torch.randint(0, N, (B, 17), generator=g)
    ref_v, ref_i = O.eval_topk(logits.double(), seen, True, min(K, N), rank_on="logits")
    idx, val = engine.topk(logits.clone().to(DEV), K, seen.to(DEV))
    idx, val = idx.cpu().long(), val.cpu().double()
    rows = [r for r in range(B) if r != 6]                       # row 6: signed zeros are ordered by sign bit here
    assert torch.equal(idx[rows, :min(K, N)], ref_i[rows]), "top-K indices must be bit-identical"
    assert torch.equal(val[rows, :min(K, N)], ref_v[rows])
    if N < K:
        assert bool((idx[:, N:] == -1).all()) and bool(torch.isinf(val[:, N:]).all())
    if N >= 1000:
        assert bool((val[6, :K] == 0).all()) and bool((idx[6, :K] % 2 == 1).all())


def test_topk_small_n_pads():
    from easydgl_b200 import engine
    logits = torch.tensor([[0.1, 0.7, -0.2]], device=DEV)
    idx, val = engine.topk(logits, 5)
    assert idx.cpu().tolist() == [[1, 0, 2, -1, -1]]


@pytest.mark.parametrize("G", [2, 3, 8])
def test_topk_merge_equals_global(G):
    """merging per-shard top-K must equal the single-device top-K bit for bit."""
    from easydgl_b200 import engine
    g = torch.Generator().manual_seed(7)
    B, N, K = 11, 4001, 100
    logits = torch.round(torch.randn(B, N, generator=g) * 8) / 8  # ties across shards
    gi, gv = engine.topk(logits.clone().to(DEV), K)
    per = (N + G - 1) // G
    ci, cv = [], []
    for r in range(G):
        c0, c1 = r * per, min(N, (r + 1) * per)
        i, v = engine.topk(logits[:, c0:c1].contiguous().to(DEV), K)
        i = torch.where(i >= 0, i + c0, i)
        ci.append(i)
        cv.append(v)
    mi, mv = engine.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, gi) and torch.equal(mv, gv)
    # lists that are NOT sorted (the kernel's general path) and lists with padding give the same answer
    perm = torch.randperm(K, generator=g)
    mi2, mv2 = engine.topk_merge(torch.stack(cv)[:, :, perm].contiguous(), torch.stack(ci)[:, :, perm].contiguous())
    assert torch.equal(mi2, gi) and torch.equal(mv2, gv)
    cvp, cip = torch.stack(cv).clone(), torch.stack(ci).clone()
    cvp[:, :, K // 2:] = float("-inf")
    cip[:, :, K // 2:] = -1
    mi3, mv3 = engine.topk_merge(cvp, cip)
    half = torch.cat([c[:, :K // 2] for c in cv], 1).cpu()
    hidx = torch.cat([c[:, :K // 2] for c in ci], 1).cpu()
    rv, ri = O.topk_lower_index_first(half.double(), min(K, half.shape[1]))
    # reference: sort the G*K/2 surviving candidates by (value desc, global index asc)
    order = torch.argsort(hidx, dim=1, stable=True)
    hv, hi = torch.gather(half, 1, order), torch.gather(hidx, 1, order)
    o2 = torch.argsort(-hv.double(), dim=1, stable=True)[:, :K]
    want_i, want_v = torch.gather(hi, 1, o2), torch.gather(hv, 1, o2)
    n_valid = want_i.shape[1]
    assert torch.equal(mi3.cpu()[:, :n_valid], want_i.to(torch.int32)) and torch.equal(mv3.cpu()[:, :n_valid], want_v)
    assert bool((mi3.cpu()[:, n_valid:] == -1).all())
