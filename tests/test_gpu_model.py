"""GPU parity of the whole forward path (model(features, False) and model.eval ranking)
against the CPU oracle, through the C ABI.  Tolerance 1e-3 relative on logits (north
star); top-K index sets identical up to near-ties at the cut (tie-aware)."""
import json
import os

import pytest
import torch

from helpers import O, assert_close, case, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _engine(cfg, W, max_batch, **kw):
    from easydgl_b200.engine import Engine
    return Engine(cfg, W, max_batch=max_batch, device=DEV, **kw)


def _log(name, rec):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity.jsonl"), "a") as fh:
        fh.write(json.dumps(dict(test=name, **rec)) + "\n")


def _check_model(name, batch, mode="parity", k=100, **over):
    cfg, inp, W = case(name, batch=batch, mode=mode, **over)
    k = min(k, cfg.num_rows)
    eng = _engine(cfg, W, batch, topk=k)
    ids, ts = inp["seqs_i"], inp["seqs_t"]
    logits = eng.forward_logits(ids.to(DEV), ts.to(DEV)).cpu()
    ref64 = O.forward(ids, ts, W, cfg, dtype=torch.float64)
    ref32 = O.forward(ids, ts, W, cfg, dtype=torch.float32)
    assert torch.equal(logits[:, 0], torch.full((batch,), -1000.0)), "column 0 must be exactly -1000 (Q11)"
    assert torch.isfinite(logits).all()
    # Rows on which the reference ITSELF is ill-conditioned are excluded from the tolerance check:
    # e.g. an all-padding CTSMA row with zero biases makes LayerNorm's input constant (variance = rounding
    # noise, Base.py:51-56), so the fp32 and fp64 oracles disagree by O(1) there.  A row is "well-posed"
    # when the fp32 and fp64 oracles agree to 1e-4.
    well = (ref32.double() - ref64)[:, 1:].abs().amax(1) <= 1e-4 * ref64[:, 1:].abs().max()
    assert int(well.sum()) >= batch - 3, "too many ill-conditioned rows: %s" % well
    e64 = assert_close(logits[well, 1:], ref64[well, 1:], 1e-3, name + " logits vs fp64 oracle")
    e32 = assert_close(logits[well, 1:], ref32[well, 1:], 1e-3, name + " logits vs fp32 oracle")
    # ranking
    idx, val = eng.forward_topk(ids.to(DEV), ts.to(DEV), mask_seen=True)
    idx, val = idx.cpu().long(), val.cpu()
    masked64 = O.mask_seen_logits(ref64, ids)
    abs_err = float((logits.double() - ref64)[well].abs().max())
    res = O.topk_set_compare(idx[well], masked64[well], k, tau=4 * abs_err)
    # the returned values are the masked logits of the returned indices, sorted, ties by index
    mine = O.mask_seen_logits(logits, ids)
    assert torch.equal(val, torch.gather(mine, 1, idx)), "top-K values must be the kernel's own masked logits"
    ri_v, ri_i = O.eval_topk(logits, ids, True, k, rank_on="logits")
    assert torch.equal(idx, ri_i), "top-K must be the exact ranking of the kernel's own logits"
    assert res["bad"] == 0, res
    # literal reference ranking (softmax then top_k, Base.py:164,181) agrees with ranking logits in fp64
    _, lit = O.eval_topk(ref64, ids, True, k, rank_on="probs")
    _, lg = O.eval_topk(ref64, ids, True, k, rank_on="logits")
    finite = torch.isfinite(torch.gather(masked64, 1, lg)).all(dim=1)
    _log(name, dict(batch=batch, well_posed_rows=int(well.sum()), rel_err_fp64=e64, rel_err_fp32=e32, abs_err=abs_err, topk=res,
                    probs_vs_logits_rank_equal=bool(torch.equal(lit[finite], lg[finite]))))
    return eng, cfg, inp, W, logits, idx, val


@pytest.mark.parametrize("name", ["easy_a", "easy_b", "easy_c", "easy_d", "ctsma_a", "ctsma_b"])
def test_forward_small_parity_weights(name):
    _check_model(name, batch=8)


@pytest.mark.parametrize("name", ["easy_b", "ctsma_b"])
def test_forward_reference_initialisers(name):
    """Weights drawn with the reference's own initialisers (glorot / N(0,0.02) / zeros / ones)."""
    _check_model(name, batch=8, mode="reference")


def test_forward_reference_default_length():
    """The reference default --seqslen=30 (main.py:38) -> L=31, a non-multiple-of-8 length."""
    _check_model("easy_b", batch=5, seqslen=30)


def test_forward_onehot_marks():
    _check_model("easy_a", batch=6, onehot=True)


def test_forward_c1_config():
    """BASELINE.json configs[0]: EasyDGL d=64 L=100 18K items B=32 h=8 (the reference's CPU case)."""
    _check_model("C1", batch=32)


def test_forward_c3_shape_small_batch():
    """BASELINE.json configs[2] shape (CTSMA d=64 L=100 h=4 blocks=2) at a small batch."""
    _check_model("C3", batch=16)


def test_forward_c2_shape_small_batch():
    """BASELINE.json configs[1] shape (EasyDGL d=128 L=100 h=8) at a small batch."""
    _check_model("C2", batch=16)


def test_batch_invariance_and_determinism():
    """Sequences are independent (every op incl. LayerNorm is per sample): a row's logits must not
    depend on its batch-mates, and two runs must agree bit for bit."""
    cfg, inp, W = case("easy_b", batch=12)
    eng = _engine(cfg, W, 12)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    a = eng.forward_logits(ids, ts)
    b = eng.forward_logits(ids, ts)
    assert torch.equal(a, b)
    c = eng.forward_logits(ids[3:7].contiguous(), ts[3:7].contiguous())
    assert torch.equal(a[3:7], c)


@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_topk_equals_single(G):
    """Column-sharded logits + per-shard top-K + merge == single-device top-K, bit for bit."""
    from easydgl_b200 import engine
    cfg, inp, W = case("easy_b", batch=10)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    single = _engine(cfg, W, 10)
    gi, gv = single.forward_topk(ids, ts, True)
    y = single.encode(ids, ts)
    ci, cv = [], []
    for r in range(G):
        sh = _engine(cfg, W, 10, shard_rank=r, shard_world=G)
        i, v = sh.logits_topk(y, ids)
        ci.append(i)
        cv.append(v)
    mi, mv = engine.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, gi) and torch.equal(mv, gv)
    # the interleaved exchange layout ([rows, 2, K], strides 2K) used by easydgl_b200/sharded.py
    K = single.K
    buf = torch.empty((G, 10, 2, K), dtype=torch.int32, device=DEV)
    for r in range(G):
        sh = _engine(cfg, W, 10, shard_rank=r, shard_world=G)
        sh.logits_topk(y, ids, out=(buf[r, :, 0], buf[r, :, 1].view(torch.float32)), out_stride=2 * K)
    from easydgl_b200.sharded import _merge_packed_cuda
    pi, pv = _merge_packed_cuda(buf, 0, 10)
    assert torch.equal(pi, gi) and torch.equal(pv, gv)


def test_host_entry_point_matches_device_entry_point():
    cfg, inp, W = case("easy_a", batch=7)
    eng = _engine(cfg, W, 7)
    ids, ts = inp["seqs_i"], inp["seqs_t"]
    gi, gv = eng.forward_topk(ids.to(DEV), ts.to(DEV), True)
    K = eng.K
    hi = torch.empty((7, K), dtype=torch.int32).pin_memory()
    hv = torch.empty((7, K), dtype=torch.float32).pin_memory()
    eng.forward_topk_host(ids.pin_memory(), ts.pin_memory(), hi, hv, True)
    assert torch.equal(hi, gi.cpu()) and torch.equal(hv, gv.cpu())


def test_error_conventions():
    """shape / state errors surface as exceptions, not silent garbage."""
    from easydgl_b200._lib import EdglError
    cfg, inp, W = case("easy_a", batch=4)
    eng = _engine(cfg, W, 4)
    ids, ts = inp["seqs_i"].to(DEV), inp["seqs_t"].to(DEV)
    with pytest.raises(ValueError):
        eng.forward_logits(ids[:, :-1].contiguous(), ts)
    with pytest.raises(ValueError):
        eng.forward_logits(torch.cat([ids, ids, ids]), torch.cat([ts, ts, ts]))  # > max_batch
    with pytest.raises(AssertionError):
        eng.forward_logits(ids, ts.reshape(-1))
    W2 = dict(W)
    W2["mark_table"] = W["mark_table"] + cfg.num_events  # values no longer index mark_embs
    with pytest.raises(ValueError):
        _engine(cfg, W2, 4)
    W3 = {k: v for k, v in W.items() if k != "tr_w"}
    with pytest.raises(EdglError):
        _engine(cfg, W3, 4)


def test_golden_fixtures():
    """Committed fixtures (tests/golden/, made by tests/golden/make_golden.py from the fp64 oracle)."""
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    files = sorted(f for f in os.listdir(gdir) if f.endswith(".pt"))
    assert files, "no golden fixtures"
    for f in files:
        g = torch.load(os.path.join(gdir, f))
        cfg = synth.make_config(**g["cfg"])
        eng = _engine(cfg, g["weights"], g["seqs_i"].shape[0], topk=g["topk_idx"].shape[1])
        logits = eng.forward_logits(g["seqs_i"].to(DEV), g["seqs_t"].to(DEV)).cpu()
        assert_close(logits[:, 1:], g["logits64"][:, 1:], 1e-3, f)
        idx, _ = eng.forward_topk(g["seqs_i"].to(DEV), g["seqs_t"].to(DEV), True)
        masked = O.mask_seen_logits(g["logits64"], g["seqs_i"])
        err = float((logits.double() - g["logits64"]).abs().max())
        res = O.topk_set_compare(idx.cpu().long(), masked, idx.shape[1], tau=4 * err)
        assert res["bad"] == 0, (f, res)


@pytest.mark.parametrize("mode", ["simt", "tc", "mma", "f16"])
def test_alternative_attention_kernels_agree(mode):
    """EDGL_ATTN=simt forces the CUDA-core attention kernel (the fallback for shapes the tensor-core
    kernels are not instantiated for); EDGL_ATTN=tc selects the tcgen05/TMEM kernel (dh=16, E=16, L<=128).
    Both must satisfy the same parity bar as the default mma.sync kernel.  Run in a subprocess because
    the switch is read once per process."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from helpers import O, case, assert_close\n"
        "from easydgl_b200.engine import Engine\n"
        "for name in ('easy_b', 'ctsma_b', 'C1', 'C2', 'C3'):\n"
        "    cfg, inp, W = case(name, batch=8)\n"
        "    eng = Engine(cfg, W, max_batch=8, device='cuda:0')\n"
        "    lg = eng.forward_logits(inp['seqs_i'].cuda(), inp['seqs_t'].cuda()).cpu()\n"
        "    ref = O.forward(inp['seqs_i'], inp['seqs_t'], W, cfg, dtype=torch.float64)\n"
        "    assert_close(lg[:, 1:], ref[:, 1:], 1e-3, name)\n"
        "print('SIMT_OK')\n" % (os.path.dirname(os.path.abspath(__file__)),
                                os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    env = dict(os.environ, EDGL_ATTN=mode)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=280)
    assert "SIMT_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("mode", ["mma", "f16", "f16-long"])
def test_attention_variant_layers(mode):
    """Layer-level (BiMAU / MAU) and model-level parity of one tensor-core attention kernel at fp32-level
    tolerance (2e-5 of max|ref| vs the fp64 oracle), with Q/K/V/T/MLP operands rescaled by up to 2^+-40: the
    scaled 3xFP16 kernel (attn_f16.cu) must hold the same accuracy as the 3xTF32 one over the whole range."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "attn_variant_check.py")
    env = dict(os.environ, EDGL_ATTN=mode.split("-")[0])
    if mode.endswith("-long"):  # the key-streaming two-pass kernel (attn_f16_long.cu) for every shape
        env["EDGL_ATTN_LONG"] = "1"
    res = subprocess.run([sys.executable, script], env=env, capture_output=True, text=True, timeout=280)
    assert "VARIANT_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


@pytest.mark.parametrize("env", [{"EDGL_TC_EPI": "staged"}, {"EDGL_GELU": "erf"}, {"EDGL_TC_NOBLO": "1"},
                                 {"EDGL_GEMM": "tf32"}],
                         ids=lambda e: "-".join("%s=%s" % kv for kv in e.items()))
def test_dense_layer_variants_agree(env):
    """The selectable implementations of the tensor-core dense layers (staged epilogue everywhere, erff GELU, in-kernel W
    split, no fp16 layer) meet the same bar as the defaults.  Subprocess: the switches are read once per process."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from helpers import O, case, assert_close\n"
        "from easydgl_b200.engine import Engine\n"
        "for name in ('easy_b', 'easy_d', 'ctsma_b', 'C2'):\n"
        "    cfg, inp, W = case(name, batch=8)\n"
        "    eng = Engine(cfg, W, max_batch=8, device='cuda:0')\n"
        "    lg = eng.forward_logits(inp['seqs_i'].cuda(), inp['seqs_t'].cuda()).cpu()\n"
        "    ref = O.forward(inp['seqs_i'], inp['seqs_t'], W, cfg, dtype=torch.float64)\n"
        "    assert_close(lg[:, 1:], ref[:, 1:], 1e-3, name)\n"
        "print('VARIANT_OK')\n" % (os.path.dirname(os.path.abspath(__file__)),
                                   os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True,
                         timeout=280)
    assert "VARIANT_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_f16_dense_layers_in_pipeline():
    """The scaled 3xFP16 dense layers (gemm_f16.cu) inside the C2 pipeline, fed by the producers' published activation
    maxima: none (mask 0), the default (QKVT), and all six - each within 1e-4 of the fp64 oracle's logits (the
    measured errors are 1.4e-6 ... 2.1e-6, the same as the 3xTF32 layers)."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_f16_gemm_mask.py")
    res = subprocess.run([sys.executable, script, "0", "1", "63"], capture_output=True, text=True, timeout=280,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lines = [l for l in res.stdout.splitlines() if l.startswith("mask")]
    assert len(lines) == 3, res.stdout[-2000:] + res.stderr[-2000:]
    for l in lines:
        assert "finite True" in l, l
        assert float(l.split("err")[1].split()[0]) < 1e-4, l


def test_long_sequence_key_streaming_path():
    """L = 300 (> every register/TMEM-resident kernel's limit) and dh = 32: the key-streaming kernel."""
    _check_model("easy_b", batch=3, seqslen=299)


def test_forward_c5_shape_small():
    """BASELINE.json configs[4] shape (EasyDGL d=256 L=512 h=8) with a small catalogue and batch."""
    _check_model("C5", batch=2, num_items=3000)
