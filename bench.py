#!/usr/bin/env python
"""bench.py - sequences/sec of the EasyDGL eval forward (+ seen-mask + top-100).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B=4096 synthetic Netflix-schema
sequences (BASELINE.json configs[1], "C2").  At N>1 every rank gets its own B sequences (weak
scaling) and the item table / logits are column-sharded (easydgl_b200/sharded.py).

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, max over
ranks; `e2e` = the same work through edgl_forward_topk_host with pinned HOST buffers (H2D + D2H
inside the timed region); `roofline` = the dominant kernel's algorithmic TFLOP/s from CUDA events
recorded on the launch stream inside the timed region; `cpu_baseline` = the fp32 CPU oracle (the
reference cannot run here: TensorFlow 2.3.4 is not installable) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from easydgl_b200 import synth  # noqa: E402

WORKLOAD = "C2"       # default headline workload; --workload C3|C4|C5 selects BASELINE.json configs[2..4]
NUM_INPUT_SETS = 4  # rotating distinct input batches
# Per-GPU batch of each workload.  C2 / C3 name a single-GPU batch (4096).  C4 / C5 name a GLOBAL batch "across
# 2/4/8 (8) B200": per GPU it is B/8, so the 8-GPU run is exactly the named configuration and smaller N are the
# weak-scaling points of the same per-GPU work (--batch overrides; --scaling strong fixes the global batch instead).
PER_GPU_BATCH = {"C1": 32, "C2": 4096, "C3": 4096, "C4": 8192 // 8, "C5": 16384 // 8}


def workload_desc(cfg, B, n_gpus, workload=None):
    return {
        "workload": "%s: %s d=%d L=%d items=%d B=%d/GPU h=%d blocks=%d E=%d; eval forward + mask_seen + top-%d"
                    % (workload or WORKLOAD, cfg.model, cfg.num_units, cfg.L, cfg.num_items, B, cfg.num_heads,
                       cfg.num_blocks, cfg.num_events, cfg.topk),
        "global_batch": B * n_gpus,
        "parallelism": "single GPU" if n_gpus == 1 else
        "batch-sharded encoder + column-sharded item table (%d shards); NCCL all-gather of [y|ids] rows, then "
        "exchange of per-shard top-K candidates + local merge" % n_gpus,
        "l2": "per-step working set ~3 GB of activations >> 126 MB L2; %d rotating input batches" % NUM_INPUT_SETS,
        "weights": "reference initialisers (glorot / N(0,0.02)), seed 9876",
    }


# ----------------------------------------------------------------------------- FLOP model
def stage_flops(cfg, B):
    """Algorithmic FLOPs (2*MAC) each stage performs PER STEP on B sequences (summed over the stage's launches:
    one per block, two for the CTSMA Q / KVT projections)."""
    L, d, h, E, N1, nb = cfg.L, cfg.num_units, cfg.num_heads, cfg.num_events, cfg.num_rows, cfg.num_blocks
    dh = d // h
    att = B * (6.0 * L * L * d + 2.0 * L * d * E * (dh + 2) + 2.0 * h * L * L * E)
    if cfg.model == "CTSMA":  # Q,K,V,T from [2d | d] inputs; FeedForward = two d x d conv1x1
        return {
            "qkvt_gemm": 2.0 * B * L * (2 * d) * 4 * d + (nb - 1) * 2.0 * B * L * d * 4 * d,
            "attention": nb * att,
            "ff1_gemm": nb * 2.0 * B * L * d * d,
            "ff2_gemm": nb * 2.0 * B * L * d * d,
            "logits_gemm": 2.0 * B * d * N1,
        }
    return {
        # block 0 runs the folded [d+E, 4d] kernel; later blocks a [d, 4d] one
        "qkvt_gemm": 2.0 * B * L * (d + E) * 4 * d + (nb - 1) * 2.0 * B * L * d * 4 * d,
        "attention": nb * att,
        "ao_gemm": nb * 2.0 * B * L * d * d,
        "ff1_gemm": nb * 4.0 * B * L * d * d,
        "ff2_gemm": nb * 4.0 * B * L * d * d,
        "tr_gemm": 2.0 * B * L * d * d,
        "logits_gemm": 2.0 * B * d * N1,
    }


def stage_bytes(cfg, B):
    """Compulsory HBM bytes per launch for the memory-bound stages."""
    L, d, E, N1, K = cfg.L, cfg.num_units, cfg.num_events, cfg.num_rows, cfg.topk
    rows = B * L
    return {
        "embed": rows * (8 + 4 + 4 * d) + rows * 4 * (d + E) + rows * (4 + E + 1),
        "ln_att": 2.0 * rows * d * 4, "ln_ff": 2.0 * rows * d * 4, "ln_out": rows * d * 4 + B * d * 4,
        "topk": B * N1 * 4 + B * K * 8,
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (pynvml, 20 ms period)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- CPU reference arm
CPU_CHUNK = 128      # sequences per oracle call (the literal 4-D gate needs 1.3 MB per sequence and head at C2)
CPU_SAMPLE = 256     # sequences per timed step of the reference arm / cpu_baseline leg


def cpu_reference(cfg, W, n_seqs, steps, warmup, threads, literal=False, chunk=CPU_CHUNK):
    """Times the fp32 CPU oracle (torch-CPU/MKL; contracted gate unless literal) on `n_seqs` sequences per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import easydgl_oracle as O
    torch.set_num_threads(threads)
    inp = synth.make_inputs(cfg, n_seqs, seed=synth.SEED + 100)
    ids, ts = inp["seqs_i"], inp["seqs_t"]

    def one_step():
        for s in range(0, n_seqs, chunk):
            logits = O.forward(ids[s:s + chunk], ts[s:s + chunk], W, cfg, dtype=torch.float32, literal=literal)
            O.eval_topk(logits, ids[s:s + chunk], True, cfg.topk, rank_on="probs")

    with torch.no_grad():
        for _ in range(warmup):
            one_step()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        dt = time.perf_counter() - t0
    return n_seqs * steps / dt, dt / steps


def cpu_c1_legs(threads):
    """the reference's own CPU-runnable case C1 (d=64, L=100, B=32) in literal mode (the two
    [hB,L,L,E] tensors of temporal.py:309-313 materialised) and contracted mode, with 1 thread (what the reference
    configures itself, main.py:167-168) and with all host threads.  A few seconds in total."""
    cfg = synth.named_config("C1")
    W = synth.make_weights(cfg, mode="reference")
    out = {}
    for name, lit, thr in (("c1_literal_1thread", True, 1), ("c1_contracted_1thread", False, 1),
                           ("c1_literal_allthreads", True, threads), ("c1_contracted_allthreads", False, threads)):
        sps, _ = cpu_reference(cfg, W, 32, 3, 1, thr, literal=lit, chunk=32)
        out[name] = round(sps, 1)
    torch.set_num_threads(threads)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    cfg = synth.named_config(wl)
    W = synth.make_weights(cfg, mode="reference")
    threads = os.cpu_count() or 1
    n_seqs = CPU_SAMPLE if cfg.L * cfg.num_units <= 200 * 128 else 32
    sps, step_s = cpu_reference(cfg, W, n_seqs, args.steps, max(args.warmup, 1), threads, chunk=min(CPU_CHUNK, n_seqs))
    sample = "%d sequences/step of the %s workload (%d chunk(s) of %d), fp32 torch-CPU oracle, contracted gate, %d threads" % (
        n_seqs, wl, max(1, n_seqs // CPU_CHUNK), min(CPU_CHUNK, n_seqs), threads)
    B = args.batch or PER_GPU_BATCH[wl]
    conf = workload_desc(cfg, B, args.gpus, wl)
    conf["reference_sample"] = ("this arm times %d sequences per step (seq/s is per sequence, so it is comparable to "
                                "the GPU arm's B=%d/GPU steps)" % (n_seqs, B))
    line = {
        "impl": "reference", "metric": "sequences/sec", "value": sps, "unit": "seq/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": conf,
        "cpu_baseline": {"value": sps, "unit": "seq/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference needs tensorflow-gpu==2.3.4 (not installable here); this arm times oracle/, the "
                "op-for-op CPU restatement of its forward path, with all host threads",
    }
    if not args.no_c1:
        line["cpu_baseline"]["c1_legs_seq_per_s"] = cpu_c1_legs(threads)
    _emit(args.out_fd, line)
    return 0


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    from easydgl_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = args.workload
    cfg = synth.named_config(wl)
    B = args.batch or PER_GPU_BATCH[wl]
    if args.scaling == "strong" and not args.batch:
        B = max(1, synth.CONFIGS[wl]["batch"] // world)
    W = synth.make_weights(cfg, mode="reference")
    eng = E.Engine(cfg, W, max_batch=B, device=dev, shard_rank=rank if world > 1 else 0,
                   shard_world=world if world > 1 else 1)
    ranker = None
    if world > 1:
        from easydgl_b200.sharded import ShardedRanker
        ranker = ShardedRanker(eng, exchange=args.exchange)

    sets = []
    for i in range(NUM_INPUT_SETS):
        inp = synth.make_inputs(cfg, B, seed=synth.SEED + 1000 * rank + i)
        sets.append((inp["seqs_i"].pin_memory(), inp["seqs_t"].pin_memory()))
    dsets = [(a.to(dev), b.to(dev)) for a, b in sets]
    idx = torch.empty((B, cfg.topk), dtype=torch.int32, device=dev)
    val = torch.empty((B, cfg.topk), dtype=torch.float32, device=dev)

    def step(i):
        a, b = dsets[i % NUM_INPUT_SETS]
        if ranker is None:
            eng.forward_topk(a, b, True, out=(idx, val))
        else:
            ranker.forward_topk(a, b, True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    eng.profile(True)
    launches0 = E.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = E.launch_count() - launches0
    prof = eng.profile_read()
    eng.profile(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e: pinned host inputs -> H2D -> forward -> top-K -> D2H, every step
    # Two result buffers: the caller keeps two batches in flight (edgl_forward_topk_host_submit / _wait), so the upload
    # of batch i+1 and the download of batch i-1 overlap the kernels of batch i.  Every step still uploads its own
    # inputs from pinned host memory and downloads its own result inside the timed region.
    h_idx = [torch.empty((B, cfg.topk), dtype=torch.int32).pin_memory() for _ in range(2)]
    h_val = [torch.empty((B, cfg.topk), dtype=torch.float32).pin_memory() for _ in range(2)]

    def run_e2e(n_steps):
        if ranker is None:
            prev = None
            for i in range(n_steps):
                a, b = sets[i % NUM_INPUT_SETS]
                slot = eng.forward_topk_host_submit(a, b, h_idx[i % 2], h_val[i % 2], True)
                if prev is not None:
                    eng.forward_topk_host_wait(prev)
                prev = slot
            if prev is not None:
                eng.forward_topk_host_wait(prev)
        else:
            for i in range(n_steps):
                a, b = sets[i % NUM_INPUT_SETS]
                da, db = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
                ri, rv = ranker.forward_topk(da, db, True)
                h_idx[0].copy_(ri, non_blocking=True)
                h_val[0].copy_(rv, non_blocking=True)
                torch.cuda.current_stream().synchronize()

    run_e2e(min(args.warmup, 3))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)  # host-synchronous calls: wall clock bounds the device span
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = B * cfg.L * 8 + B * cfg.ts_len * 4
    d2h = B * cfg.topk * 8

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
        fl, by = stage_flops(cfg, B), stage_bytes(cfg, B)
        stages = {}
        total_stage_ms = sum(v[0] for v in prof.values()) or 1.0
        for name, (sms, cnt) in prof.items():
            per_step = sms / args.steps  # all launches of the stage in one step (one per block; logits chunks)
            rec = {"ms": round(per_step, 4), "launches_per_step": cnt / args.steps,
                   "share": round(sms / total_stage_ms, 4)}
            # sharded: per rank G*B rows x N1/G columns = the same FLOPs/bytes as unsharded
            if name in fl:
                rec["tflops"] = round(fl[name] / (per_step * 1e-3) / 1e12, 3)
            if name in by:
                rec["gbs"] = round(by[name] / (per_step * 1e-3) / 1e9, 1)
            stages[name] = rec
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
        roof = None
        traffic = None
        try:  # DRAM bytes per launch of the same kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh)["stages"].get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        if dom in fl:
            ach = fl[dom] / ((prof[dom][0] / args.steps) * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": round(ach, 3), "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": round(ach / tf_peak, 5), "traffic": traffic, "peak_source": peak_src,
                    "note": "algorithmic fp32 FLOPs of the kernel / CUDA-event time; peak is the measured dense bf16 "
                            "tcgen05 rate. The attention core runs a scaled 3xFP16 split on mma.sync (3 tensor-core "
                            "products per algorithmic product: tensor ceiling = peak/3 even on tcgen05) because the "
                            "top-K parity bar needs fp32-level accuracy, and it is bound by issue slots and the MUFU "
                            "(256 sigmoids per row and head), not by the tensor pipe: see stages[*] / "
                            "profiles/traffic.json for issue, XU and tensor pipe utilisation (DESIGN.md 4); "
                            "traffic = ncu dram bytes/launch (profiles/traffic.json)"}
        elif dom in by:
            hb = peaks.get("hbm_gbs", 6650.0)
            ach = by[dom] / ((prof[dom][0] / args.steps) * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": hb, "unit": "GB/s",
                    "frac": round(ach / hb, 4), "traffic": traffic, "peak_source": peak_src}
        line = {
            "metric": "sequences/sec", "value": value, "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_desc(cfg, B, world, wl),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "seq/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof, "stages": stages,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            n_seqs = CPU_SAMPLE if cfg.L * cfg.num_units <= 200 * 128 else 32
            sps, step_s = cpu_reference(cfg, W, n_seqs, args.cpu_steps, 1, threads, chunk=min(CPU_CHUNK, n_seqs))
            line["cpu_baseline"] = {
                "value": sps, "unit": "seq/s", "cores": threads, "kind": "port",
                "sample": "%d steps x %d sequences of the %s workload (%.1f s), fp32 torch-CPU oracle, contracted gate"
                          % (args.cpu_steps, n_seqs, wl, step_s * args.cpu_steps)}
            if not args.no_c1:
                line["cpu_baseline"]["c1_legs_seq_per_s"] = cpu_c1_legs(threads)
        _emit(args.out_fd, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    # Libraries (NCCL's version banner, torchrun) may write to fd 1: park the real stdout and route fd 1 to
    # stderr for the whole run, so that stdout carries exactly ONE line - the JSON result.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rc = _main(real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    return rc


def _emit(fd, line):
    os.write(fd, (json.dumps(line) + "\n").encode())


def _main(out_fd):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE.json configs[1..4]; C2 is the headline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed per-GPU batch (PER_GPU_BATCH); strong: the workload's batch divided over the ranks")
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 literal / 1-thread CPU legs")
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--exchange", default="all_to_all", choices=["all_to_all", "all_gather", "p2p"],
                    help="candidate exchange of the column-sharded multi-GPU path")
    args = ap.parse_args()
    args.out_fd = out_fd
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
