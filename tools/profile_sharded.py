#!/usr/bin/env python
"""Per-phase CUDA-event timing of the column-sharded step (run under torchrun); rank 0 prints a table."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easydgl_b200 import synth  # noqa: E402
from easydgl_b200.engine import Engine  # noqa: E402
from easydgl_b200.sharded import ShardedRanker, _merge_packed_cuda  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.named_config("C2")
    B = 4096
    W = synth.make_weights(cfg)
    eng = Engine(cfg, W, max_batch=B, device=dev, shard_rank=rank, shard_world=world)
    inp = synth.make_inputs(cfg, B, seed=5 + rank)
    ids, ts = inp["seqs_i"].to(dev), inp["seqs_t"].to(dev)
    G, d, L, K = world, eng.d, eng.L, eng.K
    Wd = d + 2 * L
    mine = torch.empty((B, Wd), dtype=torch.float32, device=dev)
    allp = torch.empty((G * B, Wd), dtype=torch.float32, device=dev)
    cand = torch.empty((G * B, 2, K), dtype=torch.int32, device=dev)
    recv = torch.empty((G, B, 2, K), dtype=torch.int32, device=dev)
    names = ["encode", "pack", "all_gather", "logits_topk", "all_to_all", "merge"]
    acc = {n: 0.0 for n in names}
    iters = 20
    for it in range(iters + 3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier()
        ev[0].record()
        eng.encode_packed(ids, ts, mine)
        ev[1].record()
        ev[2].record()
        dist.all_gather_into_tensor(allp, mine)
        ev[3].record()
        eng.logits_topk(allp[:, :d], allp.view(torch.int64)[:, d // 2:], out=(cand[:, 0], cand[:, 1].view(torch.float32)),
                        out_stride=2 * K)
        ev[4].record()
        dist.all_to_all_single(recv.view(G * B, 2, K), cand)
        ev[5].record()
        _merge_packed_cuda(recv, 0, B)
        ev[6].record()
        torch.cuda.synchronize()
        if it >= 3:
            for i, n in enumerate(names):
                acc[n] += ev[i].elapsed_time(ev[i + 1]) / iters
    if rank == 0:
        print("PHASES world=%d " % world + " ".join("%s=%.3f" % (n, acc[n]) for n in names) +
              " total=%.3f" % sum(acc.values()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
