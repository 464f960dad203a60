#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
the column-sharded NCCL path must return, for each rank's own sequences, exactly the top-K that a
single unsharded engine returns.  Prints SHARDED_OK <world> on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_nccl.py [C2|C4]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easydgl_b200 import synth  # noqa: E402
from easydgl_b200.engine import Engine  # noqa: E402
from easydgl_b200.sharded import ShardedRanker  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.named_config(name)
    W = synth.make_weights(cfg, mode="parity")
    inp = synth.make_inputs(cfg, B, seed=77 + rank, edge_cases=True)
    ids, ts = inp["seqs_i"].to(dev), inp["seqs_t"].to(dev)
    single = Engine(cfg, W, max_batch=B, device=dev)
    want_i, want_v = single.forward_topk(ids, ts, True)
    shard = Engine(cfg, W, max_batch=B, device=dev, shard_rank=rank, shard_world=world)
    got_i, got_v = ShardedRanker(shard).forward_topk(ids, ts, True)
    ok = torch.equal(got_i, want_i) and torch.equal(got_v, want_v)
    ag_i, ag_v = ShardedRanker(shard, exchange="all_gather").forward_topk(ids, ts, True)
    ok = ok and torch.equal(ag_i, want_i) and torch.equal(ag_v, want_v)
    pr = ShardedRanker(shard, exchange="p2p")
    for _ in range(3):  # several epochs through the same peer buffers
        pi, pv = pr.forward_topk(ids, ts, True)
        ok = ok and torch.equal(pi, want_i) and torch.equal(pv, want_v)
    pi2, _ = pr.forward_topk(ids, ts, False)
    ok = ok and torch.equal(pi2, single.forward_topk(ids, ts, False)[0])
    got_i2, _ = ShardedRanker(shard).forward_topk(ids, ts, False)
    want_i2, _ = single.forward_topk(ids, ts, False)
    ok = ok and torch.equal(got_i2, want_i2)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK %d %s B=%d" % (world, name, B) if int(flag.item()) == 1 else "SHARDED_MISMATCH")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
