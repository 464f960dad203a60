#!/bin/bash
# does the all-to-all of the candidates get faster with more NCCL point-to-point channels?  (8 GPUs, phase profile only)
set -u
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516"
echo "default:"; timeout 200 $TR tools/profile_sharded.py 2>&1 | grep PHASES
echo "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32:"
NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 timeout 200 $TR tools/profile_sharded.py 2>&1 | grep PHASES
