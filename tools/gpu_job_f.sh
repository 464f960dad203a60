#!/bin/bash
# multi-GPU job: parity of the sharded path + bench at N GPUs (N = $1)
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_sharded_nccl.py C2 512 2>&1 | grep -E "SHARDED|Error|error" | head -5
timeout 300 $TR tools/check_sharded_nccl.py C4 512 2>&1 | grep -E "SHARDED|Error|error" | head -5
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_bench_C2_${N}gpu.json 2> gpurun_out/r2f_bench_C2_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench_C2_${N}gpu.json").read())
print("N=${N} ms/step %.3f"%d["ms_per_step"], "value %.0f"%d["value"], "e2e %.3f"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("   ", {k:v["ms"] for k,v in d["stages"].items()})
PY
tail -2 gpurun_out/r2f_bench_C2_${N}gpu.err
