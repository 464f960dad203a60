#!/bin/bash
# 8-GPU (or $1-GPU) bench: micro-batch pipelining on / off
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
for M in 1 2; do
  EDGL_MICRO=$M timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_C2_${N}gpu_m$M.json 2> gpurun_out/r2g_bench_C2_${N}gpu_m$M.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_C2_${N}gpu_m$M.json").read())
    print("N=${N} micro=$M ms/step %.3f"%d["ms_per_step"], "value %.0f"%d["value"], "e2e %.3f"%d["e2e"]["ms_per_step"])
    print("   ", {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e: print("failed", e)
PY
done
EDGL_MICRO=1 EDGL_TOPK_WARP=0 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_C2_${N}gpu_m1_nowarp.json 2>/dev/null
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_C2_${N}gpu_m1_nowarp.json").read())
    print("N=${N} micro=1 CTA-topk ms/step %.3f"%d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if k in ("topk","logits_gemm","mask_seen")})
except Exception as e: print("failed", e)
PY
