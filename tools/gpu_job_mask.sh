#!/bin/bash
# which dense layers gain from the scaled 3xFP16 tcgen05 GEMM at the large shapes (EDGL_F16_MASK bit i = layer i)
set -u
mkdir -p gpurun_out
for W in C5 C4 C2; do
  for M in 1 63; do
    ST=10; if [ "$W" = "C5" ]; then ST=3; fi
    EDGL_F16_MASK=$M timeout 600 python bench.py --workload $W --steps $ST --warmup 3 --no-cpu > gpurun_out/r2l_mask${M}_${W}.json 2>/dev/null
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_mask${M}_${W}.json").read())
    print("${W} mask=${M} ms/step %.3f"%d["ms_per_step"], {k:v["ms"] for k,v in d["stages"].items() if "gemm" in k})
except Exception as e: print("${W} ${M} failed", e)
PY
  done
done
