#!/bin/bash
# N-GPU job: per-phase profile of the sharded step + bench lines:  bash tools/gpu_job_scale2.sh N "C2 C4" TAG
set -u
N=${1:-8}; WLS=${2:-"C2"}; TAG=${3:-r2s}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
timeout 300 $TR tools/profile_sharded.py 2>&1 | grep PHASES
timeout 300 $TR tools/check_sharded_nccl.py C2 512 2>&1 | grep -E "SHARDED|Error|error" | head -3
for W in $WLS; do
  ST=20; if [ "$W" = "C5" ]; then ST=3; fi
  timeout 600 $TR bench.py --gpus $N --workload $W --steps $ST --warmup 3 --no-cpu > gpurun_out/${TAG}_${W}_${N}gpu.json 2> gpurun_out/${TAG}_${W}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${W}_${N}gpu.json").read().strip().splitlines()[-1])
    print("${W} N=${N} ms/step %.3f value %.0f e2e %.3f"%(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), {k:v["ms"] for k,v in d["stages"].items() if k in ("logits_gemm","mask_seen","topk")})
except Exception as e: print("${W} failed", e)
PY
done
