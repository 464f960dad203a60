#!/bin/bash
# scaling points: bash tools/gpu_job_scale.sh N "C2 C4 C5"
set -u
N=${1:-8}
WLS=${2:-"C2 C4"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
if [ "$N" = "1" ]; then TR="python"; fi
for W in $WLS; do
  ST=20; if [ "$W" = "C5" ]; then ST=3; fi
  timeout 600 $TR bench.py --gpus $N --workload $W --steps $ST --warmup 3 --no-cpu > gpurun_out/r2_scale_${W}_${N}gpu.json 2> gpurun_out/r2_scale_${W}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_scale_${W}_${N}gpu.json").read())
    print("${W} N=${N} B/GPU=%s ms/step %.3f value %.0f e2e %.3f"%(d["config"]["global_batch"]//${N}, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("${W} N=${N} failed", e)
PY
  tail -2 gpurun_out/r2_scale_${W}_${N}gpu.err | cut -c1-300
done
