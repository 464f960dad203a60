#!/bin/bash
# GPU test suite + single-GPU bench lines:  bash tools/gpu_job_test_bench.sh TAG "C2 C3 C4" [pytest-args]
set -u
TAG=${1:-r2x}
WLS=${2:-"C2"}
PT=${3:-"-x -q"}
mkdir -p gpurun_out
if [ "$PT" != "skip" ]; then
  timeout 1200 python -m pytest tests -m gpu $PT > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -6 gpurun_out/${TAG}_pytest.log
fi
for W in $WLS; do
  ST=20; if [ "$W" = "C5" ]; then ST=3; fi
  timeout 600 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_${W}.json 2> gpurun_out/${TAG}_bench_${W}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${W}.json").read())
    print("${W} ms/step %.3f"%d["ms_per_step"], "value %.0f"%d["value"], "e2e %.3f"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
    print("   ", {k:v["ms"] for k,v in d["stages"].items()})
except Exception as e: print("${W} failed", e)
PY
  tail -2 gpurun_out/${TAG}_bench_${W}.err
done
