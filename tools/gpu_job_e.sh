#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_facade.py -m gpu -x -q -k "topk or host_entry or sharded" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/r2e_bench_C2.json 2> gpurun_out/r2e_bench_C2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e_bench_C2.json").read())
print("ms/step %.3f"%d["ms_per_step"], "e2e %.3f"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
PY
tail -3 gpurun_out/r2e_bench_C2.err
