#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -6 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench_C2.json 2> gpurun_out/r2d_bench_C2.err
EDGL_LN_FUSE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench_C2_nofuse.json 2> gpurun_out/r2d_bench_C2_nofuse.err
python - <<'PY'
import json
for f in ("r2d_bench_C2","r2d_bench_C2_nofuse"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read())
        print(f, "ms/step %.3f"%d["ms_per_step"], "e2e %.3f"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
        print("   ", {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e: print(f, "failed", e)
PY
tail -3 gpurun_out/r2d_bench_C2.err
