#!/bin/bash
# Round-2 final measurements on ONE GPU: full GPU test suite, bench lines of every workload (the C2 line with its CPU
# leg), the reference arm, the ncu launch list of the bench command and an `ncu --set full` capture of one step.
#   bash tools/gpu_job_final.sh TAG
set -u
TAG=${1:-r2_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err
for W in C3 C4 C5; do
  ST=20; if [ "$W" = "C5" ]; then ST=3; fi
  timeout 600 python bench.py --workload $W --steps $ST --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_${W}.json 2> gpurun_out/${TAG}_bench_${W}.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
python - <<PY
import json
for W in ("C2","C3","C4","C5"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%W).read().strip().splitlines()[-1])
        print(W, "ms/step %.3f"%d["ms_per_step"], "value %.0f"%d["value"], "e2e %.3f"%d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "roofline", d["roofline"]["frac"], "cpu", (d.get("cpu_baseline") or {}).get("value"))
        print("   ", {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e: print(W, "failed", e)
try:
    print("reference arm", open("gpurun_out/${TAG}_reference_arm.json").read()[:400])
except Exception as e: print(e)
PY
# ncu: launch list of the bench command (per-launch durations), then one full step with --set full
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"attention|gemm_tc|gemm_f16|embed_kernel|layernorm|ln_finalize|topk_kernel|mask_seen" --launch-skip 39 -c 13 \
  -f -o gpurun_out/${TAG}_step python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ls -la gpurun_out/${TAG}_step.ncu-rep gpurun_out/${TAG}_launches.csv
