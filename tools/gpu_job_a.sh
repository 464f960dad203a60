#!/bin/bash
# round-2 job A: GPU test suite + baseline bench lines for C2..C5 (before the round-2 kernels)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_C2.json 2> gpurun_out/r2a_bench_C2.err
timeout 300 python bench.py --workload C3 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a_bench_C3.json 2> gpurun_out/r2a_bench_C3.err
timeout 300 python bench.py --workload C4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_C4.json 2> gpurun_out/r2a_bench_C4.err
timeout 400 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu --batch 512 > gpurun_out/r2a_bench_C5.json 2> gpurun_out/r2a_bench_C5.err
for f in C2 C3 C4 C5; do echo "== $f"; cut -c1-400 gpurun_out/r2a_bench_$f.json; tail -2 gpurun_out/r2a_bench_$f.err; done
