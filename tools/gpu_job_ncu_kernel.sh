#!/bin/bash
# ncu --set full of ONE launch of a kernel inside a bench run:  bash tools/gpu_job_ncu_kernel.sh TAG WORKLOAD KERNEL_REGEX [skip]
set -u
TAG=$1; W=$2; K=$3; SKIP=${4:-3}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --launch-skip $SKIP -c 1 -k regex:$K -o gpurun_out/${TAG} -f \
  python bench.py --workload $W --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | cut -c1-300
