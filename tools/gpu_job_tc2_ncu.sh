#!/bin/bash
# ncu --set full of one launch of the tcgen05 attention kernel at the C2 shape (+ one-slot timing)
set -u
TAG=${1:-r2_tc2}
mkdir -p gpurun_out
EDGL_TC2_ONESLOT=1 timeout 120 tools/attn_selftest 4096 8 100 0 1 1 > gpurun_out/${TAG}_oneslot.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none --launch-skip 5 -c 1 -k regex:attention_tc2 -o gpurun_out/${TAG} -f tools/attn_selftest 4096 8 100 0 1 1 > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_oneslot.log
tail -3 gpurun_out/${TAG}_ncu.log
