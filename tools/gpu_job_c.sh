#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -c 1 -f -o gpurun_out/r2_tc2_v3 tools/attn_selftest 4096 8 100 0 1 1 > gpurun_out/r2_tc2_v3_ncu.log 2>&1
tail -5 gpurun_out/r2_tc2_v3_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
