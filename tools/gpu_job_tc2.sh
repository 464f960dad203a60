#!/bin/bash
# tcgen05 attention self-test: parity shapes, then timing + per-phase cycle counters at the C2 shape
set -u
mkdir -p gpurun_out
OUT=gpurun_out/${1:-r2_tc2_selftest}.log
: > $OUT
for cfg in "3 2 100 0 1" "2 4 31 0 1" "2 2 64 1 0" "5 8 100 0 1" "4 1 13 0 0" "3 2 100 1 0" "2 2 112 0 1"; do
  echo "== $cfg" >> $OUT
  timeout 120 tools/attn_selftest $cfg >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
echo "== bench" >> $OUT
timeout 120 tools/attn_selftest 4096 8 100 0 1 1 >> $OUT 2>&1
timeout 120 tools/attn_selftest 4096 4 100 1 0 1 >> $OUT 2>&1
grep -E "^==|vs double|SELFTEST|BENCH|PROF|rc=|error|failed" $OUT
