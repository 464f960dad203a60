#!/bin/bash
# attention kernel self-test (attn_tc2.cu) on small shapes, then a timing at the C2 / C3 shapes
set -u
mkdir -p gpurun_out
out=gpurun_out/r2b_selftest.log
: > $out
run() { echo "== $*" >> $out; timeout 90 tools/attn_selftest "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run 3 2 100 0 1
run 3 2 100 1 0
run 2 4 31 0 1
run 2 2 64 1 0
run 5 8 100 0 1
run 4 1 13 0 0
run 4096 8 100 0 1 1
run 4096 4 100 1 0 1
cat $out
