#!/bin/bash
# one B200, a few seconds: 128-thread CTAs for the stage kernels against the default build
mkdir -p gpurun_out
: > gpurun_out/ab3.jsonl
for v in st128 fl128; do
  AFX_LIB=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_$v.so QUICK_AB_FUSE=1 timeout 12 python scripts/quick_ab.py >> gpurun_out/ab3.jsonl 2>> gpurun_out/ab3.err || echo "$v failed"
done
QUICK_AB_FUSE=1 timeout 12 python scripts/quick_ab.py >> gpurun_out/ab3.jsonl 2>> gpurun_out/ab3.err || echo "default failed"
cat gpurun_out/ab3.jsonl | cut -c1-420
