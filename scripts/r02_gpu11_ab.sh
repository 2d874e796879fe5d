#!/bin/bash
# round 2, call 11 (one B200, ~3 min): A/B of the register / occupancy variants of k_limiter<1> (stored extremes) and k_dt_grad<.,2>
# (scripts/build_variants.py) at 16M and 1M cells, one process per library; plus the new parity test of the stored extremes
mkdir -p gpurun_out
: > gpurun_out/r02k_ab_16M.jsonl; : > gpurun_out/r02k_ab_1M.jsonl; : > gpurun_out/r02k_ab.err
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "stored_limiter or first_stage_limiter" 2>&1 | tail -2
for v in default lpm9 lpm10 dtg5 dtg4 defer defer5 defer4 lpm9_defer5 default; do
  lib=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_$v.so
  [ $v = default ] && lib=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200.so
  AFX_LIB=$lib PIPE_AB_MESH=16M PIPE_AB_CONFIGS=off timeout 60 python scripts/pipe_ab.py 2>> gpurun_out/r02k_ab.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02k_ab_16M.jsonl
  AFX_LIB=$lib PIPE_AB_MESH=1M PIPE_AB_CONFIGS=off timeout 30 python scripts/pipe_ab.py 2>> gpurun_out/r02k_ab.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02k_ab_1M.jsonl
done
python - <<PY
import json
for n in ("16M", "1M"):
    for l in open("gpurun_out/r02k_ab_%s.jsonl" % n):
        d = json.loads(l)
        print(n, "%-12s" % d["variant"], "%.4f ms" % d["ms_per_iteration"], {k: round(v, 4) for k, v in d["phase_ms"].items() if v}, d["norm_last"])
PY
tail -3 gpurun_out/r02k_ab.err
