#!/bin/bash
# one B200, ~40 s: k_dt_grad tuning variants (scripts/build_variants.py), first-stage limiter inside the kernel
mkdir -p gpurun_out
: > gpurun_out/ab2.jsonl
for v in nopreload nopreload_pf nopreload_early nopreload_minb4 nopreload_t128 nopreload_t128_pf preload_pf preload_minb3 preload_minb3_pf; do
  AFX_LIB=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_$v.so QUICK_AB_FUSE=1 timeout 20 python scripts/quick_ab.py >> gpurun_out/ab2.jsonl 2>> gpurun_out/ab2.err || echo "$v failed"
done
python - <<PY
import json
for l in open("gpurun_out/ab2.jsonl"):
    d = json.loads(l); p = d["phase_ms"]
    print("%-44s ms/iter %.4f  dt_grad %.4f lim %.4f flux %.4f gather %.4f" % (d["lib"], d["ms_per_iteration"], p["dt_grad"], p["limiter"], p["flux"], p["gather_update"]))
PY
