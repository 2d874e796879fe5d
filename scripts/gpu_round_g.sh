#!/bin/bash
mkdir -p gpurun_out
run() { # name lib ctas tile
  AFX_LIB=$PWD/aeroflex_b200/lib/$2 AFX_STAGE_CTAS=$3 AFX_TILE=$4 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/g_$1.json 2> gpurun_out/g_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/g_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], d["config"]["stage_kernel"], "stage %.4f"%d["roofline"]["phase_ms_per_iteration"]["stage"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/g_$1.err").read()[-400:])
PY
}
run s256x2_T160 libafx_s256x2.so 2 160
run s256x2_T192 libafx_s256x2.so 2 192
run s256x2_T224 libafx_s256x2.so 2 224
run s192x3_T112 libafx_s192x3.so 3 112
run s192x3_T128 libafx_s192x3.so 3 128
run s384x1_T288 libafx_s384x1.so 1 288
run s384x1_T384 libafx_s384x1.so 1 384
run s512x1_T400 libaeroflex_rans_b200.so 1 400
