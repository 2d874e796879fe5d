#!/bin/bash
mkdir -p gpurun_out
run() { # name lib ctas tile
  AFX_LIB=$PWD/aeroflex_b200/lib/$2 AFX_STAGE_CTAS=$3 AFX_TILE=$4 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/j_$1.json 2> gpurun_out/j_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/j_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], d["config"]["stage_kernel"], "stage %.4f"%d["roofline"]["phase_ms_per_iteration"]["stage"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/j_$1.err").read()[-400:])
PY
}
run s768x1_T384 libafx_s768x1.so 1 384
run s768x1_T352 libafx_s768x1.so 1 352
run s768x1_T416 libafx_s768x1.so 1 416
run s1024x1_T384 libafx_s1024x1.so 1 384
run s384x2_T192 libafx_s384x2.so 2 192
run s384x2_T176 libafx_s384x2.so 2 176
