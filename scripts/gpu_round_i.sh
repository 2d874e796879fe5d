#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/i_fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -3 gpurun_out/i_fused_tests.log
run() { # name lib ctas tile
  AFX_LIB=$PWD/aeroflex_b200/lib/$2 AFX_STAGE_CTAS=$3 AFX_TILE=$4 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/i_$1.json 2> gpurun_out/i_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/i_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], d["config"]["stage_kernel"], "stage %.4f"%d["roofline"]["phase_ms_per_iteration"]["stage"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/i_$1.err").read()[-400:])
PY
}
run s512x1_T384 libaeroflex_rans_b200.so 1 384
run s512x1_T320 libaeroflex_rans_b200.so 1 320
run s256x2_T192 libafx_s256x2.so 2 192
