#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_fused.py -x -q -m gpu --timeout 60 > gpurun_out/n_fused_tests.log 2>&1
rc=$?; echo "fused tests rc=$rc"; tail -3 gpurun_out/n_fused_tests.log
[ $rc -ne 0 ] && exit 0
run() { # name ctas tile
  AFX_STAGE_CTAS=$2 AFX_TILE=$3 timeout 90 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 --fused 1 > gpurun_out/n_$1.json 2> gpurun_out/n_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/n_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], d["config"]["stage_kernel"], "stage %.4f"%d["roofline"]["phase_ms_per_iteration"]["stage"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/n_$1.err").read()[-400:])
PY
}
run x2_T192 2 192
run x2_T160 2 160
