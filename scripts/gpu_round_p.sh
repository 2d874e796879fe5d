#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/p_gpu_tests.log 2>&1
rc=$?; echo "gpu tests rc=$rc"; tail -3 gpurun_out/p_gpu_tests.log
timeout 90 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p_1M.json 2> gpurun_out/p_1M.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/p_1M.json").read().strip().splitlines()[-1])
    print("1M", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items() if v})
except Exception as e:
    print("failed", e); print(open("gpurun_out/p_1M.err").read()[-400:])
PY
