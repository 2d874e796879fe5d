#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/r_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/r_gpu_tests.log
run() { # name
  timeout 90 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 $2 $3 > gpurun_out/r_$1.json 2> gpurun_out/r_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items() if v})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r_$1.err").read()[-400:])
PY
}
AFX_PDL=1 run pdl_on
AFX_PDL=0 run pdl_off
AFX_PDL=1 run pdl_on_b
