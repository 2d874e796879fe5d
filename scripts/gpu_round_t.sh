#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu --timeout 150 > gpurun_out/t_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/t_gpu_tests.log
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1", "%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in r["phase_ms_per_iteration"].items() if v}, r["kernel"], "%.3f"%r["frac"], "iter %.3f"%r["iteration"]["frac"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/$1.err").read()[-400:])
PY
}
timeout 90 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/t_1M.json 2> gpurun_out/t_1M.err; show t_1M
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 --workload synthetic-16M-mixed-omesh > gpurun_out/t_16M.json 2> gpurun_out/t_16M.err; show t_16M
