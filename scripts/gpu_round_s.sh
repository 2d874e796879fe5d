#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu --timeout 150 > gpurun_out/s_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/s_gpu_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/s_bench_default.json 2> gpurun_out/s_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/s_bench_default.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], r["kernel"], "%.3f"%r["frac"], "traffic", r["traffic"], "iter %.3f"%r["iteration"]["frac"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
PY
