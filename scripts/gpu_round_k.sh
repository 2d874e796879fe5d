#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/k_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/k_gpu_tests.log
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d["config"]["stage_kernel"][:60], {k:round(v,4) for k,v in r["phase_ms_per_iteration"].items() if v}, "dom",r["kernel"],"frac %.3f"%r["frac"], "iter frac %.3f"%r["iteration"]["frac"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/$1.err").read()[-400:])
PY
}
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/k_1M_unfused.json 2> gpurun_out/k_1M_unfused.err; show k_1M_unfused
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --fused 1 > gpurun_out/k_1M_fused.json 2> gpurun_out/k_1M_fused.err; show k_1M_fused
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 --workload synthetic-16M-mixed-omesh > gpurun_out/k_16M_unfused.json 2> gpurun_out/k_16M_unfused.err; show k_16M_unfused
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 --workload synthetic-16M-mixed-omesh --fused 1 > gpurun_out/k_16M_fused.json 2> gpurun_out/k_16M_fused.err; show k_16M_fused
