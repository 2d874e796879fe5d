#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/f_fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -3 gpurun_out/f_fused_tests.log
for T in 256 320 384 448; do
  AFX_TILE=$T timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/f_bench_T$T.json 2> gpurun_out/f_bench_T$T.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/f_bench_T$T.json").read().strip().splitlines()[-1])
    print("T=$T", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], d["config"]["stage_kernel"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items()})
except Exception as e:
    print("T=$T failed", e); print(open("gpurun_out/f_bench_T$T.err").read()[-600:])
PY
done
AFX_FUSED=0 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 --fused 0 > gpurun_out/f_bench_unfused.json 2> gpurun_out/f_bench_unfused.err
python - <<PY
import json
d=json.loads(open("gpurun_out/f_bench_unfused.json").read().strip().splitlines()[-1])
print("unfused(hilbert)", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items()})
PY
