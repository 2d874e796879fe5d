#!/bin/bash
mkdir -p gpurun_out
python - <<PY
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size)
PY
run() { # name
  timeout 90 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/o_$1.json 2> gpurun_out/o_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/o_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items() if v})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/o_$1.err").read()[-400:])
PY
}
AFX_L2_PERSIST=1 run l2on
AFX_L2_PERSIST=0 run l2off
