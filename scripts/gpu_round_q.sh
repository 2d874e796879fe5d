#!/bin/bash
mkdir -p gpurun_out
run() { # name lib
  AFX_LIB=$PWD/aeroflex_b200/lib/$2 timeout 90 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/q_$1.json 2> gpurun_out/q_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_$1.json").read().strip().splitlines()[-1])
    print("$1", "%.3e"%d["value"], "ms/it %.4f"%d["ms_per_step"], {k:round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items() if v})
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/q_$1.err").read()[-400:])
PY
}
run base libaeroflex_rans_b200.so
run dtg256x2 libafx_dtg256x2.so
run dtg128x5 libafx_dtg128x5.so
