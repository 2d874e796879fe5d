#!/bin/bash
# calibrate the polar sweep on one GPU (8 angles), then ncu launch list of the default bench
mkdir -p gpurun_out
timeout 600 python scripts/polar_sweep_bench.py 1 -6.5 > gpurun_out/l_polar_1gpu_8alpha.json 2> gpurun_out/l_polar.err
python - <<PY
import json
d=json.loads(open("gpurun_out/l_polar_1gpu_8alpha.json").read().strip().splitlines()[-1])
print("polar", d["n_gpus"], "gpus", d["alphas"], "alphas", "%.1f s"%d["wall_s"], d["rc"], d["stderr_tail"]); print(d["polar_alpha_cl_cd_cm"][:3])
PY
