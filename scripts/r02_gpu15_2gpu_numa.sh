#!/bin/bash
# round 2, what is left of the budget (2 GPUs x ~40 s): bench.py under torchrun with the NUMA binding of the ranks, 1M cells cut in two
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload synthetic-1M-mixed-omesh --steps 20 --warmup 3 --e2e-steps 3 > gpurun_out/r02o_2gpu_numa.json 2> gpurun_out/r02o_2gpu_numa.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02o_2gpu_numa.json").read().strip().splitlines()[-1])
    print("%.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["e2e"]["ms_per_step"], d["details"].get("numa"), d.get("parity", {}).get("ok"), d.get("parity", {}).get("norm_max_rel_diff"))
except Exception as e:
    print("failed", e)
PY
tail -3 gpurun_out/r02o_2gpu_numa.err
