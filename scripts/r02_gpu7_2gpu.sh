#!/bin/bash
# round 2, call 7 (TWO B200s): where the halo hand-off spends its time (separate signalling / wait kernels timed), new defaults at 2 GPUs
mkdir -p gpurun_out
run() { name=$1; np=$2; shift 2; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $np "$@" > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err; echo "$name rc=$?"; }
run strong16M_2_sep 2 AFX_HALO_EARLY_SIGNAL=0 -- --steps 20 --warmup 5 --no-parity-probe --e2e-steps 2
run strong16M_2 2 X=1 -- --steps 20 --warmup 5 --e2e-steps 2
python - <<PY
import json, glob
for n in sorted(glob.glob("gpurun_out/r02g_*.json")):
    try:
        d=json.loads(open(n).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(n.split("r02g_")[1], d["n_gpus"], "value %.4g ms %.4f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "parity", p.get("ok"), p.get("norm_max_rel_diff"))
        print("   ", {k: [round(x,4) for x in v] for k,v in d["details"]["rank_spread"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
tail -c 400 gpurun_out/r02g_*.err | tail -12
