#!/bin/bash
# round 2, call 9 (EIGHT B200s): final tree -- multi-rank parity tests, 16M strong with the halo on a second stream (AFX_HALO_OVERLAP=1) against the default
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 200 -k "8-strict-p2p-0 or 4-strict-p2p-1 or 3-laminar or 8-fast-p2p-1" > gpurun_out/r02i_gpu_tests_8gpu.log 2>&1; echo "8-gpu tests rc=$?"; tail -3 gpurun_out/r02i_gpu_tests_8gpu.log
run() { name=$1; np=$2; shift 2; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus $np "$@" > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err; echo "$name rc=$?"; }
run strong16M_8 8 X=1 -- --steps 40 --warmup 5 --e2e-steps 2
run strong16M_8_overlap 8 AFX_HALO_OVERLAP=1 -- --steps 40 --warmup 5 --no-parity-probe --e2e-steps 2
python - <<PY
import json, glob
for n in sorted(glob.glob("gpurun_out/r02i_*.json")):
    try:
        d=json.loads(open(n).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(n.split("r02i_")[1], d["n_gpus"], "value %.4g ms %.4f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "parity", p.get("ok"), p.get("norm_max_rel_diff"))
        print("   ", {k: [round(x,4) for x in v] for k,v in d["details"]["rank_spread"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
