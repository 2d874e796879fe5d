#!/bin/bash
# round 2, call 5 (EIGHT B200s): multi-process tests with 3-8 ranks, BASELINE config 3 (16M strong) and 4 (64M weak) with the parity probe,
# Hilbert vs graph partition, flag hand-off from inside the update kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 300 -k "3- or 4- or 5- or 8- or 3] or early" > gpurun_out/r02e_gpu_tests_8gpu.log 2>&1; echo "8-gpu tests rc=$?"; tail -4 gpurun_out/r02e_gpu_tests_8gpu.log
run() { # name, nproc, extra env..., then bench args after --
  name=$1; np=$2; shift 2
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err; echo "$name rc=$?"
}
run strong16M_8 8 X=1 -- --steps 40 --warmup 5
run strong16M_8_graph 8 AFX_PARTITION=graph -- --steps 40 --warmup 5
run strong16M_8_early 8 AFX_HALO_EARLY_SIGNAL=1 -- --steps 40 --warmup 5 --no-parity-probe
run strong16M_8_nccl 8 AFX_HALO=nccl -- --steps 40 --warmup 5 --no-parity-probe
run strong16M_4 4 X=1 -- --steps 40 --warmup 5 --no-parity-probe
run weak64M_8 8 X=1 -- --steps 20 --warmup 5 --scaling weak
python - <<PY
import json, glob
for n in sorted(glob.glob("gpurun_out/r02e_*.json")):
    try:
        d=json.loads(open(n).read().strip().splitlines()[-1])
        p=d.get("parity") or {}
        print(n.split("r02e_")[1], d["config"]["workload"], d["n_gpus"], "value %.4g ms %.4f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "parity", p.get("ok"), p.get("norm_max_rel_diff"), {k: round(v,4) for k,v in d["roofline"]["phase_ms_per_iteration"].items() if v}, d["details"]["parallelism"][-60:], "setup %.0fs"%d["details"]["setup_s"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -c 600 gpurun_out/r02e_*.err | tail -40
