#!/bin/bash
# 8-GPU session: partitioned parity tests, 16M strong scaling at 2/4/8, weak scaling at 8 (1M and 8M cells per GPU), polar sweep
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/m8_multi_tests.log 2>&1
echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/m8_multi_tests.log
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$1 N=%d value %.4g ms/step %.4f e2e %.4g scaling=%s phases %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["scaling"][:6], {k: round(v,4) for k,v in r["phase_ms_per_iteration"].items() if v}))
    print("   ", d["config"]["parallelism"][:200])
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/$1.err").read()[-800:])
PY
}
tr() { # name nproc args...
  name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  show $name
}
for n in 2 4 8; do tr m8_strong16M_$n $n --steps 50 --warmup 5 --e2e-steps 3 --workload synthetic-16M-mixed-omesh; done
tr m8_weak1M_8 8 --steps 100 --warmup 10
# session-3 experiments that no multi-GPU box has timed yet: flags raised from inside the update kernel; graph partition
AFX_HALO_EARLY_SIGNAL=1 tr m8_weak1M_8_early_signal 8 --steps 100 --warmup 10
AFX_HALO_EARLY_SIGNAL=1 tr m8_strong16M_8_early_signal 8 --steps 50 --warmup 5 --e2e-steps 3 --workload synthetic-16M-mixed-omesh
AFX_PARTITION=graph tr m8_weak1M_8_graph_partition 8 --steps 100 --warmup 10
timeout 300 python scripts/polar_sweep_bench.py 8 > gpurun_out/m8_polar_8gpu.json 2> gpurun_out/m8_polar.err
python - <<PY
import json
d=json.loads(open("gpurun_out/m8_polar_8gpu.json").read().strip().splitlines()[-1])
print("polar", d["n_gpus"], "gpus", d["alphas"], "alphas", "%.1f s"%d["wall_s"], d["rc"], d["stderr_tail"])
PY
avail=$(free -g | awk '/Mem:/{print $7}')
if [ "$avail" -gt 400 ]; then tr m8_weak64M_8 8 --steps 20 --warmup 3 --e2e-steps 2 --workload synthetic-64M-mixed-omesh; else echo "skip 64M: only $avail GB of host memory available"; fi
