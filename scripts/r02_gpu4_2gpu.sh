#!/bin/bash
# round 2, call 4 (TWO B200s): the multi-process tests that need one GPU per rank (2-rank cases), the driver's 2-GPU bench line with its parity probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py tests/test_gpu_pipe.py -q -m gpu --timeout 300 > gpurun_out/r02d_gpu_tests_2gpu.log 2>&1; echo "2-gpu tests rc=$?"; tail -4 gpurun_out/r02d_gpu_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_16M_2gpu.json 2> gpurun_out/r02d_bench_16M_2gpu.err; echo "bench 2gpu rc=$?"
tail -c 1500 gpurun_out/r02d_bench_16M_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02d_bench_16M_2gpu.json").read().strip().splitlines()[-1])
    print("value %.4g ms %.4f e2e %.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("parity"), d["details"]["parallelism"], d["roofline"]["phase_ms_per_iteration"])
except Exception as e:
    print("failed", e)
PY
