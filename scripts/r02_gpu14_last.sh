#!/bin/bash
# round 2, the last 100 s of the budget: the GPU tests added after the final evidence call (Michalak limiter, sweep failure path, Arnoldi graphs),
# and the 1M-cell iteration time of the shipped library (the default kernels must not have moved)
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_edge_cases.py -q -m gpu -k "michalak or sweep_goes_on or arnoldi" 2>&1 | tail -6 > gpurun_out/r02n_new_tests.log; cat gpurun_out/r02n_new_tests.log
PIPE_AB_MESH=1M PIPE_AB_CONFIGS=off timeout 25 python scripts/pipe_ab.py > gpurun_out/r02n_1M.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r02n_1M.jsonl
