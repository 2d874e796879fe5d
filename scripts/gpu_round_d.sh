#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/d_fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -3 gpurun_out/d_fused_tests.log
AFX_TILE=320 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 6 -c 2 -o gpurun_out/d_prof_stage -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/d_ncu.log 2>&1
tail -2 gpurun_out/d_ncu.log | cut -c1-300
