#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/b_fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -3 gpurun_out/b_fused_tests.log
AFX_TILE=256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 6 -c 2 -o gpurun_out/b_prof_stage_T256 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_ncu.log 2>&1
AFX_TILE=384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 6 -c 1 -o gpurun_out/b_prof_stage_T384 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 >> gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
