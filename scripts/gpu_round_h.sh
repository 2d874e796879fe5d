#!/bin/bash
mkdir -p gpurun_out
AFX_TILE=384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 6 -c 1 -o gpurun_out/h_prof_512x1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/h_ncu.log 2>&1
AFX_LIB=$PWD/aeroflex_b200/lib/libafx_s256x2.so AFX_STAGE_CTAS=2 AFX_TILE=192 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stage -s 6 -c 1 -o gpurun_out/h_prof_256x2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 >> gpurun_out/h_ncu.log 2>&1
tail -2 gpurun_out/h_ncu.log | cut -c1-200
