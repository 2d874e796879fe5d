#!/bin/bash
# ~40 s on one B200: A/B of the k_dt_grad variants, then the two parity tests that cover them
mkdir -p gpurun_out
timeout 40 python scripts/quick_ab.py > gpurun_out/ab_default.jsonl 2> gpurun_out/ab_default.err; echo "default rc=$?"
AFX_LIB=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_nopreload.so timeout 40 python scripts/quick_ab.py > gpurun_out/ab_nopreload.jsonl 2> gpurun_out/ab_nopreload.err; echo "nopreload rc=$?"
cat gpurun_out/ab_default.jsonl gpurun_out/ab_nopreload.jsonl
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "first_stage or explicit_history or fast_mode_history" > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
