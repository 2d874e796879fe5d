#!/bin/bash
# round 2, call 3 (one B200): GPU suite (lock-step in-process halo, norm finish kernel, k_pipe v2), A/B of k_pipe variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02c_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02c_gpu_tests.log
L=$PWD/aeroflex_b200/lib
PIPE_AB_MESH=1M PIPE_AB_CONFIGS="off;15,3,2;16,3,2;15,2,1;14,4,3;16,2,2;17,2,1" timeout 200 python scripts/pipe_ab.py > gpurun_out/r02c_pipe_ab_1M.jsonl 2> gpurun_out/r02c_pipe_ab_1M.err; echo "ab 1M rc=$?"
for v in pipe_ept1 pipe_ept4 pipe_t128 pipe_r80; do
  AFX_LIB=$L/libaeroflex_rans_b200_$v.so PIPE_AB_MESH=1M PIPE_AB_CONFIGS="15,3,2;16,3,2" timeout 100 python scripts/pipe_ab.py > gpurun_out/r02c_pipe_ab_1M_$v.jsonl 2> gpurun_out/r02c_pipe_ab_1M_$v.err; echo "ab 1M $v rc=$?"
done
PIPE_AB_MESH=16M PIPE_AB_CONFIGS="off;15,3,2;16,3,2;16,2,2" timeout 400 python scripts/pipe_ab.py > gpurun_out/r02c_pipe_ab_16M.jsonl 2> gpurun_out/r02c_pipe_ab_16M.err; echo "ab 16M rc=$?"
AFX_LIB=$L/libaeroflex_rans_b200_pipe_ept4.so PIPE_AB_MESH=16M PIPE_AB_CONFIGS="15,3,2;16,3,2" timeout 300 python scripts/pipe_ab.py > gpurun_out/r02c_pipe_ab_16M_pipe_ept4.jsonl 2> gpurun_out/r02c_pipe_ab_16M_pipe_ept4.err; echo "ab 16M ept4 rc=$?"
python - <<PY
import json, glob
for n in sorted(glob.glob("gpurun_out/r02c_pipe_ab_*.jsonl")):
    for l in open(n):
        d=json.loads(l); print(n.split("r02c_pipe_ab_")[1], d["config"], "ms %.4f"%d["ms_per_iteration"], "%.4g"%d["cell_updates_per_s"], d["kernels_per_iteration"], {k: round(v,4) for k,v in d["phase_ms"].items() if v}, d["max_rel_norm_diff_vs_first"])
PY
