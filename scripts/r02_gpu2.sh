#!/bin/bash
# round 2, call 2 (one B200): GPU suite again (group cases fixed, pipelined-stage tests), A/B of k_pipe at 1M and 16M
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02b_gpu_tests.log
PIPE_AB_MESH=1M PIPE_AB_CONFIGS="off;15,2,2;16,2,2;14,3,3;15,1,1;15,3,3;16,1,1;17,1,1" timeout 200 python scripts/pipe_ab.py > gpurun_out/r02b_pipe_ab_1M.jsonl 2> gpurun_out/r02b_pipe_ab_1M.err; echo "ab 1M rc=$?"
PIPE_AB_MESH=16M PIPE_AB_CONFIGS="off;15,2,2;16,2,2;14,3,3;15,1,1;16,1,1" timeout 400 python scripts/pipe_ab.py > gpurun_out/r02b_pipe_ab_16M.jsonl 2> gpurun_out/r02b_pipe_ab_16M.err; echo "ab 16M rc=$?"
python - <<PY
import json
for n in ["r02b_pipe_ab_1M","r02b_pipe_ab_16M"]:
    for l in open("gpurun_out/%s.jsonl"%n):
        d=json.loads(l); print(d["mesh"], d["config"], "ms %.4f"%d["ms_per_iteration"], "%.4g"%d["cell_updates_per_s"], d["kernels_per_iteration"], {k: round(v,4) for k,v in d["phase_ms"].items()}, d["max_rel_norm_diff_vs_first"])
PY
tail -5 gpurun_out/r02b_pipe_ab_1M.err gpurun_out/r02b_pipe_ab_16M.err
