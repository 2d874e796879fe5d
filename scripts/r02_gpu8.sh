#!/bin/bash
# round 2, call 8 (one B200): whole GPU suite on the final tree, fast-mode deviations, limiter byte diet A/B, conf.ini polar
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 400 > gpurun_out/r02h_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02h_gpu_tests.log
timeout 200 python scripts/fast_mode_deviation.py > gpurun_out/r02h_fast_mode_deviation.json 2> gpurun_out/r02h_fast_mode_deviation.err; echo "deviation rc=$?"
for pm in 1 0; do
  AFX_LIM_PM=$pm PIPE_AB_MESH=1M PIPE_AB_CONFIGS="off" timeout 100 python scripts/pipe_ab.py > gpurun_out/r02h_pm${pm}_1M.jsonl 2>> gpurun_out/r02h_pm.err
  AFX_LIM_PM=$pm PIPE_AB_MESH=16M PIPE_AB_CONFIGS="off" timeout 200 python scripts/pipe_ab.py > gpurun_out/r02h_pm${pm}_16M.jsonl 2>> gpurun_out/r02h_pm.err
done
timeout 300 python bench.py --workload confini-polar > gpurun_out/r02h_bench_confini_polar.json 2> gpurun_out/r02h_bench_confini_polar.err; echo "polar rc=$?"
timeout 200 python bench.py --workload confini-polar --impl reference --steps 1 > gpurun_out/r02h_bench_confini_polar_reference.json 2> gpurun_out/r02h_bench_confini_polar_reference.err; echo "polar ref rc=$?"
python - <<PY
import json, glob
print(open("gpurun_out/r02h_fast_mode_deviation.json").read()[:3000])
for n in sorted(glob.glob("gpurun_out/r02h_pm*.jsonl")):
    for l in open(n):
        d=json.loads(l); print(n.split("r02h_")[1], "ms %.4f"%d["ms_per_iteration"], "%.4g"%d["cell_updates_per_s"], d["kernels_per_iteration"], {k: round(v,4) for k,v in d["phase_ms"].items() if v}, "norm_last", d["norm_last"])
for n in ["r02h_bench_confini_polar", "r02h_bench_confini_polar_reference"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["unit"], d.get("seconds"), d.get("parity"), [ (r["alpha"], r["iterations"], r["cl"]) for r in d.get("polar", [])], d.get("forces"))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02h_*.err | tail -30
