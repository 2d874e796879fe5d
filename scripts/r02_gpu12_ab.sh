#!/bin/bash
# round 2, call 12 (one B200, ~3.5 min): CTA shapes never timed before (64-thread flux / limiter CTAs, 512-thread gather, k_dt_grad 64 x 10 / 256 x 2,
# k_flux at 72 registers) against the default, one process per library, 16M and 1M cells; the polar-chain test; the bench.py code paths added today
mkdir -p gpurun_out
: > gpurun_out/r02l_ab_16M.jsonl; : > gpurun_out/r02l_ab_1M.jsonl; : > gpurun_out/r02l_ab.err
timeout 200 python -m pytest tests/test_gpu_fmg.py -q -m gpu -k "polar_chains" 2>&1 | tail -15 > gpurun_out/r02l_polar_test.log; tail -3 gpurun_out/r02l_polar_test.log
python -c "import bench; print('numa cpus:', bench.bind_to_gpu_numa_node(0))"
timeout 100 python bench.py --workload synthetic-1M-mixed-omesh --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02l_bench_1M.json 2> gpurun_out/r02l_bench_1M.err; echo "bench1M rc=$?"
for v in default f64 l64 f128x7 dtg64x10 dtg256x2 g512 default; do
  lib=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_$v.so
  [ $v = default ] && lib=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200.so
  AFX_LIB=$lib PIPE_AB_MESH=16M PIPE_AB_CONFIGS=off timeout 60 python scripts/pipe_ab.py 2>> gpurun_out/r02l_ab.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02l_ab_16M.jsonl
  AFX_LIB=$lib PIPE_AB_MESH=1M PIPE_AB_CONFIGS=off timeout 30 python scripts/pipe_ab.py 2>> gpurun_out/r02l_ab.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/r02l_ab_1M.jsonl
done
python - <<PY
import json
for n in ("16M", "1M"):
    for l in open("gpurun_out/r02l_ab_%s.jsonl" % n):
        d = json.loads(l)
        print(n, "%-10s" % d["variant"], "%.4f ms" % d["ms_per_iteration"], {k: round(v, 4) for k, v in d["phase_ms"].items() if v}, d["norm_last"])
d = json.loads(open("gpurun_out/r02l_bench_1M.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("bench1M %.4g" % d["value"], {k: round(v["frac"], 3) for k, v in r["kernels"].items()}, "iter", round(r["iteration"]["frac"], 3), "resid+upd", round(r["residual_loop_and_update"]["frac"], 3))
PY
tail -3 gpurun_out/r02l_ab.err
