#!/bin/bash
# round 2, last call (one B200): the FINAL tree -- whole GPU suite, smoke, headline bench lines (16M, 1M, confini-polar, reference arm), ncu launch
# list + full capture at 1M cells on the final kernels (the 16M capture of call 10 predates only the k_dt_grad register change)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/r02m_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02m_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r02m_bench_16M.json 2> gpurun_out/r02m_bench_16M.err; echo "bench16M rc=$?"
timeout 120 python bench.py --workload synthetic-1M-mixed-omesh --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02m_bench_1M.json 2> gpurun_out/r02m_bench_1M.err; echo "bench1M rc=$?"
timeout 120 python bench.py --workload confini-polar > gpurun_out/r02m_bench_confini_polar.json 2> gpurun_out/r02m_bench_confini_polar.err; echo "polar rc=$?"
AFX_KRY_GRAPH=0 timeout 120 python bench.py --workload confini-polar > gpurun_out/r02m_bench_confini_polar_nograph.json 2> gpurun_out/r02m_bench_confini_polar_nograph.err; echo "polar (plain launches) rc=$?"
AFX_KRY_FUSE_GIVENS=0 timeout 120 python bench.py --workload confini-polar > gpurun_out/r02m_bench_confini_polar_nofuse.json 2> gpurun_out/r02m_bench_confini_polar_nofuse.err; echo "polar (rotation in its own launch) rc=$?"
AFX_GMRES_BATCH=12 timeout 120 python bench.py --workload confini-polar > gpurun_out/r02m_bench_confini_polar_batch12.json 2> gpurun_out/r02m_bench_confini_polar_batch12.err; echo "polar (batch 12) rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_limiter|k_gather_update|k_dt_grad" -s 20 -c 9 -o gpurun_out/r02m_prof_1M -f python bench.py --workload synthetic-1M-mixed-omesh --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02m_ncu_full_1M.log 2>&1; echo "ncu full 1M rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_dt_grad" -s 3 -c 1 -o gpurun_out/r02m_prof_16M_dtgrad -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02m_ncu_full_16M_dtgrad.log 2>&1; echo "ncu dt_grad 16M rc=$?"
timeout 240 python bench.py --workload polar64 > gpurun_out/r02m_bench_polar64_1gpu.json 2> gpurun_out/r02m_bench_polar64_1gpu.err; echo "polar64 rc=$?"
# where the implicit path's time goes: per-kernel GPU time of the first 6000 launches of the polar (serialised under ncu: compare the SUM with the wall time above)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 6000 --csv --log-file gpurun_out/r02m_launches_polar.csv python bench.py --workload confini-polar > gpurun_out/r02m_ncu_list_polar.log 2>&1; echo "ncu list polar rc=$?"
python - <<PY
import json
for n in ["r02m_bench_16M", "r02m_bench_1M", "r02m_bench_confini_polar", "r02m_bench_confini_polar_nograph", "r02m_bench_confini_polar_nofuse", "r02m_bench_confini_polar_batch12", "r02m_bench_polar64_1gpu"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(n, "%.4g" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], r.get("kernel"), r.get("frac"), (r.get("iteration") or {}).get("frac"),
              {k: round(v["frac"], 3) for k, v in (r.get("kernels") or {}).items()}, d.get("parity"))
    except Exception as e:
        print(n, "failed", e)
PY
