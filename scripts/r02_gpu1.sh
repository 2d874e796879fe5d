#!/bin/bash
# round 2, call 1 (one B200): GPU suite incl. the one-GPU partition tests, smoke, headline bench (16M), 1M bench, ncu launch lists + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02a_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r02a_bench_16M.json 2> gpurun_out/r02a_bench_16M.err; echo "bench16M rc=$?"
timeout 200 python bench.py --workload synthetic-1M-mixed-omesh --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02a_bench_1M.json 2> gpurun_out/r02a_bench_1M.err; echo "bench1M rc=$?"
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02a_launches_16M.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02a_ncu_list_16M.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_limiter|k_gather_update|k_dt_grad" -s 18 -c 9 -o gpurun_out/r02a_prof_16M -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02a_ncu_full_16M.log 2>&1; echo "ncu full 16M rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_limiter|k_gather_update|k_dt_grad" -s 18 -c 9 -o gpurun_out/r02a_prof_1M -f python bench.py --workload synthetic-1M-mixed-omesh --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02a_ncu_full_1M.log 2>&1; echo "ncu full 1M rc=$?"
python - <<PY
import json
for n in ["r02a_bench_16M","r02a_bench_1M","r02a_bench_reference"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(n, "%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], r.get("kernel"), r.get("frac"), (r.get("iteration") or {}).get("frac"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_port",{}).get("value"), (d.get("details") or {}).get("setup_s"))
        print("   phases", r.get("phase_ms_per_iteration"))
    except Exception as e:
        print(n,"failed",e)
PY
ls -la gpurun_out/ | tail -15
