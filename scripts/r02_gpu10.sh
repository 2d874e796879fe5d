#!/bin/bash
# round 2, call 10 (one B200): the final tree -- whole GPU suite, smoke, headline bench lines, ncu launch list + full captures, compute-sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 400 > gpurun_out/r02j_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02j_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r02j_bench_16M.json 2> gpurun_out/r02j_bench_16M.err; echo "bench16M rc=$?"
timeout 200 python bench.py --workload synthetic-1M-mixed-omesh --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02j_bench_1M.json 2> gpurun_out/r02j_bench_1M.err; echo "bench1M rc=$?"
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/r02j_launches_16M.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02j_ncu_list_16M.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_limiter|k_gather_update|k_dt_grad|k_norm_finish" -s 20 -c 10 -o gpurun_out/r02j_prof_16M -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02j_ncu_full_16M.log 2>&1; echo "ncu full 16M rc=$?"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_group.py -q -m gpu -k "2-p2p-strict" > gpurun_out/r02j_sanitizer_memcheck_group_p2p.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02j_sanitizer_memcheck_group_p2p.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -m gpu -k "bit_identical and 192" > gpurun_out/r02j_sanitizer_racecheck_fused.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02j_sanitizer_racecheck_fused.log
timeout 150 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_pipe.py -q -m gpu -k "strict and 10-2-2" > gpurun_out/r02j_sanitizer_synccheck_pipe.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r02j_sanitizer_synccheck_pipe.log
python - <<PY
import json
for n in ["r02j_bench_16M","r02j_bench_1M","r02j_bench_reference"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(n, "%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], r.get("kernel"), r.get("frac"), (r.get("iteration") or {}).get("frac"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_port",{}).get("value"), (d.get("details") or {}).get("setup_s"), d.get("gpu_launches"))
        print("   phases", r.get("phase_ms_per_iteration"))
    except Exception as e:
        print(n,"failed",e)
PY
