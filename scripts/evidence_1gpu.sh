#!/bin/bash
# round-end evidence on one B200: GPU tests, smoke, bench lines (fast / strict / 16M / reference arm), ncu launch list and full capture
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 120 > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/f_gpu_tests.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo "bench rc=$?"
timeout 150 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "ref rc=$?"
timeout 120 python bench.py --steps 200 --warmup 10 --math strict --no-cpu-baseline > gpurun_out/f_bench_1gpu_strict.json 2> gpurun_out/f_bench_1gpu_strict.err
timeout 240 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 --workload synthetic-16M-mixed-omesh > gpurun_out/f_bench_16M.json 2> gpurun_out/f_bench_16M.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/f_ncu_list.log 2>&1
# CTA sizes session 3 could not try (scripts/build_variants.py builds them next to the default library on the CPU side first)
for v in fl64 g512; do
  [ -f aeroflex_b200/lib/libaeroflex_rans_b200_$v.so ] && AFX_LIB=$PWD/aeroflex_b200/lib/libaeroflex_rans_b200_$v.so QUICK_AB_FUSE=1 timeout 30 python scripts/quick_ab.py >> gpurun_out/f_ab_cta.jsonl 2>> gpurun_out/f_ab_cta.err
done
QUICK_AB_FUSE=1,0 timeout 30 python scripts/quick_ab.py >> gpurun_out/f_ab_cta.jsonl 2>> gpurun_out/f_ab_cta.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_flux|k_limiter|k_gather_update|k_dt_grad" -s 20 -c 8 -o gpurun_out/f_prof -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/f_ncu_full.log 2>&1
python - <<PY
import json
for n in ["f_bench_1gpu","f_bench_1gpu_strict","f_bench_16M","f_bench_reference"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(n, "%.4g"%d["value"], "ms %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], r.get("kernel"), r.get("frac"), (r.get("iteration") or {}).get("frac"), d.get("cpu_baseline",{}).get("value"), d.get("cpu_port",{}).get("value"))
    except Exception as e:
        print(n,"failed",e)
PY
ls -la gpurun_out/f_prof.ncu-rep gpurun_out/f_launches.csv
