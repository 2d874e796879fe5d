# 1/2/4/8-GPU weak-scaling series + the 16M-cell single-GPU roofline run (gpurun --gpus 8 -- bash scripts/scale_series.sh)
mkdir -p gpurun_out
set -x
nvidia-smi -L | wc -l
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err;
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$n.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=$n value %.4g ms/step %.4f e2e %.4g (%.2f ms) launches %d cells %d phases %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"]["cells"], {k: round(v,4) for k,v in r["phase_ms_per_iteration"].items()}))
    print("   ", d["config"]["parallelism"])
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
# 16M cells on one GPU (BASELINE config 3 at N=1): roofline at the target size
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --workload synthetic-16M-mixed-omesh --e2e-steps 3 > gpurun_out/bench_16M.json 2> gpurun_out/bench_16M.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_16M.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("16M value %.4g ms/step %.4f e2e %.4g dom %s frac %.3f iter_frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel"], r["frac"], r["iteration"]["frac"]))
for k,v in r["kernels"].items(): print("  ", k, round(v["kernel_ms"],4), round(v["frac"],3), round(v["share_of_iteration"],3))
PY
