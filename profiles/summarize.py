#!/usr/bin/env python
"""Turn an Nsight Compute report brought back from the GPU box (gpurun_out/*.ncu-rep) and the launch list of the same
command (gpurun_out/*launches.csv) into the tracked summaries under profiles/.

    python profiles/summarize.py gpurun_out/r01_prof.ncu-rep gpurun_out/r01_launches.csv r01
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def main(rep, launches, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    per = OrderedDict()
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        per.setdefault(k, []).append(r)
    out = OrderedDict()
    for k, rs in per.items():
        d = OrderedDict(launches_captured=len(rs))
        for m in METRICS:
            if m in col:
                vals = [float(r[col[m]].replace(",", "")) for r in rs if r[col[m]] not in ("", "n/a")]
                if vals:
                    d[m + " [" + units[col[m]] + "]"] = sum(vals) / len(vals)
        def unit(prefix, scale):  # ncu picks the unit per report (us / ms, Kbyte / Mbyte / Gbyte): normalise
            for u, f in scale.items():
                v = d.get("%s [%s]" % (prefix, u))
                if v is not None:
                    return v * f
            return None
        BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        rd = unit("dram__bytes_read.sum", BYTES); wr = unit("dram__bytes_write.sum", BYTES)
        t = unit("gpu__time_duration.sum", {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6})
        if t is not None:
            d["time_us"] = t
        if rd is not None and wr is not None:
            d["dram_traffic_per_launch_MB"] = rd + wr
            if t:
                d["dram_GBps_during_kernel"] = (rd + wr) / t * 1e3
        out[k] = d
    with open(os.path.join(HERE, tag + "_ncu_full_summary.json"), "w") as f:
        json.dump(out, f, indent=1)
    # launch list: per-kernel share of one iteration
    if launches and os.path.exists(launches):
        tot = defaultdict(float); cnt = defaultdict(int)
        with open(launches) as f:
            for r in csv.DictReader(l for l in f if l.startswith('"')):
                if r.get("Metric Name") == "gpu__time_duration.sum":
                    k = short(r["Kernel Name"])
                    tot[k] += float(r["Metric Value"].replace(",", "")); cnt[k] += 1
        total = sum(tot.values())
        lines = ["| kernel | launches | mean us | share of listed time |", "|---|---|---|---|"]
        for k in sorted(tot, key=lambda z: -tot[z]):
            lines.append("| %s | %d | %.1f | %.1f %% |" % (k, cnt[k], tot[k] / cnt[k] / 1e3, 100 * tot[k] / total))
        with open(os.path.join(HERE, tag + "_launch_shares.md"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n\n" + "\n".join(lines) + "\n")
        import shutil
        shutil.copy(launches, os.path.join(HERE, tag + "_launches.csv"))
    print(json.dumps({k: {"us": v.get("time_us"), "dram_MB": v.get("dram_traffic_per_launch_MB"),
                          "dram_GBps": v.get("dram_GBps_during_kernel")} for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else "r01")
