#!/usr/bin/env python
"""Build tuning variants of the CUDA library next to the default one (for A/B runs with AFX_LIB=...)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aeroflex_b200 as afx  # noqa: E402

VARIANTS = {
    # the k_dt_grad of sessions 1-2: one neighbour gather at a time, 80 registers, 3 CTAs per SM
    "nopreload": ["-DAFX_DTG_PRELOAD=0"],
    "nopreload_pf": ["-DAFX_DTG_PRELOAD=0", "-DAFX_DTG_DXY=1"],
    "nopreload_early": ["-DAFX_DTG_PRELOAD=0", "-DAFX_DTG_DXY=2"],
    "nopreload_minb4": ["-DAFX_DTG_PRELOAD=0", "-DAFX_DTG_MINB=4"],
    "preload_pf": ["-DAFX_DTG_PRELOAD=1", "-DAFX_DTG_DXY=1"],
    "preload_minb3": ["-DAFX_DTG_PRELOAD=1", "-DAFX_DTG_MINB=3"],
    "preload_minb3_pf": ["-DAFX_DTG_PRELOAD=1", "-DAFX_DTG_MINB=3", "-DAFX_DTG_DXY=1"],
    "fl64": ["-DAFX_FLUX_THREADS=64", "-DAFX_LIM_THREADS=64"],      # untried in session 3
    "g512": ["-DAFX_GATHER_THREADS=512"],                              # untried in session 3
    "st128": ["-DAFX_FLUX_THREADS=128", "-DAFX_LIM_THREADS=128", "-DAFX_GATHER_THREADS=128"],
    "fl128": ["-DAFX_FLUX_THREADS=128", "-DAFX_LIM_THREADS=128"],
    "nopreload_t128": ["-DAFX_DTG_PRELOAD=0", "-DAFX_DTG_THREADS=128", "-DAFX_DTG_MINB=6"],
    # pipelined stage kernel: elements per thread and item, CTA shape (rans_pipe.cuh)
    "pipe_ept1": ["-DAFX_PIPE_EPT=1"],
    "pipe_ept4": ["-DAFX_PIPE_EPT=4"],
    "pipe_t128": ["-DAFX_PIPE_THREADS=128", "-DAFX_PIPE_MINB=8", "-DAFX_PIPE_EPT=4"],
    "pipe_r80": ["-DAFX_PIPE_THREADS=256", "-DAFX_PIPE_MINB=3", "-DAFX_PIPE_EPT=2"],
    # round 2, final A/B: the stored-extremes limiter instantiation at more CTAs per SM, k_dt_grad at fewer (no spills), deferred wall-ghost stores
    "lpm9": ["-DAFX_LIM_PM_MINB=9"],
    "lpm10": ["-DAFX_LIM_PM_MINB=10"],
    "dtg5": ["-DAFX_DTG_MINB=5"],
    "dtg4": ["-DAFX_DTG_MINB=4"],
    "defer": ["-DAFX_DTG_DEFER=1"],
    "defer5": ["-DAFX_DTG_DEFER=1", "-DAFX_DTG_MINB=5"],
    "defer4": ["-DAFX_DTG_DEFER=1", "-DAFX_DTG_MINB=4"],
    "lpm9_defer5": ["-DAFX_LIM_PM_MINB=9", "-DAFX_DTG_DEFER=1", "-DAFX_DTG_MINB=5"],
    # CTA shapes never timed before the last call of round 2
    "f64": ["-DAFX_FLUX_THREADS=64"],
    "l64": ["-DAFX_LIM_THREADS=64"],
    "f128x7": ["-DAFX_FLUX_MINB=7"],
    "dtg64x10": ["-DAFX_DTG_THREADS=64", "-DAFX_DTG_MINB=10"],
    "dtg256x2": ["-DAFX_DTG_THREADS=256", "-DAFX_DTG_MINB=2"],
    "nopreload_t128_pf": ["-DAFX_DTG_PRELOAD=0", "-DAFX_DTG_THREADS=128", "-DAFX_DTG_MINB=6", "-DAFX_DTG_DXY=1"],
}
for name, defs in VARIANTS.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    out = os.path.join(ROOT, "aeroflex_b200", "lib", "libaeroflex_rans_b200_%s.so" % name)
    print(afx.build_library(force=True, defines=defs, out=out))
