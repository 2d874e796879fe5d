#!/usr/bin/env python
"""A/B of the pipelined stage kernel (k_pipe) against the three-kernel stage on one B200, no torch import.
PIPE_AB_MESH=1M|16M|64k, PIPE_AB_CONFIGS="off;15,2,2;16,2,2;..." (shift,lagF,lagU).  One JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import aeroflex_b200 as afx  # noqa: E402

MESH = {"64k": (256, 160, 64), "1M": (1024, 640, 256), "4M": (2048, 1280, 512), "16M": (4096, 2560, 1024)}
name = os.environ.get("PIPE_AB_MESH", "1M")
iters = int(os.environ.get("PIPE_AB_ITERS", {"64k": 200, "1M": 200, "4M": 60, "16M": 20}[name]))
m = afx.Mesh.synth_omesh(*MESH[name], 150.0)
bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
rng = np.random.default_rng(12345)
pert = 1 + 1e-3 * rng.uniform(-1, 1, 4 * m.N)
ref = None
for cfg in os.environ.get("PIPE_AB_CONFIGS", "off;15,2,2").split(";"):
    t0 = time.perf_counter()
    if cfg == "off":
        os.environ["AFX_PIPE"] = "0"
    else:
        sh, lf, lu = cfg.split(",")
        os.environ.update(AFX_PIPE="1", AFX_PIPE_SHIFT=sh, AFX_PIPE_LAGF=lf, AFX_PIPE_LAGU=lu)
    s = afx.GpuSolver(m, viscosity="spallart-allmaras", math=os.environ.get("PIPE_AB_MATH", "fast"))
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
    q0 = s.get_q(); q0[:4 * m.N] *= pert; s.set_q(q0)
    s.run(5, 0.9)
    l0 = s.launch_count()
    norms = s.run(iters, 0.9)
    ms = s.last_device_ms() / iters
    per_iter = (s.launch_count() - l0) / iters
    prof = s.profile_explicit(3, 0.9)
    if ref is None:
        ref = norms
    print(json.dumps({"mesh": name, "config": cfg, "pipe": s.pipe_info(), "cells": int(m.N), "ms_per_iteration": ms,
                      "cell_updates_per_s": m.N / (ms * 1e-3), "kernels_per_iteration": per_iter, "phase_ms": prof,
                      "norm_last": float(norms[-1]), "max_rel_norm_diff_vs_first": float(np.max(np.abs(norms - ref) / ref)),
                      "elapsed_s": time.perf_counter() - t0}), flush=True)
    del s
