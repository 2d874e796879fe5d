#!/usr/bin/env python
"""A/B of the k_dt_grad variants on one B200 in well under a minute (no torch import): per-phase CUDA-event times of the
explicit iteration with the first-stage limiter inside k_dt_grad (AFX_FUSE_LIM0=1) and in its own launch (=0), for the
library named by AFX_LIB (default build: neighbour preload; scripts/build_variants.py builds the earlier kernel).
One JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import aeroflex_b200 as afx  # noqa: E402

t00 = time.perf_counter()
small = os.environ.get("QUICK_AB_SMALL") == "1"
ni, nj, nq = (128, 80, 32) if small else (1024, 640, 256)
iters = 5 if small else 300
m = afx.Mesh.synth_omesh(ni, nj, nq, 150.0)
bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
rng = np.random.default_rng(12345)
pert = 1 + 1e-3 * rng.uniform(-1, 1, 4 * m.N)
ref = None
for fuse in os.environ.get("QUICK_AB_FUSE", "1,0").split(","):
    os.environ["AFX_FUSE_LIM0"] = fuse
    s = afx.GpuSolver(m, viscosity="spallart-allmaras", math="fast")
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
    q0 = s.get_q(); q0[:4 * m.N] *= pert; s.set_q(q0)
    s.run(10, 0.9)
    l0 = s.launch_count()
    norms = s.run(iters, 0.9)
    ms = s.last_device_ms() / iters
    per_iter = (s.launch_count() - l0) / iters
    prof = s.profile_explicit(5, 0.9)
    if ref is None:
        ref = norms
    print(json.dumps({"lib": os.path.basename(afx.library_path()), "fuse_lim0": int(fuse), "cells": int(m.N), "ms_per_iteration": ms,
                      "cell_updates_per_s": m.N / (ms * 1e-3), "kernels_per_iteration": per_iter, "phase_ms": prof,
                      "norm_last": float(norms[-1]), "norms_match_first_config": bool(np.allclose(norms, ref, rtol=1e-9)),
                      "elapsed_s": time.perf_counter() - t00}), flush=True)
    del s
