#!/usr/bin/env python
"""How far the fast arithmetic mode is from the strict one (= the reference, bit for bit) in the regimes where round 1 allowed
a decade of slack (limiter-constant extremes, transonic / supersonic far field), and over 100 iterations of the bench
configuration (SA / no-slip wall, 2nd order) on the 64k mesh.  Prints max relative deviations; needs one B200."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import aeroflex_b200 as afx  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_oracle_golden import REGIMES, regime_start  # noqa: E402


def dev(a, b, floor=0.0):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if floor else float(np.max(np.abs(a - b) / np.abs(b)))


out = {}
m = afx.Mesh.synth_omesh(64, 32, 8, 50.0)
for k in (0.0, 1e-3, 5.0, 1e6):
    bcs = {"farfield": ("farfield", dict(mach=0.5, angle=0.02, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    r = []
    for math in ("strict", "fast"):
        s = afx.GpuSolver(m, math=math)
        s.set_bcs(bcs); s.set_options(True, "green-gauss", k, 1.0); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q(), amp=1e-2))
        r.append((s.run(4, 0.9), s.get_q(), s.get("limiters")))
    out["limiter_k=%g" % k] = dict(norm=dev(r[1][0], r[0][0]), q=dev(r[1][1], r[0][1], 1e-3), lim=float(np.max(np.abs(r[1][2] - r[0][2]))))
d = H.load("naca0012q_coarse_euler_gg_o2")
mm = H.product_mesh(afx, d)
for tag in sorted(REGIMES):
    r = []
    for math in ("strict", "fast"):
        s = afx.GpuSolver(mm, viscosity=REGIMES[tag][3], math=math)
        s.set_q(regime_start(s, tag))
        r.append((s.run(12, 0.9), s.get_q()))
    out[tag] = dict(norm=dev(r[1][0], r[0][0]), per_iter=[float(x) for x in np.abs(r[1][0] - r[0][0]) / r[0][0]], q=dev(r[1][1], r[0][1], 1e-3))
m64 = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
r = []
for math in ("strict", "fast"):
    s = afx.GpuSolver(m64, viscosity="spallart-allmaras", math=math)
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
    s.set_q(H.synth_state(m64.N, s.get_q()))
    r.append((s.run(100, 0.9), s.get_q(), np.array(s.wall_forces("wall"))))
out["bench_config_64k_100it"] = dict(norm=dev(r[1][0], r[0][0]), q=dev(r[1][1], r[0][1], 1e-3), forces=dev(r[1][2], r[0][2]))
print(json.dumps(out, indent=1))
