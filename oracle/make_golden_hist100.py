#!/usr/bin/env python
"""Regenerate the explicit fixtures with 100-iteration residual histories (the north-star's "first 100 iterations"),
from the UNMODIFIED reference (oracle/_ref), like oracle/make_golden.py whose case definitions it reuses.
Round 1 carried 100 iterations for one of the six cases and 20-30 for the others.

    python oracle/make_golden_hist100.py          # build container only (needs /root/reference)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import make_golden as G  # noqa: E402

if __name__ == "__main__":
    slip = {"farfield": ("farfield", G.FAR), "wall": ("slip-wall", None)}
    wall = {"farfield": ("farfield", G.FAR), "wall": ("wall", None)}
    plate = {"top": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "left": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)),
             "right": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "bot0": ("slip-wall", None), "bot1": ("wall", None)}
    only = sys.argv[1:]
    cases = [
        ("naca0012_coarse_laminar_lsq_o2", "naca0012_coarse.msh", wall, "laminar", "least-squares", True, dict(full=False)),
        ("naca0012_coarse_sa_gg_o1", "naca0012_coarse.msh", wall, "spallart-allmaras", "green-gauss", False, dict(full=False)),
        ("naca0012_coarse_euler_gg_o1", "naca0012_coarse.msh", slip, "inviscid", "green-gauss", False, dict(full=False)),
        ("flat_plate_laminar_gg_o2", "flat_plate.msh", plate, "laminar", "green-gauss", True, dict(full=False, amp=0.0, cfl=1e-4)),
        ("flat_plate_sa_gg_o2", "flat_plate.msh", plate, "spallart-allmaras", "green-gauss", True, dict(full=False, amp=1e-4, cfl=1.0)),
    ]
    for tag, mesh, bcs, visc, grad, so, kw in cases:
        if only and tag not in only:
            continue
        G.explicit_case(tag, mesh, bcs, visc, grad, so, 100, **kw)
