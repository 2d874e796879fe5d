#!/usr/bin/env python
"""tests/golden/michalak_*.npz: the reference's OTHER limiter.  calc_limiters (solver.h:517-593) evaluates Michalak's function
(solver.h:557-576, physics.h:581-592) instead of Venkatakrishnan's when the headers are compiled with -DRANS_MICHALAK_LIMITER; nothing
in the reference's build defines the macro, so this is the unmodified reference headers built with that one switch
(oracle/Makefile -> oracle/_ref/libafx_ref_michalak.so).  Explicit histories on two shipped meshes, with limiter constants and
perturbation amplitudes chosen so that all three regimes of the switch (sig = 1, 0 < sig < 1, sig = 0) and both branches of the cubic occur.

    python oracle/make_golden_michalak.py        (build container only: needs /root/reference)

TEST INFRASTRUCTURE ONLY."""
import os
import sys

os.environ["AFX_REF_VARIANT"] = "michalak"  # before oracle.ref is imported: it names the library
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from oracle.make_golden import OUT, REF_MESHES, mesh_fixture, perturb, sha  # noqa: E402

assert ref.SO.endswith("libafx_ref_michalak.so")


def case(tag, mesh_file, bcs, viscosity, gradient, limiter_k, amp, n_iter=30, cfl=1.5, relax=0.9):
    rm = ref.RefMesh(REF_MESHES + mesh_file)
    d = mesh_fixture(rm)
    s = ref.RefSolver(rm, False, viscosity)
    s.set_bcs(bcs)
    s.set_options(True, gradient, limiter_k, cfl)
    s.init(); s.refill_bcs()
    q0 = perturb(s.get("q"), rm.N, amp=amp)
    s.set("q", q0)
    d["q0"] = q0
    norms = np.zeros(n_iter)
    for it in range(n_iter):
        norms[it] = s.explicit_solve(relax)
        if it == 0:
            for nm in ("q", "qW", "limiters"):
                d["it1_" + nm] = s.get(nm)
    d["norms"] = norms
    d["qN"] = s.get("q")
    d["limN"] = s.get("limiters")
    wall = [n for n in rm.patch_names if bcs[n][0] in ("wall", "slip-wall")]
    d["forces_patch"] = np.array(wall[0])
    d["forces"] = np.array(s.wall_forces(wall[0]))
    d["meta"] = np.array(repr(dict(mesh=mesh_file, bcs=bcs, viscosity=viscosity, gradient=gradient, second_order=True, limiter="michalak",
                                   limiter_k=limiter_k, n_iter=n_iter, cfl=cfl, relax=relax, amp=amp, seed=12345)))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)
    l1 = d["it1_limiters"][:4 * rm.N]
    print(tag, "N,G,E =", rm.N, rm.G, rm.E, "norms", norms[0], norms[-1], "| limiters of the first iteration: %.1f %% below 1, %.1f %% below 0.5, min %.3g"
          % (100 * np.mean(l1 < 1), 100 * np.mean(l1 < 0.5), l1.min()), "forces", d["forces"])


FAR = dict(mach=0.2, angle=1.0 * 0.01745, T=1.0, p=1.0)
case("michalak_naca0012q_coarse_euler_gg", "naca0012q_coarse.msh", {"farfield": ("farfield", FAR), "wall": ("slip-wall", None)},
     "inviscid", "green-gauss", limiter_k=0.5, amp=1e-2)
case("michalak_naca0012_coarse_laminar_lsq", "naca0012_coarse.msh", {"farfield": ("farfield", FAR), "wall": ("wall", None)},
     "laminar", "least-squares", limiter_k=0.3, amp=1e-3)
