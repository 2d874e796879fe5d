#!/usr/bin/env python
"""Converged CL/CD/CM of the UNMODIFIED reference (oracle/_ref), explicit path, driven to ||R|| / ||R_0|| <= 1e-13 -- deep
enough that the north-star force tolerance (1e-8 relative) can be ASSERTED, which the implicit sweep fixtures cannot
support (the reference's ILUT/GMRES iteration stagnates near 1e-11, oracle/make_golden.py sweep_case).
The explicit iteration is deterministic arithmetic: the strict CUDA path must reproduce the final state bit for bit
after the same number of iterations, the fast path and the implicit path must land on the same forces to 1e-8.

    python oracle/make_golden_converged.py        # build container only; ~20 minutes of CPU
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import make_golden as G  # noqa: E402
from oracle import ref  # noqa: E402

if __name__ == "__main__":
    mesh_file = "naca0012q_coarse.msh"
    bcs = {"farfield": ("farfield", G.FAR), "wall": ("slip-wall", None)}
    rm = ref.RefMesh(G.REF_MESHES + mesh_file)
    s = ref.RefSolver(rm, False, "inviscid")
    s.set_bcs(bcs)
    s.set_options(True, "green-gauss", 5.0, 1.5)
    s.init(); s.refill_bcs()
    norms, every = [], 500
    n0 = None
    it = 0
    while True:
        n = s.explicit_solve(0.9)
        it += 1
        if n0 is None:
            n0 = n
        if it % every == 0:
            norms.append(n)
            print(it, n, n / n0, s.wall_forces("wall"), flush=True)
        if (n / n0 <= 1e-13 and it % every == 0) or it >= 200000 or not np.isfinite(n):
            break
    q = s.get("q")
    d = G.mesh_fixture(rm)
    d.update(n_iter=np.array(it), every=np.array(every), norms_every=np.array(norms), norm0=np.array(n0), norm_last=np.array(n),
             sha_q=np.array(G.sha(q)), q=q, forces=np.array(s.wall_forces("wall")), forces_patch=np.array("wall"),
             meta=np.array(repr(dict(mesh=mesh_file, bcs=bcs, viscosity="inviscid", gradient="green-gauss", second_order=True, cfl=1.5, relax=0.9,
                                     start="init() + refill_bcs(), no perturbation", stop="||R||/||R_0|| <= 1e-13 at a multiple of 500 iterations"))))
    np.savez_compressed(os.path.join(G.OUT, "converged_naca0012q_coarse_explicit.npz"), **d)
    print("done", it, n / n0, d["forces"])
