#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container only (needs /root/reference for the shipped meshes
and for building oracle/_ref/libafx_ref.so).  The fixtures travel with the
repo; the GPU box and the test-suite never read /root/reference.

    python oracle/make_golden.py

Every vector is produced by the reference's own code (rans::mesh,
rans::explicitSolver / implicitSolver, flux classes, get_wall_profile) compiled
against the Eigen-API stand-in; nothing here comes from oracle/rans_oracle.c or
from the CUDA library.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF_MESHES = "/root/reference/examples/rans/"
OUT = os.path.join(ROOT, "tests", "golden")

FAR = dict(mach=0.2, angle=1.0 * 0.01745, T=1.0, p=1.0)  # rans.h:94 deg->rad literal


def sha(a):
    """sha256 of the bytes; floating-point arrays are hashed with -0.0 folded into +0.0 (the reference leaves
    -0.0 in the ghost rows of gx/gy through `*= 0`, solver.h:467-468; the sign of a zero is not a result)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a + 0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def mesh_fixture(rm):
    d = dict(x=rm.x, y=rm.y, cells=rm.cells, is_tri=rm.is_tri, b0=rm.b0, b1=rm.b1, bpatch=rm.bpatch,
             patch_names=np.array(rm.patch_names), sizes=np.array([rm.N, rm.G, rm.E], np.int64))
    for a in ("edge_cells", "enx", "eny", "elen", "ecx", "ecy", "ccx", "ccy", "area", "cell_edges", "bnd_edge"):
        d["sha_" + a] = np.array(sha(getattr(rm, a)))
    return d


def perturb(q, N, seed=12345, amp=1e-3):
    rng = np.random.default_rng(seed)
    q = q.copy()
    q[:4 * N] *= 1.0 + amp * rng.uniform(-1, 1, 4 * N)
    return q


def explicit_case(tag, mesh_file, bcs, viscosity, gradient, second_order, n_iter, cfl=1.5, relax=0.9, full=True, amp=1e-3):
    rm = ref.RefMesh(REF_MESHES + mesh_file)
    d = mesh_fixture(rm)
    s = ref.RefSolver(rm, False, viscosity)
    s.set_bcs(bcs)
    s.set_options(second_order, gradient, 5.0, cfl)
    d["uniform_residual_fresh"] = np.array(s.uniform_residual())  # qW still all-zero: no stale accumulation
    s.init(); s.refill_bcs()
    q0 = perturb(s.get("q"), rm.N, amp=amp)
    s.set("q", q0)
    d["q0"] = q0
    norms = np.zeros(n_iter)
    for it in range(n_iter):
        norms[it] = s.explicit_solve(relax)
        if it == 0:
            for nm in ("q", "qW", "gx", "gy", "limiters"):
                v = s.get(nm)
                d["sha_it1_" + nm] = np.array(sha(v))
                if full:
                    d["it1_" + nm] = v
            dt = s.get("dt")[:rm.N]
            d["sha_it1_dt"] = np.array(sha(dt))
            if full:
                d["it1_dt"] = dt
    d["norms"] = norms
    qn = s.get("q")
    d["sha_qN"] = np.array(sha(qn))
    if full:
        d["qN"] = qn
    wall = [n for n in rm.patch_names if bcs[n][0] in ("wall", "slip-wall")]
    d["forces_patch"] = np.array(wall[0])
    d["forces"] = np.array(s.wall_forces(wall[0]))
    d["meta"] = np.array(repr(dict(mesh=mesh_file, bcs=bcs, viscosity=viscosity, gradient=gradient, second_order=second_order,
                                   n_iter=n_iter, cfl=cfl, relax=relax, amp=amp, seed=12345)))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)
    print(tag, "N,G,E =", rm.N, rm.G, rm.E, "norm[0], norm[-1] =", norms[0], norms[-1], "forces", d["forces"])


def implicit_case(tag, mesh_file, bcs, viscosity, gradient, second_order, cfl=40.0):
    rm = ref.RefMesh(REF_MESHES + mesh_file)
    d = mesh_fixture(rm)
    s = ref.RefSolver(rm, True, viscosity)
    s.set_bcs(bcs)
    s.set_options(second_order, gradient, 5.0, cfl)
    s.init(); s.refill_bcs()
    q0 = perturb(s.get("q"), rm.N)
    s.set("q", q0)
    d["q0"] = q0
    d["rhs_norm"] = np.array(s.implicit_rhs())
    d["rhs"] = s.get("rhs")
    d["q_after_rhs"] = s.get("q")
    dg, o01, o10 = s.implicit_lhs()
    d["sha_diag"] = np.array(sha(dg)); d["sha_off01"] = np.array(sha(o01)); d["sha_off10"] = np.array(sha(o10))
    # a deterministic sample of blocks in full
    rng = np.random.default_rng(7)
    ci = np.sort(rng.choice(rm.N + rm.G, 64, replace=False)); ei = np.sort(rng.choice(rm.E, 64, replace=False))
    ei = np.unique(np.concatenate([ei, rm.bnd_edge[:8].astype(np.int64)]))
    d["diag_idx"] = ci; d["diag_blk"] = dg[ci]
    d["edge_idx"] = ei; d["off01_blk"] = o01[ei]; d["off10_blk"] = o10[ei]
    d["meta"] = np.array(repr(dict(mesh=mesh_file, bcs=bcs, viscosity=viscosity, gradient=gradient, second_order=second_order, cfl=cfl)))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **d)
    print(tag, "rhs norm", d["rhs_norm"])


def face_cases():
    """Single-face vectors: flux, BC ghost states and FD Jacobians of all four flux classes."""
    rng = np.random.default_rng(2024)
    g5 = ref.gas5()
    rows = []
    qfar = ref.get_conservative(0.2, 0.05, 1.0, 1.0, g5)
    for kind in (0, 1, 2, 3):
        for visc in (0, 1):
            for rep in range(12):
                th = rng.uniform(0, 2 * np.pi)
                nx, ny = np.cos(th), np.sin(th)
                mach = [0.2, 0.6, 1.4, 2.5][rep % 4]  # subsonic and supersonic, in- and outflow by the random normal
                ang = rng.uniform(-0.5, 0.5)
                qL = ref.get_conservative(mach, ang, rng.uniform(0.8, 1.2), rng.uniform(0.8, 1.2), g5)
                qL *= 1 + 1e-2 * rng.uniform(-1, 1, 4)
                if kind == 0:
                    qR = qL * (1 + 5e-2 * rng.uniform(-1, 1, 4))
                    if rep % 3 == 0:  # entropy-fix branch: normal velocity near zero / sonic
                        qR = qL.copy(); qR[1:3] += 1e-3 * rng.uniform(-1, 1, 2)
                else:
                    qR = qfar * (1 + 1e-2 * rng.uniform(-1, 1, 4))
                gx = rng.uniform(-1, 1, 4); gy = rng.uniform(-1, 1, 4)
                f = ref.flux(kind, g5, visc, nx, ny, qL, qR, gx, gy)
                v = ref.bc_vars(kind, g5, nx, ny, qL, qR)
                J = ref.fd_jacobian(kind, g5, visc, nx, ny, qL, qR, gx, gy)
                rows.append(np.concatenate([[kind, visc, nx, ny], qL, qR, gx, gy, f, v, J.ravel()]))
    rows = np.array(rows)
    cons = np.array([np.concatenate([[m, a, T, p], ref.get_conservative(m, a, T, p, g5)])
                     for (m, a, T, p) in [(0.2, 0.01745, 1, 1), (0.8, -0.2, 1.3, 0.7), (2.0, 1.0, 0.9, 2.0), (0.2, 420.0, 1, 1)]])
    np.savez_compressed(os.path.join(OUT, "faces.npz"), rows=rows, gas5=g5, conservative=cons,
                        layout=np.array("kind visc nx ny | qL[4] qR[4] gx[4] gy[4] | flux[4] vars[4] J[64 row-major 8x8]"))
    print("faces", rows.shape)


def sweep_case():
    """Converged CL/CD/CM of the reference's own implicit FMG sweep (conf.ini settings, rans.h:78-106).
    The implicit linear solver is the stand-in's GMRES/ILUT, so only the converged state is meaningful:
    the run is driven to a tight tolerance."""
    r = ref.run_sweep([REF_MESHES + "naca0012q_coarse.msh"], [1.0, 4.0], implicit=True, tolerance=1e-10, max_iterations=400)
    np.savez_compressed(os.path.join(OUT, "sweep_naca0012q_coarse.npz"), alphas=np.array([1.0, 4.0]), cl=r["cl"], cd=r["cd"], cm=r["cm"],
                        iters=r["iters"], meta=np.array("implicit, inviscid, green-gauss, second order, slip-wall, M=0.2, tol 1e-10, coarse mesh only"))
    print("sweep", r)


def prolong_case():
    """FMG prolongation of the reference (multigrid.h:100-178) between two synthetic O-meshes written by our MSH writer."""
    import tempfile
    import aeroflex_b200 as afx
    dims = ((48, 24, 8), (96, 48, 16))
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for k, (ni, nj, nq) in enumerate(dims):
            m = afx.Mesh.synth_omesh(ni, nj, nq, 60.0)
            paths.append(os.path.join(td, "m%d.msh" % k))
            m.write_msh(paths[-1])
        mc = ref.RefMesh(paths[0]); mf = ref.RefMesh(paths[1])
        rng = np.random.default_rng(99)
        qc = rng.uniform(0.5, 1.5, 4 * (mc.N + mc.G))
        qf = ref.prolongate(paths[0], paths[1], qc, 4 * (mf.N + mf.G))
    np.savez_compressed(os.path.join(OUT, "prolongation.npz"), dims=np.array(dims), far_radius=np.array(60.0), q_coarse=qc, q_fine=qf)
    print("prolongation", qf.shape, float(np.abs(qf).max()))


def mesh_only(tag, mesh_file):
    rm = ref.RefMesh(REF_MESHES + mesh_file)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **mesh_fixture(rm))
    print(tag, rm.N, rm.G, rm.E)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        prolong_case()
        mesh_only("naca0012q_mid_mesh", "naca0012q_mid.msh")
        sys.exit(0)
    slip = {"farfield": ("farfield", FAR), "wall": ("slip-wall", None)}
    wall = {"farfield": ("farfield", FAR), "wall": ("wall", None)}
    face_cases()
    explicit_case("naca0012q_coarse_euler_gg_o2", "naca0012q_coarse.msh", slip, "inviscid", "green-gauss", True, 100)
    explicit_case("naca0012_coarse_laminar_lsq_o2", "naca0012_coarse.msh", wall, "laminar", "least-squares", True, 30, full=False)
    explicit_case("naca0012_coarse_sa_gg_o1", "naca0012_coarse.msh", wall, "spallart-allmaras", "green-gauss", False, 30, full=False)
    explicit_case("naca0012_coarse_euler_gg_o1", "naca0012_coarse.msh", slip, "inviscid", "green-gauss", False, 30, full=False)
    plate = {"top": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "left": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)),
             "right": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "bot0": ("slip-wall", None), "bot1": ("wall", None)}
    # 5 patches incl. a no-slip wall; cells down to 1e-13 in area, so the laminar run needs a tiny CFL to stay
    # finite explicitly (dt has no viscous limit in the reference) and starts from the uniform state
    explicit_case("flat_plate_laminar_gg_o2", "flat_plate.msh", plate, "laminar", "green-gauss", True, 20, full=False, amp=0.0, cfl=1e-4)
    explicit_case("flat_plate_sa_gg_o2", "flat_plate.msh", plate, "spallart-allmaras", "green-gauss", True, 20, full=False, amp=1e-4, cfl=1.0)
    implicit_case("naca0012q_coarse_implicit_blocks", "naca0012q_coarse.msh", slip, "inviscid", "green-gauss", True)
    implicit_case("naca0012_coarse_implicit_laminar_blocks", "naca0012_coarse.msh", wall, "laminar", "green-gauss", True)
    sweep_case()
    prolong_case()
    mesh_only("naca0012q_mid_mesh", "naca0012q_mid.msh")
    # tests/golden/sweep_naca0012q_fmg.npz (coarse -> mid FMG at tolerance 1e-10, ~10 minutes of CPU) is made by
    # oracle/make_golden_fmg.py
