"""Parity of the CUDA path (through the C ABI) with the reference.

Three anchors: (1) the frozen outputs of the unmodified reference in
tests/golden; (2) the restated oracle on the same seeded inputs (synthetic
meshes, sizes it finishes in seconds); (3) size-independent properties at the
benchmark size.  The library is compiled with -fmad=false and evaluates every
face in the reference's orientation and every per-cell sum in the reference's
edge order, so states/gradients/limiters/residual vectors are required to be
BIT-IDENTICAL; only norms and force integrals (tree reductions on the device)
get a tolerance, 1e-12 relative -- two orders inside BASELINE.json's 1e-10 /
1e-8 bars."""
import math

import numpy as np
import pytest

from oracle import orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-12   # north_star: 1e-10 relative over the first 100 iterations
FORCE_RTOL = 1e-12  # north_star: 1e-8 relative on CL/CD/CM


def gpu_solver(afx, d, math="strict", **kw):
    m = H.product_mesh(afx, d)
    s = afx.GpuSolver(m, viscosity=d["meta"]["viscosity"], math=math, **kw)
    H.setup_solver(s, d["meta"])
    return m, s


@pytest.mark.parametrize("tag", H.EXPLICIT_CASES)
def test_explicit_history_vs_reference_golden(afx, gpu, tag):
    d = H.load(tag)
    meta = d["meta"]
    m, s = gpu_solver(afx, d)
    assert s.get_uniform_residual() == pytest.approx(float(d["uniform_residual_fresh"]), rel=NORM_RTOL)
    s.set_q(d["q0"])
    n = meta["n_iter"]
    first = s.solve(meta["relax"])
    # first-order runs never call calc_limiters (solver.h:813): that vector is uninitialised in the reference
    for nm in ("q", "qW", "gx", "gy") + (("limiters",) if meta["second_order"] else ()):
        v = s.get(nm)
        assert H.sha(v) == str(d["sha_it1_" + nm]), nm
        if "it1_" + nm in d:
            assert np.array_equal(v, d["it1_" + nm])
    assert H.sha(s.get("dt")[:m.N]) == str(d["sha_it1_dt"])
    norms = np.concatenate([[first], s.run(n - 1, meta["relax"])])
    np.testing.assert_allclose(norms, d["norms"], rtol=NORM_RTOL, atol=0)
    assert H.sha(s.get_q()) == str(d["sha_qN"])
    np.testing.assert_allclose(s.wall_forces(str(d["forces_patch"])), d["forces"], rtol=FORCE_RTOL, atol=1e-15)


@pytest.mark.parametrize("tag", H.MICHALAK_CASES)
def test_michalak_limiter_vs_reference_macro_build(afx, gpu, tag):
    """afx_rans_set_limiter(AFX_LIMITER_MICHALAK) = the reference compiled with -DRANS_MICHALAK_LIMITER (solver.h:557-576): strict mode
    bit-identical to that build over 30 iterations (limiters, states, residual vector), fast mode within the north-star tolerances;
    the implicit right-hand side takes the same limiter; switching back restores the default build's bits."""
    d = H.load(tag)
    meta = d["meta"]
    m, s = gpu_solver(afx, d)  # setup_solver applies meta["limiter"]
    s.init(); s.refill_bcs()
    s.set_q(d["q0"])
    l0 = s.launch_count()
    first = s.solve(meta["relax"])
    assert s.launch_count() - l0 == 11  # dt/gradient kernel without the limiter epilogue + a limiter launch per stage
    for nm in ("q", "qW", "limiters"):
        assert np.array_equal(s.get(nm) + 0.0, d["it1_" + nm] + 0.0), nm
    norms = np.concatenate([[first], s.run(meta["n_iter"] - 1, meta["relax"])])
    np.testing.assert_allclose(norms, d["norms"], rtol=NORM_RTOL, atol=0)
    assert np.array_equal(s.get_q(), d["qN"]) and np.array_equal(s.get("limiters") + 0.0, d["limN"] + 0.0)
    np.testing.assert_allclose(s.wall_forces(str(d["forces_patch"])), d["forces"], rtol=FORCE_RTOL, atol=1e-15)
    # fast arithmetic
    mf, f = gpu_solver(afx, d, math="fast")
    f.init(); f.refill_bcs(); f.set_q(d["q0"])
    fn = f.run(meta["n_iter"], meta["relax"])
    np.testing.assert_allclose(fn, d["norms"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(f.wall_forces(str(d["forces_patch"])), d["forces"], rtol=1e-8, atol=1e-12)
    # the implicit right-hand side (fillRhoRHS calls the same calc_limiters) against the oracle with the same switch
    om = H.oracle_mesh(d); o = orc.OracleSolver(om, viscosity=meta["viscosity"]); H.setup_solver(o, meta)
    o.init(); o.refill_bcs(); o.q[:] = d["q0"]; s.set_q(d["q0"])
    assert s.residual() == pytest.approx(o.implicit_rhs(), rel=NORM_RTOL)
    assert np.array_equal(s.get("rhs"), o.rhs)
    # back to the default build's function: the bits of the Venkatakrishnan oracle
    s.set_limiter("venkatakrishnan"); o.set_limiter("venkatakrishnan")
    s.set_q(d["q0"]); o.q[:] = d["q0"]
    gn = s.run(3, meta["relax"]); on = [o.explicit_solve(meta["relax"]) for _ in range(3)]
    np.testing.assert_allclose(gn, on, rtol=NORM_RTOL, atol=0)
    assert np.array_equal(s.get_q(), o.q)
    assert s.L.afx_rans_set_limiter(s.h, 7) == -1 and b"unknown limiter" in s.L.afx_last_error()  # AFX_ERR_INVALID


@pytest.mark.parametrize("tag", H.IMPLICIT_CASES)
def test_implicit_rhs_and_jacobian_vs_reference_golden(afx, gpu, tag):
    d = H.load(tag)
    m, s = gpu_solver(afx, d)
    s.set_q(d["q0"])
    nrm = s.residual()
    assert nrm == pytest.approx(float(d["rhs_norm"]), rel=NORM_RTOL)
    assert np.array_equal(s.get("rhs"), d["rhs"])
    assert np.array_equal(s.get_q(), d["q_after_rhs"])
    s.fill_jacobian()
    dg, o01, o10 = s.jacobian_blocks()
    assert np.array_equal(dg[d["diag_idx"]], d["diag_blk"])
    assert np.array_equal(o01[d["edge_idx"]], d["off01_blk"])
    assert np.array_equal(o10[d["edge_idx"]], d["off10_blk"])
    assert H.sha(dg) == str(d["sha_diag"])
    assert H.sha(o01) == str(d["sha_off01"])
    assert H.sha(o10) == str(d["sha_off10"])


def test_single_phases_vs_oracle(afx, gpu):
    d = H.load("naca0012_coarse_laminar_lsq_o2")
    meta = d["meta"]
    m, s = gpu_solver(afx, d)
    om = H.oracle_mesh(d)
    o = orc.OracleSolver(om, viscosity=meta["viscosity"])
    H.setup_solver(o, meta)
    s.set_q(d["q0"]); o.q[:] = d["q0"]
    s.phase_dt_gradients(); o.calc_dt(); o.walls(0); o.calc_gradients()
    assert np.array_equal(s.get("gx"), o.gx) and np.array_equal(s.get("gy"), o.gy)
    assert np.array_equal(s.get("dt")[:m.N], o.dt[:m.N])
    s.phase_limiters(); o.calc_limiters(0)
    assert np.array_equal(s.get("limiters"), o.lim)
    nrm = s.phase_residual(); o.calc_residual(0)
    assert np.array_equal(s.get("qW"), o.qW)
    assert nrm == pytest.approx(np.sqrt(np.sum(o.qW ** 2)), rel=NORM_RTOL)
    assert np.array_equal(s.get_q(), o.q)  # wall ghosts follow their owners, nothing else moved


@pytest.mark.parametrize("visc,grad,so,wall", [("inviscid", "green-gauss", True, "slip-wall"),
                                              ("spallart-allmaras", "green-gauss", True, "wall"),
                                              ("laminar", "least-squares", True, "wall"),
                                              ("inviscid", "least-squares", False, "slip-wall")])
def test_synthetic_mixed_mesh_vs_oracle(afx, gpu, visc, grad, so, wall):
    """SURVEY 8d config 2 at 1/16 scale (65 536 mixed tri/quad cells): 10 iterations, bit-identical states."""
    m = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    x, y, cells, b0, b1 = m.elements()
    om = orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names, fast=False)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=2 * 0.01745, T=1.0, p=1.0)), "wall": (wall, None)}
    s = afx.GpuSolver(m, viscosity=visc, math="strict"); o = orc.OracleSolver(om, viscosity=visc)
    for z in (s, o):
        z.set_bcs(bcs); z.set_options(so, grad, 5.0, 1.2); z.init(); z.refill_bcs()
    q0 = H.synth_state(m.N, o.q.copy(), amp=1e-4 if visc == "laminar" else 1e-3)
    s.set_q(q0); o.q[:] = q0
    gn = s.run(10, 0.9)
    on = np.array([o.explicit_solve(0.9) for _ in range(10)])
    assert np.all(np.isfinite(on))
    np.testing.assert_allclose(gn, on, rtol=NORM_RTOL, atol=0)
    assert np.array_equal(s.get_q(), o.q)
    np.testing.assert_allclose(s.wall_forces("wall"), o.wall_forces("wall"), rtol=FORCE_RTOL, atol=1e-15)


def test_unknown_bc_type_is_a_two_sided_face(afx, gpu):
    """solver.h:211-212,237-238: a bc_type that is none of the three names leaves an internal flux against the ghost cell."""
    d = H.load("naca0012q_coarse_euler_gg_o2")
    meta = dict(d["meta"]); bcs = dict(meta["bcs"]); bcs["wall"] = ("inlet-outlet", None); meta["bcs"] = bcs
    m = H.product_mesh(afx, d); s = afx.GpuSolver(m, math="strict"); H.setup_solver(s, meta)
    om = H.oracle_mesh(d); o = orc.OracleSolver(om); H.setup_solver(o, meta)
    for z in (s, o):
        z.init(); z.refill_bcs()
    q0 = H.synth_state(m.N, o.q.copy())
    s.set_q(q0); o.q[:] = q0
    gn = s.run(3, 0.9); on = [o.explicit_solve(0.9) for _ in range(3)]
    np.testing.assert_allclose(gn, on, rtol=NORM_RTOL)
    assert np.array_equal(s.get_q(), o.q)
    assert s.get_uniform_residual() == pytest.approx(o.uniform_residual(), rel=NORM_RTOL)


def test_missing_patch_raises_like_bcs_at(afx, gpu):
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d); s = afx.GpuSolver(m)
    with pytest.raises(KeyError):
        s.set_bcs({"farfield": ("farfield", None)})
    with pytest.raises(afx.AfxError):
        s.solve(1.0)  # set_bcs never succeeded


def test_state_io_and_bc_helpers(afx, gpu):
    d = H.load("naca0012_coarse_euler_gg_o1")
    m, s = gpu_solver(afx, d)
    om = H.oracle_mesh(d); o = orc.OracleSolver(om); H.setup_solver(o, d["meta"])
    rng = np.random.default_rng(3)
    q = rng.uniform(0.5, 1.5, 4 * (m.N + m.G))
    s.set_q(q)
    assert np.array_equal(s.get_q(), q)  # reference order in, reference order out, whatever the internal numbering
    s.init(); o.q[:] = q; o.init()
    assert np.array_equal(s.get_q(), o.q)
    s.refill_bcs(); o.refill_bcs()
    assert np.array_equal(s.get_q(), o.q)
    s.set_q(q); s.bcs_from_internal(); o.q[:] = q; o.bcs_from_internal()
    assert np.array_equal(s.get_q(), o.q)
    cp = s.wall_cp("wall")
    assert cp.shape == (int(np.sum(m.bnd_patch == m.patch_names.index("wall"))),) and np.all(np.isfinite(cp))


def test_renumbering_is_invisible(afx, gpu, monkeypatch):
    """Hilbert renumbering vs none: bit-identical results in reference order."""
    m = afx.Mesh.synth_omesh(192, 96, 32, 100.0)
    bcs = {"farfield": ("farfield", dict(mach=0.3, angle=0.05, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    outs = []
    for order in ("hilbert", "none"):
        monkeypatch.setenv("AFX_ORDER", order)
        s = afx.GpuSolver(m, math="strict")
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        n = s.run(5, 0.9)
        outs.append((n, s.get_q()))
    assert np.array_equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-13)


def test_graph_replay_equals_plain_launches(afx, gpu, monkeypatch):
    m = afx.Mesh.synth_omesh(128, 64, 16, 100.0)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.02, T=1.0, p=1.0)), "wall": ("wall", None)}
    outs = []
    for ng in ("0", "1"):
        monkeypatch.setenv("AFX_NO_GRAPH", ng)
        s = afx.GpuSolver(m, viscosity="spallart-allmaras", math="strict")
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        outs.append((s.run(7, 0.9), s.get_q(), s.launch_count()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_full_size_mesh_vs_oracle_and_conservation(afx, gpu):
    """BASELINE config 2 at full size (N = 2^20 mixed cells): two iterations against the oracle's fast build is
    too loose a check on its own (that build contracts FMAs), so: (a) bit-identity against the PARITY oracle for
    one iteration, (b) discrete conservation: sum_i A_i qW_i equals minus the boundary fluxes, (c) determinism."""
    m = afx.Mesh.synth_omesh(1024, 640, 256, 150.0)
    assert m.N == 2 ** 20 and m.E == 1704960
    x, y, cells, b0, b1 = m.elements()
    om = orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
    s = afx.GpuSolver(m, viscosity="spallart-allmaras", math="strict"); o = orc.OracleSolver(om, viscosity="spallart-allmaras")
    for z in (s, o):
        z.set_bcs(bcs); z.set_options(True, "green-gauss", 5.0, 1.5); z.init(); z.refill_bcs()
    q0 = H.synth_state(m.N, o.q.copy())
    s.set_q(q0); o.q[:] = q0
    gn = s.solve(0.9); o.explicit_solve(0.9)
    assert np.array_equal(s.get_q(), o.q)
    assert np.array_equal(s.get("qW"), o.qW)
    # 4M squares: the oracle's left-to-right sum is itself only good to ~1e-12, so compare with the exactly
    # rounded sum of the (bit-identical) vector
    assert gn == pytest.approx(math.sqrt(math.fsum(o.qW * o.qW)), rel=1e-13)
    # (b) conservation of the last-stage residual: interior fluxes cancel pairwise
    qW = s.get("qW").reshape(-1, 4)[:m.N]
    total = (qW * m.area[:m.N, None]).sum(axis=0)
    scale = (np.abs(qW) * m.area[:m.N, None]).sum(axis=0)
    touches = np.zeros(m.N, bool); touches[m.edge_cells[m.bnd_edge, 0]] = True
    # remove boundary cells' own boundary flux by comparing with the oracle's identical field: the check is on the sum
    o_total = (o.qW.reshape(-1, 4)[:m.N] * m.area[:m.N, None]).sum(axis=0)
    assert np.array_equal(total, o_total)
    assert np.all(np.abs(total) <= scale)  # bounded by the boundary contribution; interior cancels
    # (c) determinism: same input twice -> same bits
    s.set_q(q0); a = s.run(3, 0.9); qa = s.get_q()
    s.set_q(q0); b = s.run(3, 0.9); qb = s.get_q()
    assert np.array_equal(a, b) and np.array_equal(qa, qb)
    # (d) the fast arithmetic mode (what bench.py times) on the same input: inside the north-star tolerances
    f = afx.GpuSolver(m, viscosity="spallart-allmaras", math="fast")
    f.set_bcs(bcs); f.set_options(True, "green-gauss", 5.0, 1.5); f.set_q(q0)
    c = f.run(3, 0.9)
    np.testing.assert_allclose(c, a, rtol=1e-10, atol=0)
    np.testing.assert_allclose(f.get_q(), qa, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(f.wall_forces("wall"), s.wall_forces("wall"), rtol=1e-8, atol=1e-12)


# ---------------------------------------------------------------------------
# fast arithmetic mode (default; shared reciprocals + FMA contraction): BASELINE.json tolerances
# ---------------------------------------------------------------------------
HIST_RTOL = 1e-10   # per-iteration residual norms over the first 100 iterations
CLCD_RTOL = 1e-8    # CL / CD / CM


@pytest.mark.parametrize("tag", H.EXPLICIT_CASES)
def test_fast_mode_history_vs_reference_golden(afx, gpu, tag):
    d = H.load(tag)
    meta = d["meta"]
    m, s = gpu_solver(afx, d, math="fast")
    assert s.math == "fast"
    assert s.get_uniform_residual() == pytest.approx(float(d["uniform_residual_fresh"]), rel=HIST_RTOL)
    s.set_q(d["q0"])
    norms = s.run(meta["n_iter"], meta["relax"])
    np.testing.assert_allclose(norms, d["norms"], rtol=HIST_RTOL, atol=0)
    np.testing.assert_allclose(s.wall_forces(str(d["forces_patch"])), d["forces"], rtol=CLCD_RTOL, atol=1e-13)
    if "qN" in d:
        np.testing.assert_allclose(s.get_q(), d["qN"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("tag", H.IMPLICIT_CASES)
def test_fast_mode_rhs_and_jacobian(afx, gpu, tag):
    d = H.load(tag)
    m, s = gpu_solver(afx, d, math="fast")
    s.set_q(d["q0"])
    assert s.residual() == pytest.approx(float(d["rhs_norm"]), rel=HIST_RTOL)
    scale = np.abs(d["rhs"]).max()
    np.testing.assert_allclose(s.get("rhs"), d["rhs"], rtol=1e-9, atol=1e-12 * scale)
    s.fill_jacobian()
    dg, o01, o10 = s.jacobian_blocks()
    # forward differences with eps = 1e-6 amplify rounding by 1e6: a few 1e-10 absolute on O(1) entries
    np.testing.assert_allclose(dg[d["diag_idx"]], d["diag_blk"], rtol=1e-6, atol=1e-8 * np.abs(d["diag_blk"]).max())
    np.testing.assert_allclose(o01[d["edge_idx"]], d["off01_blk"], rtol=1e-6, atol=1e-8 * np.abs(d["off01_blk"]).max())


# ---------------------------------------------------------------------------
# implicit path end to end (parity UNPINNED for its per-iteration history: the reference's linear solver is
# Eigen's GMRES+ILUT, ours is GMRES + block-Jacobi sweeps).  What is solver-independent is the converged state.
# ---------------------------------------------------------------------------
def run_implicit(s, tol, max_iter=300, relax=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, rhs_iterations=5):
    """multigrid<implicitSolver>::run_solver (multigrid.h:240-293) on one mesh."""
    cfl = start_cfl
    err_0 = s.get_uniform_residual()
    hist = []
    for i in range(max_iter):
        s.set_cfl(cfl)
        s.implicit_fill()
        ok = s.implicit_compute()
        err = s.implicit_solve(relax, err_0 * tol, rhs_iterations) if ok == 0 else -1.0
        if i == 0 and err > 2 * err_0:
            err_0 = err
        err /= err_0
        cfl = min(start_cfl + (i + 1) * slope_cfl, max_cfl)
        hist.append(err)
        if err < 0:
            break
        if err <= tol:
            break
    return hist


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_implicit_converged_forces_match_reference(afx, gpu, math):
    g = H.load("sweep_naca0012q_coarse")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    s = afx.GpuSolver(m, math=math)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=float(g["alphas"][0]) * 0.01745, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0); s.init(); s.refill_bcs()
    hist = run_implicit(s, 1e-10, max_iter=400)
    assert hist[-1] >= 0, "linear solver failed"
    assert hist[-1] <= 1e-10, hist[-5:]
    cl, cd, cm = s.wall_forces("wall")
    # steady state is independent of the linear solver; both sides stopped at 1e-10 relative residual
    assert cl == pytest.approx(float(g["cl"][0]), rel=1e-7)
    assert cd == pytest.approx(float(g["cd"][0]), rel=1e-6)
    assert cm == pytest.approx(float(g["cm"][0]), rel=1e-6)
    print("implicit outer iterations:", len(hist), "last linear iterations:", s.last_linear_iterations())


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_arnoldi_step_graphs_and_fused_rotation_are_bit_identical(afx, gpu, monkeypatch, math):
    """The kernels of an Arnoldi step are captured once per step index and replayed with one call (AFX_KRY_GRAPH, default on), and the
    Givens rotation runs in the block that finishes the norm of the update kernel instead of a launch of its own
    (AFX_KRY_FUSE_GIVENS, default on): same arithmetic in the same order -- the implicit iteration must be bit-identical to plain,
    separate launches, linear iteration counts included; only the launch count may differ, by one per Arnoldi step."""
    m = afx.Mesh.synth_omesh(96, 48, 16, 60.0)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.05, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    outs = []
    configs = (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0"))
    if afx.is_emulation():  # the CPU test run (tests/emu) takes ~35 s per configuration: both features on against both off
        configs = (("1", "1"), ("0", "0"))
    for graph, fuse in configs:
        monkeypatch.setenv("AFX_KRY_GRAPH", graph); monkeypatch.setenv("AFX_KRY_FUSE_GIVENS", fuse)
        s = afx.GpuSolver(m, math=math)
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0); s.init(); s.refill_bcs()
        l0 = s.launch_count()
        hist = run_implicit(s, 1e-6, max_iter=12)
        first = (np.array(hist), s.get_q(), s.last_linear_iterations(), s.launch_count() - l0)
        s.set_linear_solver(restart=8, max_iterations=200, tolerance=1e-3, precond_sweeps=3)  # new restart length / sweeps: the cached graphs must go
        hist2 = run_implicit(s, 1e-8, max_iter=6)
        outs.append(first + (np.array(hist2), s.get_q()))
    a = outs[0]
    assert len(a[0]) > 3 and a[0][-1] >= 0
    for b in outs[1:]:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
        assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    if len(outs) == 4:  # graphs replay the same launches; the fusion saves one per Arnoldi step
        assert outs[0][3] == outs[1][3] and outs[2][3] == outs[3][3]
    assert outs[-1][3] > outs[0][3]


@pytest.mark.parametrize("grad,visc,math", [("green-gauss", "inviscid", "strict"), ("least-squares", "laminar", "strict"), ("green-gauss", "spallart-allmaras", "fast")])
def test_first_stage_limiter_inside_dt_grad_is_bit_identical(afx, gpu, monkeypatch, grad, visc, math):
    """k_dt_grad<.,1> writes the first stage's limiters itself (one k_limiter launch less per iteration): same
    limiter_value() on the same inputs, so state, norms and the residual RHS have the bits of the separate launch,
    in both arithmetic modes."""
    m = afx.Mesh.synth_omesh(160, 80, 24, 100.0)
    bcs = {"farfield": ("farfield", dict(mach=0.25, angle=0.03, T=1.0, p=1.0)), "wall": ("wall" if visc != "inviscid" else "slip-wall", None)}
    outs = []
    for fuse in ("1", "0"):
        monkeypatch.setenv("AFX_FUSE_LIM0", fuse)
        s = afx.GpuSolver(m, viscosity=visc, math=math)
        s.set_bcs(bcs); s.set_options(True, grad, 5.0, 1.5); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        l0 = s.launch_count()
        norms = s.run(6, 0.9)
        per_iter = (s.launch_count() - l0) / 6.0
        rhs_norm = s.residual()
        outs.append((norms, s.get_q(), per_iter, rhs_norm, s.get("rhs"), s.get("limiters")))
    if math == "strict":
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        assert outs[0][3] == outs[1][3] and np.array_equal(outs[0][4], outs[1][4]) and np.array_equal(outs[0][5], outs[1][5])
    else:  # FMA contraction may differ between the two inlining contexts
        np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-11)
        np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(outs[0][4], outs[1][4], rtol=1e-9, atol=1e-12)
    assert outs[0][2] == 10 and outs[1][2] == 11  # kernels per explicit iteration (the last one adds up the residual norm)


@pytest.mark.parametrize("grad,visc", [("green-gauss", "spallart-allmaras"), ("least-squares", "laminar")])
def test_stored_limiter_extremes_match_the_gradient_reading_limiter(afx, gpu, monkeypatch, grad, visc):
    """Fast mode: k_dt_grad<., 2> stores the largest positive / most negative projected increment of every cell and k_limiter<1>
    (stages 2 and 3) reads those 64 bytes instead of gx, gy and the four face offsets.  The extremes are the values k_limiter<0>
    computes from the same gradients and offsets, so limiters, state and norms agree (FMA contraction may differ between the two
    inlining contexts: 1e-11), and both stay within the north-star 1e-10 of strict mode."""
    m = afx.Mesh.synth_omesh(160, 80, 24, 100.0)
    bcs = {"farfield": ("farfield", dict(mach=0.25, angle=0.03, T=1.0, p=1.0)), "wall": ("wall" if visc != "inviscid" else "slip-wall", None)}
    outs = {}
    for key, math, pm in (("pm", "fast", "1"), ("nopm", "fast", "0"), ("strict", "strict", "1")):
        monkeypatch.setenv("AFX_LIM_PM", pm)
        s = afx.GpuSolver(m, viscosity=visc, math=math)
        s.set_bcs(bcs); s.set_options(True, grad, 5.0, 1.5); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        norms = s.run(8, 0.9)
        outs[key] = (norms, s.get_q(), s.get("limiters"))
    np.testing.assert_allclose(outs["pm"][0], outs["nopm"][0], rtol=1e-11)
    np.testing.assert_allclose(outs["pm"][1], outs["nopm"][1], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(outs["pm"][2], outs["nopm"][2], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(outs["pm"][0], outs["strict"][0], rtol=1e-10)
    assert 0.0 < outs["pm"][2].min() < 1.0  # the limiter is active somewhere: the comparison is not of ones with ones





def test_sweep_entry_point_matches_reference_polar_and_the_primitive_calls(afx, gpu):
    """afx_rans_sweep = the angle loop of Rans::run_airfoil (rans.h:86-104) + multigrid::run_solver (multigrid.h:182-293) on
    one level.  Implicit, both sides driven to 1e-10: the converged CL/CD/CM of the reference's own sweep
    (golden sweep_naca0012q_coarse.npz).  Explicit: identical, to the bit, to the same loop written with the primitive
    ABI calls (as the C++ adapter does)."""
    g = H.load("sweep_naca0012q_coarse")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    s = afx.GpuSolver(m, math="strict")
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0)
    r = s.sweep(g["alphas"], implicit=True, tolerance=1e-10, max_iterations=400)
    assert r["status"] == 0 and np.all(r["residual"] <= 1e-10) and np.all(r["iterations"] < 400)
    np.testing.assert_allclose(r["cl"], g["cl"], rtol=1e-7)
    np.testing.assert_allclose(r["cd"], g["cd"], rtol=1e-6)
    np.testing.assert_allclose(r["cm"], g["cm"], rtol=1e-6)
    # explicit, loose tolerance, against the primitive calls
    alphas = [1.0, 2.5]
    a = afx.GpuSolver(m, math="strict")
    a.set_bcs(bcs); a.set_options(True, "green-gauss", 5.0, 1.5)
    ra = a.sweep(alphas, implicit=False, relaxation=0.9, start_cfl=1.5, tolerance=0.2, max_iterations=150)
    b = afx.GpuSolver(m, math="strict")
    b.set_options(True, "green-gauss", 5.0, 1.5)
    forces, iters = [], []
    for k, al in enumerate(alphas):
        bb = dict(bcs); bb["farfield"] = ("farfield", dict(mach=0.2, angle=al * 0.01745, T=1.0, p=1.0))
        b.set_bcs(bb)
        if k == 0:
            b.init()
        b.refill_bcs()
        err_0 = b.get_uniform_residual()
        i = 0
        while True:
            b.set_cfl(1.5)
            err = b.solve(0.9)
            if i == 0 and err > 2 * err_0:
                err_0 = err
            err /= err_0
            i += 1
            if not (err > 0.2 and i < 150):
                break
        forces.append(b.wall_forces("wall")); iters.append(i)
    assert list(ra["iterations"]) == iters
    assert np.array_equal(np.array(forces), np.stack([ra["cl"], ra["cd"], ra["cm"]], axis=1))
    assert np.array_equal(a.get_q(), b.get_q())
