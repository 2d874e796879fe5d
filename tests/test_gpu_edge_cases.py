"""Edge cases of the hot path through the C ABI against the oracle: flow regimes that take the other branches of the
boundary states and of the Roe flux (supersonic in/outflow, transonic entropy fix), degenerate sizes (fewer cells than a
CTA, no renumbering), limiter extremes, residual-history ring wrap, and the error behaviour of a diverged state.

Strict mode is held to bit identity with the oracle (tests/helpers.py, oracle/rans_oracle.c), fast mode to the north-star
tolerances.  The same tests run on the B200 (`-m gpu`) and, through tests/test_kernel_emulation.py, against the kernel
sources under host emulation in the CPU run.
"""
import numpy as np
import pytest

from oracle import orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-12


def pair(afx, m, visc="inviscid", math="strict"):
    x, y, cells, b0, b1 = m.elements()
    om = orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names, fast=False)
    return afx.GpuSolver(m, viscosity=visc, math=math), orc.OracleSolver(om, viscosity=visc)


def start(s, o, bcs, so=True, grad="green-gauss", k=5.0, cfl=1.2, amp=1e-3, N=None):
    for z in (s, o):
        z.set_bcs(bcs); z.set_options(so, grad, k, cfl); z.init(); z.refill_bcs()
    q0 = H.synth_state(N, o.q.copy(), amp=amp)
    s.set_q(q0); o.q[:] = q0
    return q0


@pytest.mark.parametrize("mach,angle", [(0.85, 0.03), (1.6, 0.05), (1.2, -0.2)])
@pytest.mark.parametrize("wall", ["slip-wall", "wall"])
def test_transonic_and_supersonic_farfield_vs_oracle(afx, gpu, mach, angle, wall):
    """physics.h:446-530: supersonic inflow takes the far-field state, supersonic outflow the interior state; the
    transonic case walks through the entropy-fix branch of the Roe eigenvalues (physics.h:133-135)."""
    m = afx.Mesh.synth_omesh(96, 48, 16, 60.0)
    s, o = pair(afx, m, visc="spallart-allmaras" if wall == "wall" else "inviscid")
    bcs = {"farfield": ("farfield", dict(mach=mach, angle=angle, T=1.0, p=1.0)), "wall": (wall, None)}
    start(s, o, bcs, cfl=0.8, N=m.N)
    gn = s.run(6, 0.9)
    on = np.array([o.explicit_solve(0.9) for _ in range(6)])
    assert np.all(np.isfinite(on))
    np.testing.assert_allclose(gn, on, rtol=NORM_RTOL, atol=0)
    assert np.array_equal(s.get_q(), o.q)
    assert s.get_uniform_residual() == pytest.approx(o.uniform_residual(), rel=NORM_RTOL, abs=1e-13)
    assert s.residual() == pytest.approx(o.implicit_rhs(), rel=NORM_RTOL)
    assert np.array_equal(s.get("rhs"), o.rhs)


@pytest.mark.parametrize("ni,nj,nq", [(8, 4, 2), (8, 5, 0), (12, 6, 6), (16, 9, 3)])
def test_meshes_smaller_than_a_cta(afx, gpu, ni, nj, nq):
    """32 to 250 cells: one partly filled CTA per kernel, no Hilbert renumbering below 65 cells, all-quad / all-triangle /
    mixed; every cell is at most a few faces away from both boundaries."""
    m = afx.Mesh.synth_omesh(ni, nj, nq, 30.0)
    for grad in ("green-gauss", "least-squares"):
        s, o = pair(afx, m)
        bcs = {"farfield": ("farfield", dict(mach=0.3, angle=0.04, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
        start(s, o, bcs, grad=grad, cfl=1.0, N=m.N)
        gn = s.run(4, 0.9)
        on = np.array([o.explicit_solve(0.9) for _ in range(4)])
        np.testing.assert_allclose(gn, on, rtol=NORM_RTOL, atol=0)
        assert np.array_equal(s.get_q(), o.q)
        for nm, attr in (("gx", "gx"), ("gy", "gy"), ("limiters", "lim"), ("qW", "qW")):
            assert H.sha(s.get(nm)) == H.sha(getattr(o, attr)), nm


@pytest.mark.parametrize("k", [0.0, 1e-3, 1e6])
def test_limiter_constant_extremes(afx, gpu, k):
    """limiter_k = 0 turns the Venkatakrishnan function into the non-smooth limiter (K^3 a = 0), a huge one switches it off."""
    m = afx.Mesh.synth_omesh(64, 32, 8, 50.0)
    s, o = pair(afx, m)
    bcs = {"farfield": ("farfield", dict(mach=0.5, angle=0.02, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    start(s, o, bcs, k=k, cfl=1.0, amp=1e-2, N=m.N)
    gn = s.run(4, 0.9)
    on = np.array([o.explicit_solve(0.9) for _ in range(4)])
    np.testing.assert_allclose(gn, on, rtol=NORM_RTOL, atol=0)
    assert np.array_equal(s.get_q(), o.q)
    assert np.array_equal(s.get("limiters"), o.lim)
    f = afx.GpuSolver(m, math="fast")
    f.set_bcs(bcs); f.set_options(True, "green-gauss", k, 1.0); f.init(); f.refill_bcs()
    f.set_q(H.synth_state(m.N, f.get_q(), amp=1e-2))
    fn = f.run(4, 0.9)
    # the north-star tolerance, no slack (measured on the B200, scripts/fast_mode_deviation.py: norms 4e-16, states 2e-13)
    np.testing.assert_allclose(fn, on, rtol=1e-10, atol=0)
    np.testing.assert_allclose(f.get_q(), o.q, rtol=1e-10, atol=1e-13)


def test_uniform_flow_is_a_fixed_point_of_interior_cells(afx, gpu):
    """Sanity anchor that needs no oracle: on free-stream initial data every cell whose faces are all interior
    or far-field has a zero first-order residual (the face normals of a closed cell sum to zero)."""
    m = afx.Mesh.synth_omesh(64, 32, 8, 50.0)
    s = afx.GpuSolver(m, math="strict")
    s.set_bcs({"farfield": ("farfield", dict(mach=0.4, angle=0.1, T=1.0, p=1.0)), "wall": ("slip-wall", None)})
    s.set_options(False, "green-gauss", 5.0, 1.0); s.init(); s.refill_bcs()  # first order: no reconstruction across cells
    s.residual()
    rhs = s.get("rhs").reshape(-1, 4)[:m.N]
    wall = m.patch_names.index("wall")
    wall_edges = set(int(e) for e, p in zip(m.bnd_edge, m.bnd_patch) if p == wall)
    touches_wall = np.array([any(int(e) in wall_edges for e in ce[:3 if t else 4]) for ce, t in zip(m.cell_edges, m.is_tri)])
    scale = np.abs(s.get_q()[:4]).max() * m.elen.max()
    assert np.abs(rhs[~touches_wall]).max() <= 1e-12 * scale
    assert np.abs(rhs[touches_wall]).max() > 1e-6 * scale  # the slip wall turns the flow: pressure force only


def test_non_finite_state_is_a_numeric_error_not_a_crash(afx, gpu):
    """explicitSolver::solve returns a NaN norm to its caller, which stops (multigrid.h:214); the ABI reports
    AFX_ERR_NUMERIC and leaves the handle usable."""
    m = afx.Mesh.synth_omesh(32, 16, 4, 40.0)
    s = afx.GpuSolver(m, math="strict")
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.0); s.init(); s.refill_bcs()
    good = s.get_q()
    bad = good.copy(); bad[4 * 7] = np.nan
    s.set_q(bad)
    with pytest.raises(afx.AfxError) as e:
        s.run(1, 0.9)
    assert e.value.code == -3  # AFX_ERR_NUMERIC
    s.set_q(good)
    assert np.all(np.isfinite(s.run(2, 0.9)))


def test_residual_history_ring_wraps(afx, gpu):
    """The device keeps the norms of a call in a ring of 65 536 entries and run_explicit drains it in chunks of half a ring:
    one call of 40 000 iterations must return the same history and state as 25 000 + 15 000 (which wraps the ring index
    between the calls)."""
    m = afx.Mesh.synth_omesh(8, 4, 2, 30.0)
    bcs = {"farfield": ("farfield", dict(mach=0.3, angle=0.02, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    outs = []
    for split in ((40000,), (25000, 15000), (33000, 7000)):
        s = afx.GpuSolver(m, math="strict")
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.0); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        norms = np.concatenate([s.run(n, 0.9) for n in split])
        outs.append((norms, s.get_q()))
    for norms, q in outs[1:]:
        assert np.array_equal(norms, outs[0][0]) and np.array_equal(q, outs[0][1])
    assert np.all(np.isfinite(outs[0][0])) and outs[0][0][-1] < outs[0][0][0]


@pytest.mark.parametrize("tag,typ", [("inlet_outlet", "inlet-outlet"), ("unknown", "something-else")])
def test_boundary_variables_search_vs_reference_golden(afx, gpu, tag, typ):
    """solver.h:597-611 through the C ABI (AFX_BC_INLET_OUTLET): init(), the uniform-flow residual and the force rotation
    take the far-field state of the first far-field edge unless an "inlet-outlet" edge comes first."""
    g = H.load("bc_quirks")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    s = afx.GpuSolver(m, math="strict")
    s.set_bcs({"farfield": (typ, None), "wall": ("farfield", dict(mach=0.3, angle=0.05, T=1.0, p=1.0))})
    s.set_options(True, "green-gauss", 5.0, 1.2); s.init(); s.refill_bcs()
    assert np.array_equal(s.get_q(), g[tag + "_q_init"])
    assert s.get_uniform_residual() == pytest.approx(float(g[tag + "_uniform_residual"]), rel=NORM_RTOL)
    norms = s.run(3, 0.9)
    np.testing.assert_allclose(norms, g[tag + "_norms"], rtol=NORM_RTOL)
    assert np.array_equal(s.get_q(), g[tag + "_q"])


@pytest.mark.parametrize("tag", ["transonic_slip", "supersonic_slip", "supersonic_wall"])
def test_transonic_and_supersonic_histories_vs_reference_golden(afx, gpu, tag):
    """The same regimes against the reference's own histories (tests/golden/regimes.npz): strict mode to the bit, fast mode
    within the north-star tolerance."""
    from tests.test_oracle_golden import REGIMES, regime_start
    g = H.load("regimes")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    s = afx.GpuSolver(m, viscosity=REGIMES[tag][3], math="strict")
    s.set_q(regime_start(s, tag))
    norms = s.run(12, 0.9)
    np.testing.assert_allclose(norms, g[tag + "_norms"], rtol=NORM_RTOL, atol=0)
    assert H.sha(s.get_q()) == str(g[tag + "_q_sha256"])
    assert H.sha(s.get("limiters")[:4 * m.N]) == str(g[tag + "_lim_sha256"])
    f = afx.GpuSolver(m, viscosity=REGIMES[tag][3], math="fast")
    f.set_q(regime_start(f, tag))
    np.testing.assert_allclose(f.run(12, 0.9), g[tag + "_norms"], rtol=1e-10, atol=0)  # the north-star tolerance (measured: 1.4e-15)


def test_fast_mode_100_iterations_of_the_bench_configuration(afx, gpu):
    """The benchmarked arithmetic on the benchmarked settings (SA / no-slip wall, 2nd order, Green-Gauss, CFL 1.5, relaxation 0.9,
    perturbed free stream) for the north-star's 100 iterations, on the 65 536-cell mesh: residual norms within 1e-10 of strict
    mode (which is the oracle and the reference bit for bit, tests above and tests/test_gpu_parity.py), forces within 1e-8."""
    m = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}
    out = []
    for math in ("strict", "fast"):
        s = afx.GpuSolver(m, viscosity="spallart-allmaras", math=math)
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.5); s.init(); s.refill_bcs()
        s.set_q(H.synth_state(m.N, s.get_q()))
        out.append((s.run(100, 0.9), np.array(s.wall_forces("wall"))))
    np.testing.assert_allclose(out[1][0], out[0][0], rtol=1e-10, atol=0)
    np.testing.assert_allclose(out[1][1], out[0][1], rtol=1e-8, atol=0)


def test_sweep_goes_on_after_a_failed_linear_solve_like_run_airfoil(afx, gpu):
    """multigrid.h:227,285 + rans.h:92-104: when the linear solve of an angle fails (run_solver returns 1), run() hands back the
    solver of the level it ended on, run_airfoil takes ITS wall profile and goes on with the next angle.  A GMRES that is allowed
    one iteration at tolerance 1e-14 fails at once: every angle must still be filled in (forces of the untouched free-stream state on
    the coarse level, the failed iteration not counted, residual -1 / uniform-flow residual), status AFX_ERR_NUMERIC -- and with the
    default linear solver the same handles then converge, so a failure leaves nothing behind."""
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    s = afx.GpuSolver(m, math="strict")
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0)
    s.set_linear_solver(restart=2, max_iterations=1, tolerance=1e-14)
    r = s.sweep([1.0, 2.0, 3.0], implicit=True, tolerance=1e-4, max_iterations=5)
    assert r["status"] == -3
    assert list(r["iterations"]) == [0, 0, 0] and np.all(r["residual"] < 0) and np.all(np.isfinite(r["residual"]))
    assert np.all(np.isfinite(r["cl"])) and np.all(np.isfinite(r["cd"])) and np.all(np.isfinite(r["cm"]))
    # the state was never advanced: the forces are those of the free stream of the FIRST angle (init() ran once, rans.h:88), rotated
    # into each angle's wind axes by get_wall_profile
    ref = afx.GpuSolver(m, math="strict")
    for k, al in enumerate([1.0, 2.0, 3.0]):
        b = dict(bcs); b["farfield"] = ("farfield", dict(mach=0.2, angle=al * 0.01745, T=1.0, p=1.0))
        ref.set_bcs(b)
        if k == 0:
            ref.set_options(True, "green-gauss", 5.0, 40.0); ref.init()
        ref.refill_bcs()
        np.testing.assert_allclose([r["cl"][k], r["cd"][k], r["cm"][k]], ref.wall_forces("wall"), rtol=1e-12, atol=1e-15)
        assert r["residual"][k] == pytest.approx(-1.0 / ref.get_uniform_residual(), rel=1e-12)
    s.set_linear_solver()  # defaults: GMRES(30), <= 500 iterations, 1e-2, 4 sweeps
    r2 = s.sweep([3.0], implicit=True, tolerance=1e-4, max_iterations=100, reinit=False)
    assert r2["status"] == 0 and r2["residual"][0] <= 1e-4 and 0 < r2["iterations"][0] < 100
