"""The restated oracle against the frozen outputs of the unmodified reference
(tests/golden, made by oracle/make_golden.py).  CPU only.  Integer/index data
and, since neither side contracts FMAs, every floating-point vector are
compared bit for bit; norms (whose summation order inside Eigen is not
specified) to 1e-13 relative."""
import numpy as np
import pytest

from oracle import orc
from tests import helpers as H


def test_face_vectors_bit_exact():
    d = H.load("faces")
    g = orc.Gas(*d["gas5"])
    for r in d["rows"]:
        kind, visc, nx, ny = int(r[0]), int(r[1]), r[2], r[3]
        qL, qR, gx, gy = r[4:8], r[8:12], r[12:16], r[16:20]
        f, v, J = r[20:24], r[24:28], r[28:92].reshape(8, 8)
        assert np.array_equal(orc.flux(kind, g, visc, nx, ny, qL, qR, gx, gy), f)
        assert np.array_equal(orc.bc_vars(kind, g, nx, ny, qL, qR), v)
        assert np.array_equal(orc.fd_jacobian(kind, g, visc, nx, ny, qL, qR, gx, gy), J)
    for c in d["conservative"]:
        assert np.array_equal(orc.get_conservative(c[0], c[1], c[2], c[3], g), c[4:8])


@pytest.mark.parametrize("tag", H.EXPLICIT_CASES + H.IMPLICIT_CASES)
def test_mesh_arrays_match_reference(tag):
    d = H.load(tag)
    m = H.oracle_mesh(d)
    assert (m.N, m.G, m.E) == tuple(int(v) for v in d["sizes"])
    for a in H.MESH_ARRAYS:
        assert H.sha(getattr(m, a)) == str(d["sha_" + a]), a


@pytest.mark.parametrize("tag", H.EXPLICIT_CASES)
def test_explicit_history_matches_reference(tag):
    d = H.load(tag)
    meta = d["meta"]
    m = H.oracle_mesh(d)
    s = orc.OracleSolver(m, viscosity=meta["viscosity"])
    H.setup_solver(s, meta)
    assert s.uniform_residual() == pytest.approx(float(d["uniform_residual_fresh"]), rel=1e-13)
    s.q[:] = d["q0"]
    norms = np.zeros(meta["n_iter"])
    for it in range(meta["n_iter"]):
        norms[it] = s.explicit_solve(meta["relax"])
        if it == 0:
            for nm, attr in (("q", "q"), ("qW", "qW"), ("gx", "gx"), ("gy", "gy"), ("limiters", "lim")):
                assert H.sha(getattr(s, attr)) == str(d["sha_it1_" + nm]), nm
                if "it1_" + nm in d:
                    assert np.array_equal(getattr(s, attr), d["it1_" + nm])
            assert H.sha(s.dt[:m.N]) == str(d["sha_it1_dt"])
    assert np.all(np.isfinite(norms))
    np.testing.assert_allclose(norms, d["norms"], rtol=1e-13, atol=0)
    assert H.sha(s.q) == str(d["sha_qN"])
    np.testing.assert_allclose(s.wall_forces(str(d["forces_patch"])), d["forces"], rtol=1e-14, atol=1e-16)


@pytest.mark.parametrize("tag", H.MICHALAK_CASES)
def test_michalak_limiter_history_matches_the_reference_macro_build(tag):
    """calc_limiters under RANS_MICHALAK_LIMITER (solver.h:557-576, physics.h:581-592): the restatement against the unmodified
    headers compiled with that macro -- limiters, states and residuals bit for bit over 30 iterations, with 11-14 % of the limiter
    values below one (the smooth switch and both branches of the cubic are exercised)."""
    d = H.load(tag)
    meta = d["meta"]
    m = H.oracle_mesh(d)
    s = orc.OracleSolver(m, viscosity=meta["viscosity"])
    H.setup_solver(s, meta)
    s.init(); s.refill_bcs()
    s.q[:] = d["q0"]
    norms = np.zeros(meta["n_iter"])
    for it in range(meta["n_iter"]):
        norms[it] = s.explicit_solve(meta["relax"])
        if it == 0:
            for nm, attr in (("q", "q"), ("qW", "qW"), ("limiters", "lim")):
                assert np.array_equal(getattr(s, attr) + 0.0, d["it1_" + nm] + 0.0), nm
    l1 = d["it1_limiters"][:4 * m.N]
    assert 0.05 < np.mean(l1 < 1) < 0.5 and l1.min() == 0.0 and np.any((l1 > 0) & (l1 < 1))
    np.testing.assert_allclose(norms, d["norms"], rtol=1e-13, atol=0)
    assert np.array_equal(s.q, d["qN"]) and np.array_equal(s.lim + 0.0, d["limN"] + 0.0)
    np.testing.assert_allclose(s.wall_forces(str(d["forces_patch"])), d["forces"], rtol=1e-14, atol=1e-16)
    # and it is a different function: the default build's limiter on the same state differs
    s.set_limiter("venkatakrishnan")
    s.calc_limiters(0)
    assert not np.array_equal(s.lim, d["limN"])


@pytest.mark.parametrize("tag", H.IMPLICIT_CASES)
def test_implicit_rhs_and_jacobian_blocks_match_reference(tag):
    d = H.load(tag)
    meta = d["meta"]
    m = H.oracle_mesh(d)
    s = orc.OracleSolver(m, viscosity=meta["viscosity"])
    H.setup_solver(s, meta)
    s.q[:] = d["q0"]
    nrm = s.implicit_rhs()
    assert nrm == pytest.approx(float(d["rhs_norm"]), rel=1e-13)
    assert np.array_equal(s.rhs, d["rhs"])
    assert np.array_equal(s.q, d["q_after_rhs"])
    dg, o01, o10 = s.implicit_lhs()
    assert H.sha(dg) == str(d["sha_diag"])
    assert H.sha(o01) == str(d["sha_off01"])
    assert H.sha(o10) == str(d["sha_off10"])
    assert np.array_equal(dg[d["diag_idx"]], d["diag_blk"])
    assert np.array_equal(o01[d["edge_idx"]], d["off01_blk"])


def test_sanity_anchors_independent_of_any_oracle():
    """uniform free stream on a closed mesh telescopes to zero on interior cells."""
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.oracle_mesh(d)
    s = orc.OracleSolver(m)
    H.setup_solver(s, d["meta"])
    s.init(); s.refill_bcs()
    s.s.second_order = 0
    s.calc_residual(0)
    qW = s.qW.reshape(-1, 4)[:m.N]
    touches_bnd = np.zeros(m.N, bool)
    touches_bnd[m.edge_cells[m.bnd_edge, 0]] = True
    assert np.abs(qW[~touches_bnd]).max() < 1e-9 * np.abs(qW).max() + 1e-9


@pytest.mark.parametrize("tag", ["naca0012q_coarse_euler_gg_o2", "naca0012_coarse_laminar_lsq_o2", "naca0012_coarse_sa_gg_o1", "michalak_naca0012q_coarse_euler_gg"])
def test_threaded_port_is_bit_identical_to_the_serial_restatement(tag):
    """orc_explicit_solve_omp (the multi-core CPU baseline of bench.py) = orc_explicit_solve, bit for bit."""
    d = H.load(tag)
    meta = d["meta"]
    m = H.oracle_mesh(d, fast=True)
    a = orc.OracleSolver(m, viscosity=meta["viscosity"], fast=True); b = orc.OracleSolver(m, viscosity=meta["viscosity"], fast=True)
    for s in (a, b):
        H.setup_solver(s, meta)
        s.q[:] = d["q0"]
    na = [a.explicit_solve(meta["relax"]) for _ in range(5)]
    nb = [b.explicit_solve_omp(meta["relax"]) for _ in range(5)]
    assert np.array_equal(a.q, b.q) and np.array_equal(a.qW, b.qW)
    np.testing.assert_allclose(na, nb, rtol=1e-14)


@pytest.mark.parametrize("tag,typ", [("inlet_outlet", "inlet-outlet"), ("unknown", "something-else")])
def test_boundary_variables_search_matches_reference(tag, typ):
    """solver.h:597-611: the far-field state is that of the first "farfield" boundary edge, unless an edge of the literal
    type "inlet-outlet" comes first (then the defaults).  Golden vectors from the unmodified reference headers
    (oracle/make_golden_bc_quirks.py)."""
    g = H.load("bc_quirks")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    o = orc.OracleSolver(H.oracle_mesh(d))
    o.set_bcs({"farfield": (typ, None), "wall": ("farfield", dict(mach=0.3, angle=0.05, T=1.0, p=1.0))})
    o.set_options(True, "green-gauss", 5.0, 1.2); o.init(); o.refill_bcs()
    assert np.array_equal(o.q, g[tag + "_q_init"])
    assert o.uniform_residual() == float(g[tag + "_uniform_residual"])
    norms = [o.explicit_solve(0.9) for _ in range(3)]
    assert np.array_equal(norms, g[tag + "_norms"]) and np.array_equal(o.q, g[tag + "_q"])


REGIMES = {"transonic_slip": (0.85, 0.03, "slip-wall", "inviscid"), "supersonic_slip": (1.6, 0.05, "slip-wall", "inviscid"),
           "supersonic_wall": (1.3, -0.1, "wall", "spallart-allmaras")}


def regime_start(z, tag):
    """settings and start state of oracle/make_golden_regimes.py for an OracleSolver or a GpuSolver"""
    mach, angle, wall, _ = REGIMES[tag]
    z.set_bcs({"farfield": ("farfield", dict(mach=mach, angle=angle, T=1.0, p=1.0)), "wall": (wall, None)})
    z.set_options(True, "green-gauss", 5.0, 0.8); z.init(); z.refill_bcs()
    q = (z.get_q() if hasattr(z, "get_q") else z.q).copy()
    rng = np.random.default_rng(2024)
    q[:4 * 4096] *= 1 + 1e-3 * rng.uniform(-1, 1, 4 * 4096)
    return q


@pytest.mark.parametrize("tag", sorted(REGIMES))
def test_transonic_and_supersonic_histories_match_reference(tag):
    """Entropy-fix branch of the Roe flux and the supersonic in/outflow branches of the far-field state, 12 iterations of the
    unmodified reference headers (oracle/make_golden_regimes.py): norms, final state and limiters bit for bit."""
    g = H.load("regimes")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    o = orc.OracleSolver(H.oracle_mesh(d), viscosity=REGIMES[tag][3])
    o.q[:] = regime_start(o, tag)
    norms = [o.explicit_solve(0.9) for _ in range(12)]
    assert np.array_equal(norms, g[tag + "_norms"])
    assert np.array_equal(o.q[:64], g[tag + "_q_head"])
    assert H.sha(o.q) == str(g[tag + "_q_sha256"])
    assert H.sha(o.lim[:4 * 4096]) == str(g[tag + "_lim_sha256"])


def test_oracle_follows_the_converged_reference_run():
    """tests/golden/converged_naca0012q_coarse_explicit.npz (the unmodified reference, ~90 000 explicit iterations to 1e-13): the
    oracle reproduces the first checkpoints of its residual history exactly (the whole depth is run by the GPU tests)."""
    import os
    if not os.path.exists(os.path.join(H.GOLDEN, "converged_naca0012q_coarse_explicit.npz")):
        pytest.skip("fixture not generated")
    g = H.load("converged_naca0012q_coarse_explicit")
    meta = g["meta"]
    o = orc.OracleSolver(H.oracle_mesh(g), viscosity=meta["viscosity"])
    o.set_bcs(meta["bcs"]); o.set_options(meta["second_order"], meta["gradient"], 5.0, meta["cfl"]); o.init(); o.refill_bcs()
    every = int(g["every"])
    for k in range(2):
        for _ in range(every):
            n = o.explicit_solve(meta["relax"])
        assert n == g["norms_every"][k]
