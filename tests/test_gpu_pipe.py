"""The pipelined stage kernel k_pipe (rans_pipe.cuh): one persistent kernel per Runge-Kutta stage that sweeps the mesh in
chunks with the limiter, face-flux and gather/update phases a few chunks apart (hand-overs inside the L2).  It calls the
three-kernel stage's own device functions, so in BOTH arithmetic modes the states must be bit-identical to the
three-kernel stage -- for every chunk size and phase distance, with first-order and least-squares settings, on
partitioned meshes -- and, through it, to the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

BCS = {"farfield": ("farfield", dict(mach=0.2, angle=0.03, T=1.0, p=1.0)), "wall": ("wall", None)}


def _make(afx, mesh, pipe, math="strict", visc="spallart-allmaras", grad="green-gauss", so=True, seed=5):
    s = afx.GpuSolver(mesh, viscosity=visc, math=math)
    s.set_pipelined(pipe)
    s.set_bcs(BCS); s.set_options(so, grad, 5.0, 1.4); s.init(); s.refill_bcs()
    q0 = s.get_q()
    rng = np.random.default_rng(seed)
    q0[:4 * mesh.N] *= 1 + 1e-3 * rng.uniform(-1, 1, 4 * mesh.N)
    s.set_q(q0)
    return s


@pytest.mark.parametrize("shift,lag_f,lag_u", [(8, 1, 0), (9, 1, 1), (10, 2, 2), (12, 3, 1), (15, 2, 2)])
@pytest.mark.parametrize("math", ["strict", "fast"])
def test_pipelined_stage_is_bit_identical_to_three_kernel_stage(afx, gpu, monkeypatch, shift, lag_f, lag_u, math):
    monkeypatch.setenv("AFX_PIPE_SHIFT", str(shift)); monkeypatch.setenv("AFX_PIPE_LAGF", str(lag_f)); monkeypatch.setenv("AFX_PIPE_LAGU", str(lag_u))
    mesh = afx.Mesh.synth_omesh(128, 80, 32, 150.0)  # 16 384 mixed cells: 64 chunks at shift 8, one chunk at shift 15
    a = _make(afx, mesh, 0, math); b = _make(afx, mesh, 1, math)
    info = b.pipe_info()
    assert info["active"] and info["chunk_cells"] == 1 << shift and (info["lag_flux"], info["lag_update"]) == (lag_f, lag_u)
    assert not a.pipe_info()["active"]
    if shift <= 10:
        assert info["far_faces"] > 0 and info["far_cells"] > 0  # the far pass is exercised
    na = a.run(12, 0.9); nb = b.run(12, 0.9)
    assert np.array_equal(a.get_q(), b.get_q())
    for f in ("qW", "limiters", "gx", "gy", "dt"):
        assert np.array_equal(a.get(f), b.get(f)), f
    np.testing.assert_allclose(nb, na, rtol=1e-13)  # the norm's partial sums are grouped per work item, not per CTA
    assert b.launch_count() < a.launch_count()       # 4 kernels per iteration instead of 9


@pytest.mark.parametrize("so,grad", [(False, "green-gauss"), (True, "least-squares")])
def test_pipelined_stage_variants(afx, gpu, monkeypatch, so, grad):
    monkeypatch.setenv("AFX_PIPE_SHIFT", "9")
    mesh = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    a = _make(afx, mesh, 0, "strict", "inviscid", grad, so); b = _make(afx, mesh, 1, "strict", "inviscid", grad, so)
    assert b.pipe_info()["active"]
    na = a.run(8, 0.9); nb = b.run(8, 0.9)
    assert np.array_equal(a.get_q(), b.get_q())
    np.testing.assert_allclose(nb, na, rtol=1e-13)


def test_pipelined_stage_is_not_used_for_laminar_runs(afx, gpu, monkeypatch):
    """The laminar face gradient reads the iteration-start state of both cells while the last stage writes it in place: the
    host keeps the three-kernel stage there."""
    monkeypatch.setenv("AFX_PIPE_SHIFT", "9")
    mesh = afx.Mesh.synth_omesh(64, 40, 16, 150.0)
    s = afx.GpuSolver(mesh, viscosity="laminar", math="strict")
    s.set_pipelined(1)
    assert not s.pipe_info()["active"]


def test_pipelined_stage_matches_the_reference_history(afx, gpu, monkeypatch):
    """Golden fixture of the unmodified reference (naca0012q_coarse, Euler, Green-Gauss, 2nd order): state after one
    iteration bit-identical, residual history to 1e-12 (strict)."""
    monkeypatch.setenv("AFX_PIPE_SHIFT", "8")
    g = H.load("naca0012q_coarse_euler_gg_o2")
    meta = g["meta"]
    mesh = H.product_mesh(afx, g)
    s = afx.GpuSolver(mesh, viscosity=meta["viscosity"], math="strict")
    s.set_pipelined(1)
    H.setup_solver(s, meta)
    assert s.pipe_info()["active"]
    s.set_q(g["q0"])
    first = s.solve(meta["relax"])
    for nm in ("q", "qW", "gx", "gy", "limiters"):
        assert H.sha(s.get(nm)) == str(g["sha_it1_" + nm]), nm
    norms = np.concatenate([[first], s.run(meta["n_iter"] - 1, meta["relax"])])
    np.testing.assert_allclose(norms, g["norms"], rtol=1e-12, atol=0)
    assert H.sha(s.get_q()) == str(g["sha_qN"])
    np.testing.assert_allclose(s.wall_forces(str(g["forces_patch"])), g["forces"], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("world,halo", [(2, "p2p"), (3, "staged")])
def test_pipelined_stage_on_partitions(afx, gpu, monkeypatch, world, halo):
    """Partitioned: the front cells touch ring-1 cells, which sit in the last chunks -> they are far cells; the halo push
    happens from the far pass.  Bit-identical to the single-GPU three-kernel run."""
    monkeypatch.setenv("AFX_PIPE_SHIFT", "10"); monkeypatch.setenv("AFX_PIPE", "1")
    mesh = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    one = _make(afx, mesh, 0, "strict", seed=77)
    q0 = one.get_q()
    ref_norms = one.run(10, 0.9)
    Q = one.get_q().reshape(-1, 4)
    group = afx.Group(world)
    parts = [afx.Partition(mesh, world, r) for r in range(world)]
    solvers = [None] * world

    def make(r):
        def f():
            solvers[r] = afx.GpuSolver(parts[r], viscosity="spallart-allmaras", math="strict", device=r % gpu, group=group)
        return f
    afx.run_ranks([make(r) for r in range(world)])
    if halo == "p2p":
        blobs = [s.p2p_export() for s in solvers]
        for s in solvers:
            s.p2p_connect(blobs)

    def work(r):
        def f():
            s = solvers[r]
            s.set_bcs(BCS); s.set_options(True, "green-gauss", 5.0, 1.4); s.init(); s.refill_bcs()
            assert s.pipe_info()["active"]
            s.set_q(q0)
            norms = s.run(10, 0.9)
            out = np.full(4 * (mesh.N + mesh.G), np.nan)
            s.get_q(out)
            return norms, out.reshape(-1, 4)
        return f
    res = afx.run_ranks([work(r) for r in range(world)])
    for r, (norms, out) in enumerate(res):
        own = parts[r].cell_l2g[:parts[r].n_own]
        assert np.array_equal(out[own], Q[own])
        np.testing.assert_allclose(norms, ref_norms, rtol=1e-12)
    del solvers[:]
