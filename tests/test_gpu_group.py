"""Partitioned runs on ONE GPU: N partitioned solvers of this process on device 0 (in-process group, include/afx_rans.h),
one host thread each -- the partition plan, the far-field lookup of pieces without a far-field edge, the send / receive
indexing, the front-first cell order, the split ranges of k_dt_grad / k_limiter / k_flux and, in "p2p" mode, the
peer-memory push of k_gather_update with its flag hand-off and bounded wait all run on hardware on the 1-GPU box.
(NCCL refuses two ranks on one device, so tests/test_gpu_multi.py needs one GPU per rank and skips there.)
Strict mode: every rank's owned cells are bit-identical to the single-GPU run."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BCS = {"farfield": ("farfield", dict(mach=0.2, angle=0.03, T=1.0, p=1.0)), "wall": ("wall", None)}
DIMS = (256, 160, 64)  # 65 536 mixed cells


def _single(afx, mesh, visc, grad, so, n_iter, seed=77, amp=1e-3, limiter=None, limk=5.0):
    s = afx.GpuSolver(mesh, viscosity=visc, math="strict", device=0)
    s.set_bcs(BCS); s.set_options(so, grad, limk, 1.4); s.init(); s.refill_bcs()
    if limiter:
        s.set_limiter(limiter)
    q_init = s.get_q().reshape(-1, 4).copy()
    q0 = s.get_q()
    rng = np.random.default_rng(seed)
    q0[:4 * mesh.N] *= 1 + amp * rng.uniform(-1, 1, 4 * mesh.N)
    s.set_q(q0)
    ur = s.get_uniform_residual()
    norms = s.run(n_iter, 0.9)
    out = dict(q_init=q_init, q0=q0, ur=ur, norms=norms, Q=s.get_q().reshape(-1, 4).copy(), F=np.array(s.wall_forces("wall")),
               cp=s.wall_cp("wall"))
    out["rhs_norm"] = s.residual()
    out["RHS"] = s.get("rhs").reshape(-1, 4).copy()
    return out


def _group_run(afx, mesh, world, halo, math, visc, grad, so, n_iter, q0, ndev, fused=0, limiter=None, limk=5.0):
    group = afx.Group(world)
    parts = [afx.Partition(mesh, world, r) for r in range(world)]
    solvers = [None] * world

    def make(r):
        def f():
            solvers[r] = afx.GpuSolver(parts[r], viscosity=visc, math=math, device=r % ndev, group=group)
        return f
    afx.run_ranks([make(r) for r in range(world)])
    # Peer-memory mode on ONE device: the library detects that group members share a device and lets the ranks meet on the host
    # between raising and waiting for the flags (no kernel ever spins on another rank's kernel); with one device per rank -- the
    # production layout -- the same calls run the captured graph with the in-kernel wait.
    if halo == "p2p":
        blobs = [s.p2p_export() for s in solvers]
        for s in solvers:
            s.p2p_connect(blobs)
    for s in solvers:
        assert s.halo_mode() == ("p2p" if halo == "p2p" else "nccl")  # 1 = collective (NCCL or staged in-process) halo

    def work(r):
        def f():
            s, part = solvers[r], parts[r]
            s.set_bcs(BCS); s.set_options(so, grad, limk, 1.4); s.init(); s.refill_bcs()
            if limiter:
                s.set_limiter(limiter)
            qi = np.full(4 * (mesh.N + mesh.G), np.nan)
            s.get_q(qi)
            s.set_q(q0)
            ur = s.get_uniform_residual()
            norms = s.run(n_iter, 0.9)
            out = np.full(4 * (mesh.N + mesh.G), np.nan)
            s.get_q(out)
            forces = np.array(s.wall_forces("wall"))
            rhs_norm = s.residual()
            loc = s.get("rhs").reshape(-1, 4)
            return dict(q=out.reshape(-1, 4), q_init=qi.reshape(-1, 4), ur=ur, norms=norms, forces=forces, rhs_norm=rhs_norm,
                        rhs_own=loc[:part.n_own].copy(), own=part.cell_l2g[:part.n_own].copy(), launches=s.launch_count())
        return f
    res = afx.run_ranks([work(r) for r in range(world)])
    del solvers[:]
    return res, parts


CASES = [(2, "staged", "strict"), (2, "p2p", "strict"), (3, "p2p", "strict"), (4, "staged", "strict"), (8, "p2p", "strict"),
         (4, "p2p", "fast")]


@pytest.mark.parametrize("world,halo,math", CASES)
def test_partitions_on_one_gpu_match_single_gpu(afx, gpu, world, halo, math):
    n_iter = 25
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    ref = _single(afx, mesh, "spallart-allmaras", "green-gauss", True, n_iter)
    res, parts = _group_run(afx, mesh, world, halo, math, "spallart-allmaras", "green-gauss", True, n_iter, ref["q0"], gpu)
    seen = 0
    for d in res:
        own = d["own"]
        # a piece that holds no far-field edge still initialises with the far-field state of the whole mesh (solver.h:597-631)
        assert np.array_equal(d["q_init"][own], ref["q_init"][own])
        if math == "strict":
            assert np.array_equal(d["q"][own], ref["Q"][own])  # bit-identical to the single-GPU run (itself bit-identical to the oracle)
            np.testing.assert_allclose(d["norms"], ref["norms"], rtol=1e-12)
            np.testing.assert_allclose(d["forces"], ref["F"], rtol=1e-12, atol=1e-15)
            assert d["ur"] == pytest.approx(ref["ur"], rel=1e-12)
            assert np.array_equal(d["rhs_own"], ref["RHS"][own])
            assert d["rhs_norm"] == pytest.approx(ref["rhs_norm"], rel=1e-12)
        else:
            np.testing.assert_allclose(d["q"][own], ref["Q"][own], rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(d["norms"], ref["norms"], rtol=1e-10)
            np.testing.assert_allclose(d["forces"], ref["F"], rtol=1e-8, atol=1e-12)
        assert np.array_equal(d["norms"], res[0]["norms"]) and np.array_equal(d["forces"], res[0]["forces"])  # every rank sees the same numbers
        seen += len(own)
    assert seen == mesh.N


@pytest.mark.parametrize("world,visc,grad,so,halo", [(3, "laminar", "least-squares", True, "p2p"), (2, "laminar", "green-gauss", True, "staged"),
                                                      (3, "inviscid", "green-gauss", False, "p2p"), (5, "spallart-allmaras", "least-squares", True, "staged")])
def test_partition_variants_on_one_gpu(afx, gpu, world, visc, grad, so, halo):
    """Odd rank counts, the laminar face-gradient path, least-squares gradients and first-order runs, partitioned."""
    n_iter = 8
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    ref = _single(afx, mesh, visc, grad, so, n_iter, amp=1e-4 if visc == "laminar" else 1e-3)
    res, _ = _group_run(afx, mesh, world, halo, "strict", visc, grad, so, n_iter, ref["q0"], gpu)
    for d in res:
        own = d["own"]
        assert np.array_equal(d["q"][own], ref["Q"][own])
        np.testing.assert_allclose(d["norms"], ref["norms"], rtol=1e-12)
        np.testing.assert_allclose(d["forces"], ref["F"], rtol=1e-12, atol=1e-15)
        assert np.array_equal(d["rhs_own"], ref["RHS"][own])


@pytest.mark.parametrize("world,halo", [(3, "p2p"), (2, "staged")])
def test_michalak_limiter_partitioned_on_one_gpu(afx, gpu, world, halo):
    """The reference's other limiter (afx_rans_set_limiter, solver.h:557-576) on partitioned solvers: its kernel takes the same
    interior / halo-dependent cell ranges as k_limiter; same bits as the single-GPU run, implicit right-hand side included."""
    n_iter = 8
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    ref = _single(afx, mesh, "spallart-allmaras", "green-gauss", True, n_iter, amp=1e-2, limiter="michalak", limk=0.5)
    plain = _single(afx, mesh, "spallart-allmaras", "green-gauss", True, n_iter, amp=1e-2, limk=0.5)
    assert not np.array_equal(ref["Q"], plain["Q"])  # the limiter is active: the two functions give different states
    res, _ = _group_run(afx, mesh, world, halo, "strict", "spallart-allmaras", "green-gauss", True, n_iter, ref["q0"], gpu, limiter="michalak", limk=0.5)
    for d in res:
        own = d["own"]
        assert np.array_equal(d["q"][own], ref["Q"][own])
        np.testing.assert_allclose(d["norms"], ref["norms"], rtol=1e-12)
        np.testing.assert_allclose(d["forces"], ref["F"], rtol=1e-12, atol=1e-15)
        assert np.array_equal(d["rhs_own"], ref["RHS"][own])


@pytest.mark.parametrize("partition", ["graph", "hilbert"])
def test_graph_partition_on_one_gpu(afx, gpu, monkeypatch, partition):
    """AFX_PARTITION=graph (recursive graph-growing bisection): several pieces hold no far-field edge; same bits."""
    monkeypatch.setenv("AFX_PARTITION", partition)
    n_iter = 12
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    ref = _single(afx, mesh, "spallart-allmaras", "green-gauss", True, n_iter)
    res, parts = _group_run(afx, mesh, 8, "p2p", "strict", "spallart-allmaras", "green-gauss", True, n_iter, ref["q0"], gpu)
    for d in res:
        own = d["own"]
        assert np.array_equal(d["q_init"][own], ref["q_init"][own])
        assert np.array_equal(d["q"][own], ref["Q"][own])
        np.testing.assert_allclose(d["forces"], ref["F"], rtol=1e-12, atol=1e-15)


def test_wall_cp_of_a_partition_has_no_uninitialised_entries(afx, gpu):
    """afx_rans_wall_cp on a partitioned solver: every wall edge the rank holds (those of its halo cells included) gets the
    cp of its owner cell -- the values of the single-GPU run at the same global edges."""
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    ref = _single(afx, mesh, "spallart-allmaras", "green-gauss", True, 5)
    world = 4
    group = afx.Group(world)
    parts = [afx.Partition(mesh, world, r) for r in range(world)]
    solvers = [None] * world

    def make(r):
        def f():
            solvers[r] = afx.GpuSolver(parts[r], viscosity="spallart-allmaras", math="strict", device=r % gpu, group=group)
        return f
    afx.run_ranks([make(r) for r in range(world)])
    wall = mesh.patch_names.index("wall")
    gb = np.flatnonzero(mesh.bnd_patch == wall)                 # global boundary ids of the wall edges, in order
    cp_of_edge = dict(zip(mesh.bnd_edge[gb].tolist(), ref["cp"].tolist()))

    def work(r):
        def f():
            s, part = solvers[r], parts[r]
            s.set_bcs(BCS); s.set_options(True, "green-gauss", 5.0, 1.4); s.init(); s.refill_bcs()
            s.set_q(ref["q0"])
            s.run(5, 0.9)
            return s.wall_cp("wall")
        return f
    cps = afx.run_ranks([work(r) for r in range(world)])
    total = 0
    for r, cp in enumerate(cps):
        part = parts[r]
        la = part.local_arrays()
        lb = np.flatnonzero(la["bnd_patch"] == wall)
        assert len(cp) == len(lb)
        ge = part.edge_l2g[la["bnd_edge"][lb]]
        want = np.array([cp_of_edge[int(e)] for e in ge])
        assert np.all(np.isfinite(cp))
        assert np.array_equal(cp, want)
        total += len(cp)
    assert total >= len(gb)


def test_dead_peer_returns_comm_error_instead_of_hanging(afx, gpu, monkeypatch):
    """A peer that never delivers (here: it simply does not run) must not leave an unkillable kernel: the halo wait gives up
    after AFX_HALO_TIMEOUT_MS and the run returns AFX_ERR_COMM."""
    monkeypatch.setenv("AFX_HALO_TIMEOUT_MS", "300")
    monkeypatch.setenv("AFX_HALO_LOCKSTEP", "0")  # the in-kernel wait itself, although both ranks share the device (the peer never launches anything)
    mesh = afx.Mesh.synth_omesh(64, 40, 16, 150.0)
    world = 2
    group = afx.Group(world)
    parts = [afx.Partition(mesh, world, r) for r in range(world)]
    solvers = [None] * world

    def make(r):
        def f():
            solvers[r] = afx.GpuSolver(parts[r], viscosity="inviscid", math="strict", device=r % gpu, group=group)
        return f
    afx.run_ranks([make(r) for r in range(world)])
    blobs = [s.p2p_export() for s in solvers]
    for s in solvers:
        s.p2p_connect(blobs)
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.03, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    for s in solvers:
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 1.4); s.init(); s.refill_bcs()
    with pytest.raises(afx.AfxError) as ei:
        solvers[0].L.afx_rans_run_explicit.restype = int
        norms = np.zeros(2)
        rc = solvers[0].L.afx_rans_run_explicit(solvers[0].h, 0.9, 2, norms.ctypes.data)  # rank 1 never runs
        if rc < 0:
            raise afx.AfxError(rc, solvers[0].L.afx_last_error().decode())
    assert ei.value.code == -5 and "did not deliver" in str(ei.value)
    group.abort()
    del solvers[:]
