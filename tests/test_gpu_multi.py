"""Partitioned (multi-GPU) runs: N processes, one GPU each, NCCL halo exchange -- against the single-GPU run
and the oracle.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BCS = {"farfield": ("farfield", dict(mach=0.2, angle=0.03, T=1.0, p=1.0)), "wall": ("wall", None)}
DIMS = (256, 160, 64)  # 65 536 mixed cells


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _q0(afx, mesh, visc="spallart-allmaras", grad="green-gauss", so=True):
    s = afx.GpuSolver(mesh, viscosity=visc, math="strict", device=0)
    s.set_bcs(BCS); s.set_options(so, grad, 5.0, 1.4); s.init(); s.refill_bcs()
    q = s.get_q()
    rng = np.random.default_rng(77)
    q[:4 * mesh.N] *= 1 + 1e-3 * rng.uniform(-1, 1, 4 * mesh.N)
    return s, q


def _worker(rank, world, port, n_iter, out_dir, math, halo, fused, visc="spallart-allmaras", grad="green-gauss", so=True):
    import torch.distributed as dist
    import aeroflex_b200 as afx
    if fused:
        os.environ["AFX_FUSED"] = "1"  # tiles + graph-bisection numbering at creation: the fused stage kernel pushes the halo itself
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    part = afx.Partition(mesh, world, rank)
    ids = [afx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = afx.GpuSolver(part, viscosity=visc, math=math, device=rank, nccl_id=ids[0])
    if halo == "p2p":  # NVLink peer-memory halo: all-gather the IPC blobs, map the peers' receive buffers
        blobs = [None] * world
        dist.all_gather_object(blobs, s.p2p_export())
        s.p2p_connect(blobs)
    assert s.halo_mode() == halo
    assert s.tile_info()["fused"] == bool(fused)
    s.set_bcs(BCS); s.set_options(so, grad, 5.0, 1.4); s.init(); s.refill_bcs()
    qi = np.full(4 * (mesh.N + mesh.G), np.nan)
    s.get_q(qi)  # init() + refill_bcs() of this rank's piece, in global numbering
    q0 = np.load(os.path.join(out_dir, "q0.npy"))
    s.set_q(q0)
    ur = s.get_uniform_residual()
    norms = s.run(n_iter, 0.9)
    out = np.full(4 * (mesh.N + mesh.G), np.nan)
    s.get_q(out)
    forces = s.wall_forces("wall")
    rhs_norm = s.residual()  # implicitSolver::fillRhoRHS on the partitioned state: norm all-reduced, vector per rank
    rhs = np.full(4 * (mesh.N + mesh.G), np.nan)
    loc = s.get("rhs").reshape(-1, 4)
    rhs.reshape(-1, 4)[part.cell_l2g[:part.n_own]] = loc[:part.n_own]
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), q=out, norms=norms, forces=np.array(forces), owned=part.cell_l2g[:part.n_own], ur=ur,
             q_init=qi, rhs=rhs, rhs_norm=rhs_norm)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,math,halo,fused", [(2, "strict", "nccl", 0), (2, "strict", "p2p", 0), (2, "fast", "p2p", 0),
                                                   (2, "strict", "p2p", 1), (2, "strict", "nccl", 1), (4, "strict", "p2p", 1),
                                                   (8, "strict", "p2p", 0), (8, "fast", "p2p", 1)])
def test_partitioned_run_matches_single_gpu(afx, gpu, tmp_path, world, math, halo, fused):
    if gpu < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, gpu))
    import torch.multiprocessing as mp
    n_iter = 25
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    single, q0 = _q0(afx, mesh)
    q_init = single.get_q().reshape(-1, 4)  # init() + refill_bcs() on one GPU
    np.save(tmp_path / "q0.npy", q0)
    single.set_q(q0)
    ur = single.get_uniform_residual()
    ref_norms = single.run(n_iter, 0.9)
    Q = single.get_q().reshape(-1, 4)
    F = np.array(single.wall_forces("wall"))
    del single
    mp.spawn(_worker, args=(world, _free_port(), n_iter, str(tmp_path), math, halo, fused), nprocs=world, join=True)
    seen = 0
    for r in range(world):
        d = np.load(tmp_path / ("r%d.npz" % r))
        own = d["owned"]
        got = d["q"].reshape(-1, 4)[own]
        # a rank that holds no far-field edge still initialises with the far-field state of the whole mesh (solver.h:597-631)
        assert np.array_equal(d["q_init"].reshape(-1, 4)[own], q_init[own])
        if math == "strict":
            assert np.array_equal(got, Q[own])  # bit-identical to the single-GPU strict run (itself bit-identical to the oracle)
            np.testing.assert_allclose(d["norms"], ref_norms, rtol=1e-12)
            np.testing.assert_allclose(d["forces"], F, rtol=1e-12, atol=1e-15)
            assert d["ur"] == pytest.approx(ur, rel=1e-12)
        else:
            np.testing.assert_allclose(got, Q[own], rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(d["norms"], ref_norms, rtol=1e-10)
            np.testing.assert_allclose(d["forces"], F, rtol=1e-8, atol=1e-12)
        seen += len(own)
    assert seen == mesh.N


@pytest.mark.parametrize("world,visc,grad,so,halo", [(3, "laminar", "least-squares", True, "p2p"), (2, "laminar", "green-gauss", True, "nccl"),
                                                      (3, "inviscid", "green-gauss", False, "p2p"), (5, "spallart-allmaras", "least-squares", True, "p2p")])
def test_partitioned_variants_match_single_gpu(afx, gpu, tmp_path, world, visc, grad, so, halo):
    """Odd rank counts, the laminar face-gradient path (ftij, iteration-start state), least-squares gradients and first-order
    runs in a partitioned solver; also the implicit right-hand side (afx_rans_residual) of a partitioned state."""
    if gpu < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, gpu))
    import torch.multiprocessing as mp
    n_iter = 8
    mesh = afx.Mesh.synth_omesh(*DIMS, 150.0)
    single, q0 = _q0(afx, mesh, visc, grad, so)
    if visc == "laminar":  # keep the perturbation small enough for the viscous terms at this CFL
        base = single.get_q()
        q0 = base + 0.1 * (q0 - base)
    np.save(tmp_path / "q0.npy", q0)
    single.set_q(q0)
    ref_norms = single.run(n_iter, 0.9)
    Q = single.get_q().reshape(-1, 4)
    F = np.array(single.wall_forces("wall"))
    rhs_norm = single.residual()
    RHS = single.get("rhs").reshape(-1, 4)
    del single
    mp.spawn(_worker, args=(world, _free_port(), n_iter, str(tmp_path), "strict", halo, 0, visc, grad, so), nprocs=world, join=True)
    for r in range(world):
        d = np.load(tmp_path / ("r%d.npz" % r))
        own = d["owned"]
        assert np.array_equal(d["q"].reshape(-1, 4)[own], Q[own])
        np.testing.assert_allclose(d["norms"], ref_norms, rtol=1e-12)
        np.testing.assert_allclose(d["forces"], F, rtol=1e-12, atol=1e-15)
        assert np.array_equal(d["rhs"].reshape(-1, 4)[own], RHS[own])
        assert float(d["rhs_norm"]) == pytest.approx(rhs_norm, rel=1e-12)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_early_halo_signal_matches_single_gpu(afx, gpu, tmp_path, monkeypatch, world):
    """AFX_HALO_EARLY_SIGNAL=1: the last send-layer CTA of the update kernel raises the peers' flags itself (no signalling
    launch, the hand-off overlaps the interior update).  Same data, same bits."""
    monkeypatch.setenv("AFX_HALO_EARLY_SIGNAL", "1")
    test_partitioned_run_matches_single_gpu(afx, gpu, tmp_path, world, "strict", "p2p", 0)


def _implicit_worker(rank, world, port, out_dir, alpha_deg):
    import torch.distributed as dist
    import aeroflex_b200 as afx
    from tests import helpers as H
    from tests.test_gpu_parity import run_implicit
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = H.product_mesh(afx, H.load("naca0012q_coarse_euler_gg_o2"))
    part = afx.Partition(mesh, world, rank)
    ids = [afx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = afx.GpuSolver(part, math="strict", device=rank, nccl_id=ids[0])
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=alpha_deg * 0.01745, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0); s.init(); s.refill_bcs()
    hist = run_implicit(s, 1e-10, max_iter=400)
    np.savez(os.path.join(out_dir, "i%d.npz" % rank), hist=np.array(hist), forces=np.array(s.wall_forces("wall")),
             lin=s.last_linear_iterations())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_implicit_run_converges_to_the_reference_forces(afx, gpu, tmp_path, world):
    """implicitSolver on a partitioned mesh: halo rows of the Krylov vectors come from their owners before every
    matrix-vector product / Jacobi sweep, inner products are summed over the ranks.  Per-iteration histories differ from
    the single-GPU run in the last digits (summation order); the converged CL/CD/CM are the reference's
    (golden sweep_naca0012q_coarse.npz, both sides at 1e-10)."""
    if gpu < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, gpu))
    import torch.multiprocessing as mp
    from tests import helpers as H
    g = H.load("sweep_naca0012q_coarse")
    mp.spawn(_implicit_worker, args=(world, _free_port(), str(tmp_path), float(g["alphas"][0])), nprocs=world, join=True)
    outs = [np.load(tmp_path / ("i%d.npz" % r)) for r in range(world)]
    for d in outs:
        assert 0 <= d["hist"][-1] <= 1e-10, d["hist"][-5:]
        assert np.array_equal(d["hist"], outs[0]["hist"]) and np.array_equal(d["forces"], outs[0]["forces"])  # every rank sees the same numbers
        assert d["forces"][0] == pytest.approx(float(g["cl"][0]), rel=1e-7)
        assert d["forces"][1] == pytest.approx(float(g["cd"][0]), rel=1e-6)
        assert d["forces"][2] == pytest.approx(float(g["cm"][0]), rel=1e-6)
