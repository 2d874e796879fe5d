"""Multi-GPU host logic on CPU: the partition / halo plan (pure checks) and a world_size-2 gloo run in which
every rank advances its piece with the ORACLE as the compute engine and exchanges halo states as the plan says;
the owned cells must reproduce the single-domain oracle bit for bit.  (The GPU path uses the same plan with NCCL.)"""
import os
import socket

import numpy as np
import pytest

from oracle import orc

BCS = {"farfield": ("farfield", dict(mach=0.2, angle=0.03, T=1.0, p=1.0)), "wall": ("slip-wall", None)}


def neighbours(mesh):
    N = mesh.N
    ec = mesh.edge_cells
    nb = [[] for _ in range(N)]
    for a, b in ec:
        if a < N and b < N:
            nb[a].append(int(b)); nb[b].append(int(a))
    return nb


@pytest.mark.parametrize("method", ["hilbert", "graph"])
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_plan_is_consistent(afx, monkeypatch, nranks, method):
    """AFX_PARTITION=hilbert (default): cuts of the Hilbert curve that are even in work; =graph: recursive graph bisection of the
    face-neighbour graph (METIS-style).  Same plan contract for both."""
    monkeypatch.setenv("AFX_PARTITION", method)
    mesh = afx.Mesh.synth_omesh(64, 40, 16, 40.0)
    N = mesh.N
    nb = neighbours(mesh)
    parts = [afx.Partition(mesh, nranks, r) for r in range(nranks)]
    owned = np.concatenate([p.cell_l2g[:p.n_own] for p in parts])
    assert len(owned) == N and len(np.unique(owned)) == N  # a partition of the cells
    if method == "graph":
        assert max(p.n_own for p in parts) - min(p.n_own for p in parts) <= 1
    else:  # the curve is cut by WORK: a quadrilateral weighs 1.12 triangles (its share of the faces), pieces are even in that measure
        work = [int(np.sum(np.where(mesh.is_tri[p.cell_l2g[:p.n_own]] != 0, 100, 112))) for p in parts]
        assert max(work) - min(work) <= 2 * 112
    owner = np.empty(N, int)
    for p in parts:
        owner[p.cell_l2g[:p.n_own]] = p.rank
    for p in parts:
        own = set(int(c) for c in p.cell_l2g[:p.n_own])
        d1 = {n for c in own for n in nb[c]} - own
        d2 = {n for c in d1 for n in nb[c]} - own - d1
        assert set(int(c) for c in p.cell_l2g[p.n_own:p.n_own + p.n_r1]) == d1
        assert set(int(c) for c in p.cell_l2g[p.n_own + p.n_r1:p.N]) == d2
        assert np.all(np.diff(p.edge_l2g.astype(np.int64)) > 0)  # ascending global edge ids: reference accumulation order
        # every edge of an owned or ring-1 cell is present
        need = set()
        for l in range(p.n_own + p.n_r1):
            c = p.cell_l2g[l]
            need.update(int(e) for e in mesh.cell_edges[c][:3 if mesh.is_tri[c] else 4])
        assert need == set(int(e) for e in p.edge_l2g)
        # geometry is copied bit for bit
        a = p.local_arrays()
        assert np.array_equal(a["enx"], mesh.enx[p.edge_l2g]) and np.array_equal(a["area"], mesh.area[p.cell_l2g])
        assert np.array_equal(p.cell_l2g[a["edge_cells"]], mesh.edge_cells[p.edge_l2g])
        # plan symmetry: what I send to r is what r expects from me, in the same order
        for (r, send, recv) in p.peers:
            other = [q for q in parts[r].peers if q[0] == p.rank]
            assert len(other) == 1
            assert np.array_equal(p.cell_l2g[send], parts[r].cell_l2g[other[0][2]])
            assert np.all(send < p.n_own) and np.all(recv >= p.n_own) and np.all(recv < p.N)
            assert np.all(owner[p.cell_l2g[recv]] == r)
        got = np.concatenate([rv for (_, _, rv) in p.peers]) if p.peers else np.zeros(0, int)
        assert sorted(got.tolist()) == list(range(p.n_own, p.N))  # every halo cell is filled exactly once


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, n_iter, out_dir):
    import torch
    import torch.distributed as dist
    import aeroflex_b200 as afx
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh = afx.Mesh.synth_omesh(64, 40, 16, 40.0)
    part = afx.Partition(mesh, world, rank)
    om = orc.OracleMesh.from_arrays(part.local_arrays(), part.N, part.G, part.patch_names)
    o = orc.OracleSolver(om)
    o.set_bcs(BCS); o.set_options(True, "green-gauss", 5.0, 1.3); o.init(); o.refill_bcs()
    # the same global perturbed state on every rank, restricted to the local cells
    NTg = mesh.N + mesh.G
    x, y, cells, b0, b1 = mesh.elements()
    gm = orc.OracleMesh(x, y, cells, mesh.is_tri, b0, b1, mesh.bnd_patch, mesh.patch_names)
    g = orc.OracleSolver(gm); g.set_bcs(BCS); g.set_options(True, "green-gauss", 5.0, 1.3); g.init(); g.refill_bcs()
    rng = np.random.default_rng(5)
    q0 = g.q.copy(); q0[:4 * mesh.N] *= 1 + 1e-3 * rng.uniform(-1, 1, 4 * mesh.N)
    o.q[:] = q0.reshape(NTg, 4)[part.cell_l2g].ravel()
    n_own, NL = part.n_own, part.N
    norms = []
    for it in range(n_iter):
        o.calc_dt()
        o.qk[:] = o.q
        for a in (0.25, 0.5, 1.0):
            o.walls(1); o.calc_gradients(); o.calc_limiters(1); o.calc_residual(1)
            qk = o.qk.reshape(-1, 4); q = o.q.reshape(-1, 4); qW = o.qW.reshape(-1, 4)
            qk[:n_own] = q[:n_own] + qW[:n_own] * o.dt[:n_own, None] * a * 0.9
            reqs = []
            bufs = []
            for (r, send, recv) in part.peers:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(qk[send])), r))
                b = torch.empty((len(recv), 4), dtype=torch.float64); bufs.append((recv, b))
                reqs.append(dist.irecv(b, r))
            for rq in reqs:
                rq.wait()
            for recv, b in bufs:
                qk[recv] = b.numpy()
        o.q[:] = o.qk
        t = torch.tensor([float(np.sum(o.qW.reshape(-1, 4)[:n_own] ** 2))], dtype=torch.float64)
        dist.all_reduce(t)
        norms.append(float(np.sqrt(t.item())))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), q=o.q.reshape(-1, 4)[:n_own], cells=part.cell_l2g[:n_own], norms=np.array(norms),
             ghosts_q=o.q.reshape(-1, 4)[NL:], ghosts=part.cell_l2g[NL:])
    dist.destroy_process_group()


def test_graph_partition_pieces_are_connected_and_cut_less_than_the_curve(afx, monkeypatch):
    """Each piece of the graph partition is one connected component of the cell graph, and on the stretched O-mesh its halo
    is no larger than that of the Hilbert chunks (2 ranks: 512 against 640 ring cells on the 65 536-cell mesh)."""
    mesh = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    nb = neighbours(mesh)
    halo = {}
    for method in ("hilbert", "graph"):
        monkeypatch.setenv("AFX_PARTITION", method)
        parts = [afx.Partition(mesh, 2, r) for r in range(2)]
        halo[method] = max(p.n_r1 + p.n_r2 for p in parts)
        if method == "graph":
            for p in parts:
                own = set(int(c) for c in p.cell_l2g[:p.n_own])
                seen, stack = set(), [next(iter(own))]
                while stack:
                    c = stack.pop()
                    if c in seen:
                        continue
                    seen.add(c)
                    stack.extend(n for n in nb[c] if n in own and n not in seen)
                assert seen == own
    assert halo["graph"] <= halo["hilbert"]
    monkeypatch.setenv("AFX_PARTITION", "metis")
    with pytest.raises(afx.AfxError):
        afx.Partition(mesh, 2, 0)


def test_two_rank_gloo_run_reproduces_single_domain(afx, tmp_path):
    import torch.multiprocessing as mp
    n_iter = 4
    mp.spawn(_worker, args=(2, _free_port(), n_iter, str(tmp_path)), nprocs=2, join=True)
    mesh = afx.Mesh.synth_omesh(64, 40, 16, 40.0)
    x, y, cells, b0, b1 = mesh.elements()
    gm = orc.OracleMesh(x, y, cells, mesh.is_tri, b0, b1, mesh.bnd_patch, mesh.patch_names)
    g = orc.OracleSolver(gm); g.set_bcs(BCS); g.set_options(True, "green-gauss", 5.0, 1.3); g.init(); g.refill_bcs()
    rng = np.random.default_rng(5)
    g.q[:4 * mesh.N] *= 1 + 1e-3 * rng.uniform(-1, 1, 4 * mesh.N)
    ref_norms = [g.explicit_solve(0.9) for _ in range(n_iter)]
    Q = g.q.reshape(-1, 4)
    seen = 0
    for r in range(2):
        d = np.load(tmp_path / ("rank%d.npz" % r))
        assert np.array_equal(d["q"], Q[d["cells"]])          # owned cells: bit-identical to the single-domain run
        np.testing.assert_allclose(d["norms"], ref_norms, rtol=1e-13)
        seen += len(d["cells"])
    assert seen == mesh.N
