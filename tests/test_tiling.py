"""Host side of the fused stage kernel: the renumbering (recursive graph bisection) and the shared-memory tile plan.
afx_tiling_plan runs the same host code as afx_rans_create without touching a device and verifies the plan against the
connectivity (every advanced cell owned once, local neighbour / face indices and side bits right, ring-1 cells listing
exactly the faces they share with the tile); it raises on the first inconsistency."""
import os

import numpy as np
import pytest

from tests import helpers as H


@pytest.fixture
def order_env():
    old = os.environ.get("AFX_ORDER")
    yield
    if old is None:
        os.environ.pop("AFX_ORDER", None)
    else:
        os.environ["AFX_ORDER"] = old


@pytest.mark.parametrize("order", ["graph", "hilbert"])
@pytest.mark.parametrize("tile", [32, 100, 320])
def test_tile_plan_is_consistent_on_synthetic_mixed_mesh(afx, order_env, order, tile):
    os.environ["AFX_ORDER"] = order
    m = afx.Mesh.synth_omesh(128, 80, 32, 150.0)  # 16 384 cells, quads near the wall, triangles outside
    per, smem = afx.tiling_plan(m, tile)
    nc, h1, h2, nf = per.T.astype(np.int64)
    assert nc.sum() == m.N and nc.max() <= tile and nc.min() >= 1
    assert smem > 0
    # every face is local to one tile or, if it is cut by a tile boundary, to two
    assert m.E <= nf.sum() <= 2 * m.E
    if order == "graph":  # bisection leaves are balanced and compact: nobody stages more than ~3x its own cells
        assert nc.min() >= tile // 2 - 1
        assert ((nc + h1 + h2) / nc).mean() < (3.2 if tile == 32 else 2.2)


@pytest.mark.parametrize("tag", ["naca0012q_coarse_euler_gg_o2", "naca0012_coarse_euler_gg_o1", "flat_plate_sa_gg_o2"])
def test_tile_plan_on_shipped_meshes(afx, order_env, tag):
    os.environ["AFX_ORDER"] = "graph"
    m = H.product_mesh(afx, H.load(tag))
    for tile in (64, 256):
        per, _ = afx.tiling_plan(m, tile)
        assert per[:, 0].sum() == m.N


def test_tiles_beyond_the_limits_are_cut_in_two(afx, order_env):
    os.environ["AFX_ORDER"] = "graph"
    m = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    free, smem_free = afx.tiling_plan(m, 256)
    lim = (300, 280, 10 ** 6, 10 ** 6)  # local cells, own + ring 1: tighter than most 256-cell tiles need
    cut, smem_cut = afx.tiling_plan(m, 256, limits=lim)
    assert len(cut) > len(free) and cut[:, 0].sum() == m.N
    assert (cut[:, 0] + cut[:, 1] + cut[:, 2]).max() <= lim[0] and (cut[:, 0] + cut[:, 1]).max() <= lim[1]
    assert smem_cut < smem_free


@pytest.mark.parametrize("nranks", [2, 3])
def test_tile_plan_of_partitioned_meshes(afx, order_env, nranks):
    """A rank's tiles cover exactly its owned cells; the send layer (cells some peer needs) is tiled apart from the
    interior so that it can be advanced and pushed first.  The plan is verified against the local connectivity."""
    os.environ["AFX_ORDER"] = "graph"
    mesh = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    for r in range(nranks):
        p = afx.Partition(mesh, nranks, r)
        per, smem = afx.tiling_plan(p, 96)
        assert per[:, 0].sum() == p.n_own and per[:, 0].max() <= 96
        n_front = len(np.unique(np.concatenate([send for (_, send, _) in p.peers])))
        # the first tiles hold the send layer and nothing else: some prefix of the tile sizes sums to it
        assert n_front in np.cumsum(per[:, 0])
