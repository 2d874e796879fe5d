"""The fused stage kernel (one CTA per shared-memory tile: limiter + MUSCL + flux + gather + update) against the
three-kernel stage and the oracle.  Tiles recompute ring-1 limiters and the faces they share with their neighbours,
and every per-cell sum keeps the reference's edge order, so in strict mode the STATE must not depend on the tiling:
bit-identical for every tile size, to the unfused kernels and to the CPU oracle."""
import os

import numpy as np
import pytest

from oracle import orc
from tests import helpers as H

pytestmark = pytest.mark.gpu

BCS = {"farfield": ("farfield", dict(mach=0.2, angle=2 * 0.01745, T=1.0, p=1.0)), "wall": ("wall", None)}


@pytest.fixture(autouse=True)
def fused_env():
    """the tiles are built at creation only under AFX_FUSED=1"""
    old = {k: os.environ.get(k) for k in ("AFX_FUSED", "AFX_TILE")}
    os.environ["AFX_FUSED"] = "1"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def make(afx, m, tile=None, fused=True, math="strict", visc="spallart-allmaras"):
    if tile is not None:
        os.environ["AFX_TILE"] = str(tile)
    else:
        os.environ.pop("AFX_TILE", None)
    s = afx.GpuSolver(m, viscosity=visc, math=math)
    s.set_bcs(BCS); s.set_options(True, "green-gauss", 5.0, 1.2); s.init(); s.refill_bcs()
    s.set_fused(fused)
    return s


@pytest.mark.parametrize("tile", [48, 100, 192, 320])
def test_fused_stage_is_bit_identical_to_three_kernel_stage(afx, gpu, tile):
    m = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    a = make(afx, m, tile=tile, fused=True); b = make(afx, m, fused=False)
    info = a.tile_info()
    assert info["fused"] and 0 < info["tile_cells"] <= tile and info["ctas_per_sm"] >= 1
    assert not b.tile_info()["fused"]
    q0 = H.synth_state(m.N, a.get_q())
    a.set_q(q0); b.set_q(q0)
    na = a.run(7, 0.9); nb = b.run(7, 0.9)
    np.testing.assert_allclose(na, nb, rtol=1e-12, atol=0)
    assert np.array_equal(a.get_q(), b.get_q())
    for f in ("qW", "limiters", "gx", "gy"):  # kept for the last iteration of a run
        assert np.array_equal(a.get(f), b.get(f)), f
    # one more call of a single iteration: stage 0 reads q in place of the stage buffer
    assert a.solve(0.9) == pytest.approx(b.solve(0.9), rel=1e-12)
    assert np.array_equal(a.get_q(), b.get_q())


def test_fused_stage_vs_oracle_slip_wall_and_unknown_bc(afx, gpu):
    """slip walls, far field and a two-sided boundary face (unknown bc_type, solver.h:211-212) through the fused kernel"""
    m = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    x, y, cells, b0, b1 = m.elements()
    om = orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names, fast=False)
    for wall in ("slip-wall", "inlet-outlet"):
        bcs = dict(BCS); bcs["wall"] = (wall, None)
        s = afx.GpuSolver(m, viscosity="inviscid", math="strict"); o = orc.OracleSolver(om, viscosity="inviscid")
        for z in (s, o):
            z.set_bcs(bcs); z.set_options(True, "green-gauss", 5.0, 1.0); z.init(); z.refill_bcs()
        assert s.tile_info()["fused"]
        q0 = H.synth_state(m.N, o.q.copy())
        s.set_q(q0); o.q[:] = q0
        gn = s.run(5, 0.9)
        on = np.array([o.explicit_solve(0.9) for _ in range(5)])
        np.testing.assert_allclose(gn, on, rtol=1e-12, atol=0)
        assert np.array_equal(s.get_q(), o.q), wall
        # real rows; the ghost rows of a two-sided boundary face are only ever summed into the norm (checked above)
        assert np.array_equal(s.get("qW")[:4 * m.N], o.qW[:4 * m.N]), wall


def test_fused_fast_mode_within_north_star_tolerance(afx, gpu):
    m = afx.Mesh.synth_omesh(256, 160, 64, 150.0)
    f = make(afx, m, math="fast"); r = make(afx, m, fused=False, math="strict")
    q0 = H.synth_state(m.N, r.get_q())
    f.set_q(q0); r.set_q(q0)
    nf = f.run(100, 0.9); nr = r.run(100, 0.9)
    np.testing.assert_allclose(nf, nr, rtol=1e-10, atol=0)  # north_star: 1e-10 relative over the first 100 iterations
    np.testing.assert_allclose(f.get_q(), r.get_q(), rtol=1e-10, atol=1e-11)  # states are O(1); |v| is 1e-3 in places
    np.testing.assert_allclose(f.wall_forces("wall"), r.wall_forces("wall"), rtol=1e-8, atol=1e-13)


def test_first_order_and_laminar_runs_keep_the_three_kernel_stage(afx, gpu):
    m = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    s = afx.GpuSolver(m, viscosity="laminar", math="strict")
    s.set_bcs(BCS); s.set_options(True, "green-gauss", 5.0, 1.0)
    assert not s.tile_info()["fused"]
    s = afx.GpuSolver(m, viscosity="inviscid", math="strict")
    s.set_bcs(BCS); s.set_options(False, "green-gauss", 5.0, 1.0)
    assert not s.tile_info()["fused"]
