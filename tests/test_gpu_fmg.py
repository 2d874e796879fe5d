"""FMG on the device: afx_prolongation_* (q_fine = P q_coarse without leaving the GPU) and afx_rans_sweep_fmg (the whole of
Rans::run_airfoil: per angle the full-multigrid start-up over all levels).  The prolongation weights are rebuilt here with
numpy from the reference's rule (multigrid.h:100-178) -- the product ships them from its C++ adapter (k-d tree, tested in
tests/test_cpp_host.py) -- and the results are held to the reference's own vectors (tests/golden/prolongation.npz,
sweep_naca0012q_fmg.npz)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def prolongation_csr(coarse, fine):
    """multigrid::gen_mapper (multigrid.h:100-178), all pairs: coarse cell j contributes to fine cell i when d^2 < 2 A_j with
    weight 1 / max(0.1 sqrt(A_j), d); rows normalised by their sum accumulated over ascending j (ghost cells included)."""
    xc, yc, ac = np.array(coarse.ccx), np.array(coarse.ccy), np.array(coarse.area)
    xf, yf = np.array(fine.ccx), np.array(fine.ccy)
    row_begin = np.zeros(len(xf) + 1, np.uint32)
    cols, ws = [], []
    for i in range(len(xf)):
        d2 = (xf[i] - xc) * (xf[i] - xc) + (yf[i] - yc) * (yf[i] - yc)
        j = np.flatnonzero(d2 < 2 * ac)
        si = 1 / np.maximum(0.1 * np.sqrt(ac[j]), np.sqrt(d2[j]))
        scale = 0.0
        for v in si:  # the reference's running sum, in ascending j
            scale += v
        cols.append(j.astype(np.uint32)); ws.append(si / scale)
        row_begin[i + 1] = row_begin[i] + len(j)
    return row_begin, np.concatenate(cols), np.concatenate(ws)


def test_device_prolongation_matches_reference_bits(afx, gpu):
    g = np.load(H.GOLDEN + "/prolongation.npz")
    (c_dims, f_dims), far = g["dims"], float(g["far_radius"])
    mc = afx.Mesh.synth_omesh(*[int(v) for v in c_dims], far); mf = afx.Mesh.synth_omesh(*[int(v) for v in f_dims], far)
    for math in ("strict", "fast"):  # the prolongation always runs the reference's arithmetic
        sc = afx.GpuSolver(mc, math=math); sf = afx.GpuSolver(mf, math=math)
        P = afx.Prolongation(sc, sf, *prolongation_csr(mc, mf))
        sc.set_q(g["q_coarse"])
        sf.set_q(np.full(4 * (mf.N + mf.G), 7.0))
        P.apply()
        assert np.array_equal(sf.get_q(), g["q_fine"])
    with pytest.raises(afx.AfxError):  # a column outside the coarse mesh is refused, not dereferenced
        rb, col, w = prolongation_csr(mc, mf)
        col = col.copy(); col[3] = mc.N + mc.G
        afx.Prolongation(sc, sf, rb, col, w)


@pytest.fixture(scope="module")
def fmg_levels(afx):
    dc = H.load("naca0012q_coarse_euler_gg_o2"); dm = H.load("naca0012q_mid_mesh")
    mc, mm = H.product_mesh(afx, dc), H.product_mesh(afx, dm)
    return mc, mm, prolongation_csr(mc, mm)


def test_sweep_fmg_converges_to_the_reference_polar(afx, gpu, fmg_levels):
    """coarse -> mid FMG, implicit, conf.ini settings, both sides driven to 1e-10 (golden sweep_naca0012q_fmg.npz, made by
    the unmodified reference's own run_airfoil loop)."""
    g = H.load("sweep_naca0012q_fmg")
    mc, mm, csr = fmg_levels
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    levels = [afx.GpuSolver(m, math="strict") for m in (mc, mm)]
    for s in levels:
        s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0)
    P = afx.Prolongation(levels[0], levels[1], *csr)
    r = afx.sweep_fmg(levels, [P], g["alphas"], implicit=True, tolerance=1e-10, max_iterations=400)
    assert r["status"] == 0 and np.all(r["residual"] <= 1e-10)
    np.testing.assert_allclose(r["cl"], g["cl"], rtol=1e-7)
    np.testing.assert_allclose(r["cd"], g["cd"], rtol=1e-6)
    np.testing.assert_allclose(r["cm"], g["cm"], rtol=1e-6)


def test_sweep_fmg_equals_the_loop_written_with_primitive_calls(afx, gpu, fmg_levels):
    """Explicit (deterministic arithmetic): afx_rans_sweep_fmg must give, to the bit, what the C++ adapter's multigrid<T>::run
    gives when it is written out with the primitive ABI calls and a HOST prolongation (state through get_q / set_q)."""
    mc, mm, (rb, col, w) = fmg_levels
    bcs0 = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    alphas, tol, max_it = [1.0, 3.0], 0.3, 60
    a = [afx.GpuSolver(m, math="strict") for m in (mc, mm)]
    for s in a:
        s.set_bcs(bcs0); s.set_options(True, "green-gauss", 5.0, 1.5)
    P = afx.Prolongation(a[0], a[1], rb, col, w)
    ra = afx.sweep_fmg(a, [P], alphas, implicit=False, relaxation=0.9, start_cfl=1.5, tolerance=tol, max_iterations=max_it)
    b = [afx.GpuSolver(m, math="strict") for m in (mc, mm)]
    for s in b:
        s.set_options(True, "green-gauss", 5.0, 1.5)
    forces, iters = [], []
    for k, al in enumerate(alphas):
        bb = dict(bcs0); bb["farfield"] = ("farfield", dict(mach=0.2, angle=al * 0.01745, T=1.0, p=1.0))
        for s in b:
            s.set_bcs(bb)
        if k == 0:
            b[0].init()
        b[0].refill_bcs()
        total = 0
        for lvl, s in enumerate(b):
            if lvl > 0:
                b[lvl - 1].bcs_from_internal()
                qc = b[lvl - 1].get_q().reshape(-1, 4)
                qf = np.zeros((mm.N + mm.G, 4))
                for i in range(len(qf)):  # row sums in ascending column order, from zero
                    acc = np.zeros(4)
                    for p in range(rb[i], rb[i + 1]):
                        acc = acc + w[p] * qc[col[p]]
                    qf[i] = acc
                s.set_q(qf.ravel())
                s.refill_bcs()
            err_0 = s.get_uniform_residual()
            i = 0
            while True:
                s.set_cfl(1.5)
                err = s.solve(0.9)
                if i == 0 and err > 2 * err_0:
                    err_0 = err
                err /= err_0
                i += 1
                if not (err > tol and i < max_it):
                    break
            total += i
        forces.append(b[-1].wall_forces("wall")); iters.append(total)
    assert list(ra["iterations"]) == iters
    assert np.array_equal(np.array(forces), np.stack([ra["cl"], ra["cd"], ra["cm"]], axis=1))
    assert np.array_equal(a[1].get_q(), b[1].get_q()) and np.array_equal(a[0].get_q(), b[0].get_q())


def test_polar_chains_match_the_reference_polar(afx, gpu, fmg_levels):
    """BASELINE configs[4] at the reference's own settings (conf.ini: implicit, FMG coarse -> mid, tolerance 1e-4, <= 300 iterations
    per level): the 48 angles alpha = -10 ... 13.5 deg of the 64-angle polar, as the six warm-started chains of eight that
    `bench.py --workload polar64 --gpus 8` gives to its first six ranks, against the UNMODIFIED reference's run_airfoil loop on the
    same chains (tests/golden/polar64_reference.npz, oracle/make_golden_polar64.py; the two chains beyond 14 deg are stalled
    inviscid flow on which the reference itself needs hours and hits its iteration limit).  Both sides stop at 1e-4, with different linear solvers, so the forces
    agree to a few 1e-4 -- this pins the polar a user of the reference gets, not the arithmetic (that is test_gpu_converged.py) --
    and the outer iteration counts show the GMRES + block-Jacobi step is as strong as the reference's ILUT + GMRES on this case."""
    g = np.load(H.GOLDEN + "/polar64_reference.npz")
    mc, mm, csr = fmg_levels
    bcs = {"farfield": ("farfield", dict(mach=0.2, angle=0.0, T=1.0, p=1.0)), "wall": ("slip-wall", None)}
    n = 48  # the fixture also holds the chains beyond 14 deg as far as the reference has finished them (stalled flow: its iteration
    #         runs to the 300-iteration limit per level from 16.5 deg on, and so does ours); they are not compared
    al = g["alphas"][:n]
    g = {k: g[k][:n] for k in ("cl", "cd", "cm", "iters")}
    cl, cd, cm, it = [], [], [], []
    for k in range(len(al) // 8):
        levels = [afx.GpuSolver(m, math="fast") for m in (mc, mm)]  # a fresh chain: its first angle starts from the free stream
        for s in levels:
            s.set_bcs(bcs); s.set_options(True, "green-gauss", 5.0, 40.0)
        P = afx.Prolongation(levels[0], levels[1], *csr)
        r = afx.sweep_fmg(levels, [P], al[8 * k:8 * k + 8], implicit=True, relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0,
                          tolerance=1e-4, rhs_iterations=5, max_iterations=300)
        assert r["status"] == 0 and np.all(r["residual"] <= 1e-4)
        cl += list(r["cl"]); cd += list(r["cd"]); cm += list(r["cm"]); it += list(r["iterations"])
    cl, cd, cm, it = map(np.array, (cl, cd, cm, it))
    np.testing.assert_allclose(cl, g["cl"], rtol=0, atol=1e-3)   # |CL| <= 1.2: the lift curve to 1e-3
    np.testing.assert_allclose(cd, g["cd"], rtol=0, atol=1.5e-4)
    np.testing.assert_allclose(cm, g["cm"], rtol=0, atol=1e-4)
    assert np.median(np.abs(cl - g["cl"])) < 1e-4
    # as strong as the reference's linear solve: never more than 1.5x its outer iterations (+5), and not more in total
    assert np.all(it <= 1.5 * g["iters"] + 5), (it, g["iters"])
    assert it.sum() <= 1.15 * g["iters"].sum(), (it.sum(), g["iters"].sum())
