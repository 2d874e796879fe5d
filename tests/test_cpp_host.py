"""The C++ host adapter (aeroflex_b200/host/rans_b200/*.h: rans::Settings, rans::mesh, rans::solver,
rans::multigrid, rans::Rans) -- compiled with g++ against the C-ABI library and driven like the reference's CLI."""
import os
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "aeroflex_b200", "host")

CONF = """[rans-gas]
gamma = 1.4
R = 0.71428571428

[rans-bc]
<
    name = farfield,
    type = farfield,
    T = 1,
    mach = 0.2,
    angle = 420.0,
    p = 1,
>
<name = wall, type = slip-wall>

[rans-alphas]
alpha_start = 1.0
alpha_end = %(alpha_end)s
alpha_step = 3.0

[rans-solver]
solver = %(solver)s
viscosity = inviscid
gradient = green-gauss
second_order = true
relaxation = 0.9
start_cfl = %(start_cfl)s
slope_cfl = 50.0
max_cfl = 100.0
tolerance = %(tol)s
rhs_iterations = 5
max_iterations = %(max_it)s
limiter_k = 5.0
"""


def _build(afx, src, out):
    lib = os.path.realpath(afx.library_path())  # linked by file name: AFX_LIB may name a differently tuned build
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fopenmp", "-Wall", "-Werror", "-I" + HOST, "-o", str(out), src,
                    "-L" + os.path.dirname(lib), "-l:" + os.path.basename(lib), "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return str(out)


def test_adapter_maps_the_reference_michalak_macro_onto_the_library_switch():
    """The reference selects its other limiter at compile time (#ifdef RANS_MICHALAK_LIMITER, solver.h:557); the adapter's solver class
    turns the macro into afx_rans_set_limiter(AFX_LIMITER_MICHALAK) when it creates the device solver.  Compile check, no device."""
    src = os.path.join(HOST, "rans_cli.cpp")
    for extra in ([], ["-DRANS_MICHALAK_LIMITER"]):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-fopenmp", "-Wall", "-Werror", "-I" + HOST] + extra + [src], check=True)
    pre = subprocess.run(["/usr/bin/g++", "-std=c++17", "-E", "-fopenmp", "-DRANS_MICHALAK_LIMITER", "-I" + HOST, src], check=True, capture_output=True, text=True).stdout
    assert "afx_rans_set_limiter(s, AFX_LIMITER_MICHALAK)" in pre
    pre0 = subprocess.run(["/usr/bin/g++", "-std=c++17", "-E", "-fopenmp", "-I" + HOST, src], check=True, capture_output=True, text=True).stdout
    assert "afx_rans_set_limiter(s, AFX_LIMITER_MICHALAK)" not in pre0


@pytest.fixture(scope="module")
def exes(afx, tmp_path_factory):
    d = tmp_path_factory.mktemp("cpp")
    return dict(test_host=_build(afx, os.path.join(ROOT, "tests", "cpp", "test_host.cpp"), d / "test_host"),
                cli=_build(afx, os.path.join(HOST, "rans_cli.cpp"), d / "rans_cli"), dir=d)


def test_conf_ini_contract(exes, tmp_path):
    """Same sections, keys and defaults as Settings::import_config_file (core.h:238-272), incl. the vector entries."""
    ini = tmp_path / "conf.ini"
    ini.write_text(CONF % dict(solver="implicit", tol="1e-4", max_it=300, alpha_end="7.0", start_cfl="40.0"))
    out = subprocess.run([exes["test_host"], "conf", str(ini), str(tmp_path / "rt.ini")], capture_output=True, text=True, check=True).stdout
    assert "solver=implicit viscosity=inviscid gradient=green-gauss second_order=1" in out
    assert "start_cfl=40 slope_cfl=50 max_cfl=100 tolerance=0.0001 rhs_iterations=5 max_iterations=300 limiter_k=5 alpha=1:7:3" in out
    assert "bc farfield type=farfield mach=0.20000000000000001 angle=420 T=1 p=1" in out
    assert "bc wall type=slip-wall" in out and "roundtrip 1" in out
    bad = tmp_path / "bad.ini"  # exactly two rans-bc entries are required (core.h:239)
    bad.write_text((CONF % dict(solver="implicit", tol="1e-4", max_it=300, alpha_end="7.0", start_cfl="40.0")).replace("<name = wall, type = slip-wall>", ""))
    r = subprocess.run([exes["test_host"], "conf", str(bad), str(tmp_path / "rt2.ini")], capture_output=True, text=True)
    assert r.returncode == 1 and "[RANS] Invalid number of boundary conditions" in r.stdout


def test_fmg_prolongation_matches_reference(afx, exes, tmp_path):
    """multigrid::gen_mapper (multigrid.h:100-178): same cells, same weights, same sums -> bit-identical q_fine."""
    g = np.load(os.path.join(H.GOLDEN, "prolongation.npz"))
    paths = []
    for k, (ni, nj, nq) in enumerate(g["dims"]):
        m = afx.Mesh.synth_omesh(int(ni), int(nj), int(nq), float(g["far_radius"]))
        paths.append(str(tmp_path / ("m%d.msh" % k)))
        m.write_msh(paths[-1])
    g["q_coarse"].astype(np.float64).tofile(tmp_path / "qc.bin")
    subprocess.run([exes["test_host"], "prolong", paths[0], paths[1], str(tmp_path / "qc.bin"), str(tmp_path / "qf.bin")], check=True,
                   capture_output=True)
    qf = np.fromfile(tmp_path / "qf.bin")
    assert np.array_equal(qf, g["q_fine"])


def test_fmg_prolongation_tree_search_equals_all_pairs(afx, exes, tmp_path):
    """the k-d tree search must find exactly the pairs of the reference's O(m n) loop, in the same order."""
    for k, dims in enumerate([(64, 40, 16), (128, 80, 32), (96, 60, 0)]):
        afx.Mesh.synth_omesh(*dims, 150.0).write_msh(str(tmp_path / ("p%d.msh" % k)))
    for a, b in ((0, 1), (2, 1), (1, 0)):  # coarse -> fine, quads -> mixed, and fine -> coarse for good measure
        r = subprocess.run([exes["test_host"], "prolong_check", str(tmp_path / ("p%d.msh" % a)), str(tmp_path / ("p%d.msh" % b))],
                           capture_output=True, text=True)
        assert r.returncode == 0 and "identical=1" in r.stdout, r.stdout


def test_wall_distance_tree_search_equals_all_pairs(afx, exes, tmp_path):
    """mesh::compute_wall_dist (mesh.h:794-830) through the k-d tree: bit-identical to the reference's N x G scan; no wall -> 1."""
    afx.Mesh.synth_omesh(128, 80, 32, 150.0).write_msh(str(tmp_path / "w.msh"))
    r = subprocess.run([exes["test_host"], "walldist", str(tmp_path / "w.msh"), "slip-wall"], capture_output=True, text=True)
    assert r.returncode == 0 and "identical=1" in r.stdout, r.stdout
    r = subprocess.run([exes["test_host"], "walldist", str(tmp_path / "w.msh"), "inlet"], capture_output=True, text=True)
    assert r.returncode == 0 and "identical=1" in r.stdout and "max=1 " in r.stdout, r.stdout


def test_cli_fails_loudly_without_gpu(afx, exes, tmp_path):
    if afx.device_count() > 0:
        pytest.skip("a GPU is visible")
    ini = tmp_path / "conf.ini"
    ini.write_text(CONF % dict(solver="explicit", tol="1e-4", max_it=10, alpha_end="1.0", start_cfl="1.5"))
    for tag in ("naca0012q_coarse_euler_gg_o2",):
        H.product_mesh(afx, H.load(tag)).write_msh(tmp_path / "naca0012q_coarse.msh")
    H.product_mesh(afx, H.load("naca0012q_mid_mesh")).write_msh(tmp_path / "naca0012q_mid.msh")
    r = subprocess.run([exes["cli"], "-i", str(ini), "-m", str(tmp_path) + "/", "-q"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cli_polar_sweep_matches_reference_fmg(afx, gpu, exes, tmp_path):
    """BASELINE config 1: the conf.ini airfoil case (implicit, FMG coarse -> mid, alpha sweep) through rans::Rans::solve_airfoil
    on the GPU against the reference's own run, both driven to 1e-10 so that the converged CL/CD/CM are comparable."""
    path = os.path.join(H.GOLDEN, "sweep_naca0012q_fmg.npz")
    if not os.path.exists(path):
        pytest.skip("golden FMG sweep not generated (oracle/make_golden_fmg.py)")
    gold = np.load(path)
    H.product_mesh(afx, H.load("naca0012q_coarse_euler_gg_o2")).write_msh(tmp_path / "naca0012q_coarse.msh")
    H.product_mesh(afx, H.load("naca0012q_mid_mesh")).write_msh(tmp_path / "naca0012q_mid.msh")
    ini = tmp_path / "conf.ini"
    ini.write_text(CONF % dict(solver="implicit", tol="1e-10", max_it=400, alpha_end="4.0", start_cfl="40.0"))
    r = subprocess.run([exes["cli"], "-i", str(ini), "-m", str(tmp_path) + "/", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    rows = [[float(v) for v in l.split()[1:]] for l in r.stdout.splitlines() if l.startswith("POLAR")]
    assert [row[0] for row in rows] == [1.0, 4.0]
    for row, cl, cd, cm in zip(rows, gold["cl"], gold["cd"], gold["cm"]):
        assert row[1] == pytest.approx(cl, rel=1e-6)
        assert row[2] == pytest.approx(cd, rel=1e-5)
        assert row[3] == pytest.approx(cm, rel=1e-5)


@pytest.mark.gpu
def test_cli_explicit_mode_runs(afx, gpu, exes, tmp_path):
    H.product_mesh(afx, H.load("naca0012q_coarse_euler_gg_o2")).write_msh(tmp_path / "naca0012q_coarse.msh")
    H.product_mesh(afx, H.load("naca0012q_mid_mesh")).write_msh(tmp_path / "naca0012q_mid.msh")
    ini = tmp_path / "conf.ini"
    ini.write_text(CONF % dict(solver="explicit", tol="1e-3", max_it=2000, alpha_end="1.0", start_cfl="1.5"))
    r = subprocess.run([exes["cli"], "-i", str(ini), "-m", str(tmp_path) + "/", "-q"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    rows = [[float(v) for v in l.split()[1:]] for l in r.stdout.splitlines() if l.startswith("POLAR")]
    assert len(rows) == 1 and 0.05 < rows[0][1] < 0.2  # CL at 1 degree


@pytest.mark.gpu
def test_cli_writes_reference_layout_vtu(afx, gpu, exes, tmp_path):
    """save() (post.h:58-180): one <airfoil>_<alpha>.vtu per angle with the reference's arrays, sized by the mid mesh."""
    H.product_mesh(afx, H.load("naca0012q_coarse_euler_gg_o2")).write_msh(tmp_path / "naca0012q_coarse.msh")
    mid = H.product_mesh(afx, H.load("naca0012q_mid_mesh"))
    mid.write_msh(tmp_path / "naca0012q_mid.msh")
    ini = tmp_path / "conf.ini"
    ini.write_text(CONF % dict(solver="implicit", tol="1e-4", max_it=50, alpha_end="1.0", start_cfl="40.0"))
    out = tmp_path / "vtu"
    out.mkdir()
    r = subprocess.run([exes["cli"], "-i", str(ini), "-m", str(tmp_path) + "/", "-q", "--vtu", str(out) + "/"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    files = sorted(os.listdir(out))
    assert files == ["naca0012q_1.000000.vtu"]  # std::to_string(alpha), rans.h:103
    txt = (out / files[0]).read_text()
    assert 'NumberOfCells="%d"' % mid.N in txt
    for name in ("connectivity", "offsets", "types", "Wall Distance", "Mach", "Density", "Pressure", "Temperature", "Velocity"):
        assert 'Name="%s"' % name in txt
    mach = txt.split('Name="Mach" Format="ascii">')[1].split("</DataArray>")[0].split()
    assert len(mach) == mid.N and 0.0 < min(map(float, mach)) and max(map(float, mach)) < 1.0
