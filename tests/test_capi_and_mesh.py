"""CPU checks of the product's host side: the C-ABI library loads without a
GPU and exports every symbol include/afx_rans.h declares; the host mesh builder
reproduces the reference's arrays (edge order, orientation, ghosts) bit for bit."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(afx):
    hdr = open(os.path.join(ROOT, "include", "afx_rans.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(afx_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    lib = ctypes.CDLL(afx.library_path())
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(afx.EXPORTED_SYMBOLS) == declared
    afx.load_library().afx_version.restype = ctypes.c_char_p
    assert b"sm_100a" in afx.load_library().afx_version()


def test_no_gpu_means_loud_failure_not_fallback(afx):
    if afx.device_count() > 0:
        pytest.skip("a GPU is visible here")
    d = H.load("naca0012q_coarse_euler_gg_o2")
    m = H.product_mesh(afx, d)
    with pytest.raises(afx.AfxError) as e:
        afx.GpuSolver(m)
    assert e.value.code == -2  # AFX_ERR_CUDA


@pytest.mark.parametrize("tag", H.EXPLICIT_CASES[:1] + H.EXPLICIT_CASES[1:2] + H.EXPLICIT_CASES[4:5])
def test_host_mesh_matches_reference(afx, tag):
    d = H.load(tag)
    m = H.product_mesh(afx, d)
    assert (m.N, m.G, m.E) == tuple(int(v) for v in d["sizes"])
    for a in H.MESH_ARRAYS:
        assert H.sha(getattr(m, a)) == str(d["sha_" + a]), a
    assert np.array_equal(m.bnd_patch, d["bpatch"])


def test_msh_round_trip_and_oracle_agreement(afx, tmp_path):
    m = afx.Mesh.synth_omesh(96, 40, 16, 50.0)
    assert m.N == 96 * 16 + 2 * 96 * 24 and m.G == 192
    assert m.E == (4 * 96 * 16 + 3 * 2 * 96 * 24 + 192) // 2
    p = tmp_path / "synth.msh"
    m.write_msh(p)
    m2 = afx.Mesh.read_msh(p)
    for a in H.MESH_ARRAYS:
        assert np.array_equal(getattr(m, a), getattr(m2, a)), a
    assert m2.patch_names == ["wall", "farfield"]
    # independent construction by the oracle (hash map instead of sort)
    x, y, cells, b0, b1 = m.elements()
    om = H.orc.OracleMesh(x, y, cells, m.is_tri, b0, b1, m.bnd_patch, m.patch_names)
    for a in H.MESH_ARRAYS:
        assert np.array_equal(getattr(m, a), getattr(om, a)), a
    assert m.area.min() > 0


def test_benchmark_mesh_sizes_match_survey(afx):
    """SURVEY 8d config 2 at 1/64 scale keeps the tri/quad proportions; the full sizes are N=2^20, E=1704960."""
    m = afx.Mesh.synth_omesh(128, 80, 32, 150.0)
    assert m.N == 2 ** 14 and m.G == 256 and m.E == (4 * 128 * 32 + 3 * 2 * 128 * 48 + 256) // 2


def test_mesh_errors(afx, tmp_path):
    x = np.array([0., 1., 1., 0.]); y = np.array([0., 0., 1., 1.])
    cells = np.array([[0, 1, 2, 3]], np.uint32)
    with pytest.raises(afx.AfxError) as e:  # boundary segment that is not an edge: "invalid edge ref" (mesh.h:757)
        afx.Mesh.from_elements(x, y, cells, [0], [0], [2], [0], ["wall"])
    assert e.value.code == -1 and "invalid edge ref" in str(e.value)
    with pytest.raises(afx.AfxError) as e:
        afx.Mesh.read_msh(tmp_path / "missing.msh")
    assert e.value.code == -4
    bad = tmp_path / "bad.msh"
    bad.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(afx.AfxError):
        afx.Mesh.read_msh(bad)


def test_reference_reader_accepts_our_msh_writer(afx, tmp_path):
    """The CPU reference arm of bench.py feeds synthetic meshes to rans::mesh through this writer."""
    from oracle import ref
    if not os.path.exists(ref.SO):
        pytest.skip("oracle/_ref not built")
    m = afx.Mesh.synth_omesh(64, 24, 8, 30.0)
    p = tmp_path / "s.msh"
    m.write_msh(p)
    rm = ref.RefMesh(str(p))
    for a in H.MESH_ARRAYS:
        assert np.array_equal(getattr(m, a), getattr(rm, a)), a
    assert [m.patch_names[i] for i in m.bnd_patch] == rm.bnd_names


def test_no_entry_point_crashes_on_null_arguments(afx):
    """The ABI is called from C and C++ hosts: a null handle comes back as an error code (AFX_ERR_INVALID for the solver entry points,
    0 / -1 / NULL for the small accessors), never as a crash.  Every function include/afx_rans.h declares is called with nulls, each
    in a child process so that a missing check fails this test instead of taking the test run down."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, re, sys
sys.path.insert(0, %r)
import aeroflex_b200 as afx
L = C.CDLL(afx.library_path())  # a fresh handle: no argtypes, so that nulls can be passed everywhere
hdr = re.sub(r"/\*.*?\*/", "", open(%r).read(), flags=re.S)
solver = sorted(set(re.findall(r"\bint\s+(afx_rans_[a-z0-9_]+)\s*\(\s*afx_rans\s*\*\s*s\b", hdr)))
assert len(solver) >= 35, solver
every = sorted(set(re.findall(r"\b(afx_[a-z0-9_]+)\s*\(", hdr)) - {"afx_pinned_alloc"})
bad = []
for n in every:
    print(n, flush=True)  # the last name printed is the one that crashed
    f = getattr(L, n)
    f.restype = C.c_int
    rc = f(*([None] + [C.c_void_p(0)] * 13))  # nulls / zeros for everything
    if n in solver and rc != -1:
        bad.append((n, rc))
print("checked", len(every), "bad", bad)
assert not bad
''' % (ROOT, os.path.join(ROOT, "include", "afx_rans.h"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-1500:])
