"""Shared helpers: golden fixtures -> oracle objects / product objects."""
import ast
import hashlib
import os

import numpy as np

from oracle import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    """sha256 of the bytes; floating-point arrays are hashed with -0.0 folded into +0.0 (the reference leaves
    -0.0 in the ghost rows of gx/gy through `*= 0`, solver.h:467-468; the sign of a zero is not a result)."""
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a + 0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def load(tag):
    d = dict(np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False))
    d["patch_names"] = [str(s) for s in d["patch_names"]] if "patch_names" in d else []
    if "meta" in d:
        try:
            d["meta"] = ast.literal_eval(str(d["meta"]))
        except Exception:
            d["meta"] = str(d["meta"])
    return d


def oracle_mesh(d, fast=False):
    return orc.OracleMesh(d["x"], d["y"], d["cells"], d["is_tri"], d["b0"], d["b1"], d["bpatch"], d["patch_names"], fast=fast)


def product_mesh(afx, d):
    return afx.Mesh.from_elements(d["x"], d["y"], d["cells"], d["is_tri"], d["b0"], d["b1"], d["bpatch"], d["patch_names"])


MESH_ARRAYS = ("edge_cells", "enx", "eny", "elen", "ecx", "ecy", "ccx", "ccy", "area", "cell_edges", "bnd_edge")

EXPLICIT_CASES = ["naca0012q_coarse_euler_gg_o2", "naca0012_coarse_laminar_lsq_o2", "naca0012_coarse_sa_gg_o1",
                  "naca0012_coarse_euler_gg_o1", "flat_plate_laminar_gg_o2", "flat_plate_sa_gg_o2"]
IMPLICIT_CASES = ["naca0012q_coarse_implicit_blocks", "naca0012_coarse_implicit_laminar_blocks"]
# the reference headers compiled with -DRANS_MICHALAK_LIMITER (oracle/make_golden_michalak.py)
MICHALAK_CASES = ["michalak_naca0012q_coarse_euler_gg", "michalak_naca0012_coarse_laminar_lsq"]


def setup_solver(s, meta, cfl=None):
    """Apply a fixture's settings to an OracleSolver or a GpuSolver (same interface)."""
    s.set_bcs(meta["bcs"])
    s.set_options(meta["second_order"], meta["gradient"], meta.get("limiter_k", 5.0), meta["cfl"] if cfl is None else cfl)
    if "limiter" in meta:
        s.set_limiter(meta["limiter"])


def synth_state(mesh_N, q_uniform, seed=12345, amp=1e-3):
    rng = np.random.default_rng(seed)
    q = q_uniform.copy()
    q[:4 * mesh_N] *= 1.0 + amp * rng.uniform(-1, 1, 4 * mesh_N)
    return q
