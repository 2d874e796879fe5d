"""ctypes binding of the restated CPU oracle (oracle/rans_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(aeroflex_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

INTERNAL, FARFIELD, SLIPWALL, WALL = 0, 1, 2, 3
GREEN_GAUSS, LEAST_SQUARES = 0, 1
KIND_OF = {"farfield": FARFIELD, "slip-wall": SLIPWALL, "wall": WALL, "inlet-outlet": 4}  # 4: solver.h:603-606
VISC_OF = {"inviscid": 0, "laminar": 1, "spallart-allmaras": 2}


class Gas(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gamma", "R", "mu_L", "Pr_L", "cp")]

    @staticmethod
    def default(gamma=1.4, R=0.71428571428, mu_L=1e-5, Pr_L=0.72, cp=1.0):
        return Gas(gamma, R, mu_L, Pr_L, cp)


class BVars(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mach", "angle", "T", "p")]


class Mesh(C.Structure):
    _fields_ = [("N", C.c_uint32), ("G", C.c_uint32), ("E", C.c_uint32),
                ("edge_cells", C.POINTER(C.c_uint32)),
                ("enx", C.POINTER(C.c_double)), ("eny", C.POINTER(C.c_double)), ("elen", C.POINTER(C.c_double)),
                ("ecx", C.POINTER(C.c_double)), ("ecy", C.POINTER(C.c_double)),
                ("ccx", C.POINTER(C.c_double)), ("ccy", C.POINTER(C.c_double)), ("area", C.POINTER(C.c_double)),
                ("cell_edges", C.POINTER(C.c_uint32)), ("is_tri", C.POINTER(C.c_uint8)),
                ("bnd_edge", C.POINTER(C.c_uint32)), ("bnd_patch", C.POINTER(C.c_int32))]


class Solver(C.Structure):
    _fields_ = [("m", Mesh), ("g", Gas),
                ("edge_kind", C.POINTER(C.c_uint8)), ("bnd_kind", C.POINTER(C.c_uint8)),
                ("bnd_vars", C.POINTER(BVars)),
                ("viscous_type", C.c_int), ("visc_not_inviscid", C.c_int), ("second_order", C.c_int),
                ("gradient_scheme", C.c_int), ("limiter_k", C.c_double), ("cfl", C.c_double)] + \
               [(n, C.POINTER(C.c_double)) for n in ("q", "qk", "qW", "gx", "gy", "lim", "qmin", "qmax", "rhs", "dt", "lsq")] + \
               [("cf_sorted", C.POINTER(C.c_uint32)), ("fluxbuf", C.POINTER(C.c_double)), ("limiter_kind", C.c_int)]


_libs = {}


def build():
    subprocess.run(["make", "-s", "-C", HERE, os.path.join(HERE, "liborc.so"), os.path.join(HERE, "liborc_fast.so")],
                   check=True)


def lib(fast=False):
    """liborc.so (parity build, no FMA) or liborc_fast.so (-O3 -march=native -fopenmp, timing)."""
    name = "liborc_fast.so" if fast else "liborc.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_pressure.restype = C.c_double
    L.orc_pressure.argtypes = [f64p, C.c_double]
    L.orc_flux_internal.argtypes = [C.POINTER(Gas), C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, f64p]
    L.orc_bc_vars.argtypes = [C.c_int, C.POINTER(Gas), C.c_double, C.c_double, f64p, f64p, f64p]
    L.orc_flux.argtypes = [C.c_int, C.POINTER(Gas), C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, f64p]
    L.orc_fd_jacobian.argtypes = [C.c_int, C.POINTER(Gas), C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, f64p]
    L.orc_get_conservative.argtypes = [C.POINTER(BVars), C.POINTER(Gas), f64p]
    L.orc_mesh_build.restype = C.c_int
    L.orc_mesh_build.argtypes = [C.POINTER(Mesh), C.c_uint32, f64p, f64p, C.c_uint32, u32p, u8p, C.c_uint32, u32p, u32p, i32p]
    L.orc_mesh_free.argtypes = [C.POINTER(Mesh)]
    L.orc_solver_init.restype = C.c_int
    L.orc_solver_init.argtypes = [C.POINTER(Solver), C.POINTER(Mesh), C.POINTER(Gas), C.c_int]
    L.orc_solver_free.argtypes = [C.POINTER(Solver)]
    L.orc_set_bcs.argtypes = [C.POINTER(Solver), C.c_int, u8p, C.POINTER(BVars)]
    L.orc_set_gradient_scheme.argtypes = [C.POINTER(Solver), C.c_int]
    for n in ("orc_init_field", "orc_refill_bcs", "orc_bcs_from_internal", "orc_calc_dt", "orc_calc_gradients"):
        getattr(L, n).argtypes = [C.POINTER(Solver)]
    L.orc_set_walls_from_internal.argtypes = [C.POINTER(Solver), C.POINTER(C.c_double)]
    L.orc_calc_limiters.argtypes = [C.POINTER(Solver), C.POINTER(C.c_double)]
    L.orc_explicit_residual.argtypes = [C.POINTER(Solver), C.POINTER(C.c_double)]
    L.orc_explicit_solve.restype = C.c_double
    L.orc_explicit_solve.argtypes = [C.POINTER(Solver), C.c_double]
    if hasattr(L, "orc_explicit_solve_omp"):
        L.orc_explicit_solve_omp.restype = C.c_double
        L.orc_explicit_solve_omp.argtypes = [C.POINTER(Solver), C.c_double]
    L.orc_implicit_rhs.restype = C.c_double
    L.orc_implicit_rhs.argtypes = [C.POINTER(Solver)]
    L.orc_uniform_residual.restype = C.c_double
    L.orc_uniform_residual.argtypes = [C.POINTER(Solver)]
    L.orc_implicit_lhs.argtypes = [C.POINTER(Solver), f64p, f64p, f64p]
    L.orc_boundary_variables.argtypes = [C.POINTER(Solver), C.POINTER(BVars)]
    L.orc_wall_forces.restype = C.c_int
    L.orc_wall_forces.argtypes = [C.POINTER(Solver), C.c_int, f64p]
    _libs[name] = L
    return L


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


class OracleMesh:
    """Reference-layout geometry built by orc_mesh_build (mesh.h semantics)."""

    def __init__(self, x, y, cells, is_tri, b0, b1, bpatch, patch_names, fast=False):
        self.L = lib(fast)
        self.m = Mesh()
        x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
        cells = np.ascontiguousarray(cells, np.uint32).reshape(-1, 4)
        is_tri = np.ascontiguousarray(is_tri, np.uint8)
        b0 = np.ascontiguousarray(b0, np.uint32); b1 = np.ascontiguousarray(b1, np.uint32)
        bpatch = np.ascontiguousarray(bpatch, np.int32)
        self.inputs = dict(x=x, y=y, cells=cells, is_tri=is_tri, b0=b0, b1=b1, bpatch=bpatch)
        self.patch_names = list(patch_names)
        rc = self.L.orc_mesh_build(C.byref(self.m), len(x), x, y, len(cells), cells, is_tri, len(b0),
                                   b0 if len(b0) else np.zeros(1, np.uint32), b1 if len(b1) else np.zeros(1, np.uint32),
                                   bpatch if len(bpatch) else np.zeros(1, np.int32))
        if rc:
            raise ValueError("invalid edge ref: a boundary segment is not an edge of the mesh")
        m = self.m
        self.N, self.G, self.E = m.N, m.G, m.E
        NT = m.N + m.G
        self.edge_cells = _view(m.edge_cells, 2 * m.E, np.uint32).reshape(-1, 2)
        self.enx = _view(m.enx, m.E, np.float64); self.eny = _view(m.eny, m.E, np.float64)
        self.elen = _view(m.elen, m.E, np.float64)
        self.ecx = _view(m.ecx, m.E, np.float64); self.ecy = _view(m.ecy, m.E, np.float64)
        self.ccx = _view(m.ccx, NT, np.float64); self.ccy = _view(m.ccy, NT, np.float64)
        self.area = _view(m.area, NT, np.float64)
        self.cell_edges = _view(m.cell_edges, 4 * m.N, np.uint32).reshape(-1, 4)
        self.is_tri = _view(m.is_tri, NT, np.uint8)
        self.bnd_edge = _view(m.bnd_edge, m.G, np.uint32)
        self.bnd_patch = _view(m.bnd_patch, m.G, np.int32)

    def __del__(self):
        try:
            if getattr(self, "_owned", True):
                self.L.orc_mesh_free(C.byref(self.m))
        except Exception:
            pass

    @classmethod
    def from_arrays(cls, a, N, G, patch_names, fast=False):
        """Wrap ready-made reference-layout arrays (e.g. one rank's piece of a partitioned mesh) without rebuilding."""
        self = cls.__new__(cls)
        self.L = lib(fast)
        self._owned = False
        self.patch_names = list(patch_names)
        E = len(a["enx"])
        keep = dict(edge_cells=np.ascontiguousarray(a["edge_cells"], np.uint32).reshape(-1, 2),
                    cell_edges=np.ascontiguousarray(a["cell_edges"], np.uint32).reshape(-1, 4),
                    is_tri=np.ascontiguousarray(np.concatenate([a["is_tri"], np.ones(G, np.uint8)]), np.uint8),
                    bnd_edge=np.ascontiguousarray(a["bnd_edge"], np.uint32), bnd_patch=np.ascontiguousarray(a["bnd_patch"], np.int32))
        for n in ("enx", "eny", "elen", "ecx", "ecy", "ccx", "ccy", "area"):
            keep[n] = np.ascontiguousarray(a[n], np.float64)
        self._keep = keep
        m = Mesh()
        m.N, m.G, m.E = N, G, E
        for n, t in (("edge_cells", C.c_uint32), ("cell_edges", C.c_uint32), ("is_tri", C.c_uint8), ("bnd_edge", C.c_uint32),
                     ("bnd_patch", C.c_int32), ("enx", C.c_double), ("eny", C.c_double), ("elen", C.c_double), ("ecx", C.c_double),
                     ("ecy", C.c_double), ("ccx", C.c_double), ("ccy", C.c_double), ("area", C.c_double)):
            setattr(m, n, keep[n].ctypes.data_as(C.POINTER(t)))
            setattr(self, n, keep[n])
        self.m = m
        self.N, self.G, self.E = N, G, E
        return self


class OracleSolver:
    """Mirrors rans::solver / explicitSolver / implicitSolver (solver.h) on the oracle."""

    def __init__(self, mesh, gas=None, viscosity="inviscid", fast=False):
        self.L = lib(fast)
        self.mesh = mesh
        self.gas = gas or Gas.default()
        self.s = Solver()
        if self.L.orc_solver_init(C.byref(self.s), C.byref(mesh.m), C.byref(self.gas), VISC_OF[viscosity]):
            raise MemoryError("orc_solver_init")
        n4 = 4 * (mesh.N + mesh.G)
        for n in ("q", "qk", "qW", "gx", "gy", "lim", "rhs"):
            setattr(self, n, _view(getattr(self.s, n), n4, np.float64))
        self.dt = _view(self.s.dt, mesh.N + mesh.G, np.float64)
        self.bcs = {}

    def __del__(self):
        try:
            self.L.orc_solver_free(C.byref(self.s))
        except Exception:
            pass

    # bcs: {patch name: (type string, dict(mach, angle, T, p))}, as rans::Settings::bcs
    def set_bcs(self, bcs):
        self.bcs = dict(bcs)
        names = self.mesh.patch_names
        kinds = np.zeros(max(len(names), 1), np.uint8)
        vars_ = (BVars * max(len(names), 1))()
        for i, nm in enumerate(names):
            if nm not in bcs:
                raise KeyError(nm)  # bcs.at(name) throws std::out_of_range in the reference
            typ, v = bcs[nm]
            kinds[i] = KIND_OF.get(typ, INTERNAL)
            v = v or {}
            vars_[i] = BVars(v.get("mach", 0.2), v.get("angle", 0.0), v.get("T", 1.0), v.get("p", 1.0))
        self.L.orc_set_bcs(C.byref(self.s), len(names), kinds, vars_)

    def set_options(self, second_order=True, gradient="green-gauss", limiter_k=5.0, cfl=1.0):
        self.s.second_order = int(bool(second_order))
        self.L.orc_set_gradient_scheme(C.byref(self.s), GREEN_GAUSS if gradient == "green-gauss" else LEAST_SQUARES)
        self.s.limiter_k = limiter_k
        self.s.cfl = cfl

    def set_cfl(self, cfl):
        self.s.cfl = cfl

    def set_limiter(self, name="venkatakrishnan"):
        """"venkatakrishnan": the reference's default build; "michalak": its RANS_MICHALAK_LIMITER build (solver.h:557-576)."""
        self.s.limiter_kind = {"venkatakrishnan": 0, "michalak": 1}[name]

    def init(self): self.L.orc_init_field(C.byref(self.s))
    def refill_bcs(self): self.L.orc_refill_bcs(C.byref(self.s))
    def bcs_from_internal(self): self.L.orc_bcs_from_internal(C.byref(self.s))
    def calc_dt(self): self.L.orc_calc_dt(C.byref(self.s))
    def walls(self, which=0): self.L.orc_set_walls_from_internal(C.byref(self.s), self.s.qk if which else self.s.q)
    def calc_gradients(self): self.L.orc_calc_gradients(C.byref(self.s))
    def calc_limiters(self, which=0): self.L.orc_calc_limiters(C.byref(self.s), self.s.qk if which else self.s.q)
    def calc_residual(self, which=0): self.L.orc_explicit_residual(C.byref(self.s), self.s.qk if which else self.s.q)
    def explicit_solve(self, relaxation=1.0): return self.L.orc_explicit_solve(C.byref(self.s), relaxation)
    def explicit_solve_omp(self, relaxation=1.0): return self.L.orc_explicit_solve_omp(C.byref(self.s), relaxation)
    def implicit_rhs(self): return self.L.orc_implicit_rhs(C.byref(self.s))
    def uniform_residual(self): return self.L.orc_uniform_residual(C.byref(self.s))

    def implicit_lhs(self):
        NT, E = self.mesh.N + self.mesh.G, self.mesh.E
        d = np.zeros(16 * NT); o01 = np.zeros(16 * E); o10 = np.zeros(16 * E)
        self.L.orc_implicit_lhs(C.byref(self.s), d, o01, o10)
        return d.reshape(NT, 4, 4), o01.reshape(E, 4, 4), o10.reshape(E, 4, 4)

    def wall_forces(self, patch_name):
        out = np.zeros(3)
        p = self.mesh.patch_names.index(patch_name) if patch_name in self.mesh.patch_names else -1
        self.L.orc_wall_forces(C.byref(self.s), p, out)
        return tuple(out)  # cl, cd, cm


def flux(kind, gas, viscous_type, nx, ny, qL, qR, gx=None, gy=None, fast=False):
    z = np.zeros(4)
    f = np.zeros(4)
    lib(fast).orc_flux(kind, C.byref(gas), viscous_type, nx, ny, np.ascontiguousarray(qL, np.float64),
                       np.ascontiguousarray(qR, np.float64), z if gx is None else np.ascontiguousarray(gx, np.float64),
                       z if gy is None else np.ascontiguousarray(gy, np.float64), f)
    return f


def bc_vars(kind, gas, nx, ny, qL, qbc):
    r = np.zeros(4)
    lib().orc_bc_vars(kind, C.byref(gas), nx, ny, np.ascontiguousarray(qL, np.float64), np.ascontiguousarray(qbc, np.float64), r)
    return r


def fd_jacobian(kind, gas, viscous_type, nx, ny, qL, qR, gx=None, gy=None):
    z = np.zeros(4)
    J = np.zeros(64)
    lib().orc_fd_jacobian(kind, C.byref(gas), viscous_type, nx, ny, np.ascontiguousarray(qL, np.float64),
                          np.ascontiguousarray(qR, np.float64), z if gx is None else np.ascontiguousarray(gx, np.float64),
                          z if gy is None else np.ascontiguousarray(gy, np.float64), J)
    return J.reshape(8, 8)


def get_conservative(mach, angle, T, p, gas):
    q = np.zeros(4)
    v = BVars(mach, angle, T, p)
    lib().orc_get_conservative(C.byref(v), C.byref(gas), q)
    return q
