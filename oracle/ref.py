"""ctypes binding of oracle/_ref/libafx_ref.so -- the UNMODIFIED AeroFLEX
reference headers compiled against the Eigen-API stand-in (oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Used to pin the restated oracle, to generate
tests/golden/ and as the CPU "reference" arm of bench.py.  Available wherever
the prebuilt .so is (it travels to the GPU box with the tree); rebuilt only
where /root/reference exists.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# AFX_REF_VARIANT=michalak (set before the import): the same unmodified headers compiled with -DRANS_MICHALAK_LIMITER, the
# reference's compile-time switch to its other limiter (solver.h:557-576); used by oracle/make_golden_michalak.py only
VARIANT = os.environ.get("AFX_REF_VARIANT", "")
SO = os.path.join(HERE, "_ref", "libafx_ref%s.so" % ("_" + VARIANT if VARIANT else ""))

f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

_lib = None


def available():
    return os.path.exists(SO) or os.path.isdir("/root/reference/src/rans/include")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
    L = C.CDLL(SO)
    L.ref_last_error.restype = C.c_char_p
    L.ref_mesh_load.restype = C.c_void_p
    L.ref_mesh_load.argtypes = [C.c_char_p]
    L.ref_mesh_free.argtypes = [C.c_void_p]
    L.ref_mesh_sizes.argtypes = [C.c_void_p, u32p]
    L.ref_mesh_inputs.argtypes = [C.c_void_p, f64p, f64p, u32p, u8p, u32p, u32p]
    L.ref_mesh_arrays.argtypes = [C.c_void_p, u32p, u32p] + [f64p] * 8 + [u32p, u32p]
    L.ref_mesh_boundary_name.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_int]
    L.ref_solver_new.restype = C.c_void_p
    L.ref_solver_new.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double, C.c_double]
    L.ref_solver_free.argtypes = [C.c_void_p]
    L.ref_solver_set_gas.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.ref_solver_add_bc.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p] + [C.c_double] * 4
    L.ref_solver_apply_bcs.argtypes = [C.c_void_p]
    L.ref_solver_set_options.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double, C.c_double]
    L.ref_solver_set_cfl.argtypes = [C.c_void_p, C.c_double]
    for n in ("ref_solver_init", "ref_solver_refill_bcs", "ref_solver_bcs_from_internal", "ref_solver_calc_dt"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.ref_solver_uniform_residual.restype = C.c_double
    L.ref_solver_uniform_residual.argtypes = [C.c_void_p]
    L.ref_solver_get.restype = C.c_long
    L.ref_solver_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    L.ref_solver_set.restype = C.c_long
    L.ref_solver_set.argtypes = [C.c_void_p, C.c_char_p, f64p]
    for n in ("ref_solver_walls", "ref_solver_calc_gradients", "ref_solver_calc_limiters", "ref_solver_calc_residual"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_int]
    L.ref_solver_explicit_solve.restype = C.c_double
    L.ref_solver_explicit_solve.argtypes = [C.c_void_p, C.c_double]
    L.ref_solver_implicit_rhs.restype = C.c_double
    L.ref_solver_implicit_rhs.argtypes = [C.c_void_p]
    L.ref_solver_implicit_lhs.argtypes = [C.c_void_p, f64p, f64p, f64p]
    L.ref_solver_implicit_step.restype = C.c_double
    L.ref_solver_implicit_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    L.ref_solver_wall_forces.argtypes = [C.c_void_p, C.c_char_p, f64p]
    L.ref_flux.argtypes = [C.c_int, f64p, C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, f64p]
    L.ref_vars.argtypes = [C.c_int, f64p, C.c_double, C.c_double, f64p, f64p, f64p]
    L.ref_fd_jacobian.argtypes = [C.c_int, f64p, C.c_int, C.c_double, C.c_double, f64p, f64p, f64p, f64p, f64p]
    L.ref_get_conservative.argtypes = [C.c_double] * 4 + [f64p, f64p]
    L.ref_run_sweep.restype = C.c_int
    L.ref_run_sweep.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_char_p,
                                C.c_int, f64p, f64p, f64p, f64p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int]
    L.ref_prolongate.argtypes = [C.c_char_p, C.c_char_p, f64p, f64p]
    _lib = L
    return L


def gas5(gamma=1.4, R=0.71428571428, mu_L=1e-5, Pr_L=0.72, cp=1.0):
    return np.array([gamma, R, mu_L, Pr_L, cp], np.float64)


class RefMesh:
    """rans::mesh(filename) -- the reference's own MSH 4.1 reader and metrics."""

    def __init__(self, path):
        L = lib()
        self.h = L.ref_mesh_load(path.encode())
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())
        sz = np.zeros(5, np.uint32)
        L.ref_mesh_sizes(self.h, sz)
        self.nn, self.N, self.G, self.E, _ = (int(v) for v in sz)
        NT = self.N + self.G
        self.x = np.zeros(self.nn); self.y = np.zeros(self.nn)
        self.cells = np.zeros((self.N, 4), np.uint32); is_tri = np.zeros(self.N, np.uint8)
        self.b0 = np.zeros(self.G, np.uint32); self.b1 = np.zeros(self.G, np.uint32)
        L.ref_mesh_inputs(self.h, self.x, self.y, self.cells, is_tri, self.b0, self.b1)
        self.is_tri = is_tri
        self.edge_cells = np.zeros((self.E, 2), np.uint32); self.edge_nodes = np.zeros((self.E, 2), np.uint32)
        for n in ("enx", "eny", "elen", "ecx", "ecy"):
            setattr(self, n, np.zeros(self.E))
        for n in ("ccx", "ccy", "area"):
            setattr(self, n, np.zeros(NT))
        self.cell_edges = np.zeros((self.N, 4), np.uint32); self.bnd_edge = np.zeros(self.G, np.uint32)
        L.ref_mesh_arrays(self.h, self.edge_cells, self.edge_nodes, self.enx, self.eny, self.elen, self.ecx, self.ecy,
                          self.ccx, self.ccy, self.area, self.cell_edges, self.bnd_edge)
        buf = C.create_string_buffer(256)
        names = []
        for b in range(self.G):
            L.ref_mesh_boundary_name(self.h, b, buf, 256)
            names.append(buf.value.decode())
        self.bnd_names = names
        self.patch_names = sorted(set(names))
        self.bpatch = np.array([self.patch_names.index(n) for n in names], np.int32)

    def __del__(self):
        try:
            lib().ref_mesh_free(self.h)
        except Exception:
            pass


class RefSolver:
    """rans::explicitSolver / rans::implicitSolver of the reference itself."""

    def __init__(self, mesh, implicit=False, viscosity="inviscid", gamma=1.4, R=0.71428571428,
                 mu_L=1e-5, Pr_L=0.72, cp=1.0):
        L = lib()
        self.mesh = mesh
        self.h = L.ref_solver_new(mesh.h, int(implicit), viscosity.encode(), gamma, R)
        if not self.h:
            raise RuntimeError(L.ref_last_error().decode())
        L.ref_solver_set_gas(self.h, mu_L, Pr_L, cp)
        self.n4 = 4 * (mesh.N + mesh.G)

    def __del__(self):
        try:
            lib().ref_solver_free(self.h)
        except Exception:
            pass

    def set_bcs(self, bcs):
        L = lib()
        for nm, (typ, v) in bcs.items():
            v = v or {}
            L.ref_solver_add_bc(self.h, nm.encode(), typ.encode(), v.get("mach", 0.2), v.get("angle", 0.0),
                                v.get("T", 1.0), v.get("p", 1.0))
        if L.ref_solver_apply_bcs(self.h):
            raise KeyError(L.ref_last_error().decode())

    def set_options(self, second_order=True, gradient="green-gauss", limiter_k=5.0, cfl=1.0):
        if lib().ref_solver_set_options(self.h, int(second_order), gradient.encode(), limiter_k, cfl):
            raise RuntimeError(lib().ref_last_error().decode())

    def set_cfl(self, cfl): lib().ref_solver_set_cfl(self.h, cfl)
    def init(self): lib().ref_solver_init(self.h)
    def refill_bcs(self): lib().ref_solver_refill_bcs(self.h)
    def bcs_from_internal(self): lib().ref_solver_bcs_from_internal(self.h)
    def calc_dt(self): lib().ref_solver_calc_dt(self.h)
    def walls(self, which=0): lib().ref_solver_walls(self.h, which)
    def calc_gradients(self, which=0): lib().ref_solver_calc_gradients(self.h, which)
    def calc_limiters(self, which=0): lib().ref_solver_calc_limiters(self.h, which)
    def calc_residual(self, which=0): lib().ref_solver_calc_residual(self.h, which)
    def explicit_solve(self, relaxation=1.0): return lib().ref_solver_explicit_solve(self.h, relaxation)
    def implicit_rhs(self): return lib().ref_solver_implicit_rhs(self.h)
    def implicit_step(self, relaxation, tol, rhs_iterations): return lib().ref_solver_implicit_step(self.h, relaxation, tol, rhs_iterations)
    def uniform_residual(self): return lib().ref_solver_uniform_residual(self.h)

    def get(self, name):
        n = lib().ref_solver_get(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(n)
        lib().ref_solver_get(self.h, name.encode(), out.ctypes.data_as(C.c_void_p))
        return out

    def set(self, name, v):
        if lib().ref_solver_set(self.h, name.encode(), np.ascontiguousarray(v, np.float64)) < 0:
            raise KeyError(name)

    def implicit_lhs(self):
        NT, E = self.mesh.N + self.mesh.G, self.mesh.E
        d = np.zeros(16 * NT); o01 = np.zeros(16 * E); o10 = np.zeros(16 * E)
        if lib().ref_solver_implicit_lhs(self.h, d, o01, o10):
            raise RuntimeError(lib().ref_last_error().decode())
        return d.reshape(NT, 4, 4), o01.reshape(E, 4, 4), o10.reshape(E, 4, 4)

    def wall_forces(self, patch):
        out = np.zeros(3)
        if lib().ref_solver_wall_forces(self.h, patch.encode(), out):
            raise RuntimeError(lib().ref_last_error().decode())
        return tuple(out)


def flux(kind, g5, viscous_type, nx, ny, qL, qR, gx=None, gy=None):
    z = np.zeros(4); f = np.zeros(4)
    lib().ref_flux(kind, g5, viscous_type, nx, ny, np.ascontiguousarray(qL, np.float64), np.ascontiguousarray(qR, np.float64),
                   z if gx is None else np.ascontiguousarray(gx, np.float64), z if gy is None else np.ascontiguousarray(gy, np.float64), f)
    return f


def bc_vars(kind, g5, nx, ny, qL, qbc):
    r = np.zeros(4)
    lib().ref_vars(kind, g5, nx, ny, np.ascontiguousarray(qL, np.float64), np.ascontiguousarray(qbc, np.float64), r)
    return r


def fd_jacobian(kind, g5, viscous_type, nx, ny, qL, qR, gx=None, gy=None):
    z = np.zeros(4); J = np.zeros(64)
    lib().ref_fd_jacobian(kind, g5, viscous_type, nx, ny, np.ascontiguousarray(qL, np.float64), np.ascontiguousarray(qR, np.float64),
                          z if gx is None else np.ascontiguousarray(gx, np.float64), z if gy is None else np.ascontiguousarray(gy, np.float64), J)
    return J.reshape(8, 8)


def get_conservative(mach, angle, T, p, g5):
    q = np.zeros(4)
    lib().ref_get_conservative(mach, angle, T, p, g5, q)
    return q


def run_sweep(mesh_paths, alphas_deg, implicit=True, viscosity="inviscid", gradient="green-gauss", second_order=True,
              relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, tolerance=1e-4, rhs_iterations=5,
              max_iterations=300, limiter_k=5.0, gamma=1.4, R=0.71428571428, mach=0.2, T=1.0, p=1.0,
              wall_type="slip-wall", quiet=True):
    """The reference's own FMG + alpha sweep (multigrid.h + rans.h:78-106) on the given .msh files."""
    n = len(alphas_deg)
    al = np.ascontiguousarray(alphas_deg, np.float64)
    cl = np.zeros(n); cd = np.zeros(n); cm = np.zeros(n)
    iters = (C.c_int * n)(); secs = C.c_double(0)
    paths = (C.c_char_p * len(mesh_paths))(*[p_.encode() for p_ in mesh_paths])
    rc = lib().ref_run_sweep(len(mesh_paths), paths, int(implicit), viscosity.encode(), gradient.encode(), int(second_order),
                             relaxation, start_cfl, slope_cfl, max_cfl, tolerance, rhs_iterations, max_iterations,
                             limiter_k, gamma, R, mach, T, p, wall_type.encode(), n, al, cl, cd, cm, iters,
                             C.byref(secs), int(quiet))
    if rc < 0:
        raise RuntimeError(lib().ref_last_error().decode())
    return dict(cl=cl, cd=cd, cm=cm, iters=np.array(list(iters)), seconds=secs.value, done=rc)


def prolongate(coarse_path, fine_path, q_coarse, n_fine4):
    """The reference's own FMG prolongation (multigrid.h:100-178): q_fine = mapper * q_coarse."""
    out = np.zeros(n_fine4)
    if lib().ref_prolongate(coarse_path.encode(), fine_path.encode(), np.ascontiguousarray(q_coarse, np.float64), out):
        raise RuntimeError(lib().ref_last_error().decode())
    return out
