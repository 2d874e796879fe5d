"""ctypes binding of include/afx_rans.h (libaeroflex_rans_b200.so).

``GpuSolver`` mirrors the public surface of the reference's ``rans::solver``
(src/rans/include/rans/solver.h:104-175: set_bcs, set_cfl, set_second_order,
set_gradient_scheme, set_limiter_k, init, refill_bcs, bcs_from_internal, get_q,
get_uniform_residual, fill, compute, solve) with the same argument meaning and
error behaviour (unknown patch -> KeyError like ``bcs.at``; numeric failure ->
negative return from solve).  All arithmetic happens in the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIBNAME = "libaeroflex_rans_b200.so"

BC_KINDS = {"farfield": 1, "slip-wall": 2, "wall": 3, "inlet-outlet": 4}  # solver.h:220-227, 603-606; anything else -> 0 (internal)
VISCOSITY = {"inviscid": 0, "laminar": 1, "spallart-allmaras": 2}  # core.h:176 (spelling is the reference's)
GRADIENT = {"green-gauss": 0, "least-squares": 1}               # core.h:175
FIELDS = {"q": 0, "qW": 1, "gx": 2, "gy": 3, "limiters": 4, "dt": 5, "rhs": 6}

EXPORTED_SYMBOLS = [
    "afx_last_error", "afx_version", "afx_device_count", "afx_pinned_alloc", "afx_pinned_free",
    "afx_mesh_read_msh", "afx_mesh_from_elements", "afx_mesh_synth_omesh", "afx_mesh_free", "afx_mesh_get_desc",
    "afx_mesh_n_nodes", "afx_mesh_n_patches", "afx_mesh_patch_name", "afx_mesh_patch_id", "afx_mesh_get_elements",
    "afx_mesh_write_msh",
    "afx_partition_create", "afx_partition_free", "afx_partition_get_desc", "afx_partition_info", "afx_partition_cell_l2g",
    "afx_partition_edge_l2g", "afx_partition_peer", "afx_nccl_unique_id", "afx_rans_create_partitioned",
    "afx_rans_p2p_export", "afx_rans_p2p_connect", "afx_rans_halo_mode",
    "afx_group_create", "afx_group_free", "afx_group_abort", "afx_rans_create_partitioned_group",
    "afx_rans_create", "afx_rans_destroy", "afx_rans_set_bcs", "afx_rans_set_options", "afx_rans_set_limiter", "afx_rans_set_cfl",
    "afx_rans_set_math_mode", "afx_rans_get_math_mode", "afx_rans_set_fused", "afx_rans_tile_info", "afx_rans_set_pipelined", "afx_rans_pipe_info", "afx_tiling_plan", "afx_tiling_plan_partition",
    "afx_rans_init", "afx_rans_refill_bcs", "afx_rans_bcs_from_internal", "afx_rans_set_q", "afx_rans_get_q",
    "afx_rans_set_q_local", "afx_rans_get_q_local", "afx_rans_get_field", "afx_rans_boundary_variables", "afx_rans_uniform_residual", "afx_rans_step_explicit",
    "afx_rans_run_explicit", "afx_rans_phase_dt_gradients", "afx_rans_phase_limiters", "afx_rans_phase_residual",
    "afx_rans_residual", "afx_rans_fill_jacobian", "afx_rans_get_jacobian_blocks", "afx_rans_step_implicit", "afx_rans_compute",
    "afx_rans_set_linear_solver", "afx_rans_last_linear_iterations",
    "afx_rans_wall_forces", "afx_rans_wall_cp", "afx_rans_sweep", "afx_rans_sweep_fmg", "afx_prolongation_create", "afx_prolongation_free",
    "afx_prolongation_apply", "afx_rans_last_device_ms", "afx_rans_launch_count",
    "afx_rans_profile_explicit", "afx_rans_profile_halo_ms",
]


class AfxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("afx error %d: %s" % (code, msg))
        self.code = code


class Gas(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("gamma", "R", "mu_L", "Pr_L", "cp")]


class BVars(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mach", "angle", "T", "p")]


class SweepSettings(C.Structure):
    _fields_ = [("implicit", C.c_int), ("relaxation", C.c_double), ("start_cfl", C.c_double), ("slope_cfl", C.c_double),
                ("max_cfl", C.c_double), ("tolerance", C.c_double), ("rhs_iterations", C.c_int), ("max_iterations", C.c_int)]


class MeshDesc(C.Structure):
    _fields_ = [("n_cells", C.c_uint32), ("n_ghost", C.c_uint32), ("n_edges", C.c_uint32),
                ("edges_cells", C.POINTER(C.c_uint32)),
                ("edges_nx", C.POINTER(C.c_double)), ("edges_ny", C.POINTER(C.c_double)),
                ("edges_len", C.POINTER(C.c_double)), ("edges_cx", C.POINTER(C.c_double)),
                ("edges_cy", C.POINTER(C.c_double)),
                ("cells_cx", C.POINTER(C.c_double)), ("cells_cy", C.POINTER(C.c_double)),
                ("cells_area", C.POINTER(C.c_double)),
                ("cells_edges", C.POINTER(C.c_uint32)), ("cells_is_tri", C.POINTER(C.c_uint8)),
                ("boundary_edges", C.POINTER(C.c_uint32)), ("boundary_patch", C.POINTER(C.c_int32))]


def library_path():
    """AFX_LIB overrides the path (used to A/B differently tuned builds of the same sources)."""
    return os.environ.get("AFX_LIB") or os.path.join(LIBDIR, LIBNAME)


NVCC_COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ccbin", "/usr/bin/g++",
               "-Xcompiler", "-fPIC,-fopenmp,-O3"]
# (source, object, extra flags): the kernels are compiled twice, once per arithmetic mode
UNITS = [("rans_kernels_tu.cu", "kernels_strict.o", ["-DAFX_FAST=0", "-fmad=false"]),
         ("rans_kernels_tu.cu", "kernels_fast.o", ["-DAFX_FAST=1", "-fmad=true"]),
         ("rans_solver.cu", "rans_solver.o", ["-fmad=false"]),
         ("mesh_host.cpp", "mesh_host.o", []),
         ("ordering.cpp", "ordering.o", []),
         ("tiling.cpp", "tiling.o", []),
         ("partition.cpp", "partition.o", []),
         ("mesh_capi.cpp", "mesh_capi.o", [])]


def build_library(force=False, verbose=False, defines=(), out=None):
    """nvcc cross-compiles the library for sm_100a (works without a GPU).  `defines` (e.g. ["-DAFX_FLUX_MINB=4"])
    and `out` build a tuning variant next to the default library."""
    from concurrent.futures import ThreadPoolExecutor
    variant = out is not None
    out = out or os.path.join(LIBDIR, LIBNAME)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp", ".cuh", ".h"))] + \
        [os.path.join(ROOT, "include", "afx_rans.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(LIBDIR, exist_ok=True)
    bdir = os.path.join(PKG, "build", os.path.basename(out) if variant else "default")
    os.makedirs(bdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

    def cc(unit):
        src, obj, extra = unit
        cmd = [nvcc] + NVCC_COMMON + extra + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", os.path.join(bdir, obj)]
        subprocess.run(cmd, check=True, cwd=CSRC)
    with ThreadPoolExecutor(len(UNITS)) as ex:
        list(ex.map(cc, UNITS))
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-shared", "-o", out] +
                   [os.path.join(bdir, u[1]) for u in UNITS] + ["-lgomp", "-ldl"], check=True)
    return out


_lib = None


def is_emulation():
    """True if the loaded library is the host emulation of tests/emu (bench.py and smoke() refuse to run on it)."""
    return b"EMULATION" in load_library().afx_version()


def load_library():
    """Load the CUDA library; raises (never falls back) if it is absent and cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        build_library()
    L = C.CDLL(path)
    L.afx_last_error.restype = C.c_char_p
    L.afx_version.restype = C.c_char_p
    if b"EMULATION" in L.afx_version() and os.environ.get("AFX_ALLOW_EMULATION") != "tests":
        # tests/emu builds the kernel sources for the host so that the CPU test run can execute them; that library is
        # test infrastructure and must never stand in for the CUDA library
        raise RuntimeError("%s is the host emulation of the kernel sources (tests/emu): it is loaded by the test run only "
                           "(AFX_ALLOW_EMULATION=tests); the product has no CPU path" % path)
    vp, dp, u32p, u8p, i32p = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    L.afx_pinned_alloc.restype = C.c_void_p
    L.afx_pinned_alloc.argtypes = [C.c_size_t]
    L.afx_pinned_free.argtypes = [vp]
    L.afx_mesh_read_msh.argtypes = [C.POINTER(vp), C.c_char_p]
    L.afx_mesh_from_elements.argtypes = [C.POINTER(vp), C.c_uint32, vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp,
                                         C.c_int, C.POINTER(C.c_char_p)]
    L.afx_mesh_synth_omesh.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
    L.afx_mesh_free.argtypes = [vp]
    L.afx_mesh_get_desc.argtypes = [vp, C.POINTER(MeshDesc)]
    L.afx_mesh_n_nodes.restype = C.c_uint32
    L.afx_mesh_n_nodes.argtypes = [vp]
    L.afx_mesh_n_patches.argtypes = [vp]
    L.afx_mesh_patch_name.restype = C.c_char_p
    L.afx_mesh_patch_name.argtypes = [vp, C.c_int]
    L.afx_mesh_patch_id.argtypes = [vp, C.c_char_p]
    L.afx_mesh_get_elements.argtypes = [vp] * 6
    L.afx_mesh_write_msh.argtypes = [vp, C.c_char_p]
    L.afx_rans_create.argtypes = [C.POINTER(vp), C.POINTER(MeshDesc), C.POINTER(Gas), C.c_int, C.c_int]
    L.afx_partition_create.argtypes = [C.POINTER(vp), C.POINTER(MeshDesc), C.c_int, C.c_int]
    L.afx_partition_free.argtypes = [vp]
    L.afx_partition_get_desc.argtypes = [vp, C.POINTER(MeshDesc)]
    L.afx_partition_info.argtypes = [vp, u32p]
    L.afx_partition_cell_l2g.restype = u32p
    L.afx_partition_cell_l2g.argtypes = [vp]
    L.afx_partition_edge_l2g.restype = u32p
    L.afx_partition_edge_l2g.argtypes = [vp]
    L.afx_partition_peer.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(u32p), u32p, C.POINTER(u32p), u32p]
    L.afx_nccl_unique_id.argtypes = [C.c_char_p]
    L.afx_rans_p2p_export.argtypes = [vp, vp, C.POINTER(C.c_size_t)]
    L.afx_rans_p2p_connect.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int]
    L.afx_rans_halo_mode.argtypes = [vp]
    L.afx_rans_create_partitioned.argtypes = [C.POINTER(vp), vp, C.POINTER(Gas), C.c_int, C.c_int, C.c_char_p]
    L.afx_group_create.argtypes = [C.POINTER(vp), C.c_int]
    L.afx_group_free.argtypes = [vp]
    L.afx_group_abort.argtypes = [vp]
    L.afx_rans_create_partitioned_group.argtypes = [C.POINTER(vp), vp, C.POINTER(Gas), C.c_int, C.c_int, vp]
    L.afx_rans_destroy.argtypes = [vp]
    L.afx_rans_set_bcs.argtypes = [vp, C.c_int, vp, C.POINTER(BVars)]
    L.afx_rans_set_options.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.afx_rans_set_limiter.argtypes = [vp, C.c_int]
    L.afx_rans_set_cfl.argtypes = [vp, C.c_double]
    L.afx_rans_set_math_mode.argtypes = [vp, C.c_int]
    L.afx_tiling_plan.argtypes = [C.POINTER(MeshDesc), C.c_uint32, vp, C.POINTER(C.c_uint32), vp, C.c_uint32, C.POINTER(C.c_uint64)]
    L.afx_tiling_plan_partition.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint32), vp, C.c_uint32, C.POINTER(C.c_uint64)]
    L.afx_rans_set_fused.argtypes = [vp, C.c_int]
    L.afx_rans_tile_info.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.afx_rans_set_pipelined.argtypes = [vp, C.c_int]
    L.afx_rans_pipe_info.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.afx_rans_get_math_mode.argtypes = [vp]
    for n in ("afx_rans_init", "afx_rans_refill_bcs", "afx_rans_bcs_from_internal", "afx_rans_phase_dt_gradients",
              "afx_rans_phase_limiters", "afx_rans_fill_jacobian"):
        getattr(L, n).argtypes = [vp]
    L.afx_rans_set_q.argtypes = [vp, vp]
    L.afx_rans_get_q.argtypes = [vp, vp]
    L.afx_rans_get_field.argtypes = [vp, C.c_int, vp]
    L.afx_rans_set_q_local.argtypes = [vp, vp]
    L.afx_rans_get_q_local.argtypes = [vp, vp]
    L.afx_rans_boundary_variables.argtypes = [vp, C.POINTER(BVars)]
    L.afx_rans_uniform_residual.argtypes = [vp, dp]
    L.afx_rans_step_explicit.argtypes = [vp, C.c_double, dp]
    L.afx_rans_run_explicit.argtypes = [vp, C.c_double, C.c_int, vp]
    L.afx_rans_phase_residual.argtypes = [vp, dp]
    L.afx_rans_residual.argtypes = [vp, dp]
    L.afx_rans_get_jacobian_blocks.argtypes = [vp, vp, vp, vp]
    L.afx_rans_step_implicit.argtypes = [vp, C.c_double, C.c_double, C.c_int, dp]
    L.afx_rans_compute.argtypes = [vp]
    L.afx_rans_set_linear_solver.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int]
    L.afx_rans_last_linear_iterations.argtypes = [vp]
    L.afx_rans_wall_forces.argtypes = [vp, C.c_int, vp]
    L.afx_rans_wall_cp.argtypes = [vp, C.c_int, vp]
    L.afx_rans_sweep.argtypes = [vp, C.POINTER(SweepSettings), C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.afx_prolongation_create.argtypes = [C.POINTER(vp), vp, vp, vp, vp, vp]
    L.afx_prolongation_free.argtypes = [vp]
    L.afx_prolongation_apply.argtypes = [vp]
    L.afx_rans_sweep_fmg.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int, C.POINTER(SweepSettings), C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.afx_rans_last_device_ms.argtypes = [vp, dp]
    L.afx_rans_launch_count.restype = C.c_int64
    L.afx_rans_launch_count.argtypes = [vp]
    L.afx_rans_profile_explicit.argtypes = [vp, C.c_double, C.c_int, vp]
    L.afx_rans_profile_halo_ms.argtypes = [vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc < 0:
        raise AfxError(rc, load_library().afx_last_error().decode())
    return rc


def device_count():
    return load_library().afx_device_count()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pinned_array(n, dtype=np.float64):
    """numpy array over cudaMallocHost memory (kept alive by the returned array's base)."""
    L = load_library()
    nbytes = int(n) * np.dtype(dtype).itemsize
    p = L.afx_pinned_alloc(nbytes)
    if not p:
        raise AfxError(-2, L.afx_last_error().decode())
    buf = (C.c_char * nbytes).from_address(p)
    a = np.frombuffer(buf, dtype=dtype, count=int(n))
    _pinned_keep.append((p, buf))
    return a


_pinned_keep = []


class Mesh:
    """Host mesh in the reference's layout (rans::mesh, mesh.h:198-273)."""

    def __init__(self, handle):
        self.L = load_library()
        self.h = handle
        self.d = MeshDesc()
        _check(self.L.afx_mesh_get_desc(self.h, C.byref(self.d)))
        d = self.d
        self.N, self.G, self.E = d.n_cells, d.n_ghost, d.n_edges
        self.patch_names = [self.L.afx_mesh_patch_name(self.h, p).decode() for p in range(self.L.afx_mesh_n_patches(self.h))]

    @classmethod
    def read_msh(cls, path):
        L = load_library()
        h = C.c_void_p()
        _check(L.afx_mesh_read_msh(C.byref(h), str(path).encode()))
        return cls(h)

    @classmethod
    def from_elements(cls, x, y, cells, is_tri, b0, b1, bpatch, patch_names):
        L = load_library()
        x = np.ascontiguousarray(x, np.float64); y = np.ascontiguousarray(y, np.float64)
        cells = np.ascontiguousarray(cells, np.uint32); is_tri = np.ascontiguousarray(is_tri, np.uint8)
        b0 = np.ascontiguousarray(b0, np.uint32); b1 = np.ascontiguousarray(b1, np.uint32)
        bpatch = np.ascontiguousarray(bpatch, np.int32)
        names = (C.c_char_p * max(len(patch_names), 1))(*[n.encode() for n in patch_names])
        h = C.c_void_p()
        _check(L.afx_mesh_from_elements(C.byref(h), len(x), _ptr(x), _ptr(y), len(is_tri), _ptr(cells), _ptr(is_tri),
                                        len(b0), _ptr(b0), _ptr(b1), _ptr(bpatch), len(patch_names), names))
        return cls(h)

    @classmethod
    def synth_omesh(cls, ni, nj, n_quad_layers, far_radius=150.0):
        L = load_library()
        h = C.c_void_p()
        _check(L.afx_mesh_synth_omesh(C.byref(h), ni, nj, n_quad_layers, far_radius))
        return cls(h)

    def __del__(self):
        try:
            self.L.afx_mesh_free(self.h)
        except Exception:
            pass

    def _arr(self, ptr, n, shape=None):
        a = np.ctypeslib.as_array(ptr, shape=(n,)) if n else np.zeros(0)
        return a.reshape(shape) if shape else a

    # numpy views of the reference arrays (borrowed from the C++ object)
    @property
    def edge_cells(self): return self._arr(self.d.edges_cells, 2 * self.E, (-1, 2))
    @property
    def enx(self): return self._arr(self.d.edges_nx, self.E)
    @property
    def eny(self): return self._arr(self.d.edges_ny, self.E)
    @property
    def elen(self): return self._arr(self.d.edges_len, self.E)
    @property
    def ecx(self): return self._arr(self.d.edges_cx, self.E)
    @property
    def ecy(self): return self._arr(self.d.edges_cy, self.E)
    @property
    def ccx(self): return self._arr(self.d.cells_cx, self.N + self.G)
    @property
    def ccy(self): return self._arr(self.d.cells_cy, self.N + self.G)
    @property
    def area(self): return self._arr(self.d.cells_area, self.N + self.G)
    @property
    def cell_edges(self): return self._arr(self.d.cells_edges, 4 * self.N, (-1, 4))
    @property
    def is_tri(self): return self._arr(self.d.cells_is_tri, self.N)
    @property
    def bnd_edge(self): return self._arr(self.d.boundary_edges, self.G)
    @property
    def bnd_patch(self): return self._arr(self.d.boundary_patch, self.G)

    def elements(self):
        nn = self.L.afx_mesh_n_nodes(self.h)
        x = np.zeros(nn); y = np.zeros(nn)
        cells = np.zeros((self.N, 4), np.uint32); b0 = np.zeros(self.G, np.uint32); b1 = np.zeros(self.G, np.uint32)
        self.L.afx_mesh_get_elements(self.h, _ptr(x), _ptr(y), _ptr(cells), _ptr(b0), _ptr(b1))
        return x, y, cells, b0, b1

    def write_msh(self, path):
        _check(self.L.afx_mesh_write_msh(self.h, str(path).encode()))


class Partition:
    """One rank's piece of a mesh: owned | ring 1 | ring 2 cells + its boundary ghosts, and the exchange plan."""

    def __init__(self, mesh, nranks, rank):
        self.L = load_library()
        self.global_mesh = mesh
        self.h = C.c_void_p()
        _check(self.L.afx_partition_create(C.byref(self.h), C.byref(mesh.d), nranks, rank))
        info = (C.c_uint32 * 8)()
        self.L.afx_partition_info(self.h, info)
        self.n_own, self.n_r1, self.n_r2, self.n_bc, self.n_edges, self.n_peers, self.rank, self.nranks = (int(v) for v in info)
        self.d = MeshDesc()
        _check(self.L.afx_partition_get_desc(self.h, C.byref(self.d)))
        self.N, self.G, self.E = self.d.n_cells, self.d.n_ghost, self.d.n_edges
        self.patch_names = list(mesh.patch_names)
        self.cell_l2g = np.ctypeslib.as_array(self.L.afx_partition_cell_l2g(self.h), shape=(self.N + self.G,)).copy()
        self.edge_l2g = np.ctypeslib.as_array(self.L.afx_partition_edge_l2g(self.h), shape=(self.E,)).copy()
        self.peers = []
        for i in range(self.n_peers):
            r = C.c_int(); ns = C.c_uint32(); nr = C.c_uint32()
            sp = C.POINTER(C.c_uint32)(); rp = C.POINTER(C.c_uint32)()
            _check(self.L.afx_partition_peer(self.h, i, C.byref(r), C.byref(sp), C.byref(ns), C.byref(rp), C.byref(nr)))
            send = np.ctypeslib.as_array(sp, shape=(ns.value,)).copy() if ns.value else np.zeros(0, np.uint32)
            recv = np.ctypeslib.as_array(rp, shape=(nr.value,)).copy() if nr.value else np.zeros(0, np.uint32)
            self.peers.append((r.value, send, recv))

    def __del__(self):
        try:
            self.L.afx_partition_free(self.h)
        except Exception:
            pass

    def local_arrays(self):
        """numpy copies of the local mesh arrays (reference layout) -- what an independent solver needs."""
        d, NT, E, N = self.d, self.N + self.G, self.E, self.N
        a = lambda p, n: np.ctypeslib.as_array(p, shape=(n,)).copy()
        return dict(edge_cells=a(d.edges_cells, 2 * E).reshape(-1, 2), enx=a(d.edges_nx, E), eny=a(d.edges_ny, E), elen=a(d.edges_len, E),
                    ecx=a(d.edges_cx, E), ecy=a(d.edges_cy, E), ccx=a(d.cells_cx, NT), ccy=a(d.cells_cy, NT), area=a(d.cells_area, NT),
                    cell_edges=a(d.cells_edges, 4 * N).reshape(-1, 4), is_tri=a(d.cells_is_tri, N),
                    bnd_edge=a(d.boundary_edges, self.G) if self.G else np.zeros(0, np.uint32),
                    bnd_patch=a(d.boundary_patch, self.G) if self.G else np.zeros(0, np.int32))


class Prolongation:
    """FMG prolongation between two GpuSolvers on the device (CSR: row_begin[n_fine+1], col, w in the meshes' reference order)."""

    def __init__(self, coarse, fine, row_begin, col, w):
        self.L = load_library()
        self.coarse, self.fine = coarse, fine
        rb = np.ascontiguousarray(row_begin, np.uint32); c = np.ascontiguousarray(col, np.uint32); ww = np.ascontiguousarray(w, np.float64)
        self.h = C.c_void_p()
        _check(self.L.afx_prolongation_create(C.byref(self.h), coarse.h, fine.h, _ptr(rb), _ptr(c), _ptr(ww)))

    def apply(self):
        _check(self.L.afx_prolongation_apply(self.h))

    def __del__(self):
        try:
            self.L.afx_prolongation_free(self.h)
        except Exception:
            pass


def sweep_fmg(levels, prolongations, alphas_deg, implicit=True, relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, tolerance=1e-4,
              rhs_iterations=5, max_iterations=300, farfield="farfield", wall="wall", reinit=True):
    """Rans::run_airfoil over several mesh levels, on the device (afx_rans_sweep_fmg)."""
    L = load_library()
    al = np.ascontiguousarray(alphas_deg, dtype=np.float64)
    n = len(al)
    st = SweepSettings(int(bool(implicit)), relaxation, start_cfl, slope_cfl, max_cfl, tolerance, rhs_iterations, max_iterations)
    cl, cd, cm, res = (np.full(n, np.nan) for _ in range(4))
    it = np.zeros(n, np.int32)
    names = levels[0].mesh.patch_names
    lv = (C.c_void_p * len(levels))(*[s.h for s in levels])
    pr = (C.c_void_p * max(1, len(prolongations)))(*[p.h for p in prolongations])
    rc = L.afx_rans_sweep_fmg(lv, pr, len(levels), C.byref(st), names.index(farfield), names.index(wall), _ptr(al), n, int(bool(reinit)),
                              _ptr(cl), _ptr(cd), _ptr(cm), _ptr(it), _ptr(res))
    if rc not in (0, -3):
        _check(rc)
    return dict(cl=cl, cd=cd, cm=cm, iterations=it, residual=res, status=rc)


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(load_library().afx_nccl_unique_id(buf))
    return buf.raw


def tiling_plan(mesh, tile_cells, limits=None):
    """Host-only: the shared-memory tiles the fused stage kernel would use for `mesh` (verified against the connectivity).
    Returns (per_tile[n_tiles, 4] = own, ring-1, state-only cells and local faces, shared-memory bytes)."""
    L = load_library()
    cap = mesh.N // 8 + 64
    n = C.c_uint32()
    smem = C.c_uint64()
    per = np.zeros((cap, 4), dtype=np.uint32)
    lim = np.asarray(limits, dtype=np.uint32) if limits is not None else None
    limp = lim.ctypes.data_as(C.c_void_p) if lim is not None else None
    if isinstance(mesh, Partition):
        _check(L.afx_tiling_plan_partition(mesh.h, tile_cells, limp, C.byref(n), per.ctypes.data_as(C.c_void_p), cap, C.byref(smem)))
    else:
        _check(L.afx_tiling_plan(C.byref(mesh.d), tile_cells, limp, C.byref(n), per.ctypes.data_as(C.c_void_p), cap, C.byref(smem)))
    return per[:n.value].copy(), int(smem.value)


class Group:
    """In-process communicator for N partitioned solvers of this process (one host thread per solver): stands where NCCL
    stands, so that all ranks can share one device."""

    def __init__(self, nranks):
        self.L = load_library()
        self.h = C.c_void_p()
        self.nranks = nranks
        _check(self.L.afx_group_create(C.byref(self.h), nranks))

    def abort(self):
        self.L.afx_group_abort(self.h)

    def __del__(self):
        try:
            self.L.afx_group_free(self.h)
        except Exception:
            pass


def run_ranks(fns):
    """Run one callable per rank, each on its own host thread (ctypes releases the GIL inside the library), and re-raise
    the first failure.  The collective calls of an in-process group meet this way."""
    import threading
    errs = [None] * len(fns)
    outs = [None] * len(fns)

    def go(i):
        try:
            outs[i] = fns[i]()
        except BaseException as e:  # noqa: BLE001 -- handed to the caller
            errs[i] = e
    ts = [threading.Thread(target=go, args=(i,)) for i in range(len(fns))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return outs


class GpuSolver:
    """rans::solver on one B200 through the C ABI."""

    def __init__(self, mesh, gas=None, viscosity="inviscid", device=0, math=None, nccl_id=None, group=None):
        """math: "strict" (bit-identical to the CPU reference), "fast" (default; shared reciprocals + FMA) or None
        (library default / AFX_MATH).  group: an in-process Group instead of NCCL (mesh = Partition)."""
        self.L = load_library()
        self.mesh = mesh
        g = gas or {}
        self.gas = Gas(g.get("gamma", 1.4), g.get("R", 0.71428571428), g.get("mu_L", 1e-5), g.get("Pr_L", 0.72), g.get("cp", 1.0))
        if viscosity not in VISCOSITY:
            raise KeyError(viscosity)
        self.h = C.c_void_p()
        if isinstance(mesh, Partition):
            # one piece of a partitioned mesh: mesh = Partition, nccl_id = the 128-byte id made on rank 0
            self.partition = mesh
            self.group = group  # keeps the communicator alive as long as a member
            if group is not None:
                _check(self.L.afx_rans_create_partitioned_group(C.byref(self.h), mesh.h, C.byref(self.gas), VISCOSITY[viscosity], device, group.h))
            else:
                _check(self.L.afx_rans_create_partitioned(C.byref(self.h), mesh.h, C.byref(self.gas), VISCOSITY[viscosity], device, nccl_id))
            mesh = mesh.global_mesh
            self.mesh = mesh
        else:
            _check(self.L.afx_rans_create(C.byref(self.h), C.byref(mesh.d), C.byref(self.gas), VISCOSITY[viscosity], device))
        self.n4 = 4 * (mesh.N + mesh.G)
        self.bcs = {}
        if math is not None:
            _check(self.L.afx_rans_set_math_mode(self.h, {"strict": 0, "fast": 1}[math]))

    @property
    def math(self):
        return "strict" if self.L.afx_rans_get_math_mode(self.h) == 0 else "fast"

    def __del__(self):
        try:
            if self.h:
                self.L.afx_rans_destroy(self.h)
        except Exception:
            pass

    # bcs: {patch name: (type string, dict(mach, angle, T, p) or None)} like rans::Settings::bcs
    def set_bcs(self, bcs):
        names = self.mesh.patch_names
        used = set(int(p) for p in np.unique(self.mesh.bnd_patch))
        kinds = np.zeros(max(len(names), 1), np.uint8)
        vars_ = (BVars * max(len(names), 1))()
        for i, nm in enumerate(names):
            if nm not in bcs:
                if i in used:
                    raise KeyError(nm)  # std::out_of_range from bcs.at() in the reference
                continue
            typ, v = bcs[nm]
            v = v or {}
            kinds[i] = BC_KINDS.get(typ, 0)
            vars_[i] = BVars(v.get("mach", 0.2), v.get("angle", 0.0), v.get("T", 1.0), v.get("p", 1.0))
        self.bcs = dict(bcs)
        _check(self.L.afx_rans_set_bcs(self.h, len(names), _ptr(kinds), vars_))

    def set_options(self, second_order=True, gradient="green-gauss", limiter_k=5.0, cfl=None):
        _check(self.L.afx_rans_set_options(self.h, int(bool(second_order)), GRADIENT[gradient], limiter_k))
        if cfl is not None:
            self.set_cfl(cfl)

    def set_limiter(self, name="venkatakrishnan"):
        """calc_limiters' function: "venkatakrishnan" (the reference's default build) or "michalak" (its RANS_MICHALAK_LIMITER build)."""
        _check(self.L.afx_rans_set_limiter(self.h, {"venkatakrishnan": 0, "michalak": 1}[name]))

    def set_cfl(self, cfl): _check(self.L.afx_rans_set_cfl(self.h, cfl))
    def init(self): _check(self.L.afx_rans_init(self.h))
    def refill_bcs(self): _check(self.L.afx_rans_refill_bcs(self.h))
    def bcs_from_internal(self): _check(self.L.afx_rans_bcs_from_internal(self.h))

    def set_q(self, q):
        q = np.ascontiguousarray(q, np.float64)
        if q.size != self.n4:
            raise ValueError("q must have 4*(N+G) entries")
        _check(self.L.afx_rans_set_q(self.h, _ptr(q)))

    def get_q(self, out=None):
        out = np.empty(self.n4) if out is None else out
        _check(self.L.afx_rans_get_q(self.h, _ptr(out)))
        return out

    # peer-memory halo (CUDA IPC): blob = p2p_export(); all-gather the blobs in rank order; p2p_connect(blobs)
    def p2p_export(self):
        n = C.c_size_t()
        _check(self.L.afx_rans_p2p_export(self.h, None, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(self.L.afx_rans_p2p_export(self.h, buf, C.byref(n)))
        return buf.raw

    def p2p_connect(self, blobs):
        joined = b"".join(blobs)
        _check(self.L.afx_rans_p2p_connect(self.h, joined, len(blobs[0]), len(blobs)))

    def halo_mode(self):
        return {0: "none", 1: "nccl", 2: "p2p"}[self.L.afx_rans_halo_mode(self.h)]

    # a partitioned solver's own piece (owned | ring 1 | ring 2 | ghosts), no host-side gather through the global vector
    def local_size4(self):
        p = getattr(self, "partition", None)
        return 4 * (p.N + p.G) if p is not None else self.n4

    def set_q_local(self, q):
        q = np.ascontiguousarray(q, np.float64)
        if q.size != self.local_size4():
            raise ValueError("local state has the wrong length")
        _check(self.L.afx_rans_set_q_local(self.h, _ptr(q)))

    def get_q_local(self, out=None):
        out = np.empty(self.local_size4()) if out is None else out
        _check(self.L.afx_rans_get_q_local(self.h, _ptr(out)))
        return out

    def get(self, name):
        out = np.empty(self.mesh.N + self.mesh.G if name == "dt" else self.n4)
        _check(self.L.afx_rans_get_field(self.h, FIELDS[name], _ptr(out)))
        return out

    def get_uniform_residual(self):
        v = C.c_double()
        _check(self.L.afx_rans_uniform_residual(self.h, C.byref(v)))
        return v.value

    # explicitSolver: fill() and compute() are no-ops returning 0 (solver.h:738-739)
    def fill(self): return None
    def compute(self): return 0

    def solve(self, relaxation=1.0, tol=0.0, rhs_iterations=5):
        """explicitSolver::solve: one iteration, returns ||qW||_2 (or -1 on numeric failure)."""
        v = C.c_double()
        rc = self.L.afx_rans_step_explicit(self.h, relaxation, C.byref(v))
        if rc == -3:
            return -1.0
        _check(rc)
        return v.value

    def run(self, n_iter, relaxation=1.0):
        norms = np.zeros(n_iter)
        _check(self.L.afx_rans_run_explicit(self.h, relaxation, n_iter, _ptr(norms)))
        return norms

    def phase_dt_gradients(self): _check(self.L.afx_rans_phase_dt_gradients(self.h))
    def phase_limiters(self): _check(self.L.afx_rans_phase_limiters(self.h))

    def phase_residual(self):
        v = C.c_double()
        _check(self.L.afx_rans_phase_residual(self.h, C.byref(v)))
        return v.value

    def residual(self):
        v = C.c_double()
        _check(self.L.afx_rans_residual(self.h, C.byref(v)))
        return v.value

    def fill_jacobian(self): _check(self.L.afx_rans_fill_jacobian(self.h))

    # implicitSolver surface (solver.h:966-968): fill() / compute() / solve(relaxation, tol, rhs_iterations)
    def implicit_fill(self): self.fill_jacobian()

    def implicit_compute(self):
        rc = self.L.afx_rans_compute(self.h)
        if rc == -3:
            return -1
        _check(rc)
        return 0

    def implicit_solve(self, relaxation=1.0, tol=0.0, rhs_iterations=5):
        v = C.c_double()
        rc = self.L.afx_rans_step_implicit(self.h, relaxation, tol, rhs_iterations, C.byref(v))
        if rc == -3:
            return -1.0
        _check(rc)
        return v.value

    def set_linear_solver(self, restart=30, max_iterations=500, tolerance=1e-2, precond_sweeps=4):
        _check(self.L.afx_rans_set_linear_solver(self.h, restart, max_iterations, tolerance, precond_sweeps))

    def last_linear_iterations(self):
        return int(self.L.afx_rans_last_linear_iterations(self.h))

    def jacobian_blocks(self):
        NT, E = self.mesh.N + self.mesh.G, self.mesh.E
        d = np.zeros((NT, 4, 4)); o01 = np.zeros((E, 4, 4)); o10 = np.zeros((E, 4, 4))
        _check(self.L.afx_rans_get_jacobian_blocks(self.h, _ptr(d), _ptr(o01), _ptr(o10)))
        return d, o01, o10

    def wall_forces(self, patch_name):
        out = np.zeros(3)
        p = self.mesh.patch_names.index(patch_name) if patch_name in self.mesh.patch_names else -1
        _check(self.L.afx_rans_wall_forces(self.h, p, _ptr(out)))
        return tuple(out)  # cl, cd, cm

    def sweep(self, alphas_deg, implicit=True, relaxation=0.9, start_cfl=40.0, slope_cfl=50.0, max_cfl=100.0, tolerance=1e-4,
              rhs_iterations=5, max_iterations=300, farfield="farfield", wall="wall", reinit=True):
        """Rans::run_airfoil's angle loop on this mesh level (rans.h:86-104): returns dict(cl, cd, cm, iterations, residual)."""
        al = np.ascontiguousarray(alphas_deg, dtype=np.float64)
        n = len(al)
        st = SweepSettings(int(bool(implicit)), relaxation, start_cfl, slope_cfl, max_cfl, tolerance, rhs_iterations, max_iterations)
        cl, cd, cm, res = (np.full(n, np.nan) for _ in range(4))
        it = np.zeros(n, np.int32)
        names = self.mesh.patch_names
        rc = self.L.afx_rans_sweep(self.h, C.byref(st), names.index(farfield), names.index(wall), _ptr(al), n, int(bool(reinit)),
                                   _ptr(cl), _ptr(cd), _ptr(cm), _ptr(it), _ptr(res))
        if rc not in (0, -3):
            _check(rc)
        return dict(cl=cl, cd=cd, cm=cm, iterations=it, residual=res, status=rc)

    def wall_cp(self, patch_name):
        p = self.mesh.patch_names.index(patch_name)
        n = _check(self.L.afx_rans_wall_cp(self.h, p, None))
        cp = np.zeros(n)
        _check(self.L.afx_rans_wall_cp(self.h, p, _ptr(cp)))
        return cp

    def last_device_ms(self):
        v = C.c_double()
        self.L.afx_rans_last_device_ms(self.h, C.byref(v))
        return v.value

    def launch_count(self):
        return int(self.L.afx_rans_launch_count(self.h))

    def profile_explicit(self, n_iter=5, relaxation=1.0):
        out = np.zeros(6)
        _check(self.L.afx_rans_profile_explicit(self.h, relaxation, n_iter, _ptr(out)))
        return dict(dt_grad=out[0], limiter=out[1], flux=out[2], gather_update=out[3], halo_exchange=out[4], stage=out[5])

    def profile_halo_ms(self):
        out = np.zeros(2)
        _check(self.L.afx_rans_profile_halo_ms(self.h, _ptr(out)))
        return dict(signal=out[0], wait_scatter=out[1])

    def set_fused(self, on):
        """One fused kernel per Runge-Kutta stage (default) or limiter / flux / gather+update as three kernels."""
        _check(self.L.afx_rans_set_fused(self.h, 1 if on else 0))

    def set_pipelined(self, on):
        """One persistent kernel per Runge-Kutta stage sweeping L2-resident chunks, or limiter / flux / gather+update as three kernels."""
        _check(self.L.afx_rans_set_pipelined(self.h, 1 if on else 0))

    def pipe_info(self):
        out = (C.c_uint64 * 8)()
        _check(self.L.afx_rans_pipe_info(self.h, out))
        return dict(active=bool(out[0]), chunk_cells=int(out[1]), chunks=int(out[2]), items=int(out[3]), far_faces=int(out[4]), far_cells=int(out[5]),
                    lag_flux=int(out[6] >> 32), lag_update=int(out[6] & 0xFFFFFFFF), ctas=int(out[7]))

    def tile_info(self):
        out = (C.c_uint64 * 8)()
        _check(self.L.afx_rans_tile_info(self.h, out))
        return dict(fused=bool(out[0]), tiles=out[1], tile_cells=out[2], smem_bytes=out[3], ctas_per_sm=out[4], max_local_cells=out[5],
                    max_local_faces=out[6], local_cells=out[7])
