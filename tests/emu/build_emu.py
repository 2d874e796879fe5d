"""TEST INFRASTRUCTURE ONLY: build the host emulation of the CUDA library.

The product's own sources (aeroflex_b200/csrc) are copied into tests/emu/_build with three mechanical rewrites --

  * ``kernel<<<grid, block, smem, stream>>>(args)``  ->  ``AFX_EMU_LAUNCH((kernel), grid, block, smem, stream, args)``
  * the handful of inline-PTX statements -> their emulation (exact-text table below: the build FAILS if the product
    text changes, so the table cannot silently go stale)
  * the dynamic shared-memory declaration -> the emulator's per-block buffer

-- and compiled with g++ against tests/emu/include/cuda_runtime.h.  The result, tests/emu/_build/libaeroflex_rans_emu.so,
exports the same C ABI and is loaded ONLY by tests/test_kernel_emulation.py (through AFX_LIB in a subprocess).  The
package, bench.py and smoke() never load it; the product library still fails loudly without a GPU.

strict mode: g++ -ffp-contract=off evaluates the kernels' expressions in IEEE double precision exactly as nvcc
-fmad=false does, so the emulated strict library must be BIT-IDENTICAL to the oracle.  fast mode: the hardware
reciprocal seeds are emulated (upper word of the exact value) and g++ contracts FMAs differently from nvcc, so only the
north-star tolerance applies.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "aeroflex_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libaeroflex_rans_emu.so")

LAUNCH = re.compile(r"(\bk_\w+(?:<[^<>;()]*>)?)\s*<<<(.+?)>>>\s*\(")

# file -> [(exact product text, emulation text, expected count)]
REWRITES = {
    "rans_physics.cuh": [
        ('asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));', "x = ::afx_emu::rcp_seed(a);", 1),
        ('asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));', "y = ::afx_emu::rsqrt_seed(a);", 1),
    ],
    "rans_kernels.cuh": [
        ('asm volatile("griddepcontrol.launch_dependents;" ::: "memory");', "", 1),
        ('asm volatile("griddepcontrol.wait;" ::: "memory");', "", 1),
        ('asm volatile("prefetch.global.L2 [%0];" ::"l"(p));', "(void)p;", 1),
        ('asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));', "t = ::afx_emu::global_timer_ns();", 1),
        ('asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));', "v = *p;", 1),
    ],
    "rans_pipe.cuh": [
        ('asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");', "v = __atomic_load_n(p, __ATOMIC_ACQUIRE);", 1),
    ],
    "rans_solver.cu": [  # the library says what it is: the package refuses to load it outside the test run, bench.py and smoke() always
        ('return "aeroflex_rans_b200 0.1 (sm_100a)";', 'return "aeroflex_rans_b200 0.1 (HOST EMULATION of the sm_100a sources -- tests only, not a product path)";', 1),
    ],
    "nccl_dl.h": [  # bind the emulated NCCL of emu_runtime.cpp instead of dlopen("libnccl.so.2")
        ('void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // an NCCL the process already has', "void* h = dlopen(::afx_emu::self_path(), RTLD_NOW | RTLD_LOCAL);", 1),
        ('sym(h, "nccl', 'sym(h, "afx_emu_nccl', 10),
    ],
    "rans_stage.cuh": [
        ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', "", 1),
        ("extern __shared__ __align__(128) unsigned char smem[];", "unsigned char* smem = ::afx_emu::g_dyn_smem;", 1),
    ],
}
# rans_stage.cuh: everything between these two lines is PTX wrappers -> emu_stage_ptx.h
STAGE_BEGIN = "// ---- copy-engine and mbarrier primitives (PTX) ----"
STAGE_END = "struct TileView {"

SOURCES = ["rans_types.h", "rans_physics.cuh", "rans_kernels.cuh", "rans_krylov.cuh", "rans_stage.cuh", "rans_pipe.cuh", "rans_kernels_tu.cu",
           "rans_solver.cu", "nccl_dl.h", "ordering.h", "partition.h", "tiling.h", "mesh_host.h"]
HOST_CPP = ["mesh_host.cpp", "ordering.cpp", "tiling.cpp", "partition.cpp", "mesh_capi.cpp"]


def transform(name, text):
    for old, new, count in REWRITES.get(name, []):
        if text.count(old) != count:
            raise RuntimeError("build_emu: %s: expected %d x %r, found %d -- update tests/emu/build_emu.py" % (name, count, old, text.count(old)))
        text = text.replace(old, new)
    if name == "rans_stage.cuh":
        a, b = text.find(STAGE_BEGIN), text.find(STAGE_END)
        if a < 0 or b < a:
            raise RuntimeError("build_emu: rans_stage.cuh: PTX primitive block not found")
        text = text[:a] + '#include "emu_stage_ptx.h"\n' + text[b:]
    out, pos = [], 0
    for m in LAUNCH.finditer(text):
        out.append(text[pos:m.start()])
        out.append("AFX_EMU_LAUNCH((%s), %s, " % (m.group(1), m.group(2)))
        pos = m.end()
    out.append(text[pos:])
    text = "".join(out)
    if "<<<" in text or re.search(r"\basm\b", text):
        raise RuntimeError("build_emu: %s still holds a launch or inline assembly the emulation does not cover" % name)
    text = text.replace('"../../include/afx_rans.h"', '"%s"' % os.path.join(ROOT, "include", "afx_rans.h"))
    return "// GENERATED by tests/emu/build_emu.py from aeroflex_b200/csrc/%s -- host emulation, tests only\n" % name + text


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("build_emu.py", "emu_runtime.cpp")] + \
        [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    for name in SOURCES:
        with open(os.path.join(CSRC, name)) as f:
            text = transform(name, f.read())
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
    extra_defs = os.environ.get("AFX_EMU_DEFINES", "").split()  # e.g. "-DAFX_FLUX_THREADS=128": emulate a tuning variant
    # the Krylov reductions use a fixed grid sized for 148 SMs: a handful of blocks is the same code path on the host's cores
    common = ["g++", "-DAFX_KRY_BLOCKS=8"] + extra_defs + ["-std=c++17", "-O3", "-march=native", "-g", "-fPIC", "-fopenmp", "-fno-strict-aliasing", "-I", os.path.join(HERE, "include"), "-I", OUT, "-I", CSRC,
              "-Wno-unknown-pragmas", "-Wno-attributes"]
    strict = ["-ffp-contract=off"]
    units = [("rans_kernels_tu.cu", "kernels_strict.o", ["-x", "c++", "-DAFX_FAST=0"] + strict, OUT),
             ("rans_kernels_tu.cu", "kernels_fast.o", ["-x", "c++", "-DAFX_FAST=1", "-ffp-contract=fast", "-mfma"], OUT),
             ("rans_solver.cu", "rans_solver.o", ["-x", "c++"] + strict, OUT),
             (os.path.join(HERE, "emu_runtime.cpp"), "emu_runtime.o", strict, HERE)]
    units += [(s, s.replace(".cpp", ".o"), strict, CSRC) for s in HOST_CPP]

    def cc(u):
        src, obj, extra, d = u
        cmd = common + extra + ["-c", os.path.join(d, src), "-o", os.path.join(OUT, obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build_emu: compiling %s failed" % src)
    with ThreadPoolExecutor(len(units)) as ex:
        list(ex.map(cc, units))
    subprocess.run(["g++", "-shared", "-o", LIB] + [os.path.join(OUT, u[1]) for u in units] + ["-lgomp", "-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
