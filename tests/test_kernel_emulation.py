"""The product's kernel and solver SOURCES under host emulation (tests/emu): CPU-side coverage of the device code.

tests/emu/build_emu.py compiles aeroflex_b200/csrc/*.cu|*.cuh with g++ against an emulated CUDA runtime (every CUDA
thread a fiber, real __syncthreads / warp-shuffle rendezvous, asynchronous copies deferred to the kernel's own waits,
stream capture and graph replay) into tests/emu/_build/libaeroflex_rans_emu.so, and this file runs the single-GPU
parity tests of test_gpu_parity.py / test_gpu_fused.py against that library in a subprocess (AFX_LIB).  In strict
mode the emulated kernels must be BIT-IDENTICAL to the reference's golden vectors and to the oracle, exactly like the
CUDA build on a B200.

This is test infrastructure, not a CPU path of the product: the package never loads the emulation library by itself
(checked below), bench.py and smoke() never see it, and the numbers that count are the `-m gpu` runs on the B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


def run_gpu_tests_under_emulation(emu_lib, files, select, timeout=900, extra_env=None):
    env = dict(os.environ, AFX_LIB=emu_lib, AFX_ALLOW_EMULATION="tests")
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k", select] + [os.path.join(ROOT, "tests", f) for f in files]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).strip().splitlines()[-25:])
    assert r.returncode == 0, "parity tests failed under host emulation:\n" + tail
    return tail


def test_package_never_picks_the_emulation_library(afx, emu_lib):
    assert "AFX_LIB" not in os.environ, "the CPU test run itself must use the CUDA library"
    assert os.path.realpath(afx.library_path()) != os.path.realpath(emu_lib)
    assert os.path.dirname(os.path.realpath(emu_lib)).startswith(os.path.join(ROOT, "tests", "emu"))
    assert not afx.is_emulation()
    # outside the test run the package refuses the emulation library, and bench.py / smoke() refuse it always
    code = "import aeroflex_b200 as a; a.load_library()"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, AFX_LIB=emu_lib), capture_output=True, text=True)
    assert r.returncode != 0 and "host emulation" in r.stderr
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                       env=dict(os.environ, AFX_LIB=emu_lib, AFX_ALLOW_EMULATION="tests"), capture_output=True, text=True)
    assert r.returncode != 0 and "only the CUDA library is measured" in (r.stderr + r.stdout)


def test_three_kernel_stage_sources_match_reference_under_emulation(emu_lib):
    # explicit histories / single phases / synthetic mixed meshes / RHS and Jacobian blocks / BC handling / graph replay
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_parity.py"], "not implicit_converged and not full_size and not sweep_entry and not (arnoldi and fast)")  # the implicit end-to-end cases take ~1 min each: GPU only
    assert " passed" in tail


def test_fused_stage_kernel_sources_under_emulation_with_deferred_copies(emu_lib):
    # the bulk-copy / cp.async pipeline of k_stage: copies land only at the waits the kernel itself executes
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_fused.py"], "not fast_mode")
    assert " passed" in tail


def test_one_cta_walks_many_tiles_under_emulation(emu_lib):
    # one emulated SM: a single persistent CTA pair walks every tile, so every buffer hand-over of the pipeline is exercised
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_fused.py"], "bit_identical and (48 or 192)", extra_env={"AFX_EMU_SMS": "1"})
    assert " passed" in tail


def test_edge_cases_under_emulation(emu_lib):
    # supersonic / transonic boundary branches, meshes smaller than a CTA, limiter extremes, NaN handling
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_edge_cases.py"], "not ring_wraps and not 100_iterations")  # 120 000 iterations / 100 iterations on 64k cells: GPU only
    assert " passed" in tail


MULTI_ENV = {"AFX_EMU_DEVICES": "8", "OMP_NUM_THREADS": "2", "OMP_WAIT_POLICY": "passive"}


def test_partitioned_runs_under_emulation(emu_lib):
    """Ranks are processes, peer memory is a shared mapping of the other process's "device" allocation, NCCL is a
    shared-memory segment: the halo push from the update kernel, the flag hand-off, the split launches of the NCCL mode
    (captured across two streams) and the partition-aware k_dt_grad / k_limiter ranges run for real.  Strict mode:
    bit-identical to the single-device run."""
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_multi.py"], "2-strict-nccl-0 or 8-strict-p2p-0", extra_env=MULTI_ENV)
    assert "2 passed" in tail  # at 8 ranks some pieces hold no far-field edge: init() must still use the whole mesh's far-field state


def test_overlapped_peer_memory_halo_under_emulation(emu_lib):
    # AFX_HALO_OVERLAP=1: send layer first, exchange on the halo stream under the interior update, interior limiter first
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_multi.py"], "2-strict-p2p-0", extra_env=dict(MULTI_ENV, AFX_HALO_OVERLAP="1"))
    assert "1 passed" in tail


def test_graph_partitioned_run_under_emulation(emu_lib):
    # AFX_PARTITION=graph: recursive graph bisection instead of Hilbert chunks; same bit-identity to the single-device run
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_multi.py"], "4-strict-p2p-1", extra_env=dict(MULTI_ENV, AFX_PARTITION="graph"))
    assert "1 passed" in tail


def test_partitioned_variants_under_emulation(emu_lib):
    # 3 ranks, laminar face gradients + least squares; 3 ranks first order; partitioned implicit right-hand side
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_multi.py"], "variants and 3-laminar", extra_env=MULTI_ENV)
    assert "1 passed" in tail


def test_cpp_adapter_cli_under_emulation(emu_lib):
    # the C++ host adapter end to end (rans::Rans::solve_airfoil -> multigrid<gpuSolver> FMG, explicit) linked against the emulation
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_cpp_host.py"], "explicit_mode")
    assert "1 passed" in tail


def test_early_halo_signal_under_emulation(emu_lib):
    # AFX_HALO_EARLY_SIGNAL=1 (opt-in): flags raised by the last send-layer CTA of the update kernel, 4 ranks as processes
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_multi.py"], "early_halo and 4", extra_env=MULTI_ENV)
    assert "1 passed" in tail



def test_in_process_group_under_emulation(emu_lib):
    """tests/test_gpu_group.py (N partitioned solvers of one process on one device: what the 1-GPU box runs on hardware):
    staged halo between host rendezvous, peer-memory push with plain pointers, the bounded halo wait giving up on a dead
    peer (AFX_ERR_COMM), cp of a partition's wall edges."""
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_group.py"], "2-p2p-strict or 4-staged-strict or dead_peer or wall_cp or 3-inviscid or (michalak and staged)",
                                         extra_env={"AFX_EMU_DEVICES": "1", "OMP_NUM_THREADS": "2", "OMP_WAIT_POLICY": "passive"})
    assert "6 passed" in tail


def test_pipelined_stage_kernel_under_emulation(emu_lib):
    """tests/test_gpu_pipe.py: the persistent chunk-sweeping stage kernel (work items claimed from a device counter, per-chunk
    completion counters, far lists) -- its CTAs run on several host threads here, so the claim / wait / publish protocol is
    exercised for real; bit-identical to the three-kernel stage and to the reference's golden history."""
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_pipe.py"], "(strict and (8-1-0 or 10-2-2)) or variants or laminar or reference_history or 2-p2p",
                                         extra_env={"AFX_EMU_DEVICES": "1", "OMP_NUM_THREADS": "4", "OMP_WAIT_POLICY": "passive"})
    assert " passed" in tail and "failed" not in tail


def test_device_fmg_under_emulation(emu_lib):
    """tests/test_gpu_fmg.py: device-resident prolongation (bit-identical to the reference's product) and afx_rans_sweep_fmg against
    the same loop written with primitive calls and a host prolongation."""
    tail = run_gpu_tests_under_emulation(emu_lib, ["test_gpu_fmg.py"], "not converges_to_the_reference and not polar_chains")  # 48 implicit FMG angles: GPU only
    assert "2 passed" in tail
