"""aeroflex_b200 -- B200 (sm_100a) drop-in for AeroFLEX's src/rans hot path.

The product is the C-ABI shared library ``lib/libaeroflex_rans_b200.so``
(hand-written CUDA kernels, ``include/afx_rans.h``) plus the C++ adapter headers
under ``host/`` that mirror ``rans::solver`` / ``rans::multigrid`` / ``rans::Rans``.
This Python package is a thin ctypes view of that same C ABI for the test and
benchmark harness; it adds no compute of its own and has NO CPU fallback: every
solver call goes to the CUDA library and raises if it is missing or no B200 is
visible.
"""
from .capi import (AfxError, Mesh, GpuSolver, build_library, library_path, load_library,  # noqa: F401
                   BC_KINDS, VISCOSITY, GRADIENT, device_count, is_emulation, EXPORTED_SYMBOLS, pinned_array, Partition, nccl_unique_id, tiling_plan, Group, run_ranks, Prolongation, sweep_fmg)

__all__ = ["AfxError", "Mesh", "GpuSolver", "build_library", "library_path", "load_library", "device_count"]
