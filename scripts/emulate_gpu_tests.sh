#!/bin/bash
# Run `-m gpu` tests against the kernel sources under host emulation (tests/emu) on a machine without a GPU.
#   scripts/emulate_gpu_tests.sh                                  # the whole list except the 120 000-iteration ring test (~15 min on 8 cores)
#   scripts/emulate_gpu_tests.sh tests/test_gpu_fused.py -k 192    # any pytest selection
# Test infrastructure only: the package refuses this library without AFX_ALLOW_EMULATION=tests; bench.py and smoke() refuse it always.
cd "$(dirname "$0")/.." || exit 1
lib=$(python tests/emu/build_emu.py) || exit 1
export AFX_LIB=$lib AFX_ALLOW_EMULATION=tests AFX_EMU_DEVICES=${AFX_EMU_DEVICES:-8} OMP_WAIT_POLICY=passive
if [ $# -eq 0 ]; then set -- tests -k "not ring_wraps"; fi
exec python -m pytest -m gpu -q "$@"
