// host emulation (tests only): NVTX ranges are no-ops
#pragma once
static inline int nvtxRangePushA(const char*) { return 0; }
static inline int nvtxRangePop() { return 0; }
