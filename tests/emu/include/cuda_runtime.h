// cuda_runtime.h (HOST EMULATION SHIM) -- TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py compiles the product's own kernel and solver sources (aeroflex_b200/csrc/*.cu, *.cuh) with g++
// against this header instead of the CUDA toolkit's, so that the CPU test suite can execute the real kernel code --
// indexing, slot order, ghost handling, tile staging, graph capture, the C ABI above it -- on a machine without a GPU.
// Every CUDA thread is a fiber, __syncthreads() and the warp shuffles are real rendezvous points, asynchronous copies
// land as late as the kernel's own waits allow.  Nothing here is part of the product: the package never loads the
// emulation library, bench.py and smoke() never see it, and libaeroflex_rans_b200.so still fails loudly without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <cstddef>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define AFX_HOST_EMU 1

// ---- language extensions -------------------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local

struct uint2 { unsigned int x, y; };
struct uint3 { unsigned int x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned int x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned int x, unsigned int y, unsigned int z, unsigned int w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ---- the fiber scheduler (emu_runtime.cpp) -------------------------------------------------------------------------
namespace afx_emu {
extern thread_local unsigned char* g_dyn_smem;  // dynamic shared memory of the running block
void sync_block();                              // __syncthreads
unsigned long long warp_exchange(unsigned long long v, int src_lane);  // all live lanes rendezvous; returns lane src_lane's v
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
// asynchronous copies: queued at issue, landed at the wait that covers them
void mbar_init(uint32_t bar);
void mbar_queue(uint32_t bar, void* dst, const void* src, uint32_t bytes);
void mbar_expect(uint32_t bar, uint32_t bytes);
void mbar_wait(uint32_t bar, uint32_t parity);
void cpasync_queue(void* dst, const void* src, uint32_t bytes);
void cpasync_commit();
void cpasync_wait(int leave_pending);
void smem_access(uint32_t offset, uint32_t bytes, bool write);  // race check of a shared-memory access (emu_runtime.cpp)
double rcp_seed(double a);    // MUFU.RCP64H: the upper word of 1/a, lower word zero
double rsqrt_seed(double a);  // MUFU.RSQ64H
const char* self_path();      // file name of the emulation library (nccl_dl.h binds the emulated NCCL entry points in it)
}  // namespace afx_emu

// set by the scheduler whenever a fiber is resumed (one fiber runs at a time per host thread)
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

static inline void __syncthreads() { ::afx_emu::sync_block(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
#include <time.h>
#include <sched.h>
static inline void __nanosleep(unsigned) { sched_yield(); }
namespace afx_emu { static inline unsigned long long global_timer_ns() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec; } }
[[noreturn]] static inline void __trap() { abort(); }
static inline double __shfl_down_sync(unsigned, double v, int delta)
{
    unsigned long long b;
    memcpy(&b, &v, 8);
    const int lane = (int)(threadIdx.x & 31u);
    const int src = lane + delta < 32 ? lane + delta : lane;
    b = ::afx_emu::warp_exchange(b, src);
    memcpy(&v, &b, 8);
    return v;
}
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }

// ---- runtime API ---------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801, cudaErrorNoDevice = 100, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
struct afx_emu_stream;
struct afx_emu_graph;
typedef afx_emu_stream* cudaStream_t;
struct afx_emu_event { double t_ms; void* captured_in; };
typedef afx_emu_event* cudaEvent_t;
typedef afx_emu_graph* cudaGraph_t;
typedef afx_emu_graph* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal, cudaStreamCaptureModeThreadLocal };
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum cudaLimit { cudaLimitPersistingL2CacheSize = 6 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaAccessProperty { cudaAccessPropertyNormal, cudaAccessPropertyStreaming, cudaAccessPropertyPersisting };
enum cudaStreamAttrID { cudaStreamAttributeAccessPolicyWindow = 1 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 6 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    int persistingL2CacheMaxSize, accessPolicyMaxWindowSize;
};
struct cudaAccessPolicyWindow { void* base_ptr; size_t num_bytes; float hitRatio; cudaAccessProperty hitProp, missProp; };
union cudaStreamAttrValue { cudaAccessPolicyWindow accessPolicyWindow; };
struct cudaFuncAttributes { size_t sharedSizeBytes; int numRegs; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int d);
cudaError_t cudaDeviceSetLimit(cudaLimit l, size_t v);
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi);
cudaError_t afx_emu_malloc(void** p, size_t bytes);
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return afx_emu_malloc(reinterpret_cast<void**>(p), bytes); }
cudaError_t afx_emu_malloc_host(void** p, size_t bytes);
template <class T>
static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return afx_emu_malloc_host(reinterpret_cast<void**>(p), bytes); }
cudaError_t cudaFree(void* p);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemcpyPeerAsync(void* dst, int dst_dev, const void* src, int src_dev, size_t n, cudaStream_t st);
cudaError_t cudaStreamIsCapturing(cudaStream_t st, cudaStreamCaptureStatus* s);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaMemset(void* dst, int v, size_t n);
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* st, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags);
cudaError_t cudaStreamSetAttribute(cudaStream_t st, cudaStreamAttrID id, const cudaStreamAttrValue* v);
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode m);
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* g);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t g);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e);
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t st);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
template <class F>
static inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { a->sharedSizeBytes = 512; a->numRegs = 0; return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* nb, F, int, size_t) { *nb = 2; return cudaSuccess; }

// a launch = a closure run by the fiber scheduler now, or recorded while its stream is being captured
namespace afx_emu {
void submit(cudaStream_t st, std::function<void()> op);
template <class... KArgs, class... Args>
static inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    std::tuple<std::decay_t<KArgs>...> held(std::forward<Args>(args)...);  // by value, like kernel parameters
    submit(st, [kernel, grid, block, smem, held]() {
        run_grid(grid, block, smem, [&]() { std::apply(kernel, held); });
    });
}
}  // namespace afx_emu
template <class... KArgs, class... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args)
{
    ::afx_emu::launch(kernel, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cfg->stream, std::forward<Args>(args)...);
    return cudaSuccess;
}
#define AFX_EMU_LAUNCH(kernel, grid, block, smem, st, ...) ::afx_emu::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
