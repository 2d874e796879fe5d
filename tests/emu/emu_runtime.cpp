// emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY: the host-emulation runtime behind tests/emu/include/cuda_runtime.h.
//
// * run_grid(): every CUDA thread of a block is a fiber with its own stack (a 7-instruction x86-64 context switch);
//   the scheduler runs the fibers of one block round-robin, each until it finishes or reaches a rendezvous:
//   __syncthreads() releases when every live thread of the block waits, a warp shuffle when every live lane of the warp
//   waits.  A pass without progress is a deadlock (divergent barrier) and aborts with a message.  Blocks of a grid are
//   spread over the host's cores; atomics are real atomics.
// * Asynchronous copies are DEFERRED: cp.async.bulk lands when the first thread waits on the phase of its mbarrier,
//   cp.async when the issuing thread's wait_group covers it -- a read before the wait sees the old bytes.
// * Streams execute in submission order; a capturing stream records closures and a graph launch replays them.
#include <cuda_runtime.h>
#include <omp.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <map>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

extern "C" void afx_emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl afx_emu_switch
.type afx_emu_switch,@function
afx_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size afx_emu_switch, .-afx_emu_switch
)");

namespace afx_emu {

thread_local unsigned char* g_dyn_smem = nullptr;

namespace {

constexpr size_t STACK_BYTES = 128 * 1024;
constexpr unsigned MAX_THREADS = 1024;
constexpr size_t DYN_SMEM_BYTES = 232448;
enum State : int { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };

struct Copy { void* dst; const void* src; uint32_t bytes; };
struct Fiber {
    void* sp = nullptr;
    State state = DONE;
    std::vector<std::vector<Copy>> groups;  // committed cp.async groups, oldest first
    std::vector<Copy> open;                 // copies of the group being built
};
struct MBar { uint32_t phase = 0; std::vector<Copy> pending; };

struct Worker {  // one per host thread
    char* stacks = nullptr;
    unsigned char* smem = nullptr;
    Fiber fib[MAX_THREADS];
    void* sched_sp = nullptr;
    unsigned cur = 0, nthreads = 0;
    const std::function<void()>* body = nullptr;
    unsigned long long slot[MAX_THREADS];
    std::map<uint32_t, MBar> bars;
    void ensure()
    {
        if (stacks) return;
        stacks = static_cast<char*>(mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (stacks == MAP_FAILED) { perror("afx_emu: mmap"); abort(); }
        smem = static_cast<unsigned char*>(aligned_alloc(1024, DYN_SMEM_BYTES));
    }
};
thread_local Worker* g_w = nullptr;
Worker& worker()
{
    if (!g_w) g_w = new Worker;
    return *g_w;
}

void fiber_main()
{
    Worker& w = *g_w;
    (*w.body)();
    Fiber& f = w.fib[w.cur];
    if (!f.open.empty() || !f.groups.empty()) {  // copies never waited for are lost with the thread: land them (the data is simply unused)
        f.open.clear(); f.groups.clear();
    }
    f.state = DONE;
    void* dummy;
    afx_emu_switch(&dummy, w.sched_sp);
    abort();  // a finished fiber is never resumed
}

void yield(State s)
{
    Worker& w = *g_w;
    Fiber& f = w.fib[w.cur];
    f.state = s;
    afx_emu_switch(&f.sp, w.sched_sp);
}

void run_block(Worker& w, unsigned n)
{
    w.nthreads = n;
    w.bars.clear();
    for (unsigned t = 0; t < n; ++t) {
        Fiber& f = w.fib[t];
        char* top = w.stacks + STACK_BYTES * (size_t)(t + 1);
        void** sp = reinterpret_cast<void**>(top);
        sp[-1] = nullptr;                                  // fake return address of fiber_main
        sp[-2] = reinterpret_cast<void*>(&fiber_main);     // popped by `ret`
        for (int k = 3; k <= 8; ++k) sp[-k] = nullptr;     // rbp rbx r12 r13 r14 r15
        f.sp = sp - 8;
        f.state = RUNNABLE;
        f.groups.clear(); f.open.clear();
    }
    unsigned alive = n;
    while (alive) {
        bool progressed = false;
        for (unsigned t = 0; t < n; ++t) {
            Fiber& f = w.fib[t];
            if (f.state != RUNNABLE) continue;
            threadIdx = uint3{t % blockDim.x, (t / blockDim.x) % blockDim.y, t / (blockDim.x * blockDim.y)};
            w.cur = t;
            afx_emu_switch(&w.sched_sp, f.sp);
            progressed = true;
            if (f.state == DONE) --alive;
        }
        // release the rendezvous points that are complete
        unsigned at_block = 0;
        for (unsigned t = 0; t < n; ++t) at_block += (w.fib[t].state == WAIT_BLOCK);
        if (alive && at_block == alive) {
            for (unsigned t = 0; t < n; ++t) if (w.fib[t].state == WAIT_BLOCK) w.fib[t].state = RUNNABLE;
            progressed = true;
        }
        for (unsigned w0 = 0; w0 < n; w0 += 32) {
            unsigned live = 0, waiting = 0;
            for (unsigned t = w0; t < w0 + 32 && t < n; ++t) { live += (w.fib[t].state != DONE); waiting += (w.fib[t].state == WAIT_WARP); }
            if (live && waiting == live) {
                for (unsigned t = w0; t < w0 + 32 && t < n; ++t) if (w.fib[t].state == WAIT_WARP) w.fib[t].state = RUNNABLE;
                progressed = true;
            }
        }
        if (!progressed) {
            fprintf(stderr, "afx_emu: deadlock in block %u: %u live threads, %u at __syncthreads (divergent barrier or shuffle)\n", blockIdx.x, alive, at_block);
            abort();
        }
    }
}

}  // namespace

void sync_block() { yield(WAIT_BLOCK); }

unsigned long long warp_exchange(unsigned long long v, int src_lane)
{
    Worker& w = *g_w;
    const unsigned me = w.cur, base = me & ~31u;
    w.slot[me] = v;
    yield(WAIT_WARP);
    const unsigned long long r = w.slot[base + (unsigned)src_lane];
    yield(WAIT_WARP);
    return r;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    const unsigned n = block.x * block.y * block.z;
    if (n == 0 || n > MAX_THREADS || smem > DYN_SMEM_BYTES) { fprintf(stderr, "afx_emu: bad launch configuration\n"); abort(); }
    const long nblocks = (long)grid.x * grid.y * grid.z;
    static const bool serial = [] { const char* e = getenv("AFX_EMU_SERIAL"); return e && e[0] == '1'; }();
#pragma omp parallel if (!serial && nblocks > 1)
    {
        Worker& w = worker();
        w.ensure();
        w.body = &body;
        g_dyn_smem = w.smem;
        blockDim = block; gridDim = grid;
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < nblocks; ++b) {
            blockIdx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long)grid.x * grid.y))};
            run_block(w, n);
        }
    }
}

// ---- deferred asynchronous copies ----------------------------------------------------------------------------------
static void land(const std::vector<Copy>& cs) { for (const Copy& c : cs) memcpy(c.dst, c.src, c.bytes); }
void mbar_init(uint32_t bar) { g_w->bars[bar] = MBar{}; }
void mbar_expect(uint32_t, uint32_t) {}
void mbar_queue(uint32_t bar, void* dst, const void* src, uint32_t bytes) { g_w->bars[bar].pending.push_back(Copy{dst, src, bytes}); }
void mbar_wait(uint32_t bar, uint32_t parity)
{
    MBar& b = g_w->bars[bar];
    if (b.phase != (parity & 1u)) return;  // that phase completed earlier
    land(b.pending);
    b.pending.clear();
    b.phase ^= 1u;
}
void cpasync_queue(void* dst, const void* src, uint32_t bytes) { g_w->fib[g_w->cur].open.push_back(Copy{dst, src, bytes}); }
void cpasync_commit()
{
    Fiber& f = g_w->fib[g_w->cur];
    f.groups.push_back(std::move(f.open));
    f.open.clear();
}
void cpasync_wait(int leave_pending)
{
    Fiber& f = g_w->fib[g_w->cur];
    while ((int)f.groups.size() > leave_pending) { land(f.groups.front()); f.groups.erase(f.groups.begin()); }
}

// the hardware seeds: 1/a and 1/sqrt(a) with the lower 32 bits of the double cleared (about 20 bits)
static double upper_word(double x)
{
    unsigned long long b;
    memcpy(&b, &x, 8);
    b &= 0xFFFFFFFF00000000ull;
    memcpy(&x, &b, 8);
    return x;
}
double rcp_seed(double a) { return upper_word(1.0 / a); }
double rsqrt_seed(double a) { return upper_word(1.0 / std::sqrt(a)); }

void submit(cudaStream_t st, std::function<void()> op);

}  // namespace afx_emu

// ---- runtime API ---------------------------------------------------------------------------------------------------
struct afx_emu_graph { std::vector<std::function<void()>> ops; };
struct afx_emu_stream { afx_emu_graph* capture = nullptr; };

void afx_emu::submit(cudaStream_t st, std::function<void()> op)
{
    if (st && st->capture) st->capture->ops.push_back(std::move(op));
    else op();
}

const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "not supported by the host emulation" : "emulated CUDA error"); }
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "host emulation of sm_100a (tests only, not a GPU)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 4;  // few "SMs": the persistent kernels walk several tiles per CTA
    if (const char* e = getenv("AFX_EMU_SMS")) p->multiProcessorCount = atoi(e) > 0 ? atoi(e) : 4;
    p->persistingL2CacheMaxSize = 0; p->accessPolicyMaxWindowSize = 0;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int)
{
    if (a == cudaDevAttrMaxSharedMemoryPerBlockOptin) { *v = 232448; return cudaSuccess; }
    return cudaErrorInvalidValue;
}
cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t afx_emu_malloc(void** p, size_t bytes)
{
    *p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
    if (!*p) return cudaErrorInvalidValue;
    memset(*p, 0xCD, bytes);  // fresh device memory is not zero: poison it
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t st)
{
    afx_emu::submit(st, [=] { memmove(dst, src, n); });
    return cudaSuccess;
}
cudaError_t cudaMemset(void* dst, int v, size_t n) { memset(dst, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t st)
{
    afx_emu::submit(st, [=] { memset(dst, v, n); });
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) { *st = new afx_emu_stream; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* st, unsigned, int) { *st = new afx_emu_stream; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t st) { delete st; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t st) { return (st && st->capture) ? cudaErrorInvalidValue : cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) { return cudaSuccess; }
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode)
{
    if (!st || st->capture) return cudaErrorInvalidValue;
    st->capture = new afx_emu_graph;
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t* g)
{
    if (!st || !st->capture) return cudaErrorInvalidValue;
    *g = st->capture;
    st->capture = nullptr;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new afx_emu_graph(*g); return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
    for (auto& op : e->ops) op();
    return cudaSuccess;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new afx_emu_event{0.}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new afx_emu_event{0.}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st)
{
    afx_emu::submit(st, [=] { e->t_ms = now_ms(); });
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
