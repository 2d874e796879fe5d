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
#include <nccl.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <omp.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>

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

// Shadow of the dynamic shared memory, one entry per 4 bytes (the race check of emu_stage_ptx.h accesses):
//   * between two block barriers ("epoch") a word written by one thread may not be read or written by another;
//   * a word that an asynchronous copy (cp.async / cp.async.bulk) is still in flight to may not be touched at all.
// Words filled by a landed copy belong to nobody: whoever waited for the copy may read them.
struct Shadow { uint32_t epoch; int16_t writer, reader; uint16_t inflight; };
constexpr int16_t NOBODY = -1, MANY = -2;

struct Worker {  // one per host thread
    std::vector<Shadow> shadow;
    uint32_t epoch = 1;
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
        shadow.assign(DYN_SMEM_BYTES / 4, Shadow{0, NOBODY, NOBODY, 0});
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
    if (!f.open.empty() || !f.groups.empty()) {  // copies never waited for: a kernel must drain its cp.async groups before it exits
        fprintf(stderr, "afx_emu: thread %u of block %u exits with cp.async copies it never waited for\n", threadIdx.x, blockIdx.x);
        abort();
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
    ++w.epoch;
    for (auto& e : w.shadow) e.inflight = 0;
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
            ++w.epoch;  // accesses before and after a __syncthreads() are ordered
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

// ---- shared-memory race check ---------------------------------------------------------------------------------------
static const bool g_racecheck = [] { const char* e = getenv("AFX_EMU_RACECHECK"); return !(e && e[0] == '0'); }();
[[noreturn]] static void race(const char* what, uint32_t off, int other)
{
    fprintf(stderr, "afx_emu: shared-memory hazard: %s at byte offset %u of the dynamic shared memory (block %u, thread %u, other thread %d)\n",
            what, off, blockIdx.x, threadIdx.x, other);
    abort();
}
void smem_access(uint32_t off, uint32_t bytes, bool write)
{
    if (!g_racecheck) return;
    Worker& w = *g_w;
    const int16_t me = (int16_t)w.cur;
    for (uint32_t k = off / 4; k <= (off + bytes - 1) / 4; ++k) {
        Shadow& e = w.shadow[k];
        if (e.inflight) race(write ? "write to a word an asynchronous copy is in flight to" : "read of a word an asynchronous copy is in flight to", k * 4, -1);
        if (e.epoch != w.epoch) { e.epoch = w.epoch; e.writer = NOBODY; e.reader = NOBODY; }
        if (write) {
            if (e.writer != NOBODY && e.writer != me) race("two threads write the same word without a barrier between them", k * 4, e.writer);
            if (e.reader != NOBODY && e.reader != me) race("a thread writes a word another thread has read since the last barrier", k * 4, e.reader);
            e.writer = me;
        } else {
            if (e.writer != NOBODY && e.writer != me) race("a thread reads a word another thread has written since the last barrier", k * 4, e.writer);
            e.reader = (e.reader == NOBODY || e.reader == me) ? me : MANY;
        }
    }
}
static void smem_copy_mark(void* dst, uint32_t bytes, int delta)
{
    if (!g_racecheck || !g_w || !g_w->smem) return;
    Worker& w = *g_w;
    const ptrdiff_t off = static_cast<unsigned char*>(dst) - w.smem;
    if (off < 0 || (size_t)off + bytes > DYN_SMEM_BYTES) return;
    for (size_t k = (size_t)off / 4; k <= ((size_t)off + bytes - 1) / 4; ++k) {
        Shadow& e = w.shadow[k];
        if (delta > 0) {
            // the copy may start at once: nobody may have touched the destination since the last barrier ... unless it is the
            // issuing thread itself (program order) -- a read by ANOTHER thread in this epoch is the classic buffer-reuse bug
            if (e.epoch == w.epoch && ((e.reader != NOBODY && e.reader != (int16_t)w.cur) || (e.writer != NOBODY && e.writer != (int16_t)w.cur)))
                race("an asynchronous copy is issued into a word another thread has used since the last barrier", (uint32_t)k * 4,
                     e.reader != NOBODY ? e.reader : e.writer);
            ++e.inflight;
        } else {
            if (e.inflight) --e.inflight;
            e.epoch = 0;  // landed data belongs to nobody
        }
    }
}

// ---- deferred asynchronous copies ----------------------------------------------------------------------------------
static void land(const std::vector<Copy>& cs)
{
    for (const Copy& c : cs) { memcpy(c.dst, c.src, c.bytes); smem_copy_mark(c.dst, c.bytes, -1); }
}
void mbar_init(uint32_t bar) { g_w->bars[bar] = MBar{}; }
void mbar_expect(uint32_t, uint32_t) {}
void mbar_queue(uint32_t bar, void* dst, const void* src, uint32_t bytes)
{
    smem_copy_mark(dst, bytes, +1);
    g_w->bars[bar].pending.push_back(Copy{dst, src, bytes});
}
void mbar_wait(uint32_t bar, uint32_t parity)
{
    MBar& b = g_w->bars[bar];
    if (b.phase != (parity & 1u)) return;  // that phase completed earlier
    land(b.pending);
    b.pending.clear();
    b.phase ^= 1u;
}
void cpasync_queue(void* dst, const void* src, uint32_t bytes)
{
    smem_copy_mark(dst, bytes, +1);
    g_w->fib[g_w->cur].open.push_back(Copy{dst, src, bytes});
}
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

const char* self_path()
{
    static std::string path = [] {
        Dl_info info{};
        if (!dladdr(reinterpret_cast<void*>(&rcp_seed), &info) || !info.dli_fname) { fprintf(stderr, "afx_emu: dladdr failed\n"); abort(); }
        return std::string(info.dli_fname);
    }();
    return path.c_str();
}

}  // namespace afx_emu

// ---- runtime API ---------------------------------------------------------------------------------------------------
struct afx_emu_graph {
    std::vector<std::function<void()>> ops;
    std::vector<afx_emu_stream*> forked;  // streams that joined the capture through an event (fork / join pattern)
};
struct afx_emu_stream { afx_emu_graph* capture = nullptr; };

void afx_emu::submit(cudaStream_t st, std::function<void()> op)
{
    if (st && st->capture) st->capture->ops.push_back(std::move(op));
    else op();
}

const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "not supported by the host emulation" : "emulated CUDA error"); }
cudaError_t cudaGetLastError() { return cudaSuccess; }
// AFX_EMU_DEVICES "devices" (default 1): the multi-rank tests give every rank (= process) its own ordinal
static int emu_devices() { const char* e = getenv("AFX_EMU_DEVICES"); const int n = e ? atoi(e) : 1; return n > 0 ? n : 1; }
static thread_local int g_device = 0;
cudaError_t cudaGetDeviceCount(int* n) { *n = emu_devices(); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= emu_devices()) return cudaErrorInvalidValue; g_device = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = g_device; return cudaSuccess; }
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
// "Device" memory is memfd-backed shared mappings, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map an allocation
// of another rank's process (through /proc/<pid>/fd/<fd>) exactly like CUDA IPC maps a peer's buffer.
namespace {
struct Alloc { size_t bytes; int fd; char* map; };  // bytes = length of the whole mapping, guard page included
std::mutex g_alloc_mu;
std::map<char*, Alloc> g_allocs;   // cudaMalloc'ed blocks of this process
std::map<void*, size_t> g_opened;  // peer blocks mapped here (base -> bytes)
struct IpcHandle { int pid, fd; unsigned long long bytes, offset; };
static_assert(sizeof(IpcHandle) <= sizeof(cudaIpcMemHandle_t), "IPC handle fits");
}  // namespace
cudaError_t afx_emu_malloc(void** p, size_t bytes)
{
    // The block ENDS (up to the 256-byte allocation granularity) at a page boundary followed by an inaccessible guard page:
    // a kernel that reads or writes past the end of an array faults at once instead of touching a neighbour's bytes.
    const size_t used = std::max<size_t>((bytes + 255) & ~(size_t)255, 256);
    const size_t data = (used + 4095) & ~(size_t)4095, total = data + 4096;
    const int fd = memfd_create("afx_emu_dev", 0);
    if (fd < 0 || ftruncate(fd, (off_t)total) != 0) { if (fd >= 0) close(fd); return cudaErrorInvalidValue; }
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { close(fd); return cudaErrorInvalidValue; }
    char* start = static_cast<char*>(m) + (data - used);
    memset(m, 0xCD, data);  // fresh device memory is not zero: poison it
    mprotect(static_cast<char*>(m) + data, 4096, PROT_NONE);
    std::lock_guard<std::mutex> g(g_alloc_mu);
    g_allocs[start] = Alloc{total, fd, static_cast<char*>(m)};
    *p = start;
    return cudaSuccess;
}
cudaError_t afx_emu_malloc_host(void** p, size_t bytes)
{
    *p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
    return *p ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> g(g_alloc_mu);
    auto it = g_allocs.find(static_cast<char*>(p));
    if (it == g_allocs.end()) return cudaErrorInvalidValue;
    munmap(it->second.map, it->second.bytes);
    close(it->second.fd);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t st)
{
    afx_emu::submit(st, [=] { memmove(dst, src, n); });
    return cudaSuccess;
}
cudaError_t cudaMemcpyPeerAsync(void* dst, int, const void* src, int, size_t n, cudaStream_t st) { return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st); }
cudaError_t cudaStreamIsCapturing(cudaStream_t st, cudaStreamCaptureStatus* s) { *s = (st && st->capture) ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
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
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned)
{
    // an event recorded in a capturing stream pulls the waiting stream into the same capture: what that stream is given
    // afterwards is recorded in submission order (a valid topological order of the graph) instead of running now
    if (st && ev && ev->captured_in && !st->capture) {
        st->capture = static_cast<afx_emu_graph*>(ev->captured_in);
        st->capture->forked.push_back(st);
    }
    return cudaSuccess;
}
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
    for (afx_emu_stream* f : st->capture->forked) f->capture = nullptr;
    st->capture->forked.clear();
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
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new afx_emu_event{0., nullptr}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new afx_emu_event{0., nullptr}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st)
{
    e->captured_in = (st && st->capture) ? st->capture : nullptr;
    afx_emu::submit(st, [=] { e->t_ms = now_ms(); });
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
    std::lock_guard<std::mutex> g(g_alloc_mu);
    auto it = g_allocs.upper_bound(static_cast<char*>(p));
    if (it == g_allocs.begin()) return cudaErrorInvalidValue;
    --it;
    if (static_cast<char*>(p) >= it->second.map + it->second.bytes - 4096) return cudaErrorInvalidValue;
    IpcHandle ih{(int)getpid(), it->second.fd, it->second.bytes - 4096, (unsigned long long)(static_cast<char*>(p) - it->second.map)};
    memset(h, 0, sizeof *h);
    memcpy(h, &ih, sizeof ih);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
    IpcHandle ih;
    memcpy(&ih, &h, sizeof ih);
    char path[64];
    snprintf(path, sizeof path, "/proc/%d/fd/%d", ih.pid, ih.fd);
    const int fd = open(path, O_RDWR);
    if (fd < 0) return cudaErrorNotSupported;
    void* m = mmap(nullptr, ih.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> g(g_alloc_mu);
    g_opened[m] = ih.bytes;
    *p = static_cast<char*>(m) + ih.offset;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p)
{
    std::lock_guard<std::mutex> g(g_alloc_mu);
    for (auto it = g_opened.begin(); it != g_opened.end(); ++it)
        if (static_cast<char*>(p) >= static_cast<char*>(it->first) && static_cast<char*>(p) < static_cast<char*>(it->first) + it->second) {
            munmap(it->first, it->second);
            g_opened.erase(it);
            return cudaSuccess;
        }
    return cudaErrorInvalidValue;
}

// ---- the NCCL entry points nccl_dl.h binds (build_emu.py points it at these): ranks are processes, the "fabric" is one
// POSIX shared-memory segment per communicator (named by the unique id) holding a barrier, an all-reduce slot per rank and
// a mailbox per ordered pair of ranks.  Operations are submitted to their stream like kernels, so they are captured into
// graphs and replayed with them.
namespace {
constexpr size_t NCCL_SLOT_DOUBLES = 1 << 16;       // all-reduce chunk
constexpr size_t NCCL_MAILBOX_BYTES = 4u << 20;     // one halo message
struct NcclHeader { std::atomic<int> arrived; std::atomic<int> generation; std::atomic<int> attached; };
struct EmuComm {
    char* base; size_t bytes; int nranks, rank; std::string name;
    struct Op { bool send; void* buf; size_t bytes; int peer; };
    std::vector<Op> group;
    NcclHeader* hdr() { return reinterpret_cast<NcclHeader*>(base); }
    double* slot(int r) { return reinterpret_cast<double*>(base + 4096) + (size_t)r * NCCL_SLOT_DOUBLES; }
    char* mailbox(int src, int dst) { return base + 4096 + (size_t)nranks * NCCL_SLOT_DOUBLES * 8 + ((size_t)src * nranks + dst) * NCCL_MAILBOX_BYTES; }
    void barrier()
    {
        NcclHeader* h = hdr();
        const int gen = h->generation.load();
        if (h->arrived.fetch_add(1) + 1 == nranks) { h->arrived.store(0); h->generation.fetch_add(1); }
        else while (h->generation.load() == gen) sched_yield();
    }
};
}  // namespace
extern "C" {
ncclResult_t afx_emu_ncclGetUniqueId(ncclUniqueId* id)
{
    static std::atomic<int> counter{0};
    memset(id, 0, sizeof *id);
    snprintf(id->internal, sizeof id->internal, "/afx_emu_nccl_%d_%d_%lld", (int)getpid(), counter.fetch_add(1),
             (long long)std::chrono::steady_clock::now().time_since_epoch().count());
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank)
{
    EmuComm* c = new EmuComm;
    c->nranks = nranks; c->rank = rank; c->name = id.internal;
    c->bytes = 4096 + (size_t)nranks * NCCL_SLOT_DOUBLES * 8 + (size_t)nranks * nranks * NCCL_MAILBOX_BYTES;
    const int fd = shm_open(c->name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)c->bytes) != 0) { if (fd >= 0) close(fd); delete c; return ncclSystemError; }
    void* m = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { delete c; return ncclSystemError; }
    c->base = static_cast<char*>(m);
    c->hdr()->attached.fetch_add(1);
    c->barrier();                                   // everybody has the segment: the name can go
    if (rank == 0) shm_unlink(c->name.c_str());
    *comm = reinterpret_cast<ncclComm_t>(c);
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclCommDestroy(ncclComm_t comm)
{
    EmuComm* c = reinterpret_cast<EmuComm*>(comm);
    munmap(c->base, c->bytes);
    delete c;
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm, cudaStream_t st)
{
    if (dt != ncclDouble || op != ncclSum) return ncclInternalError;
    EmuComm* c = reinterpret_cast<EmuComm*>(comm);
    afx_emu::submit(st, [=] {
        const double* s = static_cast<const double*>(send);
        double* r = static_cast<double*>(recv);
        for (size_t o = 0; o < count; o += NCCL_SLOT_DOUBLES) {
            const size_t n = std::min(NCCL_SLOT_DOUBLES, count - o);
            memcpy(c->slot(c->rank), s + o, n * 8);
            c->barrier();
            for (size_t k = 0; k < n; ++k) {
                double acc = 0;
                for (int q = 0; q < c->nranks; ++q) acc += c->slot(q)[k];  // rank order on every rank: identical sums
                r[o + k] = acc;
            }
            c->barrier();
        }
    });
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclGroupStart() { return ncclSuccess; }
static thread_local EmuComm* g_group_comm = nullptr;
static thread_local cudaStream_t g_group_stream = nullptr;
ncclResult_t afx_emu_ncclSend(const void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t st)
{
    if (dt != ncclDouble || count * 8 > NCCL_MAILBOX_BYTES) return ncclInternalError;
    EmuComm* c = reinterpret_cast<EmuComm*>(comm);
    c->group.push_back(EmuComm::Op{true, const_cast<void*>(buf), count * 8, peer});
    g_group_comm = c; g_group_stream = st;
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclRecv(void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t st)
{
    if (dt != ncclDouble || count * 8 > NCCL_MAILBOX_BYTES) return ncclInternalError;
    EmuComm* c = reinterpret_cast<EmuComm*>(comm);
    c->group.push_back(EmuComm::Op{false, buf, count * 8, peer});
    g_group_comm = c; g_group_stream = st;
    return ncclSuccess;
}
ncclResult_t afx_emu_ncclGroupEnd()
{
    EmuComm* c = g_group_comm;
    if (!c) return ncclSuccess;  // an empty group (a rank without peers): nothing to wait for -- pairs synchronise below
    std::vector<EmuComm::Op> ops;
    ops.swap(c->group);
    cudaStream_t st = g_group_stream;
    g_group_comm = nullptr; g_group_stream = nullptr;
    // point-to-point hand-off per ordered pair: the sender fills the pair's mailbox and raises its sequence number, the
    // receiver waits for it, copies, and acknowledges; no collective barrier (ranks have different peer sets)
    afx_emu::submit(st, [c, ops] {
        struct Box { std::atomic<unsigned long long> filled, drained; };
        auto box = [&](int src, int dst) { return reinterpret_cast<Box*>(c->mailbox(src, dst)); };
        for (const auto& o : ops)
            if (o.send) {
                Box* b = box(c->rank, o.peer);
                while (b->filled.load() != b->drained.load()) sched_yield();  // previous message still unread
                memcpy(reinterpret_cast<char*>(b) + 64, o.buf, o.bytes);
                b->filled.fetch_add(1);
            }
        for (const auto& o : ops)
            if (!o.send) {
                Box* b = box(o.peer, c->rank);
                while (b->filled.load() == b->drained.load()) sched_yield();
                memcpy(o.buf, reinterpret_cast<char*>(b) + 64, o.bytes);
                b->drained.fetch_add(1);
            }
    });
    return ncclSuccess;
}
const char* afx_emu_ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }
ncclResult_t afx_emu_ncclGetVersion(int* v) { *v = 22809; return ncclSuccess; }
}  // extern "C"
