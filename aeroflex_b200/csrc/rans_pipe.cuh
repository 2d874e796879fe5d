// rans_pipe.cuh -- the pipelined stage kernel: one Runge-Kutta stage (limiter, face fluxes, gather + update) as ONE
// persistent kernel that sweeps the mesh chunk by chunk with the three phases a few chunks apart (PipeTab, rans_types.h).
//
// Why: the three-kernel stage is HBM-bound at 16M cells with every kernel at 85-97 % of the copy bandwidth on the bytes it
// really moves; only fewer bytes make it faster.  Between the kernels the limiters (32 B/cell) and the flux buffer
// (32 B/face) go out to HBM and come back, and the flux kernel re-reads the states and gradients (96 B/cell) the limiter
// kernel has just read.  Here those hand-overs happen a fraction of a millisecond apart on ~100k cells, i.e. inside the
// L2: per stage the kernel asks HBM for the static geometry, one pass over states / gradients and the new state.
// The arithmetic is the three kernels' own (limiter_cell / flux_face / gather_cell): same bits in strict mode.
//
// Scheduling: items (AFX_PIPE_THREADS consecutive cells or faces of one chunk and phase) are numbered in sweep order and
// claimed from a device counter.  An item reads only what items with SMALLER numbers wrote; those were claimed earlier, by
// CTAs that are running, so a waiting CTA always waits for running CTAs: no deadlock whatever the grid size.  Completion is
// published per chunk and phase (release: barrier, fence, atomic by one thread; acquire: one polling thread, fence, barrier);
// data produced inside the kernel is read with ld.global.cg (L2), never through a line the SM's L1 may have kept.
#pragma once
#include "rans_kernels.cuh"

#ifndef AFX_PIPE_THREADS
#define AFX_PIPE_THREADS 256
#endif
#ifndef AFX_PIPE_MINB
#define AFX_PIPE_MINB (1024 / AFX_PIPE_THREADS)
#endif
// elements (cells or faces) per thread and work item: the claim, the dependency wait and the two barriers of an item are
// paid once per AFX_PIPE_THREADS * AFX_PIPE_EPT elements
#ifndef AFX_PIPE_EPT
#define AFX_PIPE_EPT 2
#endif

namespace afx {
namespace AFX_NS {

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread waits until *c >= want.  The bound is a safety net against a scheduling bug (a hang would cost the GPU): about
// two seconds, then the error word is set and the item proceeds on whatever is there.
__device__ __forceinline__ void pipe_wait(const unsigned int* c, unsigned int want, int* err)
{
    unsigned int spins = 0;
    while (ld_acquire_u32(c) < want) {
        if (++spins < 16u) continue;
        __nanosleep(64);
        if (spins > (1u << 25)) { if (err) *reinterpret_cast<volatile int*>(err) = 1; break; }
    }
}

enum : uint32_t { PH_L = 0, PH_F = 1, PH_U = 2, PH_FARF = 3, PH_FARU = 4 };

template <int SECOND, int LAST, int HAS_L>
__global__ void __launch_bounds__(AFX_PIPE_THREADS, AFX_PIPE_MINB) k_pipe(DevMesh m, PipeTab pt, const d4* qk_in, const d4* q0, d4* qk_out,
                                                                         const d4* gx, const d4* gy, d4* lim, d4* flux, const double* dt,
                                                                         d4* qW, double alpha, const double* __restrict__ prm, GasC g,
                                                                         double limiter_k, int walls, NormOut no, PushArgs push)
{
    constexpr uint32_t T = AFX_PIPE_THREADS;
    __shared__ unsigned int s_next[2];  // the item this CTA works on next, double buffered (claimed one item ahead)
    __shared__ bool s_last_out;
    pdl_launch_dependents();
    pdl_wait();  // states, gradients, time steps and the zeroed work counters come from earlier kernels
    unsigned int* const Ldone = pt.ctr + 4;
    unsigned int* const Fdone = pt.ctr + 4 + pt.n_chunks;
    unsigned int nxt = 0;
    if (threadIdx.x == 0) s_next[0] = atomicAdd(&pt.ctr[0], 1u);
    __syncthreads();
    for (unsigned int n = 0;; ++n) {
        const unsigned int it = s_next[n & 1u];
        if (it >= pt.n_items) break;
        if (threadIdx.x == 0) nxt = atomicAdd(&pt.ctr[0], 1u);  // the next claim travels while this item is worked on
        const uint4 rec = pt.items[it];                          // same address for the whole CTA: one broadcast load
        const uint32_t ph = rec.x & 7u, chunk = rec.x >> 3, first = rec.y, cnt = rec.z;
        if (ph == PH_L) {
            if (HAS_L) {
#pragma unroll 1
                for (uint32_t k = threadIdx.x; k < cnt; k += T) {
                    const uint32_t i = first + k;
                    const LimCell c = limiter_load_static(m, i);
                    limiter_cell(m, i, c, qk_in, gx, gy, lim, limiter_k, walls);
                }
            }
        } else if (ph == PH_F || ph == PH_FARF) {
            if (HAS_L) {  // the limiters of both cells: this chunk and its two neighbours (sweep), every chunk (far pass)
                if (ph == PH_F) {
                    if (threadIdx.x < 3u) {
                        const uint32_t c = chunk + threadIdx.x;  // chunk - 1 + t
                        if (c >= 1u && c - 1u < pt.n_chunks) pipe_wait(&Ldone[c - 1u], pt.chunk_items[c - 1u].x, pt.err);
                        __threadfence();
                    }
                } else if (threadIdx.x == 0) { pipe_wait(&pt.ctr[2], pt.nL_total, pt.err); __threadfence(); }
                __syncthreads();
            }
#pragma unroll 1
            for (uint32_t k = threadIdx.x; k < cnt; k += T) {
                const uint32_t f = ph == PH_F ? first + k : pt.far_faces[first + k];
                const FaceRec fr = face_load_static(m, f);
                if (ph == PH_F && fr.fc.y < m.N) {  // a face between two chunks more than one apart waits for the far pass
                    const uint32_t ca = fr.fc.x >> pt.shift, cb = fr.fc.y >> pt.shift;
                    if (ca > cb + 1u || cb > ca + 1u) continue;
                }
                flux_face<SECOND, 0, 0, HAS_L>(m, f, fr, qk_in, q0, gx, gy, lim, flux, g, mk4(0, 0, 0, 0));
            }
        } else {  // PH_U, PH_FARU: the fluxes of the cell's faces -- lower cell in this chunk or the one before (sweep)
            if (ph == PH_U) {
                if (threadIdx.x < 2u) {
                    const uint32_t c = chunk + threadIdx.x;  // chunk - 1 + t
                    if (c >= 1u) pipe_wait(&Fdone[c - 1u], pt.chunk_items[c - 1u].y, pt.err);
                    __threadfence();
                }
            } else if (threadIdx.x == 0) { pipe_wait(&pt.ctr[3], pt.nF_total, pt.err); __threadfence(); }
            __syncthreads();
            double nrm = 0;
#pragma unroll 1
            for (uint32_t k = threadIdx.x; k < cnt; k += T) {
                uint32_t i;
                if (ph == PH_U) {
                    i = first + k;
                    if ((pt.far_mask[i >> 5] >> (i & 31u)) & 1u) continue;  // touches a far face: far pass
                } else i = pt.far_cells[first + k];
                uint32_t bnd[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) bnd[s] = m.cf[(size_t)s * m.N + i];
                nrm += gather_cell<0, LAST, 1>(m, i, bnd, flux, q0, qk_in, qk_out, dt, qW, alpha, prm, walls, push);
            }
            if (LAST) block_norm_partial(nrm, no, rec.w);
        }
        if (threadIdx.x == 0) s_next[(n + 1u) & 1u] = nxt;
        __syncthreads();  // every write of this item has been issued
        if (threadIdx.x == 0 && ph != PH_U && ph != PH_FARU && (HAS_L || ph != PH_L)) {
            // publish: the fence (cumulative over the writes the barrier has ordered before it) and the counters run while the
            // other threads are already on the next item
            __threadfence();
            if (ph == PH_L) { atomicAdd(&Ldone[chunk], 1u); atomicAdd(&pt.ctr[2], 1u); }
            else { if (ph == PH_F) atomicAdd(&Fdone[chunk], 1u); atomicAdd(&pt.ctr[3], 1u); }
        }
    }
    // the last CTA to leave zeroes the work counters for the next launch (which starts behind griddepcontrol.wait)
    if (threadIdx.x == 0) {
        __threadfence();
        s_last_out = (atomicAdd(&pt.ctr[1], 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (s_last_out) {
        for (uint32_t k = threadIdx.x; k < 4u + 2u * pt.n_chunks; k += T) pt.ctr[k] = 0u;
    }
}

namespace launch {

static int pipe_item_elems() { return AFX_PIPE_THREADS * AFX_PIPE_EPT; }
static int pipe_ctas_per_sm() { return AFX_PIPE_MINB; }

static void pipe(int second, int visc, int last, int has_l, const DevMesh& m, const PipeTab& pt, unsigned grid, const d4* qk_in, const d4* q0,
                 d4* qk_out, const d4* gx, const d4* gy, d4* lim, d4* flux, const double* dt, d4* qW, double alpha, const double* prm,
                 const GasC& g, double limiter_k, int walls, NormOut no, const PushArgs* push_in, cudaStream_t st)
{
    if (!grid || !pt.n_items) return;
    PushArgs push{};
    if (push_in) push = *push_in;
    if (no.blk_total == 0) { no.blk_off = 0; no.blk_total = pt.nU_items; }
#define AFX_PIPE(S, L, H) launch_pdl(k_pipe<S, L, H>, grid, AFX_PIPE_THREADS, st, m, pt, qk_in, q0, qk_out, gx, gy, lim, flux, dt, qW, alpha, prm, g, limiter_k, walls, no, push)
#define AFX_PIPE_LH(S) do { if (last) { if (has_l) AFX_PIPE(S, 1, 1); else AFX_PIPE(S, 1, 0); } else { if (has_l) AFX_PIPE(S, 0, 1); else AFX_PIPE(S, 0, 0); } } while (0)
    // inviscid flux form only: the laminar face gradient reads the iteration-start state of BOTH cells of a face, and the last
    // stage writes that state in place -- a later chunk's fluxes would see updated neighbours (the host keeps the three-kernel
    // stage for laminar runs)
    (void)visc;
    if (second) AFX_PIPE_LH(1);
    else AFX_PIPE_LH(0);
#undef AFX_PIPE_LH
#undef AFX_PIPE
}

}  // namespace launch
}  // namespace AFX_NS
}  // namespace afx
