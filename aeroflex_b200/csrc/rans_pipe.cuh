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
// all writes of this CTA's item are done (the caller has passed a barrier): publish
__device__ __forceinline__ void pipe_publish(unsigned int* chunk_ctr, unsigned int* total_ctr)
{
    __threadfence();
    if (chunk_ctr) atomicAdd(chunk_ctr, 1u);
    atomicAdd(total_ctr, 1u);
}

template <int SECOND, int VISC, int LAST, int HAS_L>
__global__ void __launch_bounds__(AFX_PIPE_THREADS, AFX_PIPE_MINB) k_pipe(DevMesh m, PipeTab pt, const d4* qk_in, const d4* q0, d4* qk_out,
                                                                         const d4* gx, const d4* gy, d4* lim, d4* flux, const double* dt,
                                                                         d4* qW, double alpha, const double* __restrict__ prm, GasC g,
                                                                         double limiter_k, int walls, NormOut no, PushArgs push)
{
    constexpr uint32_t T = AFX_PIPE_THREADS;
    enum : uint32_t { PH_L = 0, PH_F = 1, PH_U = 2, PH_FARF = 3, PH_FARU = 4, PH_EXIT = 5 };
    __shared__ uint32_t s_it[3];  // phase, chunk, item index inside its chunk and phase (or inside the far list)
    __shared__ bool s_last_out;
    pdl_launch_dependents();
    pdl_wait();  // states, gradients, time steps and the zeroed work counters come from earlier kernels
    unsigned int* const Ldone = pt.ctr + 4;
    unsigned int* const Fdone = pt.ctr + 4 + pt.n_chunks;
    uint32_t cursor = 0;
    unsigned int nxt = 0;
    if (threadIdx.x == 0) nxt = atomicAdd(&pt.ctr[0], 1u);
    for (;;) {
        if (threadIdx.x == 0) {
            const uint32_t it = nxt;
            uint32_t ph, chunk = 0, idx = 0;
            if (it >= pt.n_items) ph = PH_EXIT;
            else if (it < pt.n_main) {
                while (cursor + 1 < pt.n_steps && it >= pt.steps[cursor + 1].x) ++cursor;
                const uint4 sp = pt.steps[cursor];
                const uint32_t r = it - sp.x;
                if (r < sp.y) { ph = PH_L; chunk = cursor; idx = r; }
                else if (r < sp.y + sp.z) { ph = PH_F; chunk = cursor - pt.lagF; idx = r - sp.y; }
                else { ph = PH_U; chunk = cursor - pt.lagF - pt.lagU; idx = r - sp.y - sp.z; }
            } else if (it < pt.n_main + pt.n_farF_items) { ph = PH_FARF; idx = it - pt.n_main; }
            else { ph = PH_FARU; idx = it - pt.n_main - pt.n_farF_items; }
            s_it[0] = ph; s_it[1] = chunk; s_it[2] = idx;
            if (ph != PH_EXIT) nxt = atomicAdd(&pt.ctr[0], 1u);  // the next claim travels while this item is worked on
        }
        __syncthreads();
        const uint32_t ph = s_it[0], chunk = s_it[1], idx = s_it[2];
        if (ph == PH_EXIT) break;
        if (ph == PH_L) {
            if (HAS_L) {
                const uint32_t i = (chunk << pt.shift) + idx * T + threadIdx.x;
                const uint32_t hi = umin32((chunk + 1u) << pt.shift, m.n_grad);
                if (i < hi) {
                    const LimCell c = limiter_load_static(m, i);
                    limiter_cell(m, i, c, qk_in, gx, gy, lim, limiter_k, walls);
                }
                __syncthreads();
                if (threadIdx.x == 0) pipe_publish(&Ldone[chunk], &pt.ctr[2]);
            }
        } else if (ph == PH_F || ph == PH_FARF) {
            uint32_t f = 0;
            bool active;
            if (ph == PH_F) {
                f = pt.face_start[chunk] + idx * T + threadIdx.x;
                active = f < pt.face_start[chunk + 1];
            } else {
                const uint32_t k = idx * T + threadIdx.x;
                active = k < pt.n_far_faces;
                if (active) f = pt.far_faces[k];
            }
            FaceRec rec;
            if (active) {
                rec = face_load_static(m, f);
                if (ph == PH_F && rec.fc.y < m.N) {  // a face between two chunks more than one apart waits for the far pass
                    const uint32_t ca = rec.fc.x >> pt.shift, cb = rec.fc.y >> pt.shift;
                    if (ca > cb + 1u || cb > ca + 1u) active = false;
                }
            }
            if (HAS_L) {  // the limiters of both cells: this chunk and its two neighbours (sweep), every chunk (far pass)
                if (threadIdx.x == 0) {
                    if (ph == PH_F) {
                        const uint32_t c0 = chunk > 0 ? chunk - 1u : 0u, c1 = umin32(chunk + 1u, pt.n_chunks - 1u);
                        for (uint32_t c = c0; c <= c1; ++c) {
                            const uint32_t cells = umin32((c + 1u) << pt.shift, m.n_grad) - (c << pt.shift);
                            pipe_wait(&Ldone[c], (cells + T - 1u) / T, pt.err);
                        }
                    } else pipe_wait(&pt.ctr[2], pt.nL_total, pt.err);
                    __threadfence();
                }
                __syncthreads();
            }
            if (active) flux_face<SECOND, VISC, 0, HAS_L>(m, f, rec, qk_in, q0, gx, gy, lim, flux, g, mk4(0, 0, 0, 0));
            __syncthreads();
            if (threadIdx.x == 0) pipe_publish(ph == PH_F ? &Fdone[chunk] : nullptr, &pt.ctr[3]);
        } else {  // PH_U, PH_FARU
            uint32_t i = 0;
            bool active;
            if (ph == PH_U) {
                i = (chunk << pt.shift) + idx * T + threadIdx.x;
                active = i < umin32((chunk + 1u) << pt.shift, m.n_upd);
                if (active && ((pt.far_mask[i >> 5] >> (i & 31u)) & 1u)) active = false;  // touches a far face: far pass
            } else {
                const uint32_t k = idx * T + threadIdx.x;
                active = k < pt.n_far_cells;
                if (active) i = pt.far_cells[k];
            }
            uint32_t bnd[4] = {CF_NONE, CF_NONE, CF_NONE, CF_NONE};
            if (active) {
#pragma unroll
                for (int s = 0; s < 4; ++s) bnd[s] = m.cf[(size_t)s * m.N + i];
            }
            if (threadIdx.x == 0) {  // the fluxes of the cell's faces: lower cell in this chunk or the one before (sweep)
                if (ph == PH_U) {
                    const uint32_t c0 = chunk > 0 ? chunk - 1u : 0u;
                    for (uint32_t c = c0; c <= chunk; ++c) {
                        const uint32_t nf = pt.face_start[c + 1] - pt.face_start[c];
                        pipe_wait(&Fdone[c], (nf + T - 1u) / T, pt.err);
                    }
                } else pipe_wait(&pt.ctr[3], pt.nF_total, pt.err);
                __threadfence();
            }
            __syncthreads();
            double nrm = 0;
            if (active) nrm = gather_cell<0, LAST, 1>(m, i, bnd, flux, q0, qk_in, qk_out, dt, qW, alpha, prm, walls, push);
            if (LAST) block_norm_accumulate(nrm, no, ph == PH_U ? pt.u_slot0[chunk] + idx : pt.nU_near_items + idx);
        }
        __syncthreads();  // s_it is rewritten next
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

static int pipe_threads() { return AFX_PIPE_THREADS; }
static int pipe_ctas_per_sm() { return AFX_PIPE_MINB; }

static void pipe(int second, int visc, int last, int has_l, const DevMesh& m, const PipeTab& pt, unsigned grid, const d4* qk_in, const d4* q0,
                 d4* qk_out, const d4* gx, const d4* gy, d4* lim, d4* flux, const double* dt, d4* qW, double alpha, const double* prm,
                 const GasC& g, double limiter_k, int walls, NormOut no, const PushArgs* push_in, cudaStream_t st)
{
    if (!grid || !pt.n_items) return;
    PushArgs push{};
    if (push_in) push = *push_in;
    if (no.blk_total == 0) { no.blk_off = 0; no.blk_total = pt.nU_near_items + pt.n_farU_items; }
#define AFX_PIPE(S, V, L, H) launch_pdl(k_pipe<S, V, L, H>, grid, AFX_PIPE_THREADS, st, m, pt, qk_in, q0, qk_out, gx, gy, lim, flux, dt, qW, alpha, prm, g, limiter_k, walls, no, push)
#define AFX_PIPE_LH(S, V) do { if (last) { if (has_l) AFX_PIPE(S, V, 1, 1); else AFX_PIPE(S, V, 1, 0); } else { if (has_l) AFX_PIPE(S, V, 0, 1); else AFX_PIPE(S, V, 0, 0); } } while (0)
    // VISC = 0 only: the laminar face gradient reads the iteration-start state of BOTH cells of a face, and the last stage
    // writes that state in place -- a later chunk's fluxes would see updated neighbours (the host keeps the three-kernel
    // stage for laminar runs)
    (void)visc;
    if (second) AFX_PIPE_LH(1, 0);
    else AFX_PIPE_LH(0, 0);
#undef AFX_PIPE_LH
#undef AFX_PIPE
}

}  // namespace launch
}  // namespace AFX_NS
}  // namespace afx
