// rans_stage.cuh -- one Runge-Kutta stage of explicitSolver::solve (solver.h:808-822) as ONE persistent kernel on
// shared-memory tiles, fed by the copy engine.
//
// A tile is a run of consecutive (Hilbert-ordered) cells plus its ring-1 cells (their limiter is recomputed here) plus
// the cells those read (tiling.h).  One CTA per SM walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...  For each tile:
//   P1  one thread per own/ring-1 cell: min/max over the neighbours, Venkatakrishnan limiter (calc_limiters,
//       solver.h:517-593), MUSCL face states q + lim*(g.d) (solver.h:774-781) stored per local face side
//   P2  one thread per local face: Roe / boundary flux (physics.h:160-530) times the face length, stored over the
//       left face state
//   P3  one thread per own cell: owner-computes gather in slot (= ascending reference edge) order, /area, stage update
//       (solver.h:787-798, 818-821), wall ghosts of the next stage, residual norm, halo push
// The phases read shared memory only.  Their inputs form three groups -- (1) states, gradients, face offsets, areas and
// local connectivity of the cells, (2) face normals/lengths/kinds, (3) iteration-start state and time step -- and a
// group is requested as soon as the phase that reads it has finished with the previous tile: contiguous pieces by bulk
// copies (cp.async.bulk, completion on an mbarrier), the ring cells' states and gradients by 16-byte cp.async gathers.
// So the copies of tile k+1 fly under the arithmetic of tile k and the kernel is bound by the fp64 pipe / HBM, not by
// load latency.
// The limiter and flux arrays of the three-kernel stage never exist.  Faces cut by a tile boundary are evaluated by both
// tiles (identical inputs -> identical bits), ring-1 limiters are recomputed; every per-cell sum keeps the reference's
// order, so in strict mode the state is bit-identical to the unfused kernels and to the CPU reference.
#pragma once
#include <algorithm>

#include "rans_kernels.cuh"

// measured best on B200 (profiles/): two CTAs of 256 threads per SM, tiles of 192 cells
#ifndef AFX_STAGE_THREADS
#define AFX_STAGE_THREADS 256
#endif
#ifndef AFX_STAGE_MINB
#define AFX_STAGE_MINB 2
#endif

namespace afx {
namespace AFX_NS {

// ---- copy-engine and mbarrier primitives (PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // bounded: a copy that never lands (a bug) must fail the launch, not hang the device
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 24)) __trap();
    }
}
// global -> shared bulk copy; bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory accesses by 32-bit shared-window address: no generic-address arithmetic in the inner loops
__device__ __forceinline__ d4 lds_d4(uint32_t a)
{
    d4 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(v.z), "=d"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_d4(uint32_t a, const d4& v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
    asm volatile("st.shared.v2.f64 [%0+16], {%1, %2};" ::"r"(a), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double2 lds_d2(uint32_t a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_d(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_d(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, const uint4& v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

struct TileView {  // one tile's header, unpacked
    uint32_t cell0, nc, n1, n1p, nf, nh, nhp, off_halo, off_cell, off_face;
};
__device__ __forceinline__ TileView tile_view(const TileTab& tt, uint32_t tile)
{
    const uint4 a = tt.head[2 * (size_t)tile], b = tt.head[2 * (size_t)tile + 1];
    TileView v;
    v.cell0 = a.x; v.nc = a.y; v.n1 = a.y + a.z; v.n1p = (v.n1 + 1u) & ~1u; v.nh = a.z + a.w; v.nhp = (v.nh + 3u) & ~3u;
    v.nf = b.x; v.off_halo = b.y; v.off_cell = b.z; v.off_face = b.w;
    return v;
}

template <int LAST>
__global__ void __launch_bounds__(AFX_STAGE_THREADS, AFX_STAGE_MINB)
k_stage(DevMesh m, TileTab tt, const d4* __restrict__ qk_in, const d4* q0, d4* qk_out, const d4* __restrict__ gx,
        const d4* __restrict__ gy, const double* __restrict__ dt, d4* __restrict__ qW, d4* __restrict__ lim, double alpha,
        const double* __restrict__ prm, GasC g, NormOut no, PushArgs push)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const StageSmem L = stage_smem_layout(tt.max_loc, tt.max_n1, tt.max_nf, tt.max_nc, tt.max_halo);
    const uint32_t sb = smem_u32(smem);
    // shared-window addresses of the staging arrays (rans_types.h: StageSmem)
    const uint32_t a_sq = sb + L.sq, a_gx = sb + L.sgx, a_gy = sb + L.sgy, a_dxy = sb + L.sdxy, a_k3a = sb + L.sarea, a_ctab = sb + L.sctab;
    const uint32_t a_fg = sb + L.sfg, a_rec = sb + L.srec;
    const uint32_t a_q0 = sb + L.sq0, a_dt = sb + L.sdt, a_area3 = sb + L.sarea3, a_ctab3 = sb + L.sctab3, a_halo = sb + L.shalo;
    const uint32_t bar1 = sb + L.bars, bar2 = bar1 + 8, bar3 = bar1 + 16;
    const uint32_t tid = threadIdx.x;
    const bool q0_is_in = (q0 == qk_in);  // stage 0: the stage state IS the iteration-start state

    if (tid == 0) { mbar_init(bar1, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // ---- requests ----
    // group 1 of tile v (its ring ids are in halo buffer `buf`); also the ring ids of the tile after it into the other buffer
    auto request_g1 = [&](const TileView& v, uint32_t buf, bool has_next, const TileView& nx) {
        if (tid == 0) {
            const uint32_t bytes = 32u * v.nc * 3u + 64u * v.n1p + 8u * v.n1p + 16u * v.n1p + (has_next ? 4u * nx.nhp : 0u);
            mbar_expect_tx(bar1, bytes);
            bulk_g2s(a_sq, qk_in + v.cell0, 32u * v.nc, bar1);
            bulk_g2s(a_gx, gx + v.cell0, 32u * v.nc, bar1);
            bulk_g2s(a_gy, gy + v.cell0, 32u * v.nc, bar1);
            bulk_g2s(a_dxy, tt.dxy_t + 4 * (size_t)v.off_cell, 64u * v.n1p, bar1);
            bulk_g2s(a_k3a, tt.k3a_t + v.off_cell, 8u * v.n1p, bar1);
            bulk_g2s(a_ctab, tt.ctab + v.off_cell, 16u * v.n1p, bar1);
            if (has_next && nx.nhp) bulk_g2s(a_halo + (buf ^ 1u) * 4u * tt.max_halo, tt.halo + nx.off_halo, 4u * nx.nhp, bar1);
        }
        const uint32_t ids = a_halo + buf * 4u * tt.max_halo;
        for (uint32_t k = tid; k < v.nh; k += AFX_STAGE_THREADS) {  // ring cells: states, and gradients of ring 1
            const uint32_t gid = lds_u32(ids + 4u * k), l = v.nc + k;
            const char* s = reinterpret_cast<const char*>(qk_in + gid);
            cp_async16(a_sq + 32u * l, s); cp_async16(a_sq + 32u * l + 16, s + 16);
            if (l < v.n1) {
                const char* a = reinterpret_cast<const char*>(gx + gid);
                const char* b = reinterpret_cast<const char*>(gy + gid);
                cp_async16(a_gx + 32u * l, a); cp_async16(a_gx + 32u * l + 16, a + 16);
                cp_async16(a_gy + 32u * l, b); cp_async16(a_gy + 32u * l + 16, b + 16);
            }
        }
        cp_async_commit();
    };
    auto request_g2 = [&](const TileView& v) {
        if (tid == 0) {
            mbar_expect_tx(bar2, 32u * v.nf);
            bulk_g2s(a_fg, tt.fgeo_t + v.off_face, 32u * v.nf, bar2);
        }
    };
    auto request_g3 = [&](const TileView& v) {
        if (!q0_is_in && tid == 0) {
            mbar_expect_tx(bar3, 32u * v.nc);
            bulk_g2s(a_q0, q0 + v.cell0, 32u * v.nc, bar3);
        }
        for (uint32_t l = tid; l < v.nc; l += AFX_STAGE_THREADS) {
            cp_async8(a_dt + 8u * l, dt + v.cell0 + l);
            cp_async8(a_area3 + 8u * l, m.area + v.cell0 + l);
        }
        cp_async_commit();
    };

    double nrm = 0;
    uint32_t tile = blockIdx.x, it = 0;
    TileView cur{}, nxt{};
    if (tile < tt.n_tiles) {
        cur = tile_view(tt, tile);
        for (uint32_t k = tid; k < cur.nhp; k += AFX_STAGE_THREADS) sts_u32(a_halo + 4u * k, tt.halo[cur.off_halo + k]);  // first tile: ids by plain loads
        __syncthreads();
        const bool has_next = tile + gridDim.x < tt.n_tiles;
        if (has_next) nxt = tile_view(tt, tile + gridDim.x);
        request_g1(cur, 0, has_next, nxt);
        request_g2(cur);
    }
    const d4 zero = mk4(0, 0, 0, 0);
    for (; tile < tt.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1u;
        const bool has_next = tile + gridDim.x < tt.n_tiles;
        const bool has_next2 = tile + 2 * gridDim.x < tt.n_tiles;
        TileView nx2{};
        if (has_next2) nx2 = tile_view(tt, tile + 2 * gridDim.x);

        // ---- P1: limiter and MUSCL face states of the own and ring-1 cells ----
        cp_async_wait<0>();      // my gathers of group 1
        mbar_wait(bar1, par);
        __syncthreads();         // everybody's gathers; P3 of the previous tile is over
        request_g3(cur);         // group 3 of THIS tile: P3 of the previous tile read these buffers until the barrier above
        for (uint32_t l = tid; l < cur.n1; l += AFX_STAGE_THREADS) {
            const uint4 tc = lds_u4(a_ctab + 16u * l);
            const uint32_t nb[4] = {tc.x & 0xFFFFu, tc.x >> 16, tc.y & 0xFFFFu, tc.y >> 16};  // an empty slot points at l itself
            const uint32_t fs[4] = {tc.z & 0xFFFFu, tc.z >> 16, tc.w & 0xFFFFu, tc.w >> 16};
            double2 dxy[4];  // (0, 0) in empty slots
#pragma unroll
            for (int s = 0; s < 4; ++s) dxy[s] = lds_d2(a_dxy + 16u * (s * cur.n1p + l));
            const d4 gxi = lds_d4(a_gx + 32u * l), gyi = lds_d4(a_gy + 32u * l);
            const double K3a = lds_d(a_k3a + 8u * l);
            const d4 qi = lds_d4(a_sq + 32u * l);
            d4 lo = qi, hi = qi;
            unsigned valid = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                valid |= (nb[s] != l ? 1u : 0u) << s;
                const d4 qj = lds_d4(a_sq + 32u * nb[s]);  // wall ghosts hold their owner's state (previous stage / k_dt_grad)
                lo.x = dmin2(lo.x, qj.x); lo.y = dmin2(lo.y, qj.y); lo.z = dmin2(lo.z, qj.z); lo.w = dmin2(lo.w, qj.w);
                hi.x = dmax2(hi.x, qj.x); hi.y = dmax2(hi.y, qj.y); hi.z = dmax2(hi.z, qj.z); hi.w = dmax2(hi.w, qj.w);
                // the cell across an own cell's face is own, ring 1 or a ghost; a ghost cell has zero gradient and
                // limiter 1, so its face state is its state
                if (nb[s] >= cur.n1 && fs[s] != 0xFFFFu) sts_d4(a_rec + 64u * (fs[s] & 0x7FFFu) + 32u, qj);
            }
#if AFX_FAST
            valid = 0xFu;  // empty slots project to 0 and change neither pmax nor pmin
#endif
            const d4 lm = limiter_value(qi, lo, hi, gxi, gyi, dxy, valid, K3a);
            if (l < cur.nc) {  // what P3 needs after group 1 has been overwritten by the next tile
                sts_u4(a_ctab3 + 16u * l, tc);
                if (q0_is_in) sts_d4(a_q0 + 32u * l, qi);
                if (LAST && prm[2] != 0.0) lim[cur.cell0 + l] = lm;  // kept, like qW, for the last iteration of a run only
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (fs[s] == 0xFFFFu) continue;
                const double dx = dxy[s].x, dy = dxy[s].y;
                d4 r;  // solver.h:774-781
                r.x = qi.x + (gxi.x * dx + gyi.x * dy) * lm.x;
                r.y = qi.y + (gxi.y * dx + gyi.y * dy) * lm.y;
                r.z = qi.z + (gxi.z * dx + gyi.z * dy) * lm.z;
                r.w = qi.w + (gxi.w * dx + gyi.w * dy) * lm.w;
                sts_d4(a_rec + 32u * (2u * (fs[s] & 0x7FFFu) + (fs[s] >> 15)), r);
            }
        }
        __syncthreads();
        if (has_next) request_g1(nxt, par ^ 1u, has_next2, nx2); else cp_async_commit();

        // ---- P2: one flux per local face ----
        mbar_wait(bar2, par);
        for (uint32_t lf = tid; lf < cur.nf; lf += AFX_STAGE_THREADS) {
            const d4 gA = lds_d4(a_fg + 32u * lf);
            const d4 qL = lds_d4(a_rec + 64u * lf), qR = lds_d4(a_rec + 64u * lf + 32u);
            d4 fl = face_flux<0>((int)gA.w, qL, qR, zero, zero, gA.x, gA.y, g);
            fl.x *= gA.z; fl.y *= gA.z; fl.z *= gA.z; fl.w *= gA.z;
            sts_d4(a_rec + 64u * lf, fl);
            sts_d(a_rec + 64u * lf + 32u, gA.w);  // P3 needs the kind of boundary faces after group 2 has been overwritten
        }
        __syncthreads();
        if (has_next) request_g2(nxt);

        // ---- P3: gather, update, ghosts, norm ----
        cp_async_wait<1>();      // my time steps and areas (group 1 of the next tile may still be in flight)
        if (!q0_is_in) mbar_wait(bar3, par);
        for (uint32_t l = tid; l < cur.nc; l += AFX_STAGE_THREADS) {
            const uint32_t i = cur.cell0 + l;
            const uint4 tc = lds_u4(a_ctab3 + 16u * l);
            const uint32_t nb[4] = {tc.x & 0xFFFFu, tc.x >> 16, tc.y & 0xFFFFu, tc.y >> 16};
            const uint32_t fs[4] = {tc.z & 0xFFFFu, tc.z >> 16, tc.w & 0xFFFFu, tc.w >> 16};
            d4 r = zero;
            uint32_t wall_ghost[4] = {CF_NONE, CF_NONE, CF_NONE, CF_NONE};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (fs[s] == 0xFFFFu) continue;
                const uint32_t lf = fs[s] & 0x7FFFu;
                const d4 fl = lds_d4(a_rec + 64u * lf);
                if (fs[s] >> 15) { r.x += fl.x; r.y += fl.y; r.z += fl.z; r.w += fl.w; }
                else { r.x -= fl.x; r.y -= fl.y; r.z -= fl.z; r.w -= fl.w; }
                if (nb[s] >= cur.n1) {  // boundary face (rare)
                    const int kind = (int)lds_d(a_rec + 64u * lf + 32u);
                    if (LAST && kind == K_INTERNAL)  // two-sided boundary face: the ghost row of qW holds +flux
                        nrm += fl.x * fl.x + fl.y * fl.y + fl.z * fl.z + fl.w * fl.w;
                    if (kind == K_SLIPWALL || kind == K_WALL) wall_ghost[s] = tt.halo[cur.off_halo + nb[s] - cur.nc];
                }
            }
            const double A = lds_d(a_area3 + 8u * l);
#if AFX_FAST
            const double rA = fast_rcp(A);
            r.x *= rA; r.y *= rA; r.z *= rA; r.w *= rA;
#else
            r.x /= A; r.y /= A; r.z /= A; r.w /= A;
#endif
            const d4 qs = lds_d4(a_q0 + 32u * l);
            const double dti = lds_d(a_dt + 8u * l);
            const double relax = prm[1];
            d4 o;
            o.x = qs.x + r.x * dti * alpha * relax;
            o.y = qs.y + r.y * dti * alpha * relax;
            o.z = qs.z + r.z * dti * alpha * relax;
            o.w = qs.w + r.w * dti * alpha * relax;
            qk_out[i] = o;
            if (push.enabled && i < push.n_front) {  // halo push: straight into the peers' receive buffers over NVLink
                const unsigned long long pr = (*push.epoch) & 1ull;
                for (uint32_t k = push.dst_ptr[i]; k < push.dst_ptr[i + 1]; ++k) {
                    const uint32_t d = push.dst[k], p = d >> 28, slot = d & 0x0FFFFFFFu;
                    push.peer_buf[p][pr * push.peer_stride[p] + slot] = o;
                }
                __threadfence_system();
            }
            // wall ghosts of the next stage follow their owner; after the last stage the ghost keeps the state its owner
            // had when the stage started (solver.h:811 ran before the update)
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (wall_ghost[s] != CF_NONE) qk_out[wall_ghost[s]] = LAST ? (q0_is_in ? qs : qk_in[i]) : o;
            if (LAST) {
                if (prm[2] != 0.0) qW[i] = r;  // prm[2]: keep qW (only the last iteration of a run needs it)
                nrm += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
            }
        }
        cur = nxt; nxt = nx2;
    }
    cp_async_wait<0>();
    if (LAST) block_norm_partial(nrm, no, blockIdx.x);
}

// K^3 a of every tile-local cell record from its area (solver.h:551-552); run when limiter_k changes
__global__ void k_tile_k3a(const double* __restrict__ area_t, double* __restrict__ k3a_t, size_t n, double limiter_k)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) k3a_t[i] = limiter_k3a(area_t[i], limiter_k);
}

// kinds of the tiles' local faces (fgeo_t.w) from the face kinds set by set_bcs
__global__ void k_tile_face_kinds(d4* __restrict__ fgeo_t, const uint32_t* __restrict__ tile_face, const uint8_t* __restrict__ fkind, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fgeo_t[i].w = (double)fkind[tile_face[i]];
}

namespace launch {

static int stage_threads() { return AFX_STAGE_THREADS; }

static int stage_prepare(size_t smem)
{
    // opt in to the device maximum once: the attribute belongs to the function, not to a solver handle
    int dev = 0, smem_max = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cudaFuncAttributes fa0{}, fa1{};  // static shared memory (the norm reduction) counts against the same limit
    if (cudaFuncGetAttributes(&fa0, k_stage<0>) != cudaSuccess || cudaFuncGetAttributes(&fa1, k_stage<1>) != cudaSuccess) { cudaGetLastError(); return -1; }
    const int dyn_max = smem_max - (int)std::max(fa0.sharedSizeBytes, fa1.sharedSizeBytes);
    if ((int)smem > dyn_max) return -1;
    if (cudaFuncSetAttribute(k_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (cudaFuncSetAttribute(k_stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max) != cudaSuccess) { cudaGetLastError(); return -1; }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stage<0>, AFX_STAGE_THREADS, smem) != cudaSuccess) { cudaGetLastError(); return -1; }
    return nb;
}

static void stage(int last, const DevMesh& m, const TileTab& tt, unsigned grid, size_t smem, const d4* qk_in, const d4* q0, d4* qk_out,
                  const d4* gx, const d4* gy, const double* dt, d4* qW, d4* lim, double alpha, const double* prm,
                  const GasC& g, NormOut no, const PushArgs* push_in, cudaStream_t st)
{
    if (!grid || !tt.n_tiles) return;
    PushArgs push{};
    if (push_in) push = *push_in;
    if (no.blk_total == 0) { no.blk_off = 0; no.blk_total = grid; }
    if (last) k_stage<1><<<grid, AFX_STAGE_THREADS, smem, st>>>(m, tt, qk_in, q0, qk_out, gx, gy, dt, qW, lim, alpha, prm, g, no, push);
    else k_stage<0><<<grid, AFX_STAGE_THREADS, smem, st>>>(m, tt, qk_in, q0, qk_out, gx, gy, dt, qW, lim, alpha, prm, g, no, push);
}

static void tile_k3a(const double* area_t, double* k3a_t, size_t n, double limiter_k, cudaStream_t st)
{
    if (n) k_tile_k3a<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(area_t, k3a_t, n, limiter_k);
}

static void tile_face_kinds(d4* fgeo_t, const uint32_t* tile_face, const uint8_t* fkind, size_t n, cudaStream_t st)
{
    if (n) k_tile_face_kinds<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fgeo_t, tile_face, fkind, n);
}

}  // namespace launch
}  // namespace AFX_NS
}  // namespace afx
