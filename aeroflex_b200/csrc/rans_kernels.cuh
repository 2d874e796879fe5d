// rans_kernels.cuh -- sm_100a kernels of the rans hot path.
//
// Data layout in HBM (all in the solver's internal, renumbered order):
//   cell states q, qk*, gx, gy, lim, qW : AoS, one 32-byte d4 per cell (real
//       cells first, ghost cells after) -> one 256-bit load per gathered cell,
//       one full 32-byte DRAM sector per access;
//   faces: cells (uint2), geomA = {nx, ny, len, w_gg}, geomB = {d0x, d0y, d1x,
//       d1y} (face centre minus the two cell centres), kind (u8), all indexed
//       by face and read fully coalesced by the face kernel;
//   cell->face lists: cf[slot][cell], slot-major so a warp reads consecutive
//       words; slots are sorted by the ORIGINAL edge id, so per-cell sums run
//       in the order of the reference's serial edge loops (deterministic, and
//       bit-identical to the CPU result).
//   flux: one d4 per face, written once by the face kernel (each face is
//       evaluated once, in the reference's cell0 -> cell1 orientation) and
//       gathered by the two owner cells (owner-computes scatter, no atomics).
#pragma once
#include "rans_physics.cuh"

// CTA size and resident CTAs per SM asked of ptxas for the stage kernels (register cap = 65536 / (threads * n)).
// Measured on B200 at 1M cells (profiles/r01_s3_ab_stage_threads.jsonl): 128-thread CTAs for k_flux and k_limiter
// 0.3953 ms per iteration (flux phase 0.161, limiter 0.093) against 0.4030 with 256 (0.173, 0.097); k_gather_update is
// better at 256 (0.114-0.118 against 0.135 at 128).
#ifndef AFX_FLUX_THREADS
#define AFX_FLUX_THREADS 128
#endif
#ifndef AFX_LIM_THREADS
#define AFX_LIM_THREADS 128
#endif
#ifndef AFX_GATHER_THREADS
#define AFX_GATHER_THREADS 256
#endif
#ifndef AFX_FLUX_MINB
#define AFX_FLUX_MINB (1024 / AFX_FLUX_THREADS)
#endif
#ifndef AFX_LIM_MINB
#define AFX_LIM_MINB (1024 / AFX_LIM_THREADS)
#endif
// k_dt_grad tuning, measured on B200 at 1M cells with the first-stage limiter inside the kernel (profiles/r01_s3_ab_*.jsonl,
// ms per explicit iteration): 128 threads x 6 CTAs 0.405 | 256 x 3 0.411-0.417 | + L2 prefetch of the limiter's face
// offsets 0.407-0.408 | offsets loaded before the dependency wait 0.425 | 64 registers (256 x 4) 0.441 | neighbour preload:
// 128 registers (256 x 2) 0.415, + prefetch 0.405, capped at 80 registers 0.428.
//   AFX_DTG_PRELOAD=1 requests the four neighbour states together and writes the wall ghosts after the last read; it needs
//   ~128 registers to hold them without spilling a value that is still in flight, and the limiter epilogue then runs at
//   16 warps per SM: no gain over one gather at a time at 24 warps (the kernel is not bound by gather latency alone).
//   AFX_DTG_DXY: where the first-stage limiter gets its face offsets: 0 loads them in its epilogue, 1 also prefetches them
//   into L2 at kernel start, 2 loads them before the dependency wait (8 more registers through the gradient loop).
#ifndef AFX_DTG_THREADS
#define AFX_DTG_THREADS 128
#endif
#ifndef AFX_DTG_PRELOAD
#define AFX_DTG_PRELOAD 0
#endif
// Round 2 (profiles/r02k_ab_{1M,16M}.jsonl, same box, one process per build): with the stored-extremes epilogue (k_dt_grad<., 2>) the
// kernel spills 304 B per thread at 80 registers (128 x 6); at 96 registers (128 x 5, 20 warps per SM) 20 B: 0.3812 -> 0.3702 ms per
// iteration at 1M cells (dt/gradient phase 0.109 -> 0.101), 5.711 -> 5.679 at 16M; 128 x 4 (118 registers, no spills) 0.3746 / 5.794.
#ifndef AFX_DTG_MINB
#define AFX_DTG_MINB (AFX_DTG_PRELOAD ? 512 / AFX_DTG_THREADS : 640 / AFX_DTG_THREADS)
#endif
#ifndef AFX_DTG_DXY
#define AFX_DTG_DXY 0
#endif
//   AFX_DTG_DEFER=1 (without the preload): the wall-ghost stores move behind the slot loop, so no store sits between two
//   neighbour gathers and the compiler may overlap them as far as its register budget allows (same values, same bits).
#ifndef AFX_DTG_DEFER
#define AFX_DTG_DEFER 0
#endif

namespace afx {
namespace AFX_NS {

// Programmatic dependent launch: the four kernels of the explicit iteration are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the CTAs of kernel k+1 start as the last wave of kernel k
// drains, read their STATIC inputs (connectivity, geometry) and only then wait for kernel k to complete.  Anything a
// predecessor writes is read after pdl_wait() and through plain (coherent) loads: the pointers to solver state carry
// no __restrict__ in these kernels, because a dependent CTA can be resident while an older kernel still runs on its SM and
// ld.global.nc may then serve lines that were cached before the predecessor's writes.  Kernels launched without the
// attribute see both calls as no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// spectral radius c + |V.n| of one state, solver.h:329-336, split into the part that belongs to the cell -- the speed
// of sound, one square root (and in strict mode the division 0.5/rho) -- and the part that belongs to the face.  The
// split evaluates exactly the reference's expressions, so the sum is bit-identical to computing both per face.
struct CellSound {
    double c;   // sqrt(gamma p / rho)
    double r;   // fast mode: 1/rho, shared with the normal velocity
};
__device__ __forceinline__ CellSound cell_sound(const d4& q, double gam)
{
    CellSound s;
#if AFX_FAST
    s.r = fast_rcp(q.x);
    const double p = (gam - 1) * (q.w - 0.5 * s.r * (q.y * q.y + q.z * q.z));
    s.c = fast_sqrt(p * gam * s.r);
#else
    const double p = (gam - 1) * (q.w - 0.5 / q.x * (q.y * q.y + q.z * q.z));
    s.c = sqrt(p * gam / q.x);
    s.r = 0;
#endif
    return s;
}
__device__ __forceinline__ double spectral_radius(const d4& q, const CellSound& s, double nx, double ny)
{
#if AFX_FAST
    return s.c + fabs((q.y * nx + q.z * ny) * s.r);
#else
    return s.c + fabs((q.y * nx + q.z * ny) / q.x);
#endif
}

// ---- Venkatakrishnan limiter function (used by k_dt_grad<.,1>, k_limiter and the fused k_stage) ----
__device__ __forceinline__ double venkat(double dqg, double dmax, double dmin, double K3a)
{
    if (dqg > 1e-16)
        return 1 / dqg * ((dmax * dmax + K3a) * dqg + 2 * dqg * dqg * dmax) / (dmax * dmax + 2 * dqg * dqg + dmax * dqg + K3a);
    if (dqg < -1e-16)
        return 1 / dqg * ((dmin * dmin + K3a) * dqg + 2 * dqg * dqg * dmin) / (dmin * dmin + 2 * dqg * dqg + dmin * dqg + K3a);
    return 1.0;
}

// K^3 a of the Venkatakrishnan function, solver.h:551-552
__device__ __forceinline__ double limiter_k3a(double area, double limiter_k)
{
    const double Ka = limiter_k * sqrt(area);
    return Ka * Ka * Ka;
}

#if AFX_FAST
// min(1, phi(pmax), phi(pmin)) with ONE reciprocal: both denominators are positive (dm and the increment have the same
// sign), so the smaller fraction is found by cross-multiplying
__device__ __forceinline__ double venkat_pair(double pmax, double pmin, double dmax, double dmin, double K3a)
{
    double n1 = 1.0, d1 = 1.0, n2 = 1.0, d2 = 1.0;
    if (pmax > 1e-16) { n1 = dmax * dmax + K3a + 2 * pmax * dmax; d1 = dmax * dmax + 2 * pmax * pmax + dmax * pmax + K3a; }
    if (pmin < -1e-16) { n2 = dmin * dmin + K3a + 2 * pmin * dmin; d2 = dmin * dmin + 2 * pmin * pmin + dmin * pmin + K3a; }
    const bool first = n1 * d2 < n2 * d1;
    const double n = first ? n1 : n2, d = first ? d1 : d2;
    return n >= d ? 1.0 : n * fast_rcp(d);
}
#endif

#if AFX_FAST
// largest positive / most negative projected increment g . (x_f - x_c) over the cell's faces, per component.  The gradients
// are those of the iteration-start state for all three stages and the offsets are geometry: these eight numbers
// are CONSTANT over an iteration.  k_dt_grad stores them (`pm`, 64 B per cell) and the limiter kernels of stages 2 and 3 read
// them instead of gx, gy and the four face offsets (128 B per cell): same values, same bits, 120 B per cell and launch less.
__device__ __forceinline__ void projected_extremes(const d4& gxi, const d4& gyi, const double2 (&dxy)[4], unsigned valid, d4& pmax, d4& pmin)
{
    pmax = mk4(0, 0, 0, 0); pmin = mk4(0, 0, 0, 0);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!(valid & (1u << s))) continue;
        const double dx = dxy[s].x, dy = dxy[s].y;
        const double p0 = gxi.x * dx + gyi.x * dy, p1 = gxi.y * dx + gyi.y * dy, p2 = gxi.z * dx + gyi.z * dy, p3 = gxi.w * dx + gyi.w * dy;
        pmax.x = dmax2(pmax.x, p0); pmax.y = dmax2(pmax.y, p1); pmax.z = dmax2(pmax.z, p2); pmax.w = dmax2(pmax.w, p3);
        pmin.x = dmin2(pmin.x, p0); pmin.y = dmin2(pmin.y, p1); pmin.z = dmin2(pmin.z, p2); pmin.w = dmin2(pmin.w, p3);
    }
}
__device__ __forceinline__ d4 limiter_from_extremes(const d4& dmax, const d4& dmin, const d4& pmax, const d4& pmin, double K3a)
{
    d4 l;
    l.x = venkat_pair(pmax.x, pmin.x, dmax.x, dmin.x, K3a);
    l.y = venkat_pair(pmax.y, pmin.y, dmax.y, dmin.y, K3a);
    l.z = venkat_pair(pmax.z, pmin.z, dmax.z, dmin.z, K3a);
    l.w = venkat_pair(pmax.w, pmin.w, dmax.w, dmin.w, K3a);
    return l;
}
#endif

// limiter of one cell from its state, the min/max over its neighbours, its gradient and the face offsets of its slots
// (solver.h:538-592); `valid` bit s = slot s holds a face.  Shared by k_limiter and the fused k_stage.
__device__ __forceinline__ d4 limiter_value(const d4& qi, const d4& lo, const d4& hi, const d4& gxi, const d4& gyi, const double2 (&dxy)[4],
                                            unsigned valid, double K3a)
{
    const d4 dmax = mk4(hi.x - qi.x, hi.y - qi.y, hi.z - qi.z, hi.w - qi.w);
    const d4 dmin = mk4(lo.x - qi.x, lo.y - qi.y, lo.z - qi.z, lo.w - qi.w);
    d4 l = mk4(1, 1, 1, 1);
#if AFX_FAST
    // Where the limiter function is below 1 it decreases monotonically with |dqg| (d phi/d dqg < 0 for dqg > dm/2), and
    // values above 1 never survive the min with 1: the minimum over the faces is attained at the largest positive and
    // the most negative projected increment -> one pair of evaluations per component instead of one per face.
    d4 pmax, pmin;
    projected_extremes(gxi, gyi, dxy, valid, pmax, pmin);
    l = limiter_from_extremes(dmax, dmin, pmax, pmin, K3a);
#else
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!(valid & (1u << s))) continue;
        const double dx = dxy[s].x, dy = dxy[s].y;
        l.x = dmin2(l.x, venkat(gxi.x * dx + gyi.x * dy, dmax.x, dmin.x, K3a));
        l.y = dmin2(l.y, venkat(gxi.y * dx + gyi.y * dy, dmax.y, dmin.y, K3a));
        l.z = dmin2(l.z, venkat(gxi.z * dx + gyi.z * dy, dmax.z, dmin.z, K3a));
        l.w = dmin2(l.w, venkat(gxi.w * dx + gyi.w * dy, dmax.w, dmin.w, K3a));
    }
#endif
    return l;
}

// ---------------------------------------------------------------------------
// Local time step + gradients of q, one thread per real cell.
// calc_dt (solver.h:308-356), set_walls_from_internal (289-305, folded in: the
// owner writes its wall ghosts), calc_gradients Green-Gauss (428-469) or
// least-squares (470-513).  Source state is the iteration-start q.
// ---------------------------------------------------------------------------
// LIM = 1 also writes the limiters of the FIRST Runge-Kutta stage (calc_limiters, solver.h:517-593): that stage limits
// the iteration-start state, whose neighbour states this kernel has just read and whose gradient it has just computed,
// so the first k_limiter launch of the iteration -- a second pass over q, gx, gy, the neighbour table and the areas --
// is saved.  Same limiter_value() on the same inputs in the same order: the bits are those of k_limiter.
template <int GRAD, int LIM>
__global__ void __launch_bounds__(AFX_DTG_THREADS, AFX_DTG_MINB) k_dt_grad(DevMesh m, d4* q, double* dt,
                                                 d4* gx, d4* gy, const double* __restrict__ prm,
                                                 double gam, int want_grad, int walls, d4* lim, double limiter_k, d4* pm)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n_grad) return;
    pdl_launch_dependents();
    uint32_t nbv[4];
    d4 geo[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {  // independent coalesced loads first: neighbour | flags and the slot's face geometry
        nbv[s] = m.cnb[(size_t)s * m.N + i];
        geo[s] = m.cgeo[(size_t)s * m.N + i];
    }
#if AFX_DTG_DXY == 1
    if (LIM) {  // the limiter epilogue's face offsets: on their way to L2 while the gradients are computed, no register held
#pragma unroll
        for (int s = 0; s < 4; ++s) prefetch_l2(&m.cdxy[(size_t)s * m.N + i]);
    }
#elif AFX_DTG_DXY == 2
    double2 dxy_early[4];
    if (LIM) {
#pragma unroll
        for (int s = 0; s < 4; ++s) dxy_early[s] = m.cdxy[(size_t)s * m.N + i];
    }
#endif
    pdl_wait();  // the state comes from the previous kernel
    const d4 qi = q[i];
    double dsum = 0;
    d4 ax = mk4(0, 0, 0, 0), ay = mk4(0, 0, 0, 0);
    unsigned wall_slots = 0;  // slots whose neighbour is a wall ghost (it takes this cell's state)
#if AFX_DTG_PRELOAD
    // All four neighbour states are requested before any of them is used, and the wall ghosts are written after the last
    // read: a ghost store between two gathers would order them (the compiler must assume the addresses may coincide), and
    // this kernel is bound by the latency of exactly these gathers.  A wall ghost is read by its owner only.
    d4 qn[4];
    unsigned kinds = 0;  // 2 bits per slot
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t v = nbv[s];
        qn[s] = mk4(0, 0, 0, 0);
        if (v == CF_NONE) continue;
        const int kind = (v & CF_BND) ? (int)m.fkind[m.cf[(size_t)s * m.N + i] & CF_ID] : K_INTERNAL;
        const bool wall_ghost = walls && (v & CF_BND) && (kind == K_SLIPWALL || kind == K_WALL);
        kinds |= (unsigned)kind << (2 * s);
        if (wall_ghost) wall_slots |= 1u << s;
        else qn[s] = q[v & CF_ID];
    }
#endif
    const CellSound si = cell_sound(qi, gam);  // once per cell instead of once per face
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t v = nbv[s];
        if (v == CF_NONE) continue;
        const bool side = v & CF_SIDE;
        const uint32_t j = v & CF_ID;
        const d4 gA = geo[s];
#if AFX_DTG_PRELOAD
        const int kind = (int)((kinds >> (2 * s)) & 3u);
        const d4 qj = (wall_slots & (1u << s)) ? qi : qn[s];
        (void)j;
#else
        const int kind = (v & CF_BND) ? (int)m.fkind[m.cf[(size_t)s * m.N + i] & CF_ID] : K_INTERNAL;
        const bool wall_ghost = walls && (v & CF_BND) && (kind == K_SLIPWALL || kind == K_WALL);
#if AFX_DTG_DEFER
        if (wall_ghost) wall_slots |= 1u << s;
#else
        if (wall_ghost) q[j] = qi;  // ghost <- owner (set_walls_from_internal)
        if (LIM && wall_ghost) wall_slots |= 1u << s;
#endif
        const d4 qj = wall_ghost ? qi : q[j];
#endif
        const d4 qL = side ? qj : qi, qR = side ? qi : qj;  // states of cell0 / cell1
        const double nx = gA.x, ny = gA.y, len = gA.z;
        // spectral radius, solver.h:328-345
        // cell0's radius alone at one-sided (boundary) faces, where this cell is always cell0
        double eig = spectral_radius(qi, si, nx, ny);
        if (kind == K_INTERNAL) {
            const double eig_j = spectral_radius(qj, cell_sound(qj, gam), nx, ny);
            const double eig_L = side ? eig_j : eig, eig_R = side ? eig : eig_j;
            eig = (eig_L < eig_R) ? eig_R : eig_L;
        }
        dsum += eig * len;
        if (GRAD == 0 && want_grad) {  // Green-Gauss face value, solver.h:444-457
            const d4 qv = bc_vars(kind, qL, qR, nx, ny, gam);
            const double w = gA.w;
            const double f0 = (qL.x * (1.0 - w) + qv.x * w) * len;
            const double f1 = (qL.y * (1.0 - w) + qv.y * w) * len;
            const double f2 = (qL.z * (1.0 - w) + qv.z * w) * len;
            const double f3 = (qL.w * (1.0 - w) + qv.w * w) * len;
            if (!side) {
                ax.x += f0 * nx; ax.y += f1 * nx; ax.z += f2 * nx; ax.w += f3 * nx;
                ay.x += f0 * ny; ay.y += f1 * ny; ay.z += f2 * ny; ay.w += f3 * ny;
            } else {
                ax.x -= f0 * nx; ax.y -= f1 * nx; ax.z -= f2 * nx; ax.w -= f3 * nx;
                ay.x -= f0 * ny; ay.y -= f1 * ny; ay.z -= f2 * ny; ay.w -= f3 * ny;
            }
        }
    }
#if AFX_DTG_PRELOAD || AFX_DTG_DEFER
    if (wall_slots) {  // ghost <- owner (set_walls_from_internal)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (wall_slots & (1u << s)) q[nbv[s] & CF_ID] = qi;
    }
#endif
    const double A = m.area[i];
    dt[i] = prm[0] * A / dsum;  // prm[0] = cfl
    if (!want_grad) return;
    if (GRAD == 0) {
#if AFX_FAST
        const double rA = fast_rcp(A);
        ax.x *= rA; ax.y *= rA; ax.z *= rA; ax.w *= rA;
        ay.x *= rA; ay.y *= rA; ay.z *= rA; ay.w *= rA;
#else
        ax.x /= A; ax.y /= A; ax.z /= A; ax.w /= A;
        ay.x /= A; ay.y /= A; ay.z /= A; ay.w /= A;
#endif
    } else {  // least squares, rows in cellsEdges order, solver.h:471-508
        const uint32_t perm = m.lsq_perm[i];  // bits 0-7: slot of local side j (2 bits each); bits 8-10: number of sides
        const int nside = (int)(perm >> 8);
#pragma unroll
        for (int jrow = 0; jrow < 4; ++jrow) {
            if (jrow >= nside) break;
            const int s = (perm >> (2 * jrow)) & 3;
            const uint32_t cfv = m.cf[(size_t)s * m.N + i];
            const uint32_t f = cfv & CF_ID;
            const uint2 fc = m.fcells[f];
            const d4 gA = m.fgA[f];
            const int kind = m.fkind[f];
            const uint32_t j = (cfv & CF_SIDE) ? fc.x : fc.y;
            const d4 qn = q[j];
            const d4 qv = bc_vars(kind, qi, qn, gA.x, gA.y, gam);  // (q_p, q_n) whatever the orientation
            const double d0 = qi.x - qv.x, d1 = qi.y - qv.y, d2 = qi.z - qv.z, d3 = qi.w - qv.w;
            const double m0 = m.lsqM[(size_t)jrow * m.N + i], m1 = m.lsqM[(size_t)(4 + jrow) * m.N + i];
            ax.x += m0 * d0; ax.y += m0 * d1; ax.z += m0 * d2; ax.w += m0 * d3;
            ay.x += m1 * d0; ay.y += m1 * d1; ay.z += m1 * d2; ay.w += m1 * d3;
        }
    }
    gx[i] = ax;
    gy[i] = ay;
    if (LIM) {
        // min / max over the edge neighbours (solver.h:524-536) in an epilogue of its own: the neighbour states are read a
        // second time, from L1, instead of carrying eight more accumulators through the gradient loop
        double2 dxy[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#if AFX_DTG_DXY == 2
            dxy[s] = dxy_early[s];
#else
            dxy[s] = m.cdxy[(size_t)s * m.N + i];
#endif
        }
        d4 lo = qi, hi = qi;
        unsigned valid = 0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t v = nbv[s];
            if (v == CF_NONE) continue;
            valid |= 1u << s;
            if (wall_slots & (1u << s)) continue;  // the ghost holds qi: neither bound moves
            const d4 qj = q[v & CF_ID];
            lo.x = dmin2(lo.x, qj.x); lo.y = dmin2(lo.y, qj.y); lo.z = dmin2(lo.z, qj.z); lo.w = dmin2(lo.w, qj.w);
            hi.x = dmax2(hi.x, qj.x); hi.y = dmax2(hi.y, qj.y); hi.z = dmax2(hi.z, qj.z); hi.w = dmax2(hi.w, qj.w);
        }
#if AFX_FAST
        if (LIM == 2 || pm) {  // the projected extremes for the limiter kernels of the later stages (see projected_extremes); LIM == 2: known at compile time
            d4 pmax, pmin;
            projected_extremes(ax, ay, dxy, valid, pmax, pmin);
            pm[2 * (size_t)i] = pmax; pm[2 * (size_t)i + 1] = pmin;
            const d4 dmax = mk4(hi.x - qi.x, hi.y - qi.y, hi.z - qi.z, hi.w - qi.w);
            const d4 dmin = mk4(lo.x - qi.x, lo.y - qi.y, lo.z - qi.z, lo.w - qi.w);
            lim[i] = limiter_from_extremes(dmax, dmin, pmax, pmin, limiter_k3a(A, limiter_k));
            return;
        }
#endif
        if (LIM == 2) return;  // (fast mode only; the host never launches <., 2> in strict mode)
        lim[i] = limiter_value(qi, lo, hi, ax, ay, dxy, valid, limiter_k3a(A, limiter_k));
    }
}

// ---------------------------------------------------------------------------
// Venkatakrishnan limiter, one thread per real cell.  calc_limiters
// (solver.h:517-593): min/max over the edge neighbours (ghosts included) of the
// stage state, then the minimum of the limiter function over the cell's faces.
// ---------------------------------------------------------------------------
// static inputs of one cell's limiter (connectivity, geometry): loaded before anything the kernel has to wait for
struct LimCell {
    uint32_t nbs[4];
    double2 dxy[4];
    double area;
};
__device__ __forceinline__ LimCell limiter_load_static(const DevMesh& m, uint32_t i, bool want_dxy = true)
{
    LimCell c;
#pragma unroll
    for (int s = 0; s < 4; ++s) {  // independent coalesced loads
        c.nbs[s] = m.cnb[(size_t)s * m.N + i];
        c.dxy[s] = want_dxy ? m.cdxy[(size_t)s * m.N + i] : make_double2(0., 0.);
    }
    c.area = m.area[i];
    return c;
}
// the limiter of cell i from the stage state and the gradients (shared by k_limiter and the pipelined stage kernel k_pipe)
// PM: 1 = the stored projected extremes are there (compile-time: k_limiter<1> carries no registers for the other path),
//     0 = they are not, -1 = decided by `pm` at run time
template <int PM = -1>
__device__ __forceinline__ void limiter_cell(const DevMesh& m, uint32_t i, const LimCell& c, const d4* qk, const d4* gx, const d4* gy, d4* lim,
                                             double limiter_k, int walls, const d4* pm = nullptr)
{
    const uint32_t (&nbs)[4] = c.nbs;
    const double2 (&dxy)[4] = c.dxy;
    const double area_i = c.area;
    const d4 qi = qk[i];
    d4 lo = qi, hi = qi;
    unsigned valid = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (nbs[s] == CF_NONE) continue;
        const uint32_t j = nbs[s] & CF_ID;
        valid |= 1u << s;
        // a wall ghost holds its owner's state (set_walls_from_internal): no need to read it
        bool wall_ghost = false;
        if (walls && (nbs[s] & CF_BND)) {
            const int kind = m.fkind[m.cf[(size_t)s * m.N + i] & CF_ID];
            wall_ghost = (kind == K_SLIPWALL || kind == K_WALL);
        }
        const d4 qj = wall_ghost ? qi : qk[j];
        lo.x = dmin2(lo.x, qj.x); lo.y = dmin2(lo.y, qj.y); lo.z = dmin2(lo.z, qj.z); lo.w = dmin2(lo.w, qj.w);
        hi.x = dmax2(hi.x, qj.x); hi.y = dmax2(hi.y, qj.y); hi.z = dmax2(hi.z, qj.z); hi.w = dmax2(hi.w, qj.w);
    }
#if AFX_FAST
    if (PM == 1 || (PM == -1 && pm)) {  // extremes of the projected increments stored by k_dt_grad: neither the gradients nor the face offsets are read
        const d4 pmax = pm[2 * (size_t)i], pmin = pm[2 * (size_t)i + 1];
        const d4 dmax = mk4(hi.x - qi.x, hi.y - qi.y, hi.z - qi.z, hi.w - qi.w);
        const d4 dmin = mk4(lo.x - qi.x, lo.y - qi.y, lo.z - qi.z, lo.w - qi.w);
        lim[i] = limiter_from_extremes(dmax, dmin, pmax, pmin, limiter_k3a(area_i, limiter_k));
        return;
    }
#endif
    if (PM == 1) return;  // (strict mode never stores extremes; the host never launches <1> there)
    lim[i] = limiter_value(qi, lo, hi, gx[i], gy[i], dxy, valid, limiter_k3a(area_i, limiter_k));
}

#ifndef AFX_LIM_PM_MINB  // resident CTAs per SM asked for the stored-extremes instantiation (it needs fewer registers)
#define AFX_LIM_PM_MINB AFX_LIM_MINB
#endif
template <int PM>
__global__ void __launch_bounds__(AFX_LIM_THREADS, PM ? AFX_LIM_PM_MINB : AFX_LIM_MINB) k_limiter(DevMesh m, const d4* qk, const d4* gx,
                                                 const d4* gy, d4* lim, double limiter_k, int walls,
                                                 uint32_t lo1, uint32_t n1, uint32_t lo2, uint32_t n2, const d4* pm)
{
    // cells [lo1, lo1+n1) and [lo2, lo2+n2): a partitioned run limits its interior cells while the halo is in flight
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 + n2) return;
    const uint32_t i = t < n1 ? lo1 + t : lo2 + (t - n1);
    pdl_launch_dependents();
    const LimCell c = limiter_load_static(m, i, !PM);
    pdl_wait();  // stage state and gradients come from the previous kernels
    limiter_cell<PM>(m, i, c, qk, gx, gy, lim, limiter_k, walls, pm);
}

// ---- the reference's other limiter: Michalak (solver.h:557-576 under RANS_MICHALAK_LIMITER, physics.h:581-592) ----------------------
// Off in the reference's default build (nothing defines the macro); opt-in here with afx_rans_set_limiter.  Per component: a
// smooth switch sig between "no limiting" (the spread of the neighbour values is below K^3 a) and Michalak's monotone cubic of
// y = delta / (g . dx); lim = sig + (1 - sig) * phi(y), minimum over the cell's faces.  The reference's expressions in the
// reference's order in both arithmetic modes (the function is not on the benchmarked path; fast mode only allows FMA contraction).
__device__ __forceinline__ double michalak_phi(double y)  // physics.h:583-592, RANS_YT = 2
{
    if (y >= 2.0) return 1.0;
    constexpr double a = 1.0 / (2.0 * 2.0) - 2.0 / (2.0 * 2.0 * 2.0);
    constexpr double b = -3.0 / 2.0 * a * 2.0 - 0.5 / 2.0;
    return a * y * y * y + b * y * y + y;
}
__device__ __forceinline__ double michalak_one(double dqg, double dmax, double dmin, double K3a)
{
    const double dMaxMin2 = (dmax - dmin) * (dmax - dmin);
    double lim = 1.0, sig;
    if (dMaxMin2 <= K3a) sig = 1.;
    else if (dMaxMin2 <= 2 * K3a) { const double y = (dMaxMin2 / K3a - 1.0); sig = 2.0 * y * y * y - 3.0 * y * y + 1.0; }
    else sig = 0.;
    if (sig < 1.0) {
        if (dqg > 1e-14) lim = michalak_phi(dmax / dqg);
        else if (dqg < -1e-14) lim = michalak_phi(dmin / dqg);
        else lim = 1.0;
    }
    return sig + (1.0 - sig) * lim;
}
__global__ void __launch_bounds__(128) k_limiter_michalak(DevMesh m, const d4* qk, const d4* gx, const d4* gy, d4* lim, double limiter_k, int walls,
                                                          uint32_t lo1, uint32_t n1, uint32_t lo2, uint32_t n2)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n1 + n2) return;
    const uint32_t i = t < n1 ? lo1 + t : lo2 + (t - n1);
    pdl_launch_dependents();
    const LimCell c = limiter_load_static(m, i, true);
    pdl_wait();
    const d4 qi = qk[i];
    d4 lo = qi, hi = qi;
    unsigned valid = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {  // min / max over the edge neighbours, exactly as limiter_cell
        if (c.nbs[s] == CF_NONE) continue;
        const uint32_t j = c.nbs[s] & CF_ID;
        valid |= 1u << s;
        bool wall_ghost = false;
        if (walls && (c.nbs[s] & CF_BND)) {
            const int kind = m.fkind[m.cf[(size_t)s * m.N + i] & CF_ID];
            wall_ghost = (kind == K_SLIPWALL || kind == K_WALL);
        }
        const d4 qj = wall_ghost ? qi : qk[j];
        lo.x = dmin2(lo.x, qj.x); lo.y = dmin2(lo.y, qj.y); lo.z = dmin2(lo.z, qj.z); lo.w = dmin2(lo.w, qj.w);
        hi.x = dmax2(hi.x, qj.x); hi.y = dmax2(hi.y, qj.y); hi.z = dmax2(hi.z, qj.z); hi.w = dmax2(hi.w, qj.w);
    }
    const d4 gxi = gx[i], gyi = gy[i];
    const double K3a = limiter_k3a(c.area, limiter_k);
    const d4 dmax = mk4(hi.x - qi.x, hi.y - qi.y, hi.z - qi.z, hi.w - qi.w);
    const d4 dmin = mk4(lo.x - qi.x, lo.y - qi.y, lo.z - qi.z, lo.w - qi.w);
    d4 l = mk4(1, 1, 1, 1);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!(valid & (1u << s))) continue;
        const double dx = c.dxy[s].x, dy = c.dxy[s].y;
        l.x = dmin2(l.x, michalak_one(gxi.x * dx + gyi.x * dy, dmax.x, dmin.x, K3a));
        l.y = dmin2(l.y, michalak_one(gxi.y * dx + gyi.y * dy, dmax.y, dmin.y, K3a));
        l.z = dmin2(l.z, michalak_one(gxi.z * dx + gyi.z * dy, dmax.z, dmin.z, K3a));
        l.w = dmin2(l.w, michalak_one(gxi.w * dx + gyi.w * dy, dmax.w, dmin.w, K3a));
    }
    lim[i] = l;
}

// ---------------------------------------------------------------------------
// Face loop, one thread per face: MUSCL reconstruction + flux, written once.
// explicitSolver::calc_residual (solver.h:751-786) / fillRhoRHS (1097-1134) /
// get_uniform_residual (659-686, UNIFORM=1: both states are the far-field state).
// average_gradients (solver.h:359-398) feeds the laminar term with the
// iteration-start q.
// ---------------------------------------------------------------------------
// 256-bit loads that are cached in L2 only: for data written by OTHER CTAs of the running kernel (k_pipe), which a line
// this SM's L1 has kept from an earlier read would hide
__device__ __forceinline__ d4 ld_cg4(const d4* p)
{
    d4 v;
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
template <int CG>
__device__ __forceinline__ d4 ld4(const d4* p) { if (CG) return ld_cg4(p); return *p; }

// one 64-byte record per face: normal, length, both centre offsets, the two cells and the kind
struct FaceRec {
    d4 gA, gB;
    uint2 fc;
    int kind;
};
__device__ __forceinline__ FaceRec face_load_static(const DevMesh& m, uint32_t f)
{
    const d4 r0 = m.frec[2 * (size_t)f], r1 = m.frec[2 * (size_t)f + 1];
    const unsigned long long cw = (unsigned long long)__double_as_longlong(r1.w);
    FaceRec r;
    r.fc = make_uint2((uint32_t)cw & CF_ID, (uint32_t)(cw >> 32));
    r.kind = (int)(((uint32_t)cw) >> 30);
    r.gA = mk4(r0.x, r0.y, r0.z, 0.);
    r.gB = mk4(r0.w, r1.x, r1.y, r1.z);
    return r;
}
// MUSCL reconstruction + flux of face f, written once (shared by k_flux and k_pipe; CG: the limiters were written by
// other CTAs of the running kernel)
template <int SECOND, int VISC, int UNIFORM, int CG>
__device__ __forceinline__ void flux_face(const DevMesh& m, uint32_t f, const FaceRec& rec, const d4* qk, const d4* q0, const d4* gx, const d4* gy,
                                          const d4* lim, d4* flux, const GasC& g, const d4& qfar)
{
    const uint2 fc = rec.fc;
    const int kind = rec.kind;
    const d4 gA = rec.gA, gB = rec.gB;
    d4 qL, qR;
    if (UNIFORM) { qL = qfar; qR = qfar; }
    else { qL = qk[fc.x]; qR = qk[fc.y]; }
    d4 gL0, gL1, gR0, gR1;
    if ((SECOND && !UNIFORM) || VISC == 1) { gL0 = gx[fc.x]; gL1 = gy[fc.x]; gR0 = gx[fc.y]; gR1 = gy[fc.y]; }
    if (SECOND && !UNIFORM) {  // solver.h:774-781
        const d4 lL = ld4<CG>(lim + fc.x), lR = ld4<CG>(lim + fc.y);
        qL.x = qL.x + (gL0.x * gB.x + gL1.x * gB.y) * lL.x;
        qL.y = qL.y + (gL0.y * gB.x + gL1.y * gB.y) * lL.y;
        qL.z = qL.z + (gL0.z * gB.x + gL1.z * gB.y) * lL.z;
        qL.w = qL.w + (gL0.w * gB.x + gL1.w * gB.y) * lL.w;
        qR.x = qR.x + (gR0.x * gB.z + gR1.x * gB.w) * lR.x;
        qR.y = qR.y + (gR0.y * gB.z + gR1.y * gB.w) * lR.y;
        qR.z = qR.z + (gR0.z * gB.z + gR1.z * gB.w) * lR.z;
        qR.w = qR.w + (gR0.w * gB.z + gR1.w * gB.w) * lR.w;
    }
    d4 gfx = mk4(0, 0, 0, 0), gfy = mk4(0, 0, 0, 0);
    if (VISC == 1 && kind == K_INTERNAL) {  // face gradient, solver.h:369-397
        const d4 t = m.ftij[f];
        const d4 a = q0[fc.x], b = q0[fc.y];
        const double bx0 = (gL0.x + gR0.x) * 0.5, by0 = (gL1.x + gR1.x) * 0.5;
        const double bx1 = (gL0.y + gR0.y) * 0.5, by1 = (gL1.y + gR1.y) * 0.5;
        const double bx2 = (gL0.z + gR0.z) * 0.5, by2 = (gL1.z + gR1.z) * 0.5;
        const double bx3 = (gL0.w + gR0.w) * 0.5, by3 = (gL1.w + gR1.w) * 0.5;
        const double e0 = (bx0 * t.x + by0 * t.y) - (a.x - b.x) / t.z;
        const double e1 = (bx1 * t.x + by1 * t.y) - (a.y - b.y) / t.z;
        const double e2 = (bx2 * t.x + by2 * t.y) - (a.z - b.z) / t.z;
        const double e3 = (bx3 * t.x + by3 * t.y) - (a.w - b.w) / t.z;
        gfx = mk4(bx0 - e0 * t.x, bx1 - e1 * t.x, bx2 - e2 * t.x, bx3 - e3 * t.x);
        gfy = mk4(by0 - e0 * t.y, by1 - e1 * t.y, by2 - e2 * t.y, by3 - e3 * t.y);
    }
    d4 fl = face_flux<VISC>(kind, qL, qR, gfx, gfy, gA.x, gA.y, g);
    fl.x *= gA.z; fl.y *= gA.z; fl.z *= gA.z; fl.w *= gA.z;
    flux[f] = fl;
}

template <int SECOND, int VISC, int UNIFORM>
__global__ void __launch_bounds__(AFX_FLUX_THREADS, AFX_FLUX_MINB) k_flux(DevMesh m, const d4* qk, const d4* q0,
                                              const d4* gx, const d4* gy,
                                              const d4* lim, d4* flux, GasC g, d4 qfar)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.e_flux) return;
    pdl_launch_dependents();
    const FaceRec rec = face_load_static(m, f);
    pdl_wait();  // states, gradients, limiters come from the previous kernels
    flux_face<SECOND, VISC, UNIFORM, 0>(m, f, rec, qk, q0, gx, gy, lim, flux, g, qfar);
}

// ---------------------------------------------------------------------------
// Block-level deterministic sum of one double per thread -> partial[slot];
// k_norm_finish adds the partials in slot order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void block_norm_partial(double v, const NormOut& no, unsigned int slot)
{
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) no.partial[no.blk_off + slot] = t;
    }
}

// The partial sums of one phase, added in slot order by one block of its own (fixed shape: strided per thread, then a
// tree), the square root stored in the residual-history ring (solver.h:827 / 1178).  It used to be the last block of the
// producing kernel that did this, found with a fence + atomic per block -- which made EVERY block of the last update
// kernel wait ~1.5 us for its atomic to return before it could retire: k_gather_update<0,1> took 0.63 ms against 0.35 ms
// for k_gather_update<0,0> at 16M cells (ncu, profiles/r02a_16M_ncu_full_summary.json).  A 2 us launch is cheaper.
__global__ void __launch_bounds__(1024) k_norm_finish(NormOut no, unsigned int total)
{
    __shared__ double sh[32];
    pdl_launch_dependents();
    pdl_wait();  // the partial sums come from the previous kernel
    double t = 0;
    for (unsigned int b = threadIdx.x; b < total; b += blockDim.x) t += no.partial[b];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) sh[wid] = t;
    __syncthreads();
    if (wid == 0) {
        double u = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) u += __shfl_down_sync(0xffffffffu, u, o);
        if (lane == 0) { const unsigned int k = *no.norm_idx; no.norms[k % NORM_RING] = no.store_square ? u : sqrt(u); *no.norm_idx = k + 1; }
    }
}

// ---------------------------------------------------------------------------
// Owner-computes gather of the face fluxes + stage update, one thread per real
// cell.  MODE 0: explicit stage (solver.h:787-798, 818-821): qW = sum/area,
// qk_out = q + qW*dt*alpha*relax, wall ghosts of qk_out follow their owner
// (solver.h:811 of the next stage); qk_out may alias q on the last stage.
// MODE 1: implicit right-hand side (solver.h:1135-1151): rhs = sum, no
// division, no update, ghost rows zero.  MODE 2: uniform-flow residual
// (solver.h:680-689): like MODE 1 but the ghost rows of two-sided boundary
// faces count in the norm.  LAST adds the squared entries to the residual norm
// and stores the vector (qW or rhs).
// ---------------------------------------------------------------------------
// gather + update of cell i; returns the cell's share of the squared residual norm (LAST).  Shared by k_gather_update and the
// pipelined stage kernel k_pipe (CG: the fluxes were written by other CTAs of the running kernel).
template <int MODE, int LAST, int CG>
__device__ __forceinline__ double gather_cell(const DevMesh& m, uint32_t i, const uint32_t (&bnd)[4], const d4* flux, const d4* q, const d4* qk_in,
                                              d4* qk_out, const double* dt, d4* qW, double alpha, const double* __restrict__ prm, int walls,
                                              const PushArgs& push)
{
    double nrm = 0;
    d4 r = mk4(0, 0, 0, 0);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t cfv = bnd[s];
        if (cfv == CF_NONE) continue;
        const d4 fl = ld4<CG>(flux + (cfv & CF_ID));
        if (cfv & CF_SIDE) { r.x += fl.x; r.y += fl.y; r.z += fl.z; r.w += fl.w; }
        else { r.x -= fl.x; r.y -= fl.y; r.z -= fl.z; r.w -= fl.w; }
        if (LAST && MODE != 1 && (cfv & CF_BND) && m.fkind[cfv & CF_ID] == K_INTERNAL)  // two-sided boundary face: the ghost row of qW holds +flux
            nrm += fl.x * fl.x + fl.y * fl.y + fl.z * fl.z + fl.w * fl.w;
    }
    if (MODE == 0) {
        const double A = m.area[i];
#if AFX_FAST
        const double rA = fast_rcp(A);
        r.x *= rA; r.y *= rA; r.z *= rA; r.w *= rA;
#else
        r.x /= A; r.y /= A; r.z /= A; r.w /= A;
#endif
        const d4 q0 = q[i];
        const double dti = dt[i];
        d4 o;
        const double relax = prm[1];
        o.x = q0.x + r.x * dti * alpha * relax;
        o.y = q0.y + r.y * dti * alpha * relax;
        o.z = q0.z + r.z * dti * alpha * relax;
        o.w = q0.w + r.w * dti * alpha * relax;
        qk_out[i] = o;
        if (push.enabled && i < push.n_front) {  // halo push: straight into the peers' receive buffers over NVLink
            const unsigned long long par = (*push.epoch) & 1ull;
            for (uint32_t k = push.dst_ptr[i]; k < push.dst_ptr[i + 1]; ++k) {
                const uint32_t d = push.dst[k], p = d >> 28, slot = d & 0x0FFFFFFFu;
                push.peer_buf[p][par * push.peer_stride[p] + slot] = o;
            }
            // no fence per thread: the flags are raised after a system-scope fence that follows all of these stores --
            // in k_halo_signal (a later kernel of this stream) or, with the early hand-off, once per front CTA below
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t cfv = bnd[s];
            if (!walls || cfv == CF_NONE || !(cfv & CF_BND)) continue;
            const int kind = m.fkind[cfv & CF_ID];
            // wall ghosts of the next stage follow their owner; after the last stage the ghost keeps the state its
            // owner had when the stage started (solver.h:811 ran before the update)
            if (kind == K_SLIPWALL || kind == K_WALL) qk_out[m.fcells[cfv & CF_ID].y] = LAST ? qk_in[i] : o;
        }
    }
    if (LAST) {
        if (MODE != 0 || prm[2] != 0.0) qW[i] = r;  // prm[2]: keep qW (only the last iteration of a run needs it)
        nrm += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
    }
    return nrm;
}

template <int MODE, int LAST>
__global__ void __launch_bounds__(AFX_GATHER_THREADS) k_gather_update(DevMesh m, const d4* flux,
                                                       const d4* q, const d4* qk_in,
                                                       d4* qk_out, const double* dt,
                                                       d4* qW, double alpha, const double* __restrict__ prm,
                                                       int walls, NormOut no, uint32_t lo, uint32_t hi, PushArgs push)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    double nrm = 0;
    pdl_launch_dependents();
    uint32_t bnd[4] = {CF_NONE, CF_NONE, CF_NONE, CF_NONE};
    if (i < hi) {
#pragma unroll
        for (int s = 0; s < 4; ++s) bnd[s] = m.cf[(size_t)s * m.N + i];
    }
    pdl_wait();  // the fluxes come from the previous kernel
    if (i < hi) nrm = gather_cell<MODE, LAST, 0>(m, i, bnd, flux, q, qk_in, qk_out, dt, qW, alpha, prm, walls, push);
    if (MODE == 0 && push.early_signal && blockIdx.x < push.n_front_blocks) {  // uniform per CTA
        // every store of this CTA into the peers' buffers has been issued: one system-scope fence per CTA (cumulative over
        // the stores the barrier has ordered before it), not one per thread
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(push.front_done, 1u) == push.n_front_blocks - 1) {  // the whole send layer is on its way
                *push.front_done = 0;
                __threadfence_system();
                const unsigned long long e = *push.epoch_rw + 1ull;
                for (int p = 0; p < push.n_peers; ++p) *reinterpret_cast<volatile unsigned long long*>(push.peer_flag[p]) = e;
                __threadfence_system();
                *push.epoch_rw = e;  // read next by k_halo_wait_scatter (a later kernel of this stream)
            }
        }
    }
    if (LAST) block_norm_partial(nrm, no, blockIdx.x);
}

// ---------------------------------------------------------------------------
// Finite-difference flux Jacobian per face (calc_convective_jacobian,
// physics.h:533-578, called from fillRhoLHS solver.h:1021-1058 with first-order
// states and the face gradient).  Writes the four 4x4 blocks times the face
// length: J[f][0]=dF/dqL (row c0,col c0), [1]=dF/dqR (c0,c1), [2]=-dF/dqL
// (c1,c0), [3]=-dF/dqR (c1,c1).
// ---------------------------------------------------------------------------
template <int VISC>
__device__ __forceinline__ d4 jac_flux(int kind, const double* qL, const double* qR, const d4& gfx, const d4& gfy,
                                       double nx, double ny, const GasC& g)
{
    return face_flux<VISC>(kind, mk4(qL[0], qL[1], qL[2], qL[3]), mk4(qR[0], qR[1], qR[2], qR[3]), gfx, gfy, nx, ny, g);
}

template <int VISC>
__global__ void __launch_bounds__(128) k_jacobian(DevMesh m, const d4* __restrict__ q, const d4* __restrict__ gx,
                                                  const d4* __restrict__ gy, d4* __restrict__ J, GasC g)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m.E) return;
    const uint2 fc = m.fcells[f];
    const d4 gA = m.fgA[f];
    const int kind = m.fkind[f];
    const d4 a = q[fc.x], b = q[fc.y];
    d4 gfx = mk4(0, 0, 0, 0), gfy = mk4(0, 0, 0, 0);
    if (VISC == 1 && kind == K_INTERNAL) {
        const d4 t = m.ftij[f];
        const d4 gL0 = gx[fc.x], gL1 = gy[fc.x], gR0 = gx[fc.y], gR1 = gy[fc.y];
        const double bx0 = (gL0.x + gR0.x) * 0.5, by0 = (gL1.x + gR1.x) * 0.5;
        const double bx1 = (gL0.y + gR0.y) * 0.5, by1 = (gL1.y + gR1.y) * 0.5;
        const double bx2 = (gL0.z + gR0.z) * 0.5, by2 = (gL1.z + gR1.z) * 0.5;
        const double bx3 = (gL0.w + gR0.w) * 0.5, by3 = (gL1.w + gR1.w) * 0.5;
        const double e0 = (bx0 * t.x + by0 * t.y) - (a.x - b.x) / t.z;
        const double e1 = (bx1 * t.x + by1 * t.y) - (a.y - b.y) / t.z;
        const double e2 = (bx2 * t.x + by2 * t.y) - (a.z - b.z) / t.z;
        const double e3 = (bx3 * t.x + by3 * t.y) - (a.w - b.w) / t.z;
        gfx = mk4(bx0 - e0 * t.x, bx1 - e1 * t.x, bx2 - e2 * t.x, bx3 - e3 * t.x);
        gfy = mk4(by0 - e0 * t.y, by1 - e1 * t.y, by2 - e2 * t.y, by3 - e3 * t.y);
    }
    double qL[4] = {a.x, a.y, a.z, a.w}, qR[4] = {b.x, b.y, b.z, b.w};
    const d4 f0 = jac_flux<VISC>(kind, qL, qR, gfx, gfy, gA.x, gA.y, g);
    const double len = gA.z;
    double JL[4][4], JR[4][4];  // [row][col]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        {
            const double h = fmax(1e-6, fabs(qL[i]) * 1e-6);
            qL[i] += h;
            const d4 fp = jac_flux<VISC>(kind, qL, qR, gfx, gfy, gA.x, gA.y, g);
            qL[i] -= h;  // restored by subtraction, residue kept (physics.h:556-558)
            JL[0][i] = (fp.x - f0.x) / h; JL[1][i] = (fp.y - f0.y) / h; JL[2][i] = (fp.z - f0.z) / h; JL[3][i] = (fp.w - f0.w) / h;
        }
        {
            const double h = fmax(1e-6, fabs(qR[i]) * 1e-6);
            qR[i] += h;
            const d4 fp = jac_flux<VISC>(kind, qL, qR, gfx, gfy, gA.x, gA.y, g);
            qR[i] -= h;
            JR[0][i] = (fp.x - f0.x) / h; JR[1][i] = (fp.y - f0.y) / h; JR[2][i] = (fp.z - f0.z) / h; JR[3][i] = (fp.w - f0.w) / h;
        }
    }
    d4* Jf = J + (size_t)f * 16;  // 4 blocks x 4 rows
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        Jf[0 + r] = mk4(JL[r][0] * len, JL[r][1] * len, JL[r][2] * len, JL[r][3] * len);
        Jf[4 + r] = mk4(JR[r][0] * len, JR[r][1] * len, JR[r][2] * len, JR[r][3] * len);
        Jf[8 + r] = mk4(-JL[r][0] * len, -JL[r][1] * len, -JL[r][2] * len, -JL[r][3] * len);
        Jf[12 + r] = mk4(-JR[r][0] * len, -JR[r][1] * len, -JR[r][2] * len, -JR[r][3] * len);
    }
}

// Diagonal blocks: area/dt on the diagonal plus the face blocks in edge order
// (solver.h:1012-1018, 1048-1054); ghost rows are the identity (1062-1070).
__global__ void __launch_bounds__(256) k_jac_diag(DevMesh m, const double* __restrict__ J, const double* __restrict__ dt,
                                                  double* __restrict__ D)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.NT) return;
    double d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = 0;
    if (i >= m.n_upd) {  // ghost rows, and the halo rows of a partition (they belong to another rank)
        d[0] = d[5] = d[10] = d[15] = 1;
    } else {
        const double t = m.area[i] / dt[i];
        d[0] = d[5] = d[10] = d[15] = t;
        for (int s = 0; s < 4; ++s) {
            const uint32_t cfv = m.cf[(size_t)s * m.N + i];
            if (cfv == CF_NONE) continue;
            const double* b = J + (size_t)(cfv & CF_ID) * 64 + ((cfv & CF_SIDE) ? 48 : 0);
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k] += b[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) D[(size_t)i * 16 + k] = d[k];
}

// ---------------------------------------------------------------------------
// Wall forces (get_wall_profile, post.h:341-376): one block, boundary edges of
// one patch, warp-shuffle reduction of (fx, fy, -m).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wall_forces(WallArgs a, DevMesh m, const d4* __restrict__ q)
{
    const uint32_t* __restrict__ bface = a.bface; const int32_t* __restrict__ bpatch = a.bpatch;
    const uint32_t G = a.G; const int patch = a.patch;
    const double* __restrict__ bcx = a.bcx; const double* __restrict__ bcy = a.bcy;
    const double gam = a.gam, p_inf = a.p_inf, mach_inf = a.mach_inf, xmin = a.xmin, xmax = a.xmax;
    const double x_moment = a.x_moment, y_moment = a.y_moment;
    double* __restrict__ out3 = a.out3; double* __restrict__ cp_out = a.cp_out;
    __shared__ double sh[3][8];
    double fx = 0, fy = 0, mm = 0;
    for (uint32_t b = threadIdx.x; b < G; b += blockDim.x) {
        if (bpatch[b] != patch) continue;
        const uint32_t f = bface[b];
        const d4 qc = q[m.fcells[f].x];
        const d4 gA = m.fgA[f];
        const double p = (gam - 1) * (qc.w - 0.5 / qc.x * (qc.y * qc.y + qc.z * qc.z));
        const double cp = 2. / (gam * mach_inf * mach_inf) * (p / p_inf - 1.);
        if (cp_out) cp_out[b] = cp;  // halo cells carry their owners' states: every wall edge this rank holds gets its cp
        if (m.fcells[f].x >= m.n_upd) continue;  // ...but the FORCE of a halo cell's wall face belongs to another rank
        const double fxi = cp * gA.x * gA.z / (xmax - xmin);
        const double fyi = cp * gA.y * gA.z / (xmax - xmin);
        const double mi = (bcx[b] - x_moment) / (xmax - xmin) * fyi - (bcy[b] - y_moment) / (xmax - xmin) * fxi;
        fx += fxi; fy += fyi; mm -= mi;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fx += __shfl_down_sync(0xffffffffu, fx, o);
        fy += __shfl_down_sync(0xffffffffu, fy, o);
        mm += __shfl_down_sync(0xffffffffu, mm, o);
    }
    if (lane == 0) { sh[0][wid] = fx; sh[1][wid] = fy; sh[2][wid] = mm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
        out3[0] = a; out3[1] = b; out3[2] = c;
    }
}

// Halo hand-off between GPUs of one node (peer-mapped memory, no collective library on the data path) ---------------
// after the send-layer update: publish "my data of exchange #epoch+1 is in your buffer" to every peer
__global__ void k_halo_signal(SignalArgs a)
{
    __threadfence_system();
    const unsigned long long e = *a.epoch + 1ull;
    if ((int)threadIdx.x < a.n_peers) {
        *reinterpret_cast<volatile unsigned long long*>(a.peer_flag[threadIdx.x]) = e;
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.epoch = e;
}
// wait until every peer has delivered exchange #epoch, then copy the receive buffer of that parity into the halo cells
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// The wait is BOUNDED: a peer that died, threw, or chose another halo mode never raises its flag, and an unbounded spin
// would leave an unkillable kernel on every surviving GPU.  After a.timeout_ns the waiting thread records 1 + the peer's
// slot in *a.err (host-visible), gives up on all peers, and the host turns that into AFX_ERR_COMM after the run.
__global__ void __launch_bounds__(256) k_halo_wait_scatter(WaitArgs a, d4* __restrict__ field)
{
    const unsigned long long e = *a.epoch;
    if (threadIdx.x == 0) {
        bool dead = a.err && *reinterpret_cast<volatile int*>(a.err) != 0;  // an earlier exchange already failed: do not wait again
        unsigned long long t0 = 0;
        for (int p = 0; p < a.n_peers && !dead; ++p) {
            unsigned spins = 0;
            while (*reinterpret_cast<const volatile unsigned long long*>(a.flag[p]) < e) {
                if (++spins < 64u) continue;           // the flag usually arrives within a microsecond or two
                if (t0 == 0) t0 = global_timer_ns();
                __nanosleep(200);
                if (a.timeout_ns && global_timer_ns() - t0 > a.timeout_ns) {
                    if (a.err) *reinterpret_cast<volatile int*>(a.err) = 1 + p;
                    dead = true;
                    break;
                }
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_recv) return;
    const d4* src = a.recv_buf + ((e - 1ull) & 1ull) * a.n_recv + k;
    d4 v;  // written by a peer GPU: read past L1
    v.x = __ldcg(&src->x); v.y = __ldcg(&src->y); v.z = __ldcg(&src->z); v.w = __ldcg(&src->w);
    field[a.recv_idx[k]] = v;
}

// kind bits of the face records from the face kinds set by set_bcs
__global__ void k_face_record_kinds(d4* __restrict__ frec, const uint8_t* __restrict__ fkind, uint32_t E)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= E) return;
    unsigned long long cw = (unsigned long long)__double_as_longlong(frec[2 * (size_t)f + 1].w);
    cw = (cw & ~(3ull << 30)) | ((unsigned long long)(fkind[f] & 3u) << 30);
    frec[2 * (size_t)f + 1].w = __longlong_as_double((long long)cw);
}

// FMG prolongation q_fine = P q_coarse on the device (multigrid.h:100-178 builds P, multigrid.h:308/341 applies it): P has one
// weight per (fine cell, coarse cell) pair, rows in the FINE mesh's reference order, columns in the COARSE mesh's; both
// states live in their solver's internal order.  Each row is summed in ascending column order from zero, which is the order
// of the reference's sparse product, so in the strict build the result carries the reference's bits.
__global__ void __launch_bounds__(256) k_prolongate(uint32_t n_fine, const uint32_t* __restrict__ row_begin, const uint32_t* __restrict__ col,
                                                    const double* __restrict__ w, const uint32_t* __restrict__ fine_new2old,
                                                    const uint32_t* __restrict__ coarse_old2new, const d4* __restrict__ qc, d4* __restrict__ qf)
{
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_fine) return;
    const uint32_t row = fine_new2old[n];
    d4 s = mk4(0, 0, 0, 0);
    for (uint32_t p = row_begin[row]; p < row_begin[row + 1]; ++p) {
        const double wp = w[p];
        const d4 c = qc[coarse_old2new[col[p]]];
        s.x += wp * c.x; s.y += wp * c.y; s.z += wp * c.z; s.w += wp * c.w;
    }
    qf[n] = s;
}

// small utilities -----------------------------------------------------------
__global__ void k_fill_cells(d4* __restrict__ q, uint32_t n, d4 v)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[i] = v;
}
// ghost rows: dst[ghost] = src_state[b] (refill_bcs) or dst[ghost] = dst[owner] (bcs_from_internal)
__global__ void k_ghost_fill(d4* __restrict__ q, const uint32_t* __restrict__ bghost, const uint32_t* __restrict__ bowner,
                             const d4* __restrict__ bstate, uint32_t G, int from_owner)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < G) q[bghost[b]] = from_owner ? q[bowner[b]] : bstate[b];
}
// wall ghosts of cells [lo, N) follow their owners (set_walls_from_internal, solver.h:289-305).  The stage kernels do this
// for the cells they advance; a partitioned run with the fused stage kernel needs it for its ring-1 halo cells too, whose
// limiters that kernel recomputes from the staged ghost states.
__global__ void k_ghost_follow(d4* __restrict__ q, const uint32_t* __restrict__ bghost, const uint32_t* __restrict__ bowner,
                               const uint32_t* __restrict__ bface, const uint8_t* __restrict__ fkind, uint32_t G, uint32_t lo)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= G) return;
    const uint32_t o = bowner[b];
    const int kind = fkind[bface[b]];
    if (o >= lo && (kind == K_SLIPWALL || kind == K_WALL)) q[bghost[b]] = q[o];
}

// permuted copies between reference order (host layout) and internal order
__global__ void k_permute4(const d4* __restrict__ src, d4* __restrict__ dst, const uint32_t* __restrict__ idx, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_scatter4(const d4* __restrict__ src, d4* __restrict__ dst, const uint32_t* __restrict__ idx, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}
__global__ void k_permute1(const double* __restrict__ src, double* __restrict__ dst, const uint32_t* __restrict__ idx, uint32_t n, uint32_t nsrc)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = idx[i];
    dst[i] = j < nsrc ? src[j] : 0.0;
}


// ---------------------------------------------------------------------------
// Launchers (one kernel each) collected in the mode's KernelTable.
// ---------------------------------------------------------------------------
namespace launch {

inline unsigned nblk(size_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// AFX_PDL=0 turns programmatic dependent launch off (plain stream order between the kernels)
inline bool pdl_on()
{
    static const bool on = [] { const char* e = getenv("AFX_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_on() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// lim != nullptr: also write the first stage's limiters (needs want_grad)
static void dt_grad(int grad, const DevMesh& m, d4* q, double* dt, d4* gx, d4* gy, const double* prm, double gam, int want_grad,
                    int walls, d4* lim, double limiter_k, d4* pm, cudaStream_t st)
{
    const unsigned nb = nblk(m.n_grad, AFX_DTG_THREADS);
#define AFX_DTG(G, L) launch_pdl(k_dt_grad<G, L>, nb, AFX_DTG_THREADS, st, m, q, dt, gx, gy, prm, gam, want_grad, walls, lim, limiter_k, pm)
    if (lim && want_grad && AFX_FAST && pm) { if (grad == 0) AFX_DTG(0, 2); else AFX_DTG(1, 2); }
    else if (lim && want_grad) { if (grad == 0) AFX_DTG(0, 1); else AFX_DTG(1, 1); }
    else { if (grad == 0) AFX_DTG(0, 0); else AFX_DTG(1, 0); }
#undef AFX_DTG
}
static void limiter(const DevMesh& m, const d4* qk, const d4* gx, const d4* gy, d4* lim, double k, int walls, uint32_t lo1, uint32_t n1,
                    uint32_t lo2, uint32_t n2, const d4* pm, cudaStream_t st)
{
    if (n1 + n2 == 0) return;
    if (AFX_FAST && pm) launch_pdl(k_limiter<1>, nblk(n1 + n2, AFX_LIM_THREADS), AFX_LIM_THREADS, st, m, qk, gx, gy, lim, k, walls, lo1, n1, lo2, n2, pm);
    else launch_pdl(k_limiter<0>, nblk(n1 + n2, AFX_LIM_THREADS), AFX_LIM_THREADS, st, m, qk, gx, gy, lim, k, walls, lo1, n1, lo2, n2, (const d4*)nullptr);
}
static void limiter_michalak(const DevMesh& m, const d4* qk, const d4* gx, const d4* gy, d4* lim, double k, int walls, uint32_t lo1, uint32_t n1,
                             uint32_t lo2, uint32_t n2, cudaStream_t st)
{
    if (n1 + n2 == 0) return;
    launch_pdl(k_limiter_michalak, nblk(n1 + n2, 128), 128u, st, m, qk, gx, gy, lim, k, walls, lo1, n1, lo2, n2);
}
static void flux(int second, int visc, int uniform, const DevMesh& m, const d4* qk, const d4* q0, const d4* gx, const d4* gy,
                 const d4* lim, d4* fl, const GasC& g, d4 qfar, cudaStream_t st)
{
    const unsigned nb = nblk(m.e_flux, AFX_FLUX_THREADS);
#define AFX_FLUX(S, V, U) launch_pdl(k_flux<S, V, U>, nb, AFX_FLUX_THREADS, st, m, qk, q0, gx, gy, lim, fl, g, qfar)
    if (uniform) { if (visc) AFX_FLUX(0, 1, 1); else AFX_FLUX(0, 0, 1); }
    else if (second) { if (visc) AFX_FLUX(1, 1, 0); else AFX_FLUX(1, 0, 0); }
    else { if (visc) AFX_FLUX(0, 1, 0); else AFX_FLUX(0, 0, 0); }
#undef AFX_FLUX
}
static unsigned gather_blocks(uint32_t n_cells) { return nblk(n_cells, AFX_GATHER_THREADS); }
static void norm_finish(NormOut no, unsigned total, cudaStream_t st) { launch_pdl(k_norm_finish, 1u, 1024u, st, no, total); }
static void gather(int mode, int last, const DevMesh& m, uint32_t lo, uint32_t hi, const d4* fl, const d4* q, const d4* qk_in, d4* qk_out,
                   const double* dt, d4* vec_out, double alpha, const double* prm, int walls, NormOut no, const PushArgs* push_in,
                   cudaStream_t st)
{
    if (hi <= lo) return;
    PushArgs push{};
    if (push_in) push = *push_in;
    const unsigned nb = nblk(hi - lo, AFX_GATHER_THREADS);
    if (no.blk_total == 0) { no.blk_off = 0; no.blk_total = nb; }
#define AFX_G(M, L) launch_pdl(k_gather_update<M, L>, nb, AFX_GATHER_THREADS, st, m, fl, q, qk_in, qk_out, dt, vec_out, alpha, prm, walls, no, lo, hi, push)
    if (mode == 0) { if (last) AFX_G(0, 1); else AFX_G(0, 0); }
    else if (mode == 1) AFX_G(1, 1);
    else AFX_G(2, 1);
#undef AFX_G
}
static void jacobian(int visc, const DevMesh& m, const d4* q, const d4* gx, const d4* gy, d4* J, const GasC& g, cudaStream_t st)
{
    if (visc) k_jacobian<1><<<nblk(m.E, 128), 128, 0, st>>>(m, q, gx, gy, J, g);
    else k_jacobian<0><<<nblk(m.E, 128), 128, 0, st>>>(m, q, gx, gy, J, g);
}
static void jac_diag(const DevMesh& m, const d4* J, const double* dt, double* D, cudaStream_t st)
{
    k_jac_diag<<<nblk(m.NT), 256, 0, st>>>(m, reinterpret_cast<const double*>(J), dt, D);
}
static void wall_forces(const WallArgs& a, const DevMesh& m, const d4* q, cudaStream_t st) { k_wall_forces<<<1, 256, 0, st>>>(a, m, q); }
static void prolongate(uint32_t n_fine, const uint32_t* row_begin, const uint32_t* col, const double* w, const uint32_t* fine_new2old,
                       const uint32_t* coarse_old2new, const d4* qc, d4* qf, cudaStream_t st)
{
    if (n_fine) k_prolongate<<<nblk(n_fine), 256, 0, st>>>(n_fine, row_begin, col, w, fine_new2old, coarse_old2new, qc, qf);
}
static void fill_cells(d4* q, uint32_t n, d4 v, cudaStream_t st) { k_fill_cells<<<nblk(n), 256, 0, st>>>(q, n, v); }
static void ghost_fill(d4* q, const uint32_t* bghost, const uint32_t* bowner, const d4* bstate, uint32_t G, int from_owner, cudaStream_t st)
{
    k_ghost_fill<<<nblk(G), 256, 0, st>>>(q, bghost, bowner, bstate, G, from_owner);
}
static void face_record_kinds(d4* frec, const uint8_t* fkind, uint32_t E, cudaStream_t st)
{
    if (E) k_face_record_kinds<<<nblk(E), 256, 0, st>>>(frec, fkind, E);
}
static void ghost_follow(d4* q, const uint32_t* bghost, const uint32_t* bowner, const uint32_t* bface, const uint8_t* fkind, uint32_t G, uint32_t lo,
                         cudaStream_t st)
{
    if (G) k_ghost_follow<<<nblk(G), 256, 0, st>>>(q, bghost, bowner, bface, fkind, G, lo);
}
static void permute4(const d4* src, d4* dst, const uint32_t* idx, uint32_t n, cudaStream_t st) { k_permute4<<<nblk(n), 256, 0, st>>>(src, dst, idx, n); }
static void halo_signal(const SignalArgs& a, cudaStream_t st) { k_halo_signal<<<1, 32, 0, st>>>(a); }
static void halo_wait_scatter(const WaitArgs& a, d4* field, cudaStream_t st)
{
    k_halo_wait_scatter<<<nblk(a.n_recv ? a.n_recv : 1), 256, 0, st>>>(a, field);
}
static void scatter4(const d4* src, d4* dst, const uint32_t* idx, uint32_t n, cudaStream_t st) { k_scatter4<<<nblk(n), 256, 0, st>>>(src, dst, idx, n); }
static void permute1(const double* src, double* dst, const uint32_t* idx, uint32_t n, uint32_t nsrc, cudaStream_t st)
{
    k_permute1<<<nblk(n), 256, 0, st>>>(src, dst, idx, n, nsrc);
}

}  // namespace launch
}  // namespace AFX_NS
}  // namespace afx
