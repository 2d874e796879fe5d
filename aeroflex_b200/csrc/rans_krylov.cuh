// rans_krylov.cuh -- device pieces of the implicit step: block matrix-vector product on the face-block Jacobian,
// block-Jacobi smoother / preconditioner and the vector kernels of a restarted GMRES.
//
// The matrix of implicitSolver::fillRhoLHS (solver.h:979-1071) is never assembled into CSR: k_jacobian leaves, per face,
// the four 4x4 blocks (512 B, one 128-byte line each) and k_jac_diag the per-cell diagonal block, so
//   (A x)_i = D_i x_i + sum over the faces of i of  (i is cell0 ? J01_f x_c1 : J10_f x_c0),   ghost rows = identity.
// All reductions are two-stage with a fixed grid, hence deterministic.
#pragma once
#include "rans_physics.cuh"

namespace afx {
namespace AFX_NS {

#ifndef AFX_KRY_BLOCKS
#define AFX_KRY_BLOCKS 592  // 4 CTAs per SM on 148 SMs: the reduction grid
#endif
constexpr int KRY_BLOCKS = AFX_KRY_BLOCKS;

__device__ __forceinline__ d4 blk_mul(const d4* __restrict__ B, const d4& x)
{
    const d4 r0 = B[0], r1 = B[1], r2 = B[2], r3 = B[3];
    d4 y;
    y.x = r0.x * x.x + r0.y * x.y + r0.z * x.z + r0.w * x.w;
    y.y = r1.x * x.x + r1.y * x.y + r1.z * x.z + r1.w * x.w;
    y.z = r2.x * x.x + r2.y * x.y + r2.z * x.z + r2.w * x.w;
    y.w = r3.x * x.x + r3.y * x.y + r3.z * x.z + r3.w * x.w;
    return y;
}

__device__ __forceinline__ d4 row_Ax(const DevMesh& m, const d4* __restrict__ J, const d4* __restrict__ D, const d4* __restrict__ x, uint32_t i)
{
    const d4 xi = x[i];
    // ghost rows are the identity (solver.h:1062-1070); so are the halo rows of a partition, which belong to another rank and
    // are refreshed from it before they are read
    if (i >= m.n_upd) return xi;
    d4 y = blk_mul(D + (size_t)i * 4, xi);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t cfv = m.cf[(size_t)s * m.N + i];
        if (cfv == CF_NONE) continue;
        const uint32_t f = cfv & CF_ID;
        const bool side = cfv & CF_SIDE;
        const d4 t = blk_mul(J + (size_t)f * 16 + (side ? 8 : 4), x[m.cnb[(size_t)s * m.N + i] & CF_ID]);
        y.x += t.x; y.y += t.y; y.z += t.z; y.w += t.w;
    }
    return y;
}

// y = A x
__global__ void __launch_bounds__(256) k_spmv(DevMesh m, const d4* __restrict__ J, const d4* __restrict__ D, const d4* __restrict__ x, d4* __restrict__ y)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m.NT) y[i] = row_Ax(m, J, D, x, i);
}

// block-Jacobi sweep: first ? z = Dinv r : z_out = z_in + Dinv (r - A z_in)
// `stop` (may be null): the device-side flag of the Krylov iteration.  Once k_givens_step has seen the linear residual reach the
// tolerance, the kernels of the steps the host has already queued behind it return at once: the host looks at the iteration
// once per batch of steps, not once per step.
__global__ void __launch_bounds__(256) k_jacobi_sweep(DevMesh m, const d4* __restrict__ J, const d4* __restrict__ D, const d4* __restrict__ Dinv,
                                                      const d4* __restrict__ r, const d4* __restrict__ z_in, d4* __restrict__ z_out, int first,
                                                      const int* __restrict__ stop)
{
    if (stop && *stop) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.NT) return;
    const d4 ri = r[i];
    if (i >= m.n_upd) { z_out[i] = ri; return; }
    if (first) { z_out[i] = blk_mul(Dinv + (size_t)i * 4, ri); return; }
    const d4 az = row_Ax(m, J, D, z_in, i);
    const d4 d = blk_mul(Dinv + (size_t)i * 4, mk4(ri.x - az.x, ri.y - az.y, ri.z - az.z, ri.w - az.w));
    const d4 zi = z_in[i];
    z_out[i] = mk4(zi.x + d.x, zi.y + d.y, zi.z + d.z, zi.w + d.w);
}

// Dinv_i = D_i^-1 by Gauss-Jordan with partial pivoting, one thread per cell
__global__ void __launch_bounds__(128) k_invert_blocks(uint32_t n, const double* __restrict__ D, double* __restrict__ Dinv, int* __restrict__ singular)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[r][c] = D[(size_t)i * 16 + r * 4 + c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int piv = p;
        double big = fabs(a[p][p]);
#pragma unroll
        for (int r = p + 1; r < 4; ++r) if (fabs(a[r][p]) > big) { big = fabs(a[r][p]); piv = r; }
        if (big == 0.0) { *singular = 1; big = 1.0; }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r == piv && piv != p) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const double t = a[p][c]; a[p][c] = a[r][c]; a[r][c] = t; }
            }
        const double inv = 1.0 / a[p][p];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[p][c] *= inv;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r != p) {
                const double f = a[r][p];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[r][c] -= f * a[p][c];
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) Dinv[(size_t)i * 16 + r * 4 + c] = a[r][4 + c];
}

__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;  // valid in thread 0
}

// partial[j][block] = sum over the block's cells of V_j . w   (j = 0..k-1; V_j = V + j*stride)
__global__ void __launch_bounds__(256) k_multi_dot_partial(uint32_t n, const d4* __restrict__ V, size_t stride, int k, const d4* __restrict__ w,
                                                           double* __restrict__ partial)
{
    for (int j = 0; j < k; ++j) {
        const d4* __restrict__ vj = V + (size_t)j * stride;
        double acc = 0;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const d4 a = vj[i], b = w[i];
            acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
        const double t = block_sum(acc);
        if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = t;
    }
}
// out[j] = sum_b partial[j][b], one block per j
__global__ void __launch_bounds__(256) k_multi_dot_final(int nblocks, const double* __restrict__ partial, double* __restrict__ out)
{
    double acc = 0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += partial[(size_t)blockIdx.x * nblocks + b];
    const double t = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

// The same final sums without a second launch: the block that finishes last (device-wide counter) adds the partials of
// every j exactly as k_multi_dot_final does -- same strided order per thread, same block_sum tree -- so the results carry
// the same bits.  The Krylov iteration of the shipped meshes is bound by launches, not by bytes.
__device__ __forceinline__ bool dot_finish_by_last_block(int k, const double* partial, double* out, unsigned int* counter)
{
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return false;
    for (int j = 0; j < k; ++j) {
        double acc = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) acc += __ldcg(&partial[(size_t)j * gridDim.x + b]);
        const double t = block_sum(acc);
        if (threadIdx.x == 0) out[j] = t;
    }
    if (threadIdx.x == 0) *counter = 0;
    return true;  // in every thread of the last block; out[] was written by its thread 0
}

// multi_dot in one launch
__global__ void __launch_bounds__(256) k_multi_dot(uint32_t n, const d4* __restrict__ V, size_t stride, int k, const d4* __restrict__ w,
                                                   double* partial, double* out, unsigned int* counter, const int* __restrict__ stop)
{
    if (stop && *stop) return;
    for (int j = 0; j < k; ++j) {
        const d4* __restrict__ vj = V + (size_t)j * stride;
        double acc = 0;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const d4 a = vj[i], b = w[i];
            acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
        const double t = block_sum(acc);
        if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = t;
    }
    dot_finish_by_last_block(k, partial, out, counter);
}

// w -= sum_j c[j] V_j followed by ||w||^2 -> out[0], one launch (classical Gram-Schmidt step + the norm of the result):
// the update is k_multi_axpy's arithmetic per cell, the norm k_multi_dot's (same grid-stride partition, same trees)
__global__ void __launch_bounds__(256) k_axpy_norm(uint32_t n, const d4* __restrict__ V, size_t stride, int k, const double* __restrict__ c,
                                                   double sign, d4* w, double* partial, double* out, unsigned int* counter, const int* __restrict__ stop)
{
    if (stop && *stop) return;
    double acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        d4 a = w[i];
        for (int j = 0; j < k; ++j) {
            const double cj = sign * c[j];
            const d4 v = V[(size_t)j * stride + i];
            a.x += cj * v.x; a.y += cj * v.y; a.z += cj * v.z; a.w += cj * v.w;
        }
        w[i] = a;
        acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    dot_finish_by_last_block(1, partial, out, counter);
}

// r = A x and z = Dinv r (the first block-Jacobi sweep from a zero start) in one launch: k_spmv + k_jacobi_sweep(first)
__global__ void __launch_bounds__(256) k_spmv_sweep0(DevMesh m, const d4* __restrict__ J, const d4* __restrict__ D, const d4* __restrict__ Dinv,
                                                     const d4* __restrict__ x, d4* __restrict__ r, d4* __restrict__ z, const int* __restrict__ stop)
{
    if (stop && *stop) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.NT) return;
    const d4 ri = row_Ax(m, J, D, x, i);
    r[i] = ri;
    z[i] = (i >= m.n_upd) ? ri : blk_mul(Dinv + (size_t)i * 4, ri);
}
// w += sign * sum_j c[j] V_j
__global__ void __launch_bounds__(256) k_multi_axpy(uint32_t n, const d4* __restrict__ V, size_t stride, int k, const double* __restrict__ c,
                                                    double sign, d4* __restrict__ w)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d4 a = w[i];
    for (int j = 0; j < k; ++j) {
        const double cj = sign * c[j];
        const d4 v = V[(size_t)j * stride + i];
        a.x += cj * v.x; a.y += cj * v.y; a.z += cj * v.z; a.w += cj * v.w;
    }
    w[i] = a;
}
// y = x * (inv ? 1/s[0] : s[0]) with the scalar taken from device memory (sqrt applied first if root)
__global__ void __launch_bounds__(256) k_scale_from(uint32_t n, const d4* __restrict__ x, const double* __restrict__ s, int root, int inv, d4* __restrict__ y,
                                                    const int* __restrict__ stop)
{
    if (stop && *stop) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f = s[0];
    if (root) f = sqrt(f);
    if (inv) f = 1.0 / f;
    const d4 a = x[i];
    y[i] = mk4(a.x * f, a.y * f, a.z * f, a.w * f);
}
// y = a - b
__global__ void __launch_bounds__(256) k_sub(uint32_t n, const d4* a, const d4* b, d4* y)  // y may alias a or b
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d4 p = a[i], q = b[i];
    y[i] = mk4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
}
// q += relax * x   (solver.h:1187,1201)
__global__ void __launch_bounds__(256) k_axpy_state(uint32_t n, double relax, const d4* __restrict__ x, d4* __restrict__ q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const d4 a = x[i];
    d4 b = q[i];
    b.x += a.x * relax; b.y += a.y * relax; b.z += a.z * relax; b.w += a.w * relax;
    q[i] = b;
}

// ---- the small dense part of GMRES on the device --------------------------------------------------------------------
// state (doubles): H[(m+1) x m] row-major | cs[m] | sn[m] | g[m+1] | y[m] | r0, err, k_done, stop, fail   (GmresLayout)
struct GmresLayout {
    int m;
    __host__ __device__ int H() const { return 0; }
    __host__ __device__ int cs() const { return (m + 1) * m; }
    __host__ __device__ int sn() const { return cs() + m; }
    __host__ __device__ int g() const { return sn() + m; }
    __host__ __device__ int y() const { return g() + m + 1; }
    __host__ __device__ int r0() const { return y() + m; }
    __host__ __device__ int err() const { return r0() + 1; }
    __host__ __device__ int k_done() const { return r0() + 2; }
    __host__ __device__ int fail() const { return r0() + 3; }
    __host__ __device__ int total() const { return r0() + 4; }
};
// start of a cycle: g = (beta, 0, ...), k_done = 0; *stop = 0.  beta2 holds beta^2 (device); first: it is also r0^2.
__global__ void k_gmres_begin(GmresLayout L, double* st, const double* __restrict__ beta2, int first, int* stop)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double beta = sqrt(beta2[0]);
    for (int i = 0; i <= L.m; ++i) st[L.g() + i] = 0.0;
    st[L.g()] = beta;
    if (first) { st[L.r0()] = beta; st[L.fail()] = (beta == beta) ? 0.0 : 1.0; }
    st[L.k_done()] = 0.0;
    st[L.err()] = (st[L.r0()] > 0) ? beta / st[L.r0()] : 0.0;
    *stop = (!(beta == beta) || beta == 0.0) ? 1 : 0;
}
// Arnoldi step k is done (h[0..k] = V^T w, h[k+1] = ||w||^2 after orthogonalisation): new Hessenberg column, the earlier
// rotations, the new rotation, the residual estimate |g[k+1]| / r0 -- the recurrence Solver::gmres ran on the host.
__device__ __forceinline__ void givens_step_body(const GmresLayout& L, double* st, const double* h, int k, double tol, int* stop)
{
    const int m = L.m;
    double* H = st + L.H(); double* cs = st + L.cs(); double* sn = st + L.sn(); double* g = st + L.g();
    const double hn = sqrt(h[k + 1]);
    for (int i = 0; i <= k; ++i) H[i * m + k] = h[i];
    for (int i = 0; i < k; ++i) {
        const double a = H[i * m + k], c = H[(i + 1) * m + k];
        H[i * m + k] = cs[i] * a + sn[i] * c;
        H[(i + 1) * m + k] = -sn[i] * a + cs[i] * c;
    }
    const double a = H[k * m + k], den = sqrt(a * a + hn * hn);
    if (!(den == den)) { st[L.fail()] = 1.0; *stop = 1; return; }
    cs[k] = den == 0 ? 1 : a / den; sn[k] = den == 0 ? 0 : hn / den;
    H[k * m + k] = den;
    g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k];
    const double err = fabs(g[k + 1]) / st[L.r0()];
    st[L.err()] = err;
    st[L.k_done()] = (double)(k + 1);
    if (err < tol || hn == 0) *stop = 1;
}
__global__ void k_givens_step(GmresLayout L, double* st, const double* __restrict__ h, int k, double tol, int* stop)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || *stop) return;
    givens_step_body(L, st, h, k, tol, stop);
}
// k_axpy_norm (Gram-Schmidt update + ||w||^2 -> h[k]) followed, in the block that finished the norm, by the rotation of step k - 1:
// one launch instead of two.  h = [V^T w (k entries) | ||w||^2]; same arithmetic in the same order as the two kernels.
__global__ void __launch_bounds__(256) k_axpy_norm_givens(uint32_t n, const d4* __restrict__ V, size_t stride, int k, const double* c, double sign, d4* w,
                                                          double* partial, double* h, unsigned int* counter, GmresLayout L, double* st, double tol, int* stop)
{
    if (*stop) return;
    double acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        d4 a = w[i];
        for (int j = 0; j < k; ++j) {
            const double cj = sign * c[j];
            const d4 v = V[(size_t)j * stride + i];
            a.x += cj * v.x; a.y += cj * v.y; a.z += cj * v.z; a.w += cj * v.w;
        }
        w[i] = a;
        acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    const double t = block_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
    if (dot_finish_by_last_block(1, partial, h + k, counter) && threadIdx.x == 0) givens_step_body(L, st, h, k - 1, tol, stop);
}
// H y = g for the k columns done (back substitution), y -> out[0..k)
__global__ void k_gmres_solve_y(GmresLayout L, const double* __restrict__ st, double* __restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = L.m, k = (int)st[L.k_done()];
    const double* H = st + L.H(); const double* g = st + L.g();
    for (int i = k - 1; i >= 0; --i) {
        double s = g[i];
        for (int j = i + 1; j < k; ++j) s -= H[i * m + j] * out[j];
        out[i] = s / H[i * m + i];
    }
}

namespace launch {

static void gmres_begin(int m, double* st, const double* beta2, int first, int* stop, cudaStream_t s_) { k_gmres_begin<<<1, 32, 0, s_>>>(GmresLayout{m}, st, beta2, first, stop); }
static void givens_step(int m, double* st, const double* h, int k, double tol, int* stop, cudaStream_t s_) { k_givens_step<<<1, 32, 0, s_>>>(GmresLayout{m}, st, h, k, tol, stop); }
static void gmres_solve_y(int m, const double* st, double* out, cudaStream_t s_) { k_gmres_solve_y<<<1, 32, 0, s_>>>(GmresLayout{m}, st, out); }
static int gmres_state_doubles(int m) { return GmresLayout{m}.total(); }

static void spmv(const DevMesh& m, const d4* J, const double* D, const d4* x, d4* y, cudaStream_t st)
{
    k_spmv<<<(m.NT + 255) / 256, 256, 0, st>>>(m, J, reinterpret_cast<const d4*>(D), x, y);
}
static void jacobi_sweep(const DevMesh& m, const d4* J, const double* D, const double* Dinv, const d4* r, const d4* z_in, d4* z_out, int first,
                         const int* stop, cudaStream_t st)
{
    k_jacobi_sweep<<<(m.NT + 255) / 256, 256, 0, st>>>(m, J, reinterpret_cast<const d4*>(D), reinterpret_cast<const d4*>(Dinv), r, z_in, z_out, first, stop);
}
static void invert_blocks(uint32_t n, const double* D, double* Dinv, int* singular, cudaStream_t st)
{
    k_invert_blocks<<<(n + 127) / 128, 128, 0, st>>>(n, D, Dinv, singular);
}
static void multi_dot(uint32_t n, const d4* V, size_t stride, int k, const d4* w, double* partial, double* out, cudaStream_t st)
{
    k_multi_dot_partial<<<KRY_BLOCKS, 256, 0, st>>>(n, V, stride, k, w, partial);
    k_multi_dot_final<<<k, 256, 0, st>>>(KRY_BLOCKS, partial, out);
}
// one-launch forms (counter: a zeroed device word, reset by the kernel)
static void multi_dot1(uint32_t n, const d4* V, size_t stride, int k, const d4* w, double* partial, double* out, unsigned int* counter, const int* stop,
                       cudaStream_t st)
{
    k_multi_dot<<<KRY_BLOCKS, 256, 0, st>>>(n, V, stride, k, w, partial, out, counter, stop);
}
static void axpy_norm(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, double* partial, double* out,
                      unsigned int* counter, const int* stop, cudaStream_t st)
{
    k_axpy_norm<<<KRY_BLOCKS, 256, 0, st>>>(n, V, stride, k, c, sign, w, partial, out, counter, stop);
}
static void axpy_norm_givens(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, double* partial, double* h,
                             unsigned int* counter, int m, double* state, double tol, int* stop, cudaStream_t st)
{
    k_axpy_norm_givens<<<KRY_BLOCKS, 256, 0, st>>>(n, V, stride, k, c, sign, w, partial, h, counter, GmresLayout{m}, state, tol, stop);
}
static void spmv_sweep0(const DevMesh& m, const d4* J, const double* D, const double* Dinv, const d4* x, d4* r, d4* z, const int* stop, cudaStream_t st)
{
    k_spmv_sweep0<<<(m.NT + 255) / 256, 256, 0, st>>>(m, J, reinterpret_cast<const d4*>(D), reinterpret_cast<const d4*>(Dinv), x, r, z, stop);
}
static void multi_axpy(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, cudaStream_t st)
{
    k_multi_axpy<<<(n + 255) / 256, 256, 0, st>>>(n, V, stride, k, c, sign, w);
}
static void scale_from(uint32_t n, const d4* x, const double* s, int root, int inv, d4* y, const int* stop, cudaStream_t st)
{
    k_scale_from<<<(n + 255) / 256, 256, 0, st>>>(n, x, s, root, inv, y, stop);
}
static void sub(uint32_t n, const d4* a, const d4* b, d4* y, cudaStream_t st) { k_sub<<<(n + 255) / 256, 256, 0, st>>>(n, a, b, y); }
static void axpy_state(uint32_t n, double relax, const d4* x, d4* q, cudaStream_t st) { k_axpy_state<<<(n + 255) / 256, 256, 0, st>>>(n, relax, x, q); }

}  // namespace launch
}  // namespace AFX_NS
}  // namespace afx
