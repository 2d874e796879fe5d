// rans_types.h -- types shared by the host solver and the two kernel
// translation units (strict / fast arithmetic, see rans_kernels_tu.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace afx {

// 32-byte aligned 4-vector: one LDG.E.256 / STG.E.256 per cell state on sm_100a
struct __align__(32) d4 {
    double x, y, z, w;
};

struct GasC {
    double gamma, R, mu_L, Pr_L, cp;
};

enum : int { K_INTERNAL = 0, K_FARFIELD = 1, K_SLIPWALL = 2, K_WALL = 3 };

constexpr uint32_t CF_NONE = 0xFFFFFFFFu;
constexpr uint32_t CF_SIDE = 0x80000000u;  // this cell is cell1 of the face
constexpr uint32_t CF_BND = 0x40000000u;   // boundary face (cell1 is a ghost)
constexpr uint32_t CF_ID = 0x3FFFFFFFu;
constexpr unsigned int NORM_RING = 1u << 16;  // capacity of the device-side residual history ring

struct DevMesh {
    uint32_t N, G, E, NT;        // real cells, ghosts, faces, N+G
    uint32_t n_upd, n_grad;      // cells [0,n_upd) are advanced; cells [0,n_grad) get dt/gradients/limiters (halo ring 1 included)
    uint32_t e_flux;             // faces [0,e_flux) touch an advanced cell and get a flux
    const uint2* fcells;         // [E]
    const d4* fgA;               // [E] nx, ny, len, w
    const d4* fgB;               // [E] d0x, d0y, d1x, d1y
    const d4* ftij;              // [E] t0, t1, l, - (laminar face-gradient direction, solver.h:369-376)
    const uint8_t* fkind;        // [E]
    const d4* frec;              // [E][2] the face kernel's record: {nx, ny, len, d0x}, {d0y, d1x, d1y, bits(cell0 | kind << 30, cell1)}
    const uint32_t* cf;          // [4][N]
    const uint32_t* cnb;         // [4][N] the cell across slot s (same slots as cf; CF_NONE where empty) | CF_SIDE | CF_BND like cf
    const d4* cgeo;              // [4][N] {nx, ny, len, w_gg} of the face in slot s: the dt/gradient kernel reads no face record
    const double2* cdxy;         // [4][N] face centre minus this cell's centre for slot s (the fgB half this cell needs)
    const double* area;          // [NT]
    const double* lsqM;          // [8][N]  (M * dT) rows in cellsEdges order, LSQ only
    const uint16_t* lsq_perm;    // [N] bits 0-7: slot of local side j (2 bits each), bits 8-10: number of sides
};

// Device view of the shared-memory tiles (tiling.h): tile t owns cells [cell0, cell0+nc) and stages, with LOCAL 16-bit
// indices, its ring-1 cells (limiter recomputed), the state-only cells around them and every face with an end in the
// tile.  Static geometry is packed per tile so that everything but the ring cells' states arrives by bulk copies.
struct TileTab {
    const uint4* head;       // [n_tiles][2]: {cell0, nc, h1, h2}, {nf, off_halo, off_cell, off_face}
    const uint32_t* halo;    // [off_halo + k]: global ids of the local cells nc.. (ring 1, then state-only cells); padded to 4
    const uint4* ctab;       // [off_cell + l], l < nc+h1: 4 x u16 local neighbour, 4 x u16 local face | side bit (0xFFFF none)
    const double* k3a_t;     // [off_cell + l]: (limiter_k sqrt(area))^3 of the cell, tile-local order (off_cell and the per-tile count are even)
    const double2* dxy_t;    // [4*off_cell + s*n1p + l]: face centre minus cell centre of slot s (n1p = nc+h1 rounded up to even)
    const d4* fgeo_t;        // [off_face + lf]: {nx, ny, len, kind} of the local face
    uint32_t n_tiles;
    uint32_t max_loc, max_n1, max_nf, max_nc, max_halo;  // shared-memory carve-up (stage_smem_layout)
};

// Shared-memory carve-up of k_stage, byte offsets; the same function sizes the launch on the host.
struct StageSmem {
    uint32_t sq, sgx, sgy, sdxy, sarea, sctab;   // group 1: inputs of the limiter / reconstruction phase
    uint32_t sfg;                                // group 2: face geometry
    uint32_t srec;                               // face states, then fluxes
    uint32_t sq0, sdt, sarea3, sctab3;           // group 3: inputs of the gather / update phase
    uint32_t shalo;                              // [2][max_halo] global ids of the ring cells, double buffered
    uint32_t bars;                               // 3 mbarriers
    uint32_t total;
};
__host__ __device__ inline uint32_t stage_smem_take(uint32_t& o, uint32_t bytes) { const uint32_t at = o; o += (bytes + 31u) & ~31u; return at; }
__host__ __device__ inline StageSmem stage_smem_layout(uint32_t max_loc, uint32_t max_n1, uint32_t max_nf, uint32_t max_nc, uint32_t max_halo)
{
    StageSmem L;
    uint32_t o = 0;
    L.sq = stage_smem_take(o, 32u * max_loc); L.sgx = stage_smem_take(o, 32u * max_n1); L.sgy = stage_smem_take(o, 32u * max_n1);
    L.sdxy = stage_smem_take(o, 64u * max_n1); L.sarea = stage_smem_take(o, 8u * max_n1); L.sctab = stage_smem_take(o, 16u * max_n1);
    L.sfg = stage_smem_take(o, 32u * max_nf);
    L.srec = stage_smem_take(o, 64u * max_nf);
    L.sq0 = stage_smem_take(o, 32u * max_nc); L.sdt = stage_smem_take(o, 8u * max_nc); L.sarea3 = stage_smem_take(o, 8u * max_nc);
    L.sctab3 = stage_smem_take(o, 16u * max_nc);
    L.shalo = stage_smem_take(o, 8u * max_halo);
    L.bars = stage_smem_take(o, 32u);
    L.total = o;
    return L;
}

// Work list of the pipelined stage kernel k_pipe (rans_pipe.cuh).  The cells [0, n_grad) are cut into chunks of 2^shift
// consecutive cells (compact patches of the mesh along the space-filling curve); ONE persistent kernel per Runge-Kutta stage
// walks the chunks with three phases a fixed number of steps apart -- limiter of chunk k, face fluxes of chunk k - lagF,
// gather + update of chunk k - lagF - lagU -- so that what one phase writes (limiters, fluxes) and what two phases read
// (states, gradients) is still in the 126 MB L2 when the next phase asks for it.  Items (one CTA-load of cells or faces) are
// handed out in step order from a device counter; an item waits on per-chunk completion counters of the items it reads from,
// all of which were handed out before it.  Faces joining two chunks more than one apart, and the cells that touch them
// ("far", a few per cent), are done after the sweep from two index lists.
struct PipeTab {
    uint32_t shift, n_chunks;
    uint32_t lagF, lagU;
    uint32_t n_items;                    // sweep items, then far-face items, then far-cell items
    uint32_t n_far_faces, n_far_cells;
    uint32_t nL_total, nF_total;         // limiter items; flux items (sweep + far)
    uint32_t nU_items;                   // update items (sweep + far) = norm slots
    const uint4* items;                  // [n_items] {phase | chunk << 3, first element (or offset into a far list), elements, norm slot}
    const uint2* chunk_items;            // [n_chunks] {limiter items, flux items} of chunk c: what a dependent item waits for
    const uint32_t* far_faces;           // [n_far_faces]
    const uint32_t* far_cells;           // [n_far_cells]
    const uint32_t* far_mask;            // bit i: advanced cell i touches a far face
    unsigned int* ctr;                   // [0] next item, [1] CTAs that left, [2] limiter items done, [3] flux items done,
                                         // [4 + c] limiter items of chunk c done, [4 + n_chunks + c] flux items of chunk c done
    int* err;                            // host-visible: a dependency wait ran into its bound (internal error, never expected)
};

struct NormOut {
    double* partial;        // [gridDim.x]
    unsigned int* counter;  // block counter
    double* norms;          // [NORM_RING]
    unsigned int* norm_idx; // running index into norms
    int store_square;       // 1: store the sum of squares (partitioned runs add the ranks' sums before the root)
    unsigned int blk_off;   // this launch's first slot in partial[] (a phase may be split into several launches)
    unsigned int blk_total; // blocks (work items) of all launches of the phase: k_norm_finish adds that many partial sums
};

// Halo push over NVLink peer memory, fused into the update kernel: an advanced cell that a peer needs is stored straight
// into that peer's receive buffer (double-buffered by stage parity) from the kernel that computes it.
constexpr int P2P_MAX_PEERS = 8;
struct PushArgs {
    int enabled;                                  // 0: no push (single GPU, NCCL halo, or cells outside the send layer)
    uint32_t n_front;                             // cells [0, n_front) have destinations
    const uint32_t* dst_ptr;                      // [n_front+1] CSR over the send-layer cells
    const uint32_t* dst;                          // (peer slot << 28) | position in that peer's receive list
    d4* peer_buf[P2P_MAX_PEERS];                  // peer's receive buffer at my offset (parity 0), a peer-mapped pointer
    unsigned long long peer_stride[P2P_MAX_PEERS];// elements between the two parity buffers of that peer
    const unsigned long long* epoch;              // completed exchanges of this rank (device counter)
    // early hand-off (AFX_HALO_EARLY_SIGNAL=1): the send layer lives in the first n_front_blocks CTAs of the update kernel; the
    // one that finishes last raises the peers' flags itself, while the rest of the kernel still advances interior cells
    int early_signal;
    unsigned int n_front_blocks;
    unsigned int* front_done;                     // device counter, zero between launches
    unsigned long long* epoch_rw;                 // the same counter as `epoch`, advanced by the signalling CTA
    int n_peers;
    unsigned long long* peer_flag[P2P_MAX_PEERS]; // peer's flag slot for my rank (peer-mapped)
};
struct SignalArgs {
    int n_peers;
    unsigned long long* peer_flag[P2P_MAX_PEERS]; // peer's flag slot for my rank (peer-mapped)
    unsigned long long* epoch;
};
struct WaitArgs {
    int n_peers;
    const unsigned long long* flag[P2P_MAX_PEERS]; // my flag slots, one per peer I receive from
    const unsigned long long* epoch;
    const d4* recv_buf;                           // [2][n_recv]
    const uint32_t* recv_idx;                     // [n_recv] halo cells to fill
    uint32_t n_recv;
    unsigned long long timeout_ns;                // give up on a peer after this long (0: wait for ever)
    int* err;                                     // host-visible word: 1 + slot of the peer that never delivered
};

struct WallArgs {
    const uint32_t* bface; const int32_t* bpatch; uint32_t G; int patch;
    const double* bcx; const double* bcy;
    double gam, p_inf, mach_inf, xmin, xmax, x_moment, y_moment;
    double* out3; double* cp_out;
};

// Launchers of one arithmetic mode.  Every function enqueues exactly one kernel on `st`.
struct KernelTable {
    const char* name;
    // lim != nullptr (and want_grad): the kernel also writes the limiters of the first stage (state q), saving that k_limiter launch
    void (*dt_grad)(int grad_scheme, const DevMesh& m, d4* q, double* dt, d4* gx, d4* gy, const double* prm, double gam,
                    int want_grad, int walls, d4* lim, double limiter_k, d4* pm, cudaStream_t st);  // pm (fast mode, with lim): [2][cell] projected extremes
    // cells [lo1, lo1+n1) and [lo2, lo2+n2)
    void (*limiter)(const DevMesh& m, const d4* qk, const d4* gx, const d4* gy, d4* lim, double limiter_k, int walls, uint32_t lo1, uint32_t n1,
                    uint32_t lo2, uint32_t n2, const d4* pm, cudaStream_t st);  // pm != null (fast mode): k_dt_grad's stored extremes
    void (*flux)(int second, int visc, int uniform, const DevMesh& m, const d4* qk, const d4* q0, const d4* gx, const d4* gy,
                 const d4* lim, d4* flux, const GasC& g, d4 qfar, cudaStream_t st);
    // cells [lo, hi); `no` carries the block bookkeeping when the phase is split into several launches
    void (*gather)(int mode, int last, const DevMesh& m, uint32_t lo, uint32_t hi, const d4* flux, const d4* q, const d4* qk_in,
                   d4* qk_out, const double* dt, d4* vec_out, double alpha, const double* prm, int walls, NormOut no,
                   const PushArgs* push, cudaStream_t st);
    // fused stage on shared-memory tiles: limiter + MUSCL + flux + gather + update in one persistent kernel of `grid` CTAs
    void (*stage)(int last, const DevMesh& m, const TileTab& tt, unsigned grid, size_t smem, const d4* qk_in, const d4* q0,
                  d4* qk_out, const d4* gx, const d4* gy, const double* dt, d4* qW, d4* lim, double alpha, const double* prm,
                  const GasC& g, NormOut no, const PushArgs* push, cudaStream_t st);
    // pipelined stage on L2-resident chunks: (limiter +) flux + gather + update in one persistent kernel
    void (*pipe)(int second, int visc, int last, int has_l, const DevMesh& m, const PipeTab& pt, unsigned grid, const d4* qk_in, const d4* q0,
                 d4* qk_out, const d4* gx, const d4* gy, d4* lim, d4* flux, const double* dt, d4* qW, double alpha, const double* prm,
                 const GasC& g, double limiter_k, int walls, NormOut no, const PushArgs* push, cudaStream_t st);
    int (*pipe_item_elems)();  // cells or faces per work item
    int (*pipe_ctas_per_sm)();
    int (*stage_prepare)(size_t smem);  // opt in to the dynamic shared memory on the current device; resident CTAs per SM or <0
    int (*stage_threads)();
    void (*tile_k3a)(const double* area_t, double* k3a_t, size_t n, double limiter_k, cudaStream_t st);  // refresh after set_options
    void (*face_record_kinds)(d4* frec, const uint8_t* fkind, uint32_t E, cudaStream_t st);  // refresh after set_bcs
    void (*tile_face_kinds)(d4* fgeo_t, const uint32_t* tile_face, const uint8_t* fkind, size_t n, cudaStream_t st);  // refresh after set_bcs
    void (*halo_signal)(const SignalArgs& a, cudaStream_t st);               // tell the peers my send layer is in their buffers
    void (*halo_wait_scatter)(const WaitArgs& a, d4* field, cudaStream_t st); // wait for the peers, then fill my halo cells
    unsigned (*gather_blocks)(uint32_t n_cells);
    void (*norm_finish)(NormOut no, unsigned total, cudaStream_t st);  // sum the partials of a phase -> residual-history ring
    void (*jacobian)(int visc, const DevMesh& m, const d4* q, const d4* gx, const d4* gy, d4* J, const GasC& g, cudaStream_t st);
    void (*jac_diag)(const DevMesh& m, const d4* J, const double* dt, double* D, cudaStream_t st);
    void (*wall_forces)(const WallArgs& a, const DevMesh& m, const d4* q, cudaStream_t st);
    void (*prolongate)(uint32_t n_fine, const uint32_t* row_begin, const uint32_t* col, const double* w, const uint32_t* fine_new2old,
                       const uint32_t* coarse_old2new, const d4* qc, d4* qf, cudaStream_t st);
    void (*fill_cells)(d4* q, uint32_t n, d4 v, cudaStream_t st);
    void (*ghost_fill)(d4* q, const uint32_t* bghost, const uint32_t* bowner, const d4* bstate, uint32_t G, int from_owner, cudaStream_t st);
    void (*ghost_follow)(d4* q, const uint32_t* bghost, const uint32_t* bowner, const uint32_t* bface, const uint8_t* fkind, uint32_t G, uint32_t lo,
                         cudaStream_t st);  // wall ghosts of cells >= lo take their owner's state
    void (*permute4)(const d4* src, d4* dst, const uint32_t* idx, uint32_t n, cudaStream_t st);
    void (*permute1)(const double* src, double* dst, const uint32_t* idx, uint32_t n, uint32_t nsrc, cudaStream_t st);
    void (*scatter4)(const d4* src, d4* dst, const uint32_t* idx, uint32_t n, cudaStream_t st);  // dst[idx[i]] = src[i]
    // implicit step (rans_krylov.cuh)
    void (*spmv)(const DevMesh& m, const d4* J, const double* D, const d4* x, d4* y, cudaStream_t st);
    void (*jacobi_sweep)(const DevMesh& m, const d4* J, const double* D, const double* Dinv, const d4* r, const d4* z_in, d4* z_out, int first,
                         const int* stop, cudaStream_t st);
    void (*invert_blocks)(uint32_t n, const double* D, double* Dinv, int* singular, cudaStream_t st);
    void (*multi_dot)(uint32_t n, const d4* V, size_t stride, int k, const d4* w, double* partial, double* out, cudaStream_t st);  // 2 kernels
    void (*multi_axpy)(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, cudaStream_t st);
    // one-launch forms for the Arnoldi step (same arithmetic, same bits; the last block finishes the sums)
    // `stop`: device flag of the Krylov iteration (null = always run), see rans_krylov.cuh
    void (*multi_dot1)(uint32_t n, const d4* V, size_t stride, int k, const d4* w, double* partial, double* out, unsigned int* counter, const int* stop,
                       cudaStream_t st);
    void (*axpy_norm)(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, double* partial, double* out,
                      unsigned int* counter, const int* stop, cudaStream_t st);
    void (*spmv_sweep0)(const DevMesh& m, const d4* J, const double* D, const double* Dinv, const d4* x, d4* r, d4* z, const int* stop, cudaStream_t st);
    void (*scale_from)(uint32_t n, const d4* x, const double* s, int root, int inv, d4* y, const int* stop, cudaStream_t st);
    // the small dense part of GMRES on the device: Hessenberg column + Givens rotations + residual estimate, back substitution
    void (*gmres_begin)(int m, double* state, const double* beta2, int first, int* stop, cudaStream_t st);
    void (*givens_step)(int m, double* state, const double* h, int k, double tol, int* stop, cudaStream_t st);
    void (*gmres_solve_y)(int m, const double* state, double* out, cudaStream_t st);
    int (*gmres_state_doubles)(int m);
    void (*sub)(uint32_t n, const d4* a, const d4* b, d4* y, cudaStream_t st);
    void (*axpy_state)(uint32_t n, double relax, const d4* x, d4* q, cudaStream_t st);
    // axpy_norm + givens_step in one launch: the block that finishes the norm also runs the rotation (single-rank solvers: a
    // partitioned step all-reduces the norm between the two)
    void (*axpy_norm_givens)(uint32_t n, const d4* V, size_t stride, int k, const double* c, double sign, d4* w, double* partial, double* h,
                             unsigned int* counter, int m, double* state, double tol, int* stop, cudaStream_t st);
    // calc_limiters built with RANS_MICHALAK_LIMITER (solver.h:557-576): same cell ranges as `limiter`
    void (*limiter_michalak)(const DevMesh& m, const d4* qk, const d4* gx, const d4* gy, d4* lim, double k, int walls, uint32_t lo1, uint32_t n1,
                             uint32_t lo2, uint32_t n2, cudaStream_t st);
};

namespace strict { const KernelTable& table(); }  // -fmad=false, reference expression order: bit-identical to the CPU reference
namespace fast { const KernelTable& table(); }    // shared reciprocals + FMA contraction: ~1e-15 relative per face

}  // namespace afx
