// rans_solver.cu -- device state, launch sequences and the afx_rans_* C ABI.
//
// One afx_rans handle == one rans::solver of the reference on one B200.  All
// state lives in HBM in a renumbered layout (Hilbert order of the cell centres,
// faces sorted by their lower cell); the ABI speaks the reference's order.
// The explicit iteration (explicitSolver::solve, solver.h:802-828) is 10
// kernels -- dt+gradients once (the gradients of all three stages are those of
// the iteration-start state), then limiter / face flux / gather+update
// per stage -- captured once into a CUDA graph and replayed per iteration; the
// residual norm is reduced on the device into a ring that the host reads once
// per call.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; the ranges cost nothing unless a profiler injects itself

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/afx_rans.h"
#include "nccl_dl.h"
#include "ordering.h"
#include "partition.h"
#include "rans_types.h"
#include "tiling.h"

namespace afx {

thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct InvalidArg : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct NumericError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct CommError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// NVTX range over a scope (tracing): visible in nsys / ncu --nvtx timelines as "afx:<name>"
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
};

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            throw ::afx::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                            std::to_string(__LINE__) + ")");                                           \
    } while (0)

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    void alloc(size_t n_)
    {
        free();
        n = n_;
        if (n) CK(cudaMalloc(&p, n * sizeof(T)));
    }
    void zero(cudaStream_t st) { if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
    void upload(const std::vector<T>& h, cudaStream_t st)
    {
        if (h.size() != n) alloc(h.size());
        if (n) CK(cudaMemcpyAsync(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void view(T* base, size_t n_) { free(); p = base; n = n_; borrowed = true; }  // a slice of another allocation
    void free() { if (p && !borrowed) cudaFree(p); p = nullptr; n = 0; borrowed = false; }
    bool borrowed = false;
    ~DBuf() { free(); }
};

#define NK(call)                                                                                     \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess)                                                                       \
            throw ::afx::CommError(std::string(#call) + ": " + ::afx::NcclApi::get().GetErrorString(r_) + " (" + __FILE__ + ":" + \
                                   std::to_string(__LINE__) + ")");                                  \
    } while (0)

struct Solver;

// In-process communicator: the partitioned solvers of ONE host process, one host thread per handle, on one device or
// several.  It stands where NCCL stands -- the staged halo exchange and the small all-reduces -- so that a box with a
// single GPU (where NCCL refuses two ranks on one device) can run every partition plan on hardware, and a single-process
// driver of several GPUs needs no collective library.  Every collective call is a rendezvous of all ranks; a rank that
// never arrives breaks the group after `timeout_s` and every waiter gets AFX_ERR_COMM instead of hanging.
struct LocalGroup {
    explicit LocalGroup(int n_) : n(n_), member((size_t)n_, nullptr), slot((size_t)n_) {}
    const int n;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t gen = 0;
    bool broken = false;
    double timeout_s = getenv("AFX_GROUP_TIMEOUT_S") ? atof(getenv("AFX_GROUP_TIMEOUT_S")) : 120.;
    std::vector<Solver*> member;              // set by init_halo, read between barriers
    std::vector<std::vector<double>> slot;    // all-reduce operands, one per rank
    void barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) throw CommError("in-process group is broken (a rank failed or never arrived)");
        const uint64_t g = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); return; }
        const bool ok = cv.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] { return gen != g || broken; });
        if (!ok || broken) { broken = true; cv.notify_all(); throw CommError("in-process group: a rank did not reach the rendezvous"); }
    }
    void abort_group() { std::lock_guard<std::mutex> lk(mu); broken = true; cv.notify_all(); }
    // sum over the ranks in rank order on every rank: the same bits everywhere
    void allreduce_sum(int rank, double* v, int cnt)
    {
        slot[(size_t)rank].assign(v, v + cnt);
        barrier();
        for (int i = 0; i < cnt; ++i) {
            double t = 0;
            for (int r = 0; r < n; ++r) t += slot[(size_t)r][(size_t)i];
            v[i] = t;
        }
        barrier();  // nobody overwrites a slot that is still being read
    }
};

// halo exchange of one partitioned solver: one NCCL communicator (or the in-process group), packed sends, scattered receives
struct Halo {
    ncclComm_t comm = nullptr;
    std::shared_ptr<LocalGroup> grp;      // in-process communicator instead of NCCL
    int rank = 0, nranks = 1;
    struct Peer { int rank; uint32_t send_off, send_cnt, recv_off, recv_cnt; };
    std::vector<Peer> peers;
    DBuf<uint32_t> send_idx, recv_idx;  // internal cell ids, all peers back to back
    DBuf<d4> send_buf, recv_buf;
    uint32_t n_send = 0, n_recv = 0;
    // global patch extents for the force integrals
    std::vector<double> patch_xmin, patch_xmax, patch_ysum;
    std::vector<uint32_t> patch_count;
    // peer-memory path (CUDA IPC): one exported block = [flags: 1024 B][receive buffer: 2 parities x n_recv cells]
    bool p2p = false;
    void* ipc_block = nullptr;
    std::vector<void*> opened;            // peer blocks mapped into this process
    DBuf<uint32_t> dst_ptr, dst;          // CSR: send-layer cell -> (peer slot, position in its receive list)
    DBuf<unsigned long long> epoch;
    DBuf<unsigned int> front_done;
    DBuf<double> red;                     // device scratch of the all-reduces (norm chunks, forces)
    bool lockstep = false;                // in-process group with several ranks on ONE device: see halo_rendezvous()
    bool can_p2p = true;                  // this rank's plan fits the peer-memory path (agreed over all ranks at connect)
    bool early_signal = false;            // AFX_HALO_EARLY_SIGNAL=1
    PushArgs push{};
    SignalArgs sig{};
    WaitArgs wait{};
    ~Halo()
    {
        if (grp) { std::lock_guard<std::mutex> lk(grp->mu); if (rank >= 0 && rank < grp->n) grp->member[(size_t)rank] = nullptr; }
        for (void* p : opened) cudaIpcCloseMemHandle(p);
        if (ipc_block) cudaFree(ipc_block);
        if (comm) NcclApi::get().CommDestroy(comm);
    }
};

struct Solver {
    int device = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evp[16] = {};
    uint32_t N = 0, G = 0, E = 0, NT = 0;
    uint32_t n_upd = 0, n_grad = 0, e_flux = 0;  // sub-ranges of a partition (== N, N, E on one GPU)
    uint32_t n_front = 0;                        // partitioned: owned cells [0,n_front) are sent to peers and are advanced first
    cudaStream_t cs = nullptr;                   // halo stream: the exchange overlaps the interior cells' update
    cudaEvent_t ev_front = nullptr, ev_halo = nullptr;
    d4* h_stage = nullptr;                       // pinned staging of a partition's local state
    std::unique_ptr<Halo> halo;                  // null on one GPU
    std::vector<uint32_t> cell_l2g;              // partitioned: local reference-order cell -> global cell
    uint32_t n_global = 0;                       // partitioned: global N+G
    GasC gas{};
    int viscosity_model = 0, viscous_type = 0, visc_not_inviscid = 0;
    int second_order = 1, gradient_scheme = AFX_GRAD_GREEN_GAUSS;
    double limiter_k = 5., cfl = 1.;
    double relax_dev = -1, cfl_dev = -1, keep_dev = -1;

    // permutations (host)
    std::vector<uint32_t> c_old2new, c_new2old, f_old2new, f_new2old;
    // host copies needed after creation
    std::vector<uint32_t> h_fcells;           // new order [E][2]
    std::vector<uint32_t> h_bnd_face;         // [G] new face id of boundary b (reference boundary order)
    std::vector<int32_t> h_bnd_patch;         // [G]
    std::vector<uint8_t> h_bnd_kind;          // [G]
    std::vector<afx_bvars> h_bnd_vars;        // [G]
    std::vector<double> h_bcx, h_bcy;         // [G] boundary edge centres
    std::vector<int32_t> patch_order;         // patch ids by first edge in the (global) boundary list
    std::vector<uint8_t> patch_kinds;         // as given to set_bcs
    std::vector<afx_bvars> patch_vars;
    bool bcs_set = false;

    // device geometry
    DBuf<uint2> fcells; DBuf<d4> fgA, fgB, ftij, frec; DBuf<uint8_t> fkind; DBuf<uint32_t> cf, cnb; DBuf<d4> cgeo; DBuf<double2> cdxy; DBuf<double> area;
    DBuf<double> lsqM; DBuf<uint16_t> lsq_perm;
    // shared-memory tiles of the fused stage kernel (tiling.h)
    DBuf<uint4> t_head, t_ctab; DBuf<uint32_t> t_halo, t_face; DBuf<double> t_area, t_k3a; DBuf<double2> t_dxy; DBuf<d4> t_fgeo;
    TileTab tt{};
    uint32_t n_tiles = 0, tile_cells = 0;
    std::vector<uint32_t> tile_sizes;  // runs of consecutive cells that become tiles (graph bisection leaves)
    uint32_t n_front_runs = 0;
    size_t stage_smem = 0, n_tile_faces = 0, n_tile_cells = 0;
    double k3a_for = -1;  // limiter_k the tiles' K^3 a table was computed for
    void refresh_tile_k3a();
    int stage_ctas_per_sm = 0;
    unsigned stage_grid = 0;
    bool tiles_ready = false, use_fused = true;
    uint64_t tile_local_cells = 0;
    void build_tile_tables(const std::vector<uint32_t>& h_cf, const std::vector<uint32_t>& h_cnb, const std::vector<double2>& h_cdxy,
                           const std::vector<double>& h_area, const std::vector<d4>& h_gA);
    bool fused_stage() const { return tiles_ready && use_fused && second_order && viscous_type == 0 && !(halo && halo_overlap) && limiter_kind == 0; }
    void launch_stage(int s, const d4* qk_in, d4* qk_out, double alpha);
    // pipelined stage kernel on L2-resident chunks (rans_pipe.cuh); [0] without, [1] with the limiter phase
    DBuf<uint4> p_items[2]; DBuf<uint2> p_chunk_items; DBuf<uint32_t> p_far_faces, p_far_cells, p_far_mask; DBuf<unsigned int> p_ctr;
    PipeTab pipe_tab[2] = {};
    bool pipe_ready = false, use_pipe = false;
    unsigned pipe_grid = 0;
    void build_pipe_tables();
    bool pipe_stage() const { return pipe_ready && use_pipe && viscous_type == 0 && !(halo && halo_overlap) && !fused_stage() && limiter_kind == 0; }
    void launch_pipe(int s, const d4* qk_in, d4* qk_out, double alpha, bool has_l);
    int* pipe_err_word() { return reinterpret_cast<int*>(h_pinned + 49); }
    DBuf<uint32_t> bface, bghost, bowner; DBuf<int32_t> bpatch; DBuf<d4> bstate; DBuf<double> bcx, bcy;
    DBuf<uint32_t> perm_c_new2old, perm_c_old2new;
    // device state
    DBuf<d4> q, qkA, qkB, gx, gy, lim, qW, rhs, flux, stage;
    DBuf<d4> pm;    // [2][NT] fast mode: extremes of the projected increments per cell, written by k_dt_grad with the first stage's limiters
    bool pm_valid = false, use_pm = true;  // AFX_LIM_PM=0: the limiter kernels read gradients and face offsets as in strict mode
    const d4* pm_for_limiter() const { return (pm_valid && kt == &fast::table()) ? pm.p : nullptr; }
    DBuf<d4> grad;  // [gx | gy] in one allocation: one L2 access-policy window covers both
    void set_l2_window();
    DBuf<double> dt, dt_ref;
    DBuf<d4> J; DBuf<double> D;   // Jacobian face blocks [E][16] d4 and diagonal blocks [NT][16]
    // implicit step: block-Jacobi preconditioned restarted GMRES on the device
    DBuf<double> Dinv, kry_partial, kry_h, kry_state; DBuf<d4> kry_V, kry_w, kry_z, kry_t, kry_x; DBuf<int> kry_flag; DBuf<unsigned int> kry_counter;
    int gmres_restart = 30, gmres_max_iter = 500, precond_sweeps = 4;
    double gmres_tol = 1e-2;
    int last_linear_iters = 0;
    bool precond_valid = false;
    int compute_preconditioner();
    void apply_preconditioner(const d4* r, d4* z);
    void precondition_Ax(const d4* x, d4* r_buf, d4* z, const int* stop = nullptr);
    struct GmresView { int m; int r0() const { return (m + 1) * m + 2 * m + (m + 1) + m; } };  // offset of {r0, err, k_done, fail} in kry_state (GmresLayout, rans_krylov.cuh)
    // partitioned implicit step: inner products run over the owned rows and are summed over the ranks; a vector whose
    // neighbour rows are about to be read (matrix-vector product, Jacobi sweep) gets its halo rows from their owners first
    uint32_t n_dot() const { return halo ? n_upd : NT; }
    void halo_refresh(d4* field) { if (halo) exchange(field, st); }
    void allreduce_sum(double* dev, int n) { if (halo && n > 0) comm_allreduce(dev, n); }  // r_buf = A x, z = M^-1 r_buf; the product and the first sweep share a launch
    bool gmres(const d4* b, d4* x);
    double step_implicit(double relax, double tol, int rhs_iterations);
    DBuf<double> partial, norms, prm, scratch;
    DBuf<unsigned int> counters;  // [0] block counter, [1] norm ring index
    double* h_pinned = nullptr;   // pinned staging for scalars
    unsigned int norm_idx_host = 0;

    cudaGraphExec_t graph_exec = nullptr;
    // One Arnoldi step (product + sweeps, inner products, update + norm, rotation, scaling: 8 small kernels) as a graph per step
    // index, replayed with one call: the Krylov iteration of the shipped meshes is bound by launches.  Single-rank solvers only
    // (a partitioned step has exchanges and all-reduces between the kernels).  AFX_KRY_GRAPH=0 disables.
    std::vector<cudaGraphExec_t> kry_graph;
    std::vector<int> kry_graph_launches;
    bool use_kry_graph = true;
    bool fuse_givens = true;  // AFX_KRY_FUSE_GIVENS=0: the rotation in its own launch (k_givens_step), as on partitioned solvers
    void invalidate_kry_graphs() { for (auto& g : kry_graph) if (g) cudaGraphExecDestroy(g); kry_graph.clear(); kry_graph_launches.clear(); }
    int64_t graph_per_iter = 0;  // kernels of ours in one captured iteration
    bool use_graph = true;
    int64_t launches = 0;
    double last_ms = 0;
    bool jac_valid = false;

    DevMesh dm{};
    const KernelTable* kt = &fast::table();  // arithmetic mode, see afx_rans_set_math_mode
    NormOut norm_out() { return NormOut{partial.p, counters.p, norms.p, counters.p + 1, halo ? 1 : 0, 0, 0}; }
    void set_math_mode(int mode)
    {
        if (mode != AFX_MATH_STRICT && mode != AFX_MATH_FAST) throw InvalidArg("unknown math mode");
        const KernelTable* t = (mode == AFX_MATH_STRICT) ? &strict::table() : &fast::table();
        if (t != kt) { kt = t; invalidate_graph(); jac_valid = false; pm_valid = false; }
    }

    ~Solver()
    {
        cudaSetDevice(device);
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        invalidate_kry_graphs();
        if (h_pinned) cudaFreeHost(h_pinned);
        for (auto& e : evp) if (e) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_front) cudaEventDestroy(ev_front);
        if (ev_halo) cudaEventDestroy(ev_halo);
        if (h_stage) cudaFreeHost(h_stage);
        if (cs) cudaStreamDestroy(cs);
        if (st) cudaStreamDestroy(st);
    }

    void use() { CK(cudaSetDevice(device)); }
    static unsigned blocks(size_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

    struct DryRun { uint32_t tile_cells; TileLimits limits; TilePlan plan; std::string check; };  // host-only: renumber, tile, verify (no device)
    void create(const afx_mesh_desc& m, const afx_gas& g, int visc, int dev, const Partition* part = nullptr, DryRun* dry = nullptr);
    void init_halo(const Partition& part, const char* nccl_id, std::shared_ptr<LocalGroup> group = nullptr);
    void comm_allreduce(double* dev, int n);   // sum over the ranks, in place, on the solver's stream
    // Ranks of an in-process group that SHARE a device cannot rely on a kernel of one rank running while a kernel of another
    // spins on its flag (two streams may sit in one hardware queue; measured on the B200: 3- and 8-rank runs ran into the bounded
    // wait).  There the ranks meet on the host after their flags are raised and before anybody waits: the wait kernel finds its
    // flags set.  Same kernels, same buffers, same epochs -- only the overlap (and the CUDA graph) is given up.
    void halo_rendezvous(cudaStream_t s_) { if (halo && halo->lockstep) { CK(cudaStreamSynchronize(s_)); halo->grp->barrier(); } }
    bool graph_ok() const { return use_graph && !(halo && halo->grp && (!halo->p2p || halo->lockstep)); }
    void check_comm();                         // AFX_ERR_COMM if a halo wait gave up on a peer
    int* comm_err_word() { return reinterpret_cast<int*>(h_pinned + 48); }
    size_t p2p_export(void* blob);
    void p2p_connect(const void* blobs, size_t blob_size, int nranks);
    void exchange(d4* field, cudaStream_t stream);
    int prof_stage = -1;  // >= 0 while afx_rans_profile_explicit runs stage `prof_stage`: the halo hand-off is bracketed by evp[8 + 2 s], evp[9 + 2 s]
    void prof_mark(int which)
    {
        if (prof_stage < 0 || prof_stage >= 3) return;
        if (which < 2) CK(cudaEventRecord(evp[8 + 2 * prof_stage + which], st));
        else CK(cudaEventRecord(evp[2 + prof_stage], st));  // after the signalling kernel (evp[2..4])
    }
    double prof_halo_ms[2] = {0, 0};  // last afx_rans_profile_explicit: signalling kernel, wait + scatter kernel (per iteration)
    bool halo_pending = false;
    bool halo_overlap = false;  // AFX_HALO_OVERLAP=1: exchange on a second stream under the interior update (needed for the NCCL halo)
    void ensure_halo() { if (halo_pending) { CK(cudaStreamWaitEvent(st, ev_halo, 0)); halo_pending = false; } }
    void reduce_norms(double* v, int n);
    void set_bcs(int n_patch, const uint8_t* kinds, const afx_bvars* vars);
    void set_options(int so, int grad, double k);
    void push_params(double relax, bool keep_qW = true);
    void launch_dt_grad(bool want_grad, bool walls, bool with_lim = false);
    bool fuse_lim0 = true;  // AFX_FUSE_LIM0=0: keep the first stage's limiter in its own k_limiter launch
    // afx_rans_set_limiter: AFX_LIMITER_VENKATAKRISHNAN (the reference's default build) or AFX_LIMITER_MICHALAK (its
    // RANS_MICHALAK_LIMITER build, solver.h:557-576).  The Michalak function runs in its own kernel on the three-kernel stage:
    // no limiter inside k_dt_grad, no stored extremes, no fused / pipelined stage kernel.
    int limiter_kind = 0;
    void set_limiter(int kind)
    {
        if (kind != AFX_LIMITER_VENKATAKRISHNAN && kind != AFX_LIMITER_MICHALAK) throw InvalidArg("unknown limiter");
        if (kind != limiter_kind) { limiter_kind = kind; pm_valid = false; invalidate_graph(); }
    }
    bool lim0_fused() const { return fuse_lim0 && second_order && limiter_kind == 0; }
    // the first stage limits the iteration-start state: k_dt_grad can do it from the neighbour states it has just read
    bool lim0_in_dt_grad() const { return lim0_fused() && !fused_stage(); }  // also ahead of the pipelined stage
    void launch_limiter(const d4* qk);
    void launch_flux(const d4* qk, bool uniform, d4 qfar);
    template <int MODE, int LAST>
    void launch_gather(const d4* qk_in, d4* qk_out, d4* vec_out, double alpha, bool walls);
    void explicit_iteration();
    void invalidate_graph() { if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; } invalidate_kry_graphs(); }
    void run_explicit(double relax, int n_iter, double* norms_out);
    double fetch_last_norm();
    void check_norm(double v) { if (!(v == v) || std::isinf(v)) throw NumericError("residual norm is not finite"); }
    double residual_rhs();
    double uniform_residual();
    void fill_jacobian();
    void to_ref_order4(const d4* dev_new, double* host_out);
    void from_ref_order4(const double* host_in, d4* dev_new);
    void sync_ghost_rows();
    bool boundary_variables(afx_bvars* out) const;
    void conservative(const afx_bvars& v, double q4[4]) const;
};

// core.h:73-83
void Solver::conservative(const afx_bvars& b, double q4[4]) const
{
    const double c = std::sqrt(gas.gamma * gas.R * b.T);
    const double u = b.mach * c * std::cos(b.angle);
    const double v = b.mach * c * std::sin(b.angle);
    const double rho = b.p / (gas.R * b.T);
    const double rhoE = b.p / (gas.gamma - 1) + 0.5 * rho * (u * u + v * v);
    q4[0] = rho; q4[1] = rho * u; q4[2] = rho * v; q4[3] = rhoE;
}

// solver.h:597-611
bool Solver::boundary_variables(afx_bvars* out) const
{
    // The reference walks the boundary edges of the WHOLE mesh and stops at the first far-field one (or at the first
    // "inlet-outlet" one, whose variables are the defaults).  Per patch in order of first appearance that is the same
    // search, and it does not depend on which boundary edges this rank happens to hold.
    *out = afx_bvars{0.2, 0., 1., 1.};
    for (int32_t p : patch_order) {
        if (p < 0 || p >= (int32_t)patch_kinds.size()) continue;
        if (patch_kinds[p] == AFX_BC_FARFIELD) { *out = patch_vars[p]; return true; }
        if (patch_kinds[p] == AFX_BC_INLET_OUTLET) return false;
    }
    return false;
}

void Solver::create(const afx_mesh_desc& m, const afx_gas& g, int visc, int dev, const Partition* part, DryRun* dry)
{
    NvtxRange nvtx_("afx:create (renumber, tables, upload)");
    if (!dry) {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError("no CUDA device available: libaeroflex_rans_b200 has no CPU fallback");
        if (dev < 0 || dev >= ndev) throw InvalidArg("device ordinal out of range");
        device = dev;
        use();
        cudaDeviceProp prop{};
        CK(cudaGetDeviceProperties(&prop, dev));
        // the kernels are built as arch=compute_100a,code=sm_100a: arch-specific, no forward compatibility (sm_103 / sm_110 / sm_120
        // parts would fail at the first launch with "no kernel image")
        if (prop.major != 10 || prop.minor != 0)
            throw CudaError(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                            "; this library is built for sm_100a (B200) only");
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
        for (auto& e : evp) CK(cudaEventCreate(&e));
        CK(cudaMallocHost(&h_pinned, 64 * sizeof(double)));
        std::memset(h_pinned, 0, 64 * sizeof(double));
    }
    if (const char* e = getenv("AFX_NO_GRAPH")) use_graph = !(e[0] == '1');
    if (const char* e = getenv("AFX_HALO_OVERLAP")) halo_overlap = (e[0] == '1');
    if (const char* e = getenv("AFX_FUSE_LIM0")) fuse_lim0 = !(e[0] == '0');
    if (const char* e = getenv("AFX_MATH")) kt = (std::string(e) == "strict") ? &strict::table() : &fast::table();

    N = m.n_cells; G = m.n_ghost; E = m.n_edges; NT = N + G;
    if (!N || !E) throw InvalidArg("empty mesh");
    n_upd = part ? part->n_own : N;
    n_grad = part ? part->n_own + part->n_r1 : N;
    if (E > CF_ID) throw InvalidArg("too many edges for the 30-bit face index");
    gas = GasC{g.gamma, g.R, g.mu_L, g.Pr_L, g.cp};
    viscosity_model = visc;
    viscous_type = (visc == AFX_VISC_LAMINAR) ? 1 : 0;   // solver.h:203-208 ("SA" never matches)
    visc_not_inviscid = (visc != AFX_VISC_INVISCID);

    // ---- cell renumbering ----
    // Hilbert curve over rank coordinates (coalescing, compact partitions); with the fused stage kernel the advanced
    // cells are then regrouped by recursive graph bisection so that consecutive runs are compact patches of the cell
    // graph: those runs are the kernel's shared-memory tiles (AFX_ORDER=hilbert keeps the curve and cuts it evenly).
    c_old2new.resize(NT); c_new2old.resize(NT);
    const char* ord = getenv("AFX_ORDER");
    const bool hilbert = !(ord && std::string(ord) == "none");
    // AFX_FUSED=1 opts in to the fused stage kernel (measured at parity with the three-kernel stage)
    bool want_tiles = false;
    if (const char* e = getenv("AFX_FUSED")) want_tiles = (e[0] == '1');
    if (dry) want_tiles = true;
    tile_cells = dry ? dry->tile_cells : 192;
    if (const char* e = getenv("AFX_TILE")) if (!dry) tile_cells = (uint32_t)std::max(32, std::min(1024, atoi(e)));
    const bool graph_tiles = want_tiles && !(ord && (std::string(ord) == "hilbert" || std::string(ord) == "none"));
    tile_sizes.clear(); n_front_runs = 0;
    {
        std::vector<uint32_t> idx(N);
        std::iota(idx.begin(), idx.end(), 0u);
        // a partition arrives in curve order inside each class (owned | ring 1 | ring 2) and must keep its classes
        if (hilbert && N > 64 && !part) idx = hilbert_order(m.cells_cx, m.cells_cy, idx);
        if (part) {  // owned cells that some peer needs come first: they are advanced, then sent while the rest is advanced
            std::vector<uint8_t> front(N, 0);
            for (const auto& p : part->peers) for (uint32_t c : p.send) front[c] = 1;
            std::stable_partition(idx.begin(), idx.begin() + n_upd, [&](uint32_t c) { return front[c] != 0; });
            n_front = (uint32_t)std::count(front.begin(), front.begin() + n_upd, (uint8_t)1);
        }
        const bool split_front = n_front > 0 && n_front < n_upd;
        if (graph_tiles) {
            std::vector<uint32_t> nbr((size_t)4 * N, 0xFFFFFFFFu);  // cell -> neighbour cells, reference numbering
#pragma omp parallel for schedule(static)
            for (int64_t c = 0; c < (int64_t)N; ++c) {
                const uint32_t sz = m.cells_is_tri[c] ? 3u : 4u;
                for (uint32_t k = 0; k < sz; ++k) {
                    const uint32_t e = m.cells_edges[4 * (size_t)c + k];
                    if (e >= E) continue;
                    const uint32_t a = m.edges_cells[2 * (size_t)e], b = m.edges_cells[2 * (size_t)e + 1];
                    const uint32_t j = (a == (uint32_t)c) ? b : a;
                    if (j < N) nbr[4 * (size_t)c + k] = j;
                }
            }
            auto regroup = [&](uint32_t lo, uint32_t hi) {
                if (hi <= lo) return;
                std::vector<uint32_t> sub(idx.begin() + lo, idx.begin() + hi);
                sub = graph_tile_order(N, nbr.data(), std::move(sub), tile_cells, tile_sizes);
                std::copy(sub.begin(), sub.end(), idx.begin() + lo);
            };
            if (split_front) { regroup(0, n_front); n_front_runs = (uint32_t)tile_sizes.size(); regroup(n_front, n_upd); }
            else regroup(0, n_upd);
        } else if (want_tiles) {  // even cuts of the curve
            if (split_front) { tile_sizes.push_back(n_front); n_front_runs = 1; tile_sizes.push_back(n_upd - n_front); }
            else tile_sizes.push_back(n_upd);
        }
        for (uint32_t n = 0; n < N; ++n) { c_new2old[n] = idx[n]; c_old2new[idx[n]] = n; }
        // ghosts follow their owners
        std::vector<uint32_t> gb(G);
        std::iota(gb.begin(), gb.end(), 0u);
        std::stable_sort(gb.begin(), gb.end(), [&](uint32_t a, uint32_t b) {
            return c_old2new[m.edges_cells[2 * (size_t)m.boundary_edges[a]]] < c_old2new[m.edges_cells[2 * (size_t)m.boundary_edges[b]]];
        });
        for (uint32_t k = 0; k < G; ++k) {
            const uint32_t old_ghost = m.edges_cells[2 * (size_t)m.boundary_edges[gb[k]] + 1];
            if (old_ghost < N || old_ghost >= NT) throw InvalidArg("boundary edge without a ghost cell");
            c_new2old[N + k] = old_ghost; c_old2new[old_ghost] = N + k;
        }
    }
    // ---- faces sorted by their lower (new) cell ----
    f_old2new.resize(E); f_new2old.resize(E);
    {
        std::vector<uint32_t> fi(E);
        std::iota(fi.begin(), fi.end(), 0u);
        std::vector<uint32_t> key(E);
        for (uint32_t e = 0; e < E; ++e) {
            const uint32_t a = c_old2new[m.edges_cells[2 * (size_t)e]], b = c_old2new[m.edges_cells[2 * (size_t)e + 1]];
            key[e] = std::min(a, b);
        }
        std::stable_sort(fi.begin(), fi.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        for (uint32_t n = 0; n < E; ++n) { f_new2old[n] = fi[n]; f_old2new[fi[n]] = n; }
        e_flux = 0;
        while (e_flux < E && key[fi[e_flux]] < n_upd) ++e_flux;  // faces touching an advanced cell come first
    }
    // ---- face records ----
    std::vector<uint2> h_fc(E); std::vector<d4> h_gA(E), h_gB(E), h_t(E);
    h_fcells.resize(2 * (size_t)E);
    std::vector<uint8_t> is_bnd(E, 0);
    for (uint32_t b = 0; b < G; ++b) is_bnd[m.boundary_edges[b]] = 1;
    for (uint32_t n = 0; n < E; ++n) {
        const uint32_t e = f_new2old[n];
        const uint32_t o0 = m.edges_cells[2 * (size_t)e], o1 = m.edges_cells[2 * (size_t)e + 1];
        h_fc[n] = make_uint2(c_old2new[o0], c_old2new[o1]);
        h_fcells[2 * (size_t)n] = h_fc[n].x; h_fcells[2 * (size_t)n + 1] = h_fc[n].y;
        // Green-Gauss weight exactly as solver.h:436-444
        const double dxif = m.edges_cx[e] - m.cells_cx[o0], dyif = m.edges_cy[e] - m.cells_cy[o0];
        const double dif = std::sqrt(dxif * dxif + dyif * dyif);
        const double dxij = m.cells_cx[o0] - m.cells_cx[o1], dyij = m.cells_cy[o0] - m.cells_cy[o1];
        const double dij = std::sqrt(dxij * dxij + dyij * dyij);
        h_gA[n] = d4{m.edges_nx[e], m.edges_ny[e], m.edges_len[e], dif / dij};
        // reconstruction / limiter offsets, solver.h:541-542, 774-778
        h_gB[n] = d4{m.edges_cx[e] - m.cells_cx[o0], m.edges_cy[e] - m.cells_cy[o0],
                     m.edges_cx[e] - m.cells_cx[o1], m.edges_cy[e] - m.cells_cy[o1]};
        // face-gradient direction, solver.h:369-376
        double t0 = m.cells_cx[o0] + m.cells_cx[o1], t1 = m.cells_cy[o0] + m.cells_cy[o1];
        const double l = std::sqrt(t0 * t0 + t1 * t1);
        t0 /= l; t1 /= l;
        h_t[n] = d4{t0, t1, l, 0.};
    }
    // ---- cell -> face lists, slots in ascending ORIGINAL edge id ----
    std::vector<uint32_t> h_cf((size_t)4 * N, CF_NONE), h_cnb((size_t)4 * N, CF_NONE);
    std::vector<double2> h_cdxy((size_t)4 * N, make_double2(0., 0.));
    std::vector<uint16_t> h_perm(N, 0);
    std::vector<double> h_area(NT);
    for (uint32_t n = 0; n < NT; ++n) h_area[n] = m.cells_area[c_new2old[n]];
    for (uint32_t n = 0; n < n_grad; ++n) {
        const uint32_t o = c_new2old[n];
        const uint32_t sz = m.cells_is_tri[o] ? 3u : 4u;
        uint32_t es[4]; int order[4] = {0, 1, 2, 3};
        for (uint32_t k = 0; k < sz; ++k) {
            es[k] = m.cells_edges[4 * (size_t)o + k];
            if (es[k] >= E) throw InvalidArg("cellsEdges refers to a missing edge");
        }
        std::sort(order, order + sz, [&](int a, int b) { return es[a] < es[b]; });
        uint16_t perm = (uint16_t)(sz << 8);
        for (uint32_t slot = 0; slot < sz; ++slot) {
            const uint32_t e = es[order[slot]];
            uint32_t v = f_old2new[e];
            if (m.edges_cells[2 * (size_t)e] != o) v |= CF_SIDE;
            if (is_bnd[e]) v |= CF_BND;
            h_cf[(size_t)slot * N + n] = v;
            const bool is_c1 = (m.edges_cells[2 * (size_t)e] != o);
            h_cnb[(size_t)slot * N + n] = c_old2new[m.edges_cells[2 * (size_t)e + (is_c1 ? 0 : 1)]];
            // the reference's own subtraction (solver.h:541-542): face centre minus this cell's centre
            h_cdxy[(size_t)slot * N + n] = make_double2(m.edges_cx[e] - m.cells_cx[o], m.edges_cy[e] - m.cells_cy[o]);
            perm |= (uint16_t)(slot << (2 * order[slot]));  // local side order[slot] lives in this slot
        }
        h_perm[n] = perm;
    }
    // ---- boundary tables (reference boundary order) ----
    patch_order.clear();
    if (part) patch_order = part->patch_order;
    else {
        std::vector<uint8_t> seen;
        for (uint32_t b = 0; b < G; ++b) {
            const int32_t p = m.boundary_patch[b];
            if (p < 0) continue;
            if ((size_t)p >= seen.size()) seen.resize((size_t)p + 1, 0);
            if (!seen[p]) { seen[p] = 1; patch_order.push_back(p); }
        }
    }
    h_bnd_face.resize(G); h_bnd_patch.assign(m.boundary_patch, m.boundary_patch + G);
    h_bnd_kind.assign(G, 0); h_bnd_vars.assign(G, afx_bvars{0.2, 0., 1., 1.});
    h_bcx.resize(G); h_bcy.resize(G);
    std::vector<uint32_t> h_bghost(G), h_bowner(G);
    for (uint32_t b = 0; b < G; ++b) {
        const uint32_t e = m.boundary_edges[b];
        h_bnd_face[b] = f_old2new[e];
        h_bowner[b] = c_old2new[m.edges_cells[2 * (size_t)e]];
        h_bghost[b] = c_old2new[m.edges_cells[2 * (size_t)e + 1]];
        h_bcx[b] = m.edges_cx[e]; h_bcy[b] = m.edges_cy[e];
    }

    if (dry) {  // the host half only: tile the renumbered mesh and verify the plan against the connectivity
        dry->plan = build_tiles(N, n_grad, h_cf.data(), h_cnb.data(), tile_sizes, n_front_runs, tile_cells, dry->limits);
        dry->check = check_tiles(dry->plan, N, n_upd, h_cf.data(), h_cnb.data());
        return;
    }
    build_tile_tables(h_cf, h_cnb, h_cdxy, h_area, h_gA);

    {   // the face kernel's 64-byte records (kind bits: set_bcs)
        std::vector<d4> h_rec((size_t)2 * E);
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < (int64_t)E; ++n) {
            const unsigned long long cw = (unsigned long long)h_fc[n].x | ((unsigned long long)h_fc[n].y << 32);
            double cwd;
            std::memcpy(&cwd, &cw, 8);
            h_rec[2 * (size_t)n] = d4{h_gA[n].x, h_gA[n].y, h_gA[n].z, h_gB[n].x};
            h_rec[2 * (size_t)n + 1] = d4{h_gB[n].y, h_gB[n].z, h_gB[n].w, cwd};
        }
        frec.upload(h_rec, st);
    }

    // ---- upload ----
    fcells.upload(h_fc, st); fgA.upload(h_gA, st); fgB.upload(h_gB, st);
    if (viscous_type == 1) ftij.upload(h_t, st);
    std::vector<uint8_t> h_kind(E, 0);
    fkind.upload(h_kind, st);
    {   // device copy of the neighbour table carries the side / boundary flags of cf; per-slot face geometry beside it
        if (NT > CF_ID) throw InvalidArg("too many cells for the 30-bit neighbour index");
        std::vector<uint32_t> h_cnbf(h_cnb);
        std::vector<d4> h_cgeo((size_t)4 * N, d4{0., 0., 0., 0.});
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < (int64_t)h_cnbf.size(); ++k)
            if (h_cf[k] != CF_NONE) { h_cnbf[k] |= h_cf[k] & (CF_SIDE | CF_BND); h_cgeo[k] = h_gA[h_cf[k] & CF_ID]; }
        cnb.upload(h_cnbf, st); cgeo.upload(h_cgeo, st);
    }
    cf.upload(h_cf, st); cdxy.upload(h_cdxy, st); area.upload(h_area, st); lsq_perm.upload(h_perm, st);
    bface.upload(h_bnd_face, st); bghost.upload(h_bghost, st); bowner.upload(h_bowner, st);
    bpatch.upload(h_bnd_patch, st); bcx.upload(h_bcx, st); bcy.upload(h_bcy, st);
    bstate.alloc(G ? G : 1);
    perm_c_new2old.upload(c_new2old, st); perm_c_old2new.upload(c_old2new, st);

    grad.alloc(2 * (size_t)NT); grad.zero(st);
    gx.view(grad.p, NT); gy.view(grad.p + NT, NT);
    for (DBuf<d4>* b : {&q, &qkA, &qkB, &lim, &qW, &rhs, &stage}) { b->alloc(NT); b->zero(st); }
    if (const char* e = getenv("AFX_LIM_PM")) use_pm = !(e[0] == '0');
    if (const char* e = getenv("AFX_KRY_GRAPH")) use_kry_graph = !(e[0] == '0');
    if (const char* e = getenv("AFX_KRY_FUSE_GIVENS")) fuse_givens = !(e[0] == '0');
    pm.alloc(2 * (size_t)NT); pm.zero(st);
    set_l2_window();
    flux.alloc(E); flux.zero(st);
    dt.alloc(NT); dt.zero(st); dt_ref.alloc(NT);
    build_pipe_tables();
    partial.alloc(std::max<size_t>(std::max<size_t>(kt->gather_blocks(NT), stage_grid), (size_t)pipe_tab[0].nU_items) + 4);
    norms.alloc(NORM_RING); norms.zero(st);
    prm.alloc(8); prm.zero(st); counters.alloc(4); counters.zero(st); scratch.alloc(16);
    // limiters start at 1 (ghost rows keep that value, solver.h:519)
    kt->fill_cells(lim.p, NT, d4{1, 1, 1, 1}, st);
    ++launches;

    // least-squares coefficients (M * dT), rows in cellsEdges order, solver.h:402-422 + 504
    {
        std::vector<double> h_M((size_t)8 * N, 0.);
        for (uint32_t n = 0; n < n_grad; ++n) {
            const uint32_t o = c_new2old[n];
            const uint32_t sz = m.cells_is_tri[o] ? 3u : 4u;
            double d[4][2], a00 = 0, a01 = 0, a10 = 0, a11 = 0;
            for (uint32_t j = 0; j < sz; ++j) {
                const uint32_t e = m.cells_edges[4 * (size_t)o + j];
                const uint32_t nb = m.edges_cells[2 * (size_t)e] == o ? m.edges_cells[2 * (size_t)e + 1] : m.edges_cells[2 * (size_t)e];
                d[j][0] = m.cells_cx[nb] - m.cells_cx[o];
                d[j][1] = m.cells_cy[nb] - m.cells_cy[o];
                a00 += d[j][0] * d[j][0]; a01 += d[j][0] * d[j][1]; a10 += d[j][1] * d[j][0]; a11 += d[j][1] * d[j][1];
            }
            const double det = a00 * a11 - a10 * a01, inv = 1. / det;
            const double M0 = a11 * inv, M1 = -a01 * inv, M2 = -a10 * inv, M3 = a00 * inv;
            for (uint32_t j = 0; j < sz; ++j) {
                h_M[(size_t)j * N + n] = M0 * d[j][0] + M1 * d[j][1];
                h_M[(size_t)(4 + j) * N + n] = M2 * d[j][0] + M3 * d[j][1];
            }
        }
        lsqM.upload(h_M, st);
    }
    CK(cudaStreamSynchronize(st));

    dm.N = N; dm.G = G; dm.E = E; dm.NT = NT;
    dm.n_upd = n_upd; dm.n_grad = n_grad; dm.e_flux = e_flux;
    dm.fcells = fcells.p; dm.fgA = fgA.p; dm.fgB = fgB.p; dm.ftij = ftij.p; dm.fkind = fkind.p; dm.frec = frec.p;
    dm.cf = cf.p; dm.cnb = cnb.p; dm.cgeo = cgeo.p; dm.cdxy = cdxy.p; dm.area = area.p; dm.lsqM = lsqM.p; dm.lsq_perm = lsq_perm.p;
}

// The gradients are written once per iteration and read by the limiter and the flux kernel of all three stages: where
// both arrays fit the part of the 126 MB L2 that may be set aside for persisting lines, every kernel on the solver's
// stream (and every node captured from it) keeps them resident.  Larger meshes stream as before.  AFX_L2_PERSIST=0 disables.
void Solver::set_l2_window()
{
    if (const char* e = getenv("AFX_L2_PERSIST")) if (e[0] == '0') return;
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    const size_t bytes = grad.n * sizeof(d4);
    if (prop.persistingL2CacheMaxSize <= 0 || bytes == 0 || bytes > (size_t)prop.persistingL2CacheMaxSize || bytes > (size_t)prop.accessPolicyMaxWindowSize) return;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue a{};
    a.accessPolicyWindow.base_ptr = grad.p;
    a.accessPolicyWindow.num_bytes = bytes;
    a.accessPolicyWindow.hitRatio = 1.0f;
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
}

// solver::set_bcs, solver.h:200-247
void Solver::set_bcs(int n_patch, const uint8_t* kinds, const afx_bvars* vars)
{
    use();
    std::vector<uint8_t> h_kind(E, 0);
    std::vector<d4> h_state(G ? G : 1);
    if (n_patch < 0 || (n_patch > 0 && (!kinds || !vars))) throw InvalidArg("null argument");
    for (int32_t p : patch_order)  // a partition may hold no edge of a patch of the whole mesh: bcs.at still needs it
        if (p >= n_patch) throw InvalidArg("boundary patch " + std::to_string(p) + " has no boundary condition (bcs.at)");
    patch_kinds.assign(kinds, kinds + n_patch);
    patch_vars.assign(vars, vars + n_patch);
    for (uint32_t b = 0; b < G; ++b) {
        const int p = h_bnd_patch[b];
        if (p < 0 || p >= n_patch) throw InvalidArg("boundary patch " + std::to_string(p) + " has no boundary condition (bcs.at)");
        const uint8_t k = kinds[p] <= 3 ? kinds[p] : 0;  // anything else (AFX_BC_INLET_OUTLET included) keeps the internal flux
        h_bnd_kind[b] = k;
        h_bnd_vars[b] = (k == AFX_BC_FARFIELD) ? vars[p] : afx_bvars{0.2, 0., 1., 1.};  // solver.h:219-229
        h_kind[h_bnd_face[b]] = k;
        double s4[4];
        conservative(h_bnd_vars[b], s4);
        h_state[b] = d4{s4[0], s4[1], s4[2], s4[3]};
    }
    fkind.upload(h_kind, st);
    bstate.upload(h_state, st);
    kt->face_record_kinds(frec.p, fkind.p, E, st); ++launches;
    if (tiles_ready) { kt->tile_face_kinds(t_fgeo.p, t_face.p, fkind.p, n_tile_faces, st); ++launches; }
    CK(cudaStreamSynchronize(st));
    dm.fkind = fkind.p;
    bcs_set = true;
    jac_valid = false;
}

void Solver::set_options(int so, int grad, double k)
{
    if (grad != AFX_GRAD_GREEN_GAUSS && grad != AFX_GRAD_LEAST_SQUARES) throw InvalidArg("unknown gradient scheme");
    if (so != second_order || grad != gradient_scheme) invalidate_graph();
    second_order = so ? 1 : 0; gradient_scheme = grad; limiter_k = k;
    invalidate_graph();
}

void Solver::push_params(double relax, bool keep_qW)
{
    const double kq = keep_qW ? 1.0 : 0.0;
    if (relax == relax_dev && cfl == cfl_dev && kq == keep_dev) return;
    h_pinned[32] = cfl; h_pinned[33] = relax; h_pinned[34] = kq;
    CK(cudaMemcpyAsync(prm.p, h_pinned + 32, 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // the pinned slot is reused
    relax_dev = relax; cfl_dev = cfl; keep_dev = kq;
}

void Solver::launch_dt_grad(bool want_grad, bool walls, bool with_lim)
{
    ensure_halo();
    const bool lim_here = with_lim && want_grad;
    d4* pm_out = (lim_here && use_pm && kt == &fast::table()) ? pm.p : nullptr;
    kt->dt_grad(gradient_scheme == AFX_GRAD_GREEN_GAUSS ? 0 : 1, dm, q.p, dt.p, gx.p, gy.p, prm.p, gas.gamma, want_grad, walls,
                lim_here ? lim.p : nullptr, limiter_k, pm_out, st);
    pm_valid = pm_out != nullptr;  // any other gradient pass leaves the stored extremes stale
    ++launches;
}

void Solver::launch_limiter(const d4* qk)
{
    const int walls = (visc_not_inviscid || second_order) ? 1 : 0;
    if (limiter_kind == AFX_LIMITER_MICHALAK) {  // same cell ranges, its own kernel
        if (halo_pending && n_front > 0 && n_front < n_upd) {
            kt->limiter_michalak(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, n_front, n_upd - n_front, 0, 0, st);
            ensure_halo();
            kt->limiter_michalak(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, 0, n_front, n_upd, n_grad - n_upd, st);
            launches += 2;
            return;
        }
        ensure_halo();
        kt->limiter_michalak(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, 0, n_grad, 0, 0, st);
        ++launches;
        return;
    }
    if (halo_pending && n_front > 0 && n_front < n_upd) {
        // owned cells farther than two hops from any foreign cell do not see the halo: limit them while it is in flight
        kt->limiter(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, n_front, n_upd - n_front, 0, 0, pm_for_limiter(), st);
        ensure_halo();
        kt->limiter(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, 0, n_front, n_upd, n_grad - n_upd, pm_for_limiter(), st);
        launches += 2;
        return;
    }
    ensure_halo();
    kt->limiter(dm, qk, gx.p, gy.p, lim.p, limiter_k, walls, 0, n_grad, 0, 0, pm_for_limiter(), st);
    ++launches;
}

void Solver::launch_flux(const d4* qk, bool uniform, d4 qfar)
{
    ensure_halo();
    kt->flux(second_order, viscous_type, uniform ? 1 : 0, dm, qk, q.p, gx.p, gy.p, lim.p, flux.p, gas, qfar, st);
    ++launches;
}

template <int MODE, int LAST>
void Solver::launch_gather(const d4* qk_in, d4* qk_out, d4* vec_out, double alpha, bool walls)
{
    const bool split_front = n_front > 0 && n_front < n_upd;
    // LAST: the launches below leave one partial sum per block; one small kernel adds them (k_norm_finish)
    auto finish = [&](unsigned blocks) { if (LAST) { kt->norm_finish(norm_out(), blocks, st); ++launches; } };
    if (halo && MODE == 0 && halo->p2p && !(halo_overlap && split_front)) {
        // one launch advances everything and pushes the send layer into the peers' buffers; a flag hand-off and a
        // small wait+scatter kernel complete the halo on the same stream (no fork/join, no split launches)
        if (halo->early_signal) {
            // the last CTA of the send layer raises the peers' flags from inside the update kernel: the hand-off travels
            // while the interior cells are still being advanced, and the signalling launch disappears
            PushArgs pa = halo->push;
            pa.early_signal = 1;
            pa.n_front_blocks = kt->gather_blocks(n_front);
            kt->gather(MODE, LAST, dm, 0, n_upd, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, norm_out(), &pa, st);
            prof_mark(0);
            halo_rendezvous(st); kt->halo_wait_scatter(halo->wait, qk_out, st);
            prof_mark(1);
            launches += 2;
            finish(kt->gather_blocks(n_upd));
            return;
        }
        kt->gather(MODE, LAST, dm, 0, n_upd, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, norm_out(), &halo->push, st);
        prof_mark(0);
        kt->halo_signal(halo->sig, st);
        prof_mark(2);
        halo_rendezvous(st); kt->halo_wait_scatter(halo->wait, qk_out, st);
        prof_mark(1);
        launches += 3;
        finish(kt->gather_blocks(n_upd));
        return;
    }
    if (halo && MODE == 0 && split_front && !(halo->grp && !halo->p2p)) {
        // send layer first, then the exchange on the halo stream while the interior cells are advanced
        NormOut no = norm_out();
        const unsigned b0 = kt->gather_blocks(n_front), b1 = kt->gather_blocks(n_upd - n_front);
        no.blk_off = 0; no.blk_total = b0 + b1;
        if (halo->p2p) {
            // fused: the update kernel stores the send layer straight into the peers' receive buffers (NVLink), a
            // one-warp kernel raises the peers' flags, and the halo stream waits for OUR flags and fills our halo
            kt->gather(MODE, LAST, dm, 0, n_front, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, no, &halo->push, st);
            kt->halo_signal(halo->sig, st);
            CK(cudaEventRecord(ev_front, st));
            CK(cudaStreamWaitEvent(cs, ev_front, 0));
            halo_rendezvous(st); kt->halo_wait_scatter(halo->wait, qk_out, cs);
            launches += 2;
        } else {
            kt->gather(MODE, LAST, dm, 0, n_front, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, no, nullptr, st);
            CK(cudaEventRecord(ev_front, st));
            CK(cudaStreamWaitEvent(cs, ev_front, 0));
            exchange(qk_out, cs);
        }
        CK(cudaEventRecord(ev_halo, cs));
        no.blk_off = b0;
        kt->gather(MODE, LAST, dm, n_front, n_upd, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, no, nullptr, st);
        halo_pending = true;  // joined by the first kernel that reads halo cells (ensure_halo)
        launches += 2;
        finish(b0 + b1);
        return;
    }
    kt->gather(MODE, LAST, dm, 0, n_upd, flux.p, q.p, qk_in, qk_out, dt.p, vec_out, alpha, prm.p, walls ? 1 : 0, norm_out(), nullptr, st);
    ++launches;
    finish(kt->gather_blocks(n_upd));
    if (halo && MODE == 0) exchange(qk_out, st);
}

// Tiles of the fused stage kernel (AFX_FUSED=1).  AFX_TILE sets the cells per tile,
// AFX_STAGE_CTAS the resident CTAs per SM the shared-memory budget is divided by.
void Solver::build_tile_tables(const std::vector<uint32_t>& h_cf, const std::vector<uint32_t>& h_cnb, const std::vector<double2>& h_cdxy,
                               const std::vector<double>& h_area, const std::vector<d4>& h_gA)
{
    tiles_ready = false;
    if (tile_sizes.empty()) return;  // not asked for
    int ctas = 2;
    if (const char* e = getenv("AFX_STAGE_CTAS")) ctas = std::max(1, std::min(4, atoi(e)));
    int smem_max = 0;
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t budget = (size_t)(smem_max - 1024) / ctas - (ctas > 1 ? 1024 : 0);
    // staging limits of one tile: typical ratios of a compact patch (local cells 1.6x, own + ring 1 1.35x, faces 2.2x,
    // ring ids 0.7x the own cells) scaled up as far as the shared memory allows; the few tiles beyond them are cut in two
    uint32_t T = tile_cells;
    TileLimits lim;
    StageSmem L{};
    auto limits_for = [&](uint32_t t, double f) {
        TileLimits l;
        l.max_loc = (uint32_t)(1.6 * f * t); l.max_n1 = ((uint32_t)(1.35 * f * t) + 1u) & ~1u; l.max_nf = (uint32_t)(2.2 * f * t);
        l.max_halo = ((uint32_t)(0.7 * f * t) + 3u) & ~3u;
        return l;
    };
    for (;; T = T * 7 / 8) {
        double f = 1.0;
        while (f < 3.0) {
            const TileLimits l = limits_for(T, f + 0.05);
            if (stage_smem_layout(l.max_loc, l.max_n1, l.max_nf, (T + 1u) & ~1u, l.max_halo).total > budget) break;
            f += 0.05;
        }
        lim = limits_for(T, f);
        if (stage_smem_layout(lim.max_loc, lim.max_n1, lim.max_nf, (T + 1u) & ~1u, lim.max_halo).total <= budget) break;
        if (T <= 16) return;  // no tiling fits: three-kernel stage
    }
    TilePlan plan;
    try { plan = build_tiles(N, n_grad, h_cf.data(), h_cnb.data(), tile_sizes, n_front_runs, T, lim); }
    catch (const std::invalid_argument&) { return; }  // three-kernel stage
    L = stage_smem_layout(plan.max_loc, plan.max_n1, plan.max_nf, plan.max_nc, plan.max_halo);
    if (L.total > budget) return;
    static_assert(sizeof(TileHead) == 32 && sizeof(TileCell) == 16, "tile tables are read as uint4");
    tile_cells = T; n_tiles = (uint32_t)plan.head.size(); stage_smem = L.total; tile_local_cells = plan.local_cells;
    // static geometry packed per tile: face offsets and areas of the own + ring-1 cells, normals / lengths of the local faces
    const size_t n_cell_rec = plan.ctab.size();
    n_tile_cells = n_cell_rec;
    n_tile_faces = plan.face.size();
    std::vector<double> h_ta(n_cell_rec, 1.0);
    std::vector<double2> h_td(4 * n_cell_rec, make_double2(0., 0.));
    std::vector<d4> h_tf(n_tile_faces);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t t = 0; t < (int64_t)n_tiles; ++t) {
        const TileHead& h = plan.head[t];
        const uint32_t n1 = h.nc + h.h1, n1p = (n1 + 1u) & ~1u;
        for (uint32_t l = 0; l < n1; ++l) {
            const uint32_t c = l < h.nc ? h.cell0 + l : plan.halo[h.off_halo + l - h.nc];
            h_ta[h.off_cell + l] = h_area[c];
            for (int s = 0; s < 4; ++s) h_td[4 * (size_t)h.off_cell + (size_t)s * n1p + l] = h_cdxy[(size_t)s * N + c];
        }
        for (uint32_t lf = 0; lf < h.nf; ++lf) {
            const d4& a = h_gA[plan.face[h.off_face + lf]];
            h_tf[h.off_face + lf] = d4{a.x, a.y, a.z, 0.0};  // kind: set_bcs
        }
    }
    auto up = [&](auto& dbuf, const auto& host) {
        using T_ = std::remove_pointer_t<decltype(dbuf.p)>;
        dbuf.alloc(std::max<size_t>(1, host.size() * sizeof(host[0]) / sizeof(T_)));
        if (!host.empty()) CK(cudaMemcpyAsync(dbuf.p, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, st));
    };
    up(t_head, plan.head); up(t_halo, plan.halo); up(t_ctab, plan.ctab); up(t_face, plan.face); up(t_area, h_ta); up(t_dxy, h_td); up(t_fgeo, h_tf);
    t_k3a.alloc(std::max<size_t>(1, n_cell_rec));
    CK(cudaStreamSynchronize(st));
    tt = TileTab{t_head.p, t_halo.p, t_ctab.p, t_k3a.p, t_dxy.p, t_fgeo.p, n_tiles, plan.max_loc, plan.max_n1, plan.max_nf, plan.max_nc, plan.max_halo};
    const int a = strict::table().stage_prepare(stage_smem), b = fast::table().stage_prepare(stage_smem);
    if (a < 1 || b < 1) return;  // the kernel does not fit: three-kernel stage
    stage_ctas_per_sm = std::min(a, b);
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    stage_grid = std::min<unsigned>(n_tiles, (unsigned)prop.multiProcessorCount * (unsigned)stage_ctas_per_sm);
    tiles_ready = true;
}

// the tiles' K^3 a table follows limiter_k; always computed by the strict kernels (IEEE sqrt, no contraction)
void Solver::refresh_tile_k3a()
{
    if (!tiles_ready || k3a_for == limiter_k) return;
    strict::table().tile_k3a(t_area.p, t_k3a.p, n_tile_cells, limiter_k, st);
    ++launches;
    CK(cudaStreamSynchronize(st));
    k3a_for = limiter_k;
}

// Work list of the pipelined stage kernel (PipeTab, rans_types.h): chunks of 2^shift cells along the cell order, the sweep's
// steps, and the far faces / cells that are done after the sweep.  AFX_PIPE=0|1, AFX_PIPE_SHIFT, AFX_PIPE_LAGF, AFX_PIPE_LAGU.
void Solver::build_pipe_tables()
{
    pipe_ready = false;
    if (const char* e = getenv("AFX_PIPE")) use_pipe = (e[0] == '1');
    uint32_t shift = 15, lagF = 3, lagU = 2;
    if (const char* e = getenv("AFX_PIPE_SHIFT")) shift = (uint32_t)std::max(8, std::min(24, atoi(e)));
    if (const char* e = getenv("AFX_PIPE_LAGF")) lagF = (uint32_t)std::max(1, std::min(16, atoi(e)));
    if (const char* e = getenv("AFX_PIPE_LAGU")) lagU = (uint32_t)std::max(0, std::min(16, atoi(e)));
    const uint32_t IT = (uint32_t)kt->pipe_item_elems();
    const uint32_t csz = 1u << shift;
    const uint32_t n_chunks = (n_grad + csz - 1) / csz;
    if (!n_chunks || !n_upd) return;
    auto n_it = [&](uint32_t n) { return (n + IT - 1) / IT; };
    // faces are sorted by their lower cell: the faces of chunk c are a range
    std::vector<uint32_t> fs(n_chunks + 1, e_flux);
    {
        uint32_t c = 0;
        fs[0] = 0;
        for (uint32_t f = 0; f < e_flux; ++f) {
            const uint32_t lo = std::min(h_fcells[2 * (size_t)f], h_fcells[2 * (size_t)f + 1]);
            const uint32_t cf_ = lo >> shift;
            if (cf_ >= n_chunks) throw InvalidArg("face outside the chunks of the pipelined stage");
            while (c < cf_) fs[++c] = f;
        }
        while (c < n_chunks) fs[++c] = e_flux;
    }
    std::vector<uint32_t> far_f, far_c, mask((n_upd + 31) / 32 + 1, 0u);
    for (uint32_t f = 0; f < e_flux; ++f) {
        const uint32_t a = h_fcells[2 * (size_t)f], b = h_fcells[2 * (size_t)f + 1];
        if (b >= N) continue;  // boundary face: the ghost's limiter and gradient never change
        const uint32_t ca = a >> shift, cb = b >> shift;
        if (ca > cb + 1 || cb > ca + 1) {
            far_f.push_back(f);
            if (a < n_upd) mask[a >> 5] |= 1u << (a & 31);
            if (b < n_upd) mask[b >> 5] |= 1u << (b & 31);
        }
    }
    for (uint32_t i = 0; i < n_upd; ++i) if ((mask[i >> 5] >> (i & 31)) & 1u) far_c.push_back(i);
    const uint32_t n_far_f = (uint32_t)far_f.size(), n_far_c = (uint32_t)far_c.size();
    auto cellsL = [&](uint32_t c) { return std::min((c + 1) << shift, n_grad) - (c << shift); };
    auto cellsU = [&](uint32_t c) { const uint32_t lo = c << shift; return lo >= n_upd ? 0u : std::min((c + 1) << shift, n_upd) - lo; };
    std::vector<uint2> chunk_items(n_chunks);
    uint32_t nLtot = 0, nFnear = 0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        chunk_items[c] = make_uint2(n_it(cellsL(c)), n_it(fs[c + 1] - fs[c]));
        nLtot += chunk_items[c].x; nFnear += chunk_items[c].y;
    }
    // the sweep: step k = limiter items of chunk k, flux items of chunk k - lagF, update items of chunk k - lagF - lagU
    const uint32_t n_steps = n_chunks + lagF + lagU;
    uint32_t nU_items = 0;
    for (int hl = 0; hl < 2; ++hl) {
        std::vector<uint4> items;
        uint32_t slot = 0;
        auto emit = [&](uint32_t ph, uint32_t chunk, uint32_t first, uint32_t count, bool is_u) {
            for (uint32_t o = 0; o < count; o += IT) {
                items.push_back(make_uint4(ph | (chunk << 3), first + o, std::min(IT, count - o), is_u ? slot : 0u));
                if (is_u) ++slot;
            }
        };
        for (uint32_t k = 0; k < n_steps; ++k) {
            if (hl && k < n_chunks) emit(0u, k, k << shift, cellsL(k), false);
            if (k >= lagF && k - lagF < n_chunks) { const uint32_t c = k - lagF; emit(1u, c, fs[c], fs[c + 1] - fs[c], false); }
            if (k >= lagF + lagU && k - lagF - lagU < n_chunks) { const uint32_t c = k - lagF - lagU; emit(2u, c, c << shift, cellsU(c), true); }
        }
        emit(3u, 0u, 0u, n_far_f, false);
        emit(4u, 0u, 0u, n_far_c, true);
        nU_items = slot;
        p_items[hl].upload(items, st);
        PipeTab& t = pipe_tab[hl];
        t = PipeTab{};
        t.shift = shift; t.n_chunks = n_chunks; t.lagF = lagF; t.lagU = lagU;
        t.n_items = (uint32_t)items.size();
        t.n_far_faces = n_far_f; t.n_far_cells = n_far_c;
        t.nL_total = hl ? nLtot : 0u; t.nF_total = nFnear + n_it(n_far_f); t.nU_items = nU_items;
        t.items = p_items[hl].p;
    }
    if (far_f.empty()) far_f.push_back(0);
    if (far_c.empty()) far_c.push_back(0);
    p_chunk_items.upload(chunk_items, st); p_far_faces.upload(far_f, st); p_far_cells.upload(far_c, st); p_far_mask.upload(mask, st);
    p_ctr.alloc(4 + 2 * (size_t)n_chunks); p_ctr.zero(st);
    *pipe_err_word() = 0;
    for (int hl = 0; hl < 2; ++hl) {
        PipeTab& t = pipe_tab[hl];
        t.chunk_items = p_chunk_items.p; t.far_faces = p_far_faces.p; t.far_cells = p_far_cells.p; t.far_mask = p_far_mask.p;
        t.ctr = p_ctr.p; t.err = pipe_err_word();
    }
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    pipe_grid = (unsigned)prop.multiProcessorCount * (unsigned)kt->pipe_ctas_per_sm();
    pipe_ready = true;
}

// One pipelined stage: (limiter +) flux + gather + update, chunk by chunk through the L2, then the halo hand-off.
void Solver::launch_pipe(int s, const d4* qk_in, d4* qk_out, double alpha, bool has_l)
{
    ensure_halo();
    const int walls = (visc_not_inviscid || second_order) ? 1 : 0;
    const PushArgs* push = (halo && halo->p2p) ? &halo->push : nullptr;
    const PipeTab& pt = pipe_tab[has_l ? 1 : 0];
    kt->pipe(second_order, viscous_type, s == 2, has_l ? 1 : 0, dm, pt, std::min<unsigned>(pipe_grid, pt.n_items), qk_in, q.p, qk_out, gx.p, gy.p, lim.p,
             flux.p, dt.p, qW.p, alpha, prm.p, gas, limiter_k, walls, norm_out(), push, st);
    ++launches;
    if (s == 2) { kt->norm_finish(norm_out(), pt.nU_items, st); ++launches; }
    if (!halo) return;
    if (halo->p2p) {  // the update items stored the send layer into the peers' buffers: flags, then fill our halo cells
        kt->halo_signal(halo->sig, st);
        halo_rendezvous(st); kt->halo_wait_scatter(halo->wait, qk_out, st);
        launches += 2;
    } else {
        exchange(qk_out, st);
    }
}

// One fused stage: limiter + MUSCL + flux + gather + update on shared-memory tiles, then the halo hand-off.
void Solver::launch_stage(int s, const d4* qk_in, d4* qk_out, double alpha)
{
    ensure_halo();
    const PushArgs* push = (halo && halo->p2p) ? &halo->push : nullptr;
    kt->stage(s == 2, dm, tt, stage_grid, stage_smem, qk_in, q.p, qk_out, gx.p, gy.p, dt.p, qW.p, lim.p, alpha, prm.p, gas, norm_out(), push, st);
    ++launches;
    if (s == 2) { kt->norm_finish(norm_out(), stage_grid, st); ++launches; }
    if (!halo) return;
    if (halo->p2p) {  // the kernel stored the send layer into the peers' buffers: flags, then fill our halo cells
        kt->halo_signal(halo->sig, st);
        halo_rendezvous(st); kt->halo_wait_scatter(halo->wait, qk_out, st);
        launches += 2;
    } else {
        exchange(qk_out, st);
    }
    // the next stage recomputes the limiters of this rank's ring-1 halo cells from staged states: their wall ghosts follow them
    if (s < 2 && G) { kt->ghost_follow(qk_out, bghost.p, bowner.p, bface.p, fkind.p, G, n_upd, st); ++launches; }
}

// explicitSolver::solve, solver.h:802-828.  Stage 0 reads q in place of qk (they
// are equal), stage 2 writes q in place; qkA/qkB carry the intermediate stages.
void Solver::explicit_iteration()
{
    const bool grads = visc_not_inviscid || second_order;  // solver.h:810
    const bool lim0 = lim0_in_dt_grad();  // second_order => grads
    launch_dt_grad(grads, grads, lim0);
    const d4* in[3] = {q.p, qkA.p, qkB.p};
    d4* out[3] = {qkA.p, qkB.p, q.p};
    const double alpha[3] = {0.25, 0.5, 1.};  // solver.h:723
    for (int s = 0; s < 3; ++s) {
        if (fused_stage()) { launch_stage(s, in[s], out[s], alpha[s]); continue; }
        if (pipe_stage()) { launch_pipe(s, in[s], out[s], alpha[s], second_order && !(s == 0 && lim0)); continue; }
        if (second_order && !(s == 0 && lim0)) launch_limiter(in[s]);
        launch_flux(in[s], false, d4{0, 0, 0, 0});
        if (s < 2) launch_gather<0, 0>(in[s], out[s], qW.p, alpha[s], grads);
        else launch_gather<0, 1>(in[s], out[s], qW.p, alpha[s], grads);
    }
    ensure_halo();  // the halo stream joins before the iteration (and its graph) ends
}

// NCCL halo: the ring cells' states come from their owners after every stage
void Solver::init_halo(const Partition& part, const char* nccl_id, std::shared_ptr<LocalGroup> group)
{
    halo.reset(new Halo);
    Halo& h = *halo;
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&cs, cudaStreamNonBlocking, prio_hi));  // halo kernels go ahead of the bulk update
    CK(cudaEventCreateWithFlags(&ev_front, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_halo, cudaEventDisableTiming));
    CK(cudaMallocHost(&h_stage, (size_t)NT * sizeof(d4)));
    h.rank = part.rank; h.nranks = part.nranks;
    h.patch_xmin = part.patch_xmin; h.patch_xmax = part.patch_xmax; h.patch_ysum = part.patch_ysum; h.patch_count = part.patch_count;
    std::vector<uint32_t> si, ri;
    for (const auto& p : part.peers) {
        Halo::Peer q{p.rank, (uint32_t)si.size(), (uint32_t)p.send.size(), (uint32_t)ri.size(), (uint32_t)p.recv.size()};
        for (uint32_t c : p.send) si.push_back(c_old2new[c]);
        for (uint32_t c : p.recv) ri.push_back(c_old2new[c]);
        h.peers.push_back(q);
    }
    h.n_send = (uint32_t)si.size(); h.n_recv = (uint32_t)ri.size();
    if (si.empty()) si.push_back(0);
    if (ri.empty()) ri.push_back(0);
    h.send_idx.upload(si, st); h.recv_idx.upload(ri, st);
    h.send_buf.alloc(si.size()); h.recv_buf.alloc(ri.size());
    CK(cudaStreamSynchronize(st));
    h.red.alloc(64);
    if (group) {
        if (group->n != h.nranks) throw InvalidArg("the in-process group was created for another number of ranks");
        h.grp = group;
        std::lock_guard<std::mutex> lk(group->mu);
        if (group->member[(size_t)h.rank]) throw InvalidArg("rank already present in the in-process group");
        group->member[(size_t)h.rank] = this;
    } else {
        ncclUniqueId id;
        static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
        std::memcpy(&id, nccl_id, sizeof(id));
        NK(NcclApi::get().CommInitRank(&h.comm, h.nranks, id, h.rank));
    }
    cell_l2g = part.cell_l2g;
    n_global = part.n_global_cells + part.n_global_ghost;
}

void Solver::exchange(d4* field, cudaStream_t st)
{
    NvtxRange nvtx_("afx:halo exchange (collective)");
    Halo& h = *halo;
    if (h.n_send) { kt->permute4(field, h.send_buf.p, h.send_idx.p, h.n_send, st); ++launches; }
    if (h.grp) {
        // staged exchange of the in-process group: every rank packs, all meet, every rank copies its peers' packed
        // layers into its receive buffer (device to device, across devices if the handles live on several), all meet again
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CK(cudaStreamIsCapturing(st, &cap));
        if (cap != cudaStreamCaptureStatusNone) throw InvalidArg("the staged in-process halo cannot be captured into a graph");
        CK(cudaStreamSynchronize(st));
        h.grp->barrier();
        for (const auto& p : h.peers) {
            if (!p.recv_cnt) continue;
            const Solver* P = h.grp->member[(size_t)p.rank];
            if (!P || !P->halo) throw CommError("in-process group: peer rank " + std::to_string(p.rank) + " is missing");
            const Halo::Peer* mine = nullptr;
            for (const auto& q : P->halo->peers) if (q.rank == h.rank) mine = &q;
            if (!mine || mine->send_cnt != p.recv_cnt) throw CommError("halo plans of two ranks disagree");
            CK(cudaMemcpyPeerAsync(h.recv_buf.p + p.recv_off, device, P->halo->send_buf.p + mine->send_off, P->device,
                                   (size_t)p.recv_cnt * sizeof(d4), st));
        }
        if (h.n_recv) { kt->scatter4(h.recv_buf.p, field, h.recv_idx.p, h.n_recv, st); ++launches; }
        CK(cudaStreamSynchronize(st));
        h.grp->barrier();  // nobody repacks while a peer still reads
        return;
    }
    NK(NcclApi::get().GroupStart());
    for (const auto& p : h.peers) {
        if (p.send_cnt) NK(NcclApi::get().Send(h.send_buf.p + p.send_off, (size_t)p.send_cnt * 4, ncclDouble, p.rank, h.comm, st));
        if (p.recv_cnt) NK(NcclApi::get().Recv(h.recv_buf.p + p.recv_off, (size_t)p.recv_cnt * 4, ncclDouble, p.rank, h.comm, st));
    }
    NK(NcclApi::get().GroupEnd());
    if (h.n_recv) { kt->scatter4(h.recv_buf.p, field, h.recv_idx.p, h.n_recv, st); ++launches; }
}

// ---- peer-memory halo: export this rank's receive block, map the peers' ----
struct P2PBlobPeer { int32_t rank; uint32_t recv_off, recv_cnt; };
struct P2PBlob {
    cudaIpcMemHandle_t handle;
    uint32_t n_recv, n_peers;
    P2PBlobPeer peers[16];
    int64_t pid;            // exporter's process: an importer in the same process uses raw_ptr (CUDA IPC cannot map a block
    uint64_t raw_ptr;       // into the process that owns it)
    int32_t device;
    int32_t can_p2p;        // the exporter's plan fits the peer-memory path; every rank must, or all stay on the collective halo
};
constexpr size_t P2P_FLAG_BYTES = 1024;

size_t Solver::p2p_export(void* blob)
{
    if (!halo) throw InvalidArg("not a partitioned solver");
    Halo& h = *halo;
    h.can_p2p = h.peers.size() <= (size_t)P2P_MAX_PEERS && h.peers.size() <= 16;
    if (!h.ipc_block) {
        const size_t bytes = P2P_FLAG_BYTES + 2 * (size_t)std::max<uint32_t>(h.n_recv, 1) * sizeof(d4);
        CK(cudaMalloc(&h.ipc_block, bytes));
        CK(cudaMemset(h.ipc_block, 0, bytes));
        h.epoch.alloc(1);
        CK(cudaMemset(h.epoch.p, 0, sizeof(unsigned long long)));
    }
    P2PBlob b{};
    CK(cudaIpcGetMemHandle(&b.handle, h.ipc_block));
    b.n_recv = h.n_recv; b.n_peers = (uint32_t)std::min<size_t>(h.peers.size(), 16);
    b.pid = (int64_t)getpid(); b.raw_ptr = (uint64_t)(uintptr_t)h.ipc_block; b.device = device; b.can_p2p = h.can_p2p ? 1 : 0;
    for (size_t k = 0; k < h.peers.size() && k < 16; ++k) { b.peers[k].rank = h.peers[k].rank; b.peers[k].recv_off = h.peers[k].recv_off; b.peers[k].recv_cnt = h.peers[k].recv_cnt; }
    if (blob) std::memcpy(blob, &b, sizeof b);
    return sizeof b;
}

void Solver::p2p_connect(const void* blobs, size_t blob_size, int nranks)
{
    if (!halo || !halo->ipc_block) throw InvalidArg("afx_rans_p2p_export must be called first");
    Halo& h = *halo;
    if (blob_size != sizeof(P2PBlob) || nranks != h.nranks) throw InvalidArg("peer blobs do not match this build / communicator");
    use();
    const P2PBlob* all = static_cast<const P2PBlob*>(blobs);
    h.push = PushArgs{}; h.sig = SignalArgs{}; h.wait = WaitArgs{};
    // the ranks never talk about the halo mode again: it is decided here, from data every rank sees identically.  One rank
    // on the collective exchange while its peers push and spin on flags would be a silent deadlock.
    bool all_can = true;
    for (int r = 0; r < nranks; ++r) all_can = all_can && all[r].can_p2p != 0;
    h.lockstep = false;
    if (h.grp)
        for (int r = 0; r < nranks; ++r)
            for (int r2 = r + 1; r2 < nranks; ++r2)
                if (all[r].pid == all[r2].pid && all[r].device == all[r2].device) h.lockstep = true;  // the same on every rank
    if (const char* e = getenv("AFX_HALO_LOCKSTEP")) h.lockstep = h.grp && e[0] == '1';  // override (tests of the bounded wait itself)
    if (!all_can) { h.p2p = false; invalidate_graph(); return; }
    std::vector<std::vector<uint32_t>> per_cell(n_front);
    std::vector<uint32_t> si(h.n_send);
    if (h.n_send) CK(cudaMemcpy(si.data(), h.send_idx.p, (size_t)h.n_send * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < h.peers.size(); ++k) {
        const auto& pk = h.peers[k];
        const P2PBlob& pb = all[pk.rank];
        void* base = nullptr;
        if (pb.pid == (int64_t)getpid()) {  // a handle of this process (in-process group): its block is addressable as it is
            base = reinterpret_cast<void*>((uintptr_t)pb.raw_ptr);
            if (pb.device != device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(pb.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                cudaGetLastError();
            }
        } else {
            CK(cudaIpcOpenMemHandle(&base, pb.handle, cudaIpcMemLazyEnablePeerAccess));
            h.opened.push_back(base);
        }
        const P2PBlobPeer* mine = nullptr;
        for (uint32_t j = 0; j < pb.n_peers; ++j) if (pb.peers[j].rank == h.rank) mine = &pb.peers[j];
        if (!mine || mine->recv_cnt != pk.send_cnt) throw InvalidArg("halo plans of two ranks disagree");
        h.push.peer_buf[k] = reinterpret_cast<d4*>(static_cast<char*>(base) + P2P_FLAG_BYTES) + mine->recv_off;
        h.push.peer_stride[k] = pb.n_recv;
        h.sig.peer_flag[k] = reinterpret_cast<unsigned long long*>(base) + h.rank;
        h.wait.flag[k] = reinterpret_cast<const unsigned long long*>(h.ipc_block) + pk.rank;
        for (uint32_t j = 0; j < pk.send_cnt; ++j) {
            const uint32_t cell = si[pk.send_off + j];
            if (cell >= n_front) throw InvalidArg("send-layer cell outside the front range");
            per_cell[cell].push_back(((uint32_t)k << 28) | j);
        }
    }
    std::vector<uint32_t> ptr(n_front + 1, 0), dst;
    for (uint32_t c = 0; c < n_front; ++c) { ptr[c] = (uint32_t)dst.size(); dst.insert(dst.end(), per_cell[c].begin(), per_cell[c].end()); }
    ptr[n_front] = (uint32_t)dst.size();
    if (dst.empty()) dst.push_back(0);
    h.dst_ptr.upload(ptr, st); h.dst.upload(dst, st);
    CK(cudaStreamSynchronize(st));
    h.push.enabled = 1; h.push.n_front = n_front; h.push.dst_ptr = h.dst_ptr.p; h.push.dst = h.dst.p; h.push.epoch = h.epoch.p;
    h.sig.n_peers = (int)h.peers.size(); h.sig.epoch = h.epoch.p;
    h.wait.n_peers = (int)h.peers.size(); h.wait.epoch = h.epoch.p;
    h.wait.recv_buf = reinterpret_cast<const d4*>(static_cast<char*>(h.ipc_block) + P2P_FLAG_BYTES);
    h.wait.recv_idx = h.recv_idx.p; h.wait.n_recv = h.n_recv;
    double timeout_ms = 20000.;  // AFX_HALO_TIMEOUT_MS: how long a halo wait spins for a peer before it gives up (0: for ever)
    if (const char* e = getenv("AFX_HALO_TIMEOUT_MS")) timeout_ms = atof(e);
    h.wait.timeout_ns = (unsigned long long)(timeout_ms * 1e6);
    *comm_err_word() = 0;
    h.wait.err = comm_err_word();  // pinned host memory, addressable from the device (unified addressing)
    h.front_done.alloc(1);
    CK(cudaMemset(h.front_done.p, 0, sizeof(unsigned int)));
    h.push.early_signal = 0; h.push.n_front_blocks = 0; h.push.front_done = h.front_done.p; h.push.epoch_rw = h.epoch.p;
    h.push.n_peers = (int)h.peers.size();
    for (size_t k = 0; k < h.peers.size(); ++k) h.push.peer_flag[k] = h.sig.peer_flag[k];
    // default since round 2 (measured on 8 B200s, 16M cells: 0.929 against 0.941 ms per iteration; bit-identical at 2 / 4 / 8 ranks on
    // hardware): the last send-layer CTA of the update kernel raises the peers' flags itself.  AFX_HALO_EARLY_SIGNAL=0: separate launch.
    h.early_signal = true;
    if (const char* e = getenv("AFX_HALO_EARLY_SIGNAL")) h.early_signal = (e[0] != '0');
    h.p2p = true;  // every rank can: the single-launch path works for any front size (an all-front or empty send layer included)
    invalidate_graph();
}

// partitioned runs keep per-rank sums of squares: add them over the ranks, then take the root
void Solver::reduce_norms(double* v, int n)
{
    if (!halo || n <= 0) return;
    Halo& h = *halo;
    if (h.grp) {  // host values already: sum them in rank order
        h.grp->allreduce_sum(h.rank, v, n);
    } else {
        if (h.red.n < (size_t)n) h.red.alloc((size_t)n);  // persistent scratch: no allocation (a device-wide sync) per call
        CK(cudaMemcpyAsync(h.red.p, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        NK(NcclApi::get().AllReduce(h.red.p, h.red.p, (size_t)n, ncclDouble, ncclSum, h.comm, st));
        CK(cudaMemcpyAsync(v, h.red.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    for (int i = 0; i < n; ++i) v[i] = std::sqrt(v[i]);
}

// sum of a small device vector over the ranks, in place, ordered on the solver's stream
void Solver::comm_allreduce(double* dev, int n)
{
    if (!halo || n <= 0) return;
    Halo& h = *halo;
    if (!h.grp) { NK(NcclApi::get().AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, h.comm, st)); return; }
    std::vector<double> tmp((size_t)n);
    CK(cudaMemcpyAsync(tmp.data(), dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h.grp->allreduce_sum(h.rank, tmp.data(), n);
    CK(cudaMemcpyAsync(dev, tmp.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // tmp goes out of scope
}

void Solver::check_comm()
{
    if (*reinterpret_cast<volatile int*>(pipe_err_word()) != 0)
        throw CudaError("pipelined stage kernel: a dependency wait ran into its bound (internal error); the state of this solver is no longer valid");
    if (!halo || !halo->p2p) return;
    const int e = *reinterpret_cast<volatile int*>(comm_err_word());
    if (e == 0) return;
    const int slot = e - 1;
    const int peer = (slot >= 0 && slot < (int)halo->peers.size()) ? halo->peers[(size_t)slot].rank : -1;
    throw CommError("halo wait timed out: rank " + std::to_string(peer) + " did not deliver its send layer (peer dead, or on another halo mode); "
                    "the state of this solver is no longer valid");
}

double Solver::fetch_last_norm()
{
    ensure_halo();
    unsigned int* h_idx = reinterpret_cast<unsigned int*>(h_pinned + 8);
    CK(cudaMemcpyAsync(h_idx, counters.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const unsigned int k = *h_idx;
    if (k == 0) throw NumericError("no residual norm has been computed");
    CK(cudaMemcpyAsync(h_pinned, norms.p + ((k - 1) % NORM_RING), sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    norm_idx_host = k;
    check_comm();
    double v = h_pinned[0];
    reduce_norms(&v, 1);
    return v;
}

void Solver::run_explicit(double relax, int n_iter, double* norms_out)
{
    NvtxRange nvtx_("afx:run_explicit");
    use();
    if (!bcs_set) throw InvalidArg("set_bcs has not been called");
    if (n_iter <= 0) return;
    push_params(relax, n_iter == 1);  // qW is kept for the last iteration of the call only
    refresh_tile_k3a();
    h_pinned[35] = 1.0;               // constant source of the "keep qW" flag, never rewritten
    CK(cudaEventRecord(ev0, st));
    int done = 0;
    while (done < n_iter) {
        const int chunk = std::min<int>(n_iter - done, (int)NORM_RING / 2);
        // ring index before this chunk
        unsigned int* h_idx = reinterpret_cast<unsigned int*>(h_pinned + 8);
        CK(cudaMemcpyAsync(h_idx, counters.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const unsigned int k0 = *h_idx;
        if (graph_ok()) {  // the staged / lock-step in-process halo meets on the host: no capture
            if (!graph_exec) {
                cudaGraph_t g = nullptr;
                const int64_t l0 = launches;
                CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                explicit_iteration();
                CK(cudaStreamEndCapture(st, &g));
                graph_per_iter = launches - l0;
                launches = l0;
                CK(cudaGraphInstantiate(&graph_exec, g, 0));
                CK(cudaGraphDestroy(g));
            }
            const int64_t per_iter = graph_per_iter;
            for (int it = 0; it < chunk; ++it) {
                if (done + it == n_iter - 1 && keep_dev != 1.0) {
                    CK(cudaMemcpyAsync(prm.p + 2, h_pinned + 35, sizeof(double), cudaMemcpyHostToDevice, st));
                    keep_dev = 1.0;
                }
                CK(cudaGraphLaunch(graph_exec, st));
            }
            launches += (int64_t)per_iter * chunk;
        } else {
            for (int it = 0; it < chunk; ++it) {
                if (done + it == n_iter - 1 && keep_dev != 1.0) {
                    CK(cudaMemcpyAsync(prm.p + 2, h_pinned + 35, sizeof(double), cudaMemcpyHostToDevice, st));
                    keep_dev = 1.0;
                }
                explicit_iteration();
            }
        }
        CK(cudaGetLastError());
        if (norms_out) {
            // copy this chunk of the ring (it may wrap)
            std::vector<double> tmp(chunk);
            for (int it = 0; it < chunk;) {
                const unsigned int pos = (k0 + it) % NORM_RING;
                const int n = std::min<int>(chunk - it, (int)(NORM_RING - pos));
                CK(cudaMemcpyAsync(tmp.data() + it, norms.p + pos, n * sizeof(double), cudaMemcpyDeviceToHost, st));
                it += n;
            }
            CK(cudaStreamSynchronize(st));
            check_comm();
            reduce_norms(tmp.data(), chunk);
            std::copy(tmp.begin(), tmp.end(), norms_out + done);
        }
        done += chunk;
    }
    CK(cudaEventRecord(ev1, st));
    CK(cudaEventSynchronize(ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    last_ms = ms;
    jac_valid = false;
    check_comm();
}

// implicitSolver::fillRhoRHS, solver.h:1079-1152
double Solver::residual_rhs()
{
    NvtxRange nvtx_("afx:residual_rhs (fillRhoRHS)");
    use();
    if (!bcs_set) throw InvalidArg("set_bcs has not been called");
    push_params(relax_dev < 0 ? 1.0 : relax_dev);
    const bool grads = second_order || visc_not_inviscid;  // solver.h:1083
    const bool lim0 = lim0_fused();
    launch_dt_grad(grads, grads, lim0);
    if (second_order && !lim0) launch_limiter(q.p);
    launch_flux(q.p, false, d4{0, 0, 0, 0});
    launch_gather<1, 1>(q.p, nullptr, rhs.p, 0., false);
    CK(cudaGetLastError());
    const double v = fetch_last_norm();
    return v;
}

// solver::get_uniform_residual, solver.h:636-690 (accumulating from zero)
double Solver::uniform_residual()
{
    use();
    if (!bcs_set) throw InvalidArg("set_bcs has not been called");
    push_params(relax_dev < 0 ? 1.0 : relax_dev);
    afx_bvars far;
    boundary_variables(&far);
    double s4[4];
    conservative(far, s4);
    launch_flux(q.p, true, d4{s4[0], s4[1], s4[2], s4[3]});
    launch_gather<2, 1>(q.p, nullptr, qW.p, 0., false);
    CK(cudaGetLastError());
    return fetch_last_norm();
}

// implicitSolver::fillRhoLHS, solver.h:979-1071
void Solver::fill_jacobian()
{
    NvtxRange nvtx_("afx:fill_jacobian (fillRhoLHS)");
    use();
    if (!bcs_set) throw InvalidArg("set_bcs has not been called");
    push_params(relax_dev < 0 ? 1.0 : relax_dev);
    if (!J.p) { J.alloc((size_t)E * 16); D.alloc((size_t)NT * 16); }
    launch_dt_grad(visc_not_inviscid, visc_not_inviscid);  // solver.h:981-986
    kt->jacobian(viscous_type, dm, q.p, gx.p, gy.p, J.p, gas, st);
    kt->jac_diag(dm, J.p, dt.p, D.p, st);
    launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    jac_valid = true;
    precond_valid = false;
}

// implicitSolver::compute (solver.h:1160-1167): the reference factorises an ILUT here; we invert the diagonal
// blocks for the block-Jacobi smoother that preconditions GMRES.  0 ok / -1 singular block.
int Solver::compute_preconditioner()
{
    use();
    if (!jac_valid) throw InvalidArg("afx_rans_fill_jacobian has not been called for the current state");
    if (!Dinv.p) {
        invalidate_kry_graphs();  // they hold the addresses of these buffers
        Dinv.alloc((size_t)NT * 16); kry_flag.alloc(2);  // [0] singular diagonal block, [1] the Krylov iteration's stop flag
        kry_state.alloc((size_t)kt->gmres_state_doubles(gmres_restart)); kry_state.zero(st);
        kry_V.alloc((size_t)(gmres_restart + 1) * NT); kry_w.alloc(NT); kry_z.alloc(NT); kry_t.alloc(NT); kry_x.alloc(NT);
        kry_partial.alloc((size_t)(gmres_restart + 2) * 1024); kry_h.alloc(gmres_restart + 4);
        kry_counter.alloc(1); kry_counter.zero(st);
    }
    kry_flag.zero(st);
    kt->invert_blocks(NT, D.p, Dinv.p, kry_flag.p, st);
    ++launches;
    int* hflag = reinterpret_cast<int*>(h_pinned + 40);
    CK(cudaMemcpyAsync(hflag, kry_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    precond_valid = (*hflag == 0);
    return precond_valid ? 0 : -1;
}

// z = M^-1 r with M^-1 = `precond_sweeps` block-Jacobi sweeps on A z = r starting from z = 0
void Solver::apply_preconditioner(const d4* r, d4* z)
{
    d4* a = z; d4* b = kry_t.p;
    // an even number of swaps must leave the result in z
    const int sweeps = std::max(1, precond_sweeps);
    if ((sweeps - 1) % 2) std::swap(a, b);
    kt->jacobi_sweep(dm, J.p, D.p, Dinv.p, r, nullptr, a, 1, nullptr, st);
    ++launches;
    for (int k = 1; k < sweeps; ++k) {
        halo_refresh(a);
        kt->jacobi_sweep(dm, J.p, D.p, Dinv.p, r, a, b, 0, nullptr, st);
        ++launches;
        std::swap(a, b);
    }
}

void Solver::precondition_Ax(const d4* x, d4* r_buf, d4* z, const int* stop)
{
    d4* a = z; d4* b = kry_t.p;
    const int sweeps = std::max(1, precond_sweeps);
    if ((sweeps - 1) % 2) std::swap(a, b);
    kt->spmv_sweep0(dm, J.p, D.p, Dinv.p, x, r_buf, a, stop, st);  // the caller keeps the halo rows of x current
    ++launches;
    for (int k = 1; k < sweeps; ++k) {
        halo_refresh(a);
        kt->jacobi_sweep(dm, J.p, D.p, Dinv.p, r_buf, a, b, 0, stop, st);
        ++launches;
        std::swap(a, b);
    }
}

// Left-preconditioned restarted GMRES, zero initial guess, stop on ||M^-1 (b - A x)|| <= tol ||M^-1 b||
// (the criterion of Eigen::GMRES used at solver.h:886,906-910).  Arnoldi by classical Gram-Schmidt with all
// inner products of a step in one reduction.  The small dense part -- Hessenberg column, Givens rotations, residual
// estimate, back substitution -- runs on the device too (k_givens_step / k_gmres_solve_y): the host queues the Arnoldi steps
// in batches and looks at the iteration once per batch; the steps queued behind the one that reached the tolerance see
// the device-side stop flag and return at once.  (Round 1 synchronised with the host after every step.)
bool Solver::gmres(const d4* b, d4* x)
{
    NvtxRange nvtx_("afx:gmres");
    const int m = gmres_restart;
    const size_t stride = NT;
    const GmresView L{m};
    CK(cudaMemsetAsync(x, 0, (size_t)NT * sizeof(d4), st));
    last_linear_iters = 0;
    int batch = 6;
    if (const char* e = getenv("AFX_GMRES_BATCH")) batch = std::max(1, atoi(e));
    // r = M^-1 b
    d4* V0 = kry_V.p;
    const uint32_t nd = n_dot();
    int* stop = kry_flag.p + 1;
    double* hstat = h_pinned + 56;  // r0, err, k_done, fail
    auto read_status = [&] {
        CK(cudaMemcpyAsync(hstat, kry_state.p + L.r0(), 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    };
    apply_preconditioner(b, kry_w.p);
    kt->multi_dot1(nd, kry_w.p, stride, 1, kry_w.p, kry_partial.p, kry_h.p, kry_counter.p, nullptr, st); ++launches;
    allreduce_sum(kry_h.p, 1);
    kt->gmres_begin(m, kry_state.p, kry_h.p, 1, stop, st); ++launches;
    read_status();
    const double r0 = hstat[0];
    if (!(r0 == r0)) return false;
    if (r0 == 0) return true;
    while (last_linear_iters < gmres_max_iter) {
        // V0 = r / beta   (kry_h[0] holds beta^2)
        kt->scale_from(NT, kry_w.p, kry_h.p, 1, 1, V0, nullptr, st); ++launches;
        halo_refresh(V0);
        int k = 0;        // Arnoldi steps queued in this cycle
        int k_done = 0;   // ... and known to be complete
        bool done = false;
        while (k < m && last_linear_iters + (k - k_done) < gmres_max_iter && !done) {
            const int nb = std::min({batch, m - k, gmres_max_iter - last_linear_iters - (k - k_done)});
            for (int j = 0; j < nb; ++j, ++k) {
                auto arnoldi_step = [&] {
                    // w = M^-1 A v_k   (product + first sweep, the other sweeps, dots, update + norm, rotation, scaling)
                    precondition_Ax(kry_V.p + (size_t)k * stride, kry_z.p, kry_w.p, stop);
                    // h = V^T w ; w -= V h ; ||w||^2
                    kt->multi_dot1(nd, kry_V.p, stride, k + 1, kry_w.p, kry_partial.p, kry_h.p, kry_counter.p, stop, st); ++launches;
                    allreduce_sum(kry_h.p, k + 1);
                    if (!halo && fuse_givens) {  // update + norm + rotation in one launch
                        kt->axpy_norm_givens(nd, kry_V.p, stride, k + 1, kry_h.p, -1.0, kry_w.p, kry_partial.p, kry_h.p, kry_counter.p, m, kry_state.p, gmres_tol,
                                             stop, st); ++launches;
                    } else {
                        kt->axpy_norm(nd, kry_V.p, stride, k + 1, kry_h.p, -1.0, kry_w.p, kry_partial.p, kry_h.p + (k + 1), kry_counter.p, stop, st); ++launches;
                        allreduce_sum(kry_h.p + (k + 1), 1);
                        kt->givens_step(m, kry_state.p, kry_h.p, k, gmres_tol, stop, st); ++launches;
                    }
                    if (k + 1 < m) {  // v_{k+1} = w / ||w||   (kry_h[k+1] holds ||w||^2); skipped on the device once the iteration has stopped
                        kt->scale_from(NT, kry_w.p, kry_h.p + (k + 1), 1, 1, kry_V.p + (size_t)(k + 1) * stride, stop, st); ++launches;
                        halo_refresh(kry_V.p + (size_t)(k + 1) * stride);
                    }
                };
                if (use_kry_graph && use_graph && !halo) {  // the same kernels with the same arguments, captured once per step index
                    if ((int)kry_graph.size() != m) { invalidate_kry_graphs(); kry_graph.assign(m, nullptr); kry_graph_launches.assign(m, 0); }
                    if (!kry_graph[k]) {
                        cudaGraph_t g = nullptr;
                        const int64_t l0 = launches;
                        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                        arnoldi_step();
                        CK(cudaStreamEndCapture(st, &g));
                        kry_graph_launches[k] = (int)(launches - l0);
                        launches = l0;
                        CK(cudaGraphInstantiate(&kry_graph[k], g, 0));
                        CK(cudaGraphDestroy(g));
                    }
                    CK(cudaGraphLaunch(kry_graph[k], st));
                    launches += kry_graph_launches[k];
                } else {
                    arnoldi_step();
                }
            }
            read_status();
            if (hstat[3] != 0.0) return false;  // NaN in the recurrence
            const int kd = (int)hstat[2];
            last_linear_iters += kd - k_done;
            k_done = kd;
            if (kd < k || hstat[1] < gmres_tol) done = true;  // the device stopped inside the batch (tolerance or breakdown)
        }
        // x += V y with H y = g (back substitution on the device, k_done columns)
        if (k_done > 0) {
            kt->gmres_solve_y(m, kry_state.p, kry_h.p, st); ++launches;
            kt->multi_axpy(NT, kry_V.p, stride, k_done, kry_h.p, 1.0, x, st); ++launches;
        }
        if (done) return true;
        // restart: r = M^-1 (b - A x)
        halo_refresh(x);
        kt->spmv(dm, J.p, D.p, x, kry_z.p, st); ++launches;
        kt->sub(NT, b, kry_z.p, kry_z.p, st); ++launches;
        apply_preconditioner(kry_z.p, kry_w.p);
        kt->multi_dot1(nd, kry_w.p, stride, 1, kry_w.p, kry_partial.p, kry_h.p, kry_counter.p, nullptr, st); ++launches;
        allreduce_sum(kry_h.p, 1);
        kt->gmres_begin(m, kry_state.p, kry_h.p, 0, stop, st); ++launches;
        read_status();
        if (!(hstat[1] == hstat[1])) return false;
        if (hstat[1] < gmres_tol) return true;
    }
    return false;  // Eigen reports NoConvergence -> the reference returns -1 (solver.h:1184)
}

// implicitSolver::solve, solver.h:1170-1213, with the frozen Jacobian of the last fill_jacobian
double Solver::step_implicit(double relax, double tol, int rhs_iterations)
{
    use();
    if (!jac_valid) throw InvalidArg("afx_rans_fill_jacobian has not been called for the current state");
    if (!precond_valid && compute_preconditioner() != 0) return -1;
    double err = residual_rhs();
    if (!(err == err)) return -1;
    if (err < tol) return err;
    const double err_ini = err;
    if (!gmres(rhs.p, kry_x.p)) return -1;
    kt->axpy_state(n_dot(), relax, kry_x.p, q.p, st); ++launches;
    halo_refresh(q.p);
    for (int i = 0; i < rhs_iterations; ++i) {
        err = residual_rhs();
        if (!(err == err)) return -1;
        if (err < tol) return err;
        if (err > 10 * err_ini) return err;
        if (!gmres(rhs.p, kry_x.p)) return -1;
        kt->axpy_state(n_dot(), relax, kry_x.p, q.p, st); ++launches;
        halo_refresh(q.p);
    }
    err = residual_rhs();
    sync_ghost_rows();
    CK(cudaStreamSynchronize(st));
    return err;
}

void Solver::to_ref_order4(const d4* dev_new, double* host_out)
{
    NvtxRange nvtx_("afx:get (permute + D2H)");
    // stage[old] = dev_new[old2new[old]]
    kt->permute4(dev_new, stage.p, perm_c_old2new.p, NT, st);
    ++launches;
    CK(cudaMemcpyAsync(host_out, stage.p, (size_t)NT * sizeof(d4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
}

void Solver::from_ref_order4(const double* host_in, d4* dev_new)
{
    NvtxRange nvtx_("afx:set (H2D + permute)");
    CK(cudaMemcpyAsync(stage.p, host_in, (size_t)NT * sizeof(d4), cudaMemcpyHostToDevice, st));
    // dev_new[new] = stage[new2old[new]]
    kt->permute4(stage.p, dev_new, perm_c_new2old.p, NT, st);
    ++launches;
    CK(cudaStreamSynchronize(st));
}

// the stage buffers carry the same ghost rows as q (far-field ghosts never change inside an iteration)
void Solver::sync_ghost_rows()
{
    if (!G) return;
    CK(cudaMemcpyAsync(qkA.p + N, q.p + N, (size_t)G * sizeof(d4), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(qkB.p + N, q.p + N, (size_t)G * sizeof(d4), cudaMemcpyDeviceToDevice, st));
}

}  // namespace afx

// ===========================================================================
// C ABI
// ===========================================================================
struct afx_rans {
    afx::Solver s;
};

namespace {

template <class F>
int guard(F&& f)
{
    try {
        f();
        return AFX_OK;
    } catch (const afx::CudaError& e) { afx::set_error(e.what()); return AFX_ERR_CUDA; }
    catch (const afx::InvalidArg& e) { afx::set_error(e.what()); return AFX_ERR_INVALID; }
    catch (const afx::NumericError& e) { afx::set_error(e.what()); return AFX_ERR_NUMERIC; }
    catch (const afx::CommError& e) { afx::set_error(e.what()); return AFX_ERR_COMM; }
    catch (const std::exception& e) { afx::set_error(e.what()); return AFX_ERR_INVALID; }
    catch (...) { afx::set_error("unknown error"); return AFX_ERR_INVALID; }
}
// every entry point that takes a solver handle refuses a null one instead of dereferencing it
int null_handle() { afx::set_error("null solver handle"); return AFX_ERR_INVALID; }

}  // namespace

extern "C" {

const char* afx_last_error(void) { return afx::g_last_error.c_str(); }
const char* afx_version(void) { return "aeroflex_rans_b200 0.1 (sm_100a)"; }

int afx_device_count(void)
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { afx::set_error(cudaGetErrorString(e)); return AFX_ERR_CUDA; }
    return n;
}

void* afx_pinned_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { afx::set_error("cudaMallocHost failed"); return nullptr; }
    return p;
}
void afx_pinned_free(void* p) { if (p) cudaFreeHost(p); }

int afx_rans_create(afx_rans** out, const afx_mesh_desc* mesh, const afx_gas* gas, int viscosity_model, int device)
{
    if (!out || !mesh || !gas) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_rans* h = nullptr;
    const int rc = guard([&] {
        if (viscosity_model < 0 || viscosity_model > 2) throw afx::InvalidArg("viscosity model must be 0, 1 or 2");
        h = new afx_rans;
        h->s.create(*mesh, *gas, viscosity_model, device);
    });
    if (rc) { delete h; return rc; }
    *out = h;
    return AFX_OK;
}

int afx_nccl_unique_id(char out[128])
{
    if (!out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    return guard([&] {
        ncclUniqueId id;
        NK(afx::NcclApi::get().GetUniqueId(&id));
        std::memcpy(out, &id, 128);
    });
}

int afx_rans_p2p_export(afx_rans* s, void* blob, size_t* size)
{
    if (!s) return null_handle();
    return guard([&] { const size_t n = s->s.p2p_export(blob); if (size) *size = n; });
}

int afx_rans_p2p_connect(afx_rans* s, const void* blobs, size_t blob_size, int nranks)
{
    if (!s) return null_handle();
    return guard([&] { s->s.p2p_connect(blobs, blob_size, nranks); });
}

int afx_rans_halo_mode(afx_rans* s) { return !s ? null_handle() : (!s->s.halo ? 0 : (s->s.halo->p2p ? 2 : 1)); }

int afx_rans_create_partitioned(afx_rans** out, const afx_partition* part, const afx_gas* gas, int viscosity_model, int device,
                                const char nccl_id[128])
{
    if (!out || !part || !gas || !nccl_id) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_rans* h = nullptr;
    const int rc = guard([&] {
        if (viscosity_model < 0 || viscosity_model > 2) throw afx::InvalidArg("viscosity model must be 0, 1 or 2");
        h = new afx_rans;
        const afx_mesh_desc d = part->p.desc();
        h->s.create(d, *gas, viscosity_model, device, &part->p);
        h->s.init_halo(part->p, nccl_id);
    });
    if (rc) { delete h; return rc; }
    *out = h;
    return AFX_OK;
}

struct afx_group_impl;
int afx_group_create(afx_group** out, int nranks)
{
    if (!out || nranks < 1) { afx::set_error("bad argument"); return AFX_ERR_INVALID; }
    *out = reinterpret_cast<afx_group*>(new std::shared_ptr<afx::LocalGroup>(std::make_shared<afx::LocalGroup>(nranks)));
    return AFX_OK;
}
void afx_group_free(afx_group* g) { delete reinterpret_cast<std::shared_ptr<afx::LocalGroup>*>(g); }
void afx_group_abort(afx_group* g) { if (g) (*reinterpret_cast<std::shared_ptr<afx::LocalGroup>*>(g))->abort_group(); }

int afx_rans_create_partitioned_group(afx_rans** out, const afx_partition* part, const afx_gas* gas, int viscosity_model, int device,
                                      afx_group* group)
{
    if (!out || !part || !gas || !group) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_rans* h = nullptr;
    const int rc = guard([&] {
        if (viscosity_model < 0 || viscosity_model > 2) throw afx::InvalidArg("viscosity model must be 0, 1 or 2");
        h = new afx_rans;
        const afx_mesh_desc d = part->p.desc();
        h->s.create(d, *gas, viscosity_model, device, &part->p);
        h->s.init_halo(part->p, nullptr, *reinterpret_cast<std::shared_ptr<afx::LocalGroup>*>(group));
    });
    if (rc) { delete h; return rc; }
    *out = h;
    return AFX_OK;
}

void afx_rans_destroy(afx_rans* s) { delete s; }

int afx_rans_set_bcs(afx_rans* s, int n_patch, const uint8_t* patch_kind, const afx_bvars* patch_vars)
{
    if (!s) return null_handle();
    return guard([&] { s->s.set_bcs(n_patch, patch_kind, patch_vars); });
}

int afx_rans_set_options(afx_rans* s, int second_order, int gradient_scheme, double limiter_k)
{
    if (!s) return null_handle();
    return guard([&] { s->s.set_options(second_order, gradient_scheme, limiter_k); });
}

int afx_rans_set_limiter(afx_rans* s, int limiter)
{
    if (!s) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    return guard([&] { s->s.set_limiter(limiter); });
}

int afx_rans_set_math_mode(afx_rans* s, int mode)
{
    if (!s) return null_handle();
    return guard([&] { s->s.set_math_mode(mode); });
}

static void tiling_plan_impl(const afx_mesh_desc& mesh, const afx::Partition* part, uint32_t tile_cells, const uint32_t* limits, uint32_t* n_tiles,
                             uint32_t* per_tile, uint32_t cap, uint64_t* smem_bytes)
{
    if (!n_tiles) throw afx::InvalidArg("null argument");
    afx::Solver S;
    afx::Solver::DryRun dry{tile_cells, afx::TileLimits{}, {}, {}};
    if (limits) dry.limits = afx::TileLimits{limits[0], limits[1], limits[2], limits[3]};
    const afx_gas g{1.4, 1., 0., 1., 1.};
    S.create(mesh, g, 0, 0, part, &dry);
    if (!dry.check.empty()) throw afx::InvalidArg("tile plan inconsistent: " + dry.check);
    const auto& P = dry.plan;
    *n_tiles = (uint32_t)P.head.size();
    if (smem_bytes) *smem_bytes = afx::stage_smem_layout(P.max_loc, P.max_n1, P.max_nf, P.max_nc, P.max_halo).total;
    if (per_tile)
        for (uint32_t t = 0; t < P.head.size() && t < cap; ++t) {
            per_tile[4 * t] = P.head[t].nc; per_tile[4 * t + 1] = P.head[t].h1; per_tile[4 * t + 2] = P.head[t].h2; per_tile[4 * t + 3] = P.head[t].nf;
        }
}

int afx_tiling_plan(const afx_mesh_desc* mesh, uint32_t tile_cells, const uint32_t* limits, uint32_t* n_tiles, uint32_t* per_tile, uint32_t cap,
                    uint64_t* smem_bytes)
{
    return guard([&] {
        if (!mesh) throw afx::InvalidArg("null argument");
        tiling_plan_impl(*mesh, nullptr, tile_cells, limits, n_tiles, per_tile, cap, smem_bytes);
    });
}

int afx_tiling_plan_partition(const afx_partition* part, uint32_t tile_cells, const uint32_t* limits, uint32_t* n_tiles, uint32_t* per_tile,
                              uint32_t cap, uint64_t* smem_bytes)
{
    return guard([&] {
        if (!part) throw afx::InvalidArg("null argument");
        const afx_mesh_desc d = part->p.desc();
        tiling_plan_impl(d, &part->p, tile_cells, limits, n_tiles, per_tile, cap, smem_bytes);
    });
}

int afx_rans_set_fused(afx_rans* s, int on)
{
    if (!s) return null_handle();
    return guard([&] { if (s->s.use_fused != (on != 0)) { s->s.use_fused = (on != 0); s->s.invalidate_graph(); } });
}

int afx_rans_tile_info(afx_rans* s, uint64_t out[8])
{
    if (!s) return null_handle();
    const auto& S = s->s;
    out[0] = S.fused_stage() ? 1 : 0; out[1] = S.n_tiles; out[2] = S.tile_cells; out[3] = S.stage_smem; out[4] = (uint64_t)S.stage_ctas_per_sm;
    out[5] = S.tt.max_loc; out[6] = S.tt.max_nf; out[7] = S.tile_local_cells;
    return AFX_OK;
}

int afx_rans_set_pipelined(afx_rans* s, int on)
{
    return guard([&] {
        if (!s) throw afx::InvalidArg("null argument");
        if (s->s.use_pipe != (on != 0)) { s->s.use_pipe = (on != 0); s->s.invalidate_graph(); }
    });
}

int afx_rans_pipe_info(afx_rans* s, uint64_t out[8])
{
    if (!s || !out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    const auto& S = s->s;
    const afx::PipeTab& t = S.pipe_tab[1];
    out[0] = S.pipe_stage() ? 1 : 0; out[1] = 1ull << t.shift; out[2] = t.n_chunks; out[3] = t.n_items; out[4] = t.n_far_faces; out[5] = t.n_far_cells;
    out[6] = ((uint64_t)t.lagF << 32) | t.lagU; out[7] = S.pipe_grid;
    return AFX_OK;
}

int afx_rans_get_math_mode(afx_rans* s) { return !s ? null_handle() : (s->s.kt == &afx::strict::table() ? AFX_MATH_STRICT : AFX_MATH_FAST); }

int afx_rans_set_cfl(afx_rans* s, double cfl)
{
    if (!s) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    s->s.cfl = cfl;
    return AFX_OK;
}

int afx_rans_init(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        afx_bvars v;
        S.boundary_variables(&v);
        double q4[4];
        S.conservative(v, q4);
        S.kt->fill_cells(S.q.p, S.N, afx::d4{q4[0], q4[1], q4[2], q4[3]}, S.st);
        ++S.launches;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(S.st));
        S.jac_valid = false;
    });
}

int afx_rans_refill_bcs(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.bcs_set) throw afx::InvalidArg("set_bcs has not been called");
        if (S.G) {
            S.kt->ghost_fill(S.q.p, S.bghost.p, S.bowner.p, S.bstate.p, S.G, 0, S.st);
            ++S.launches;
        }
        S.sync_ghost_rows();
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(S.st));
        S.jac_valid = false;
    });
}

int afx_rans_bcs_from_internal(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (S.G) {
            S.kt->ghost_fill(S.q.p, S.bghost.p, S.bowner.p, S.bstate.p, S.G, 1, S.st);
            ++S.launches;
        }
        S.sync_ghost_rows();
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(S.st));
        S.jac_valid = false;
    });
}

int afx_rans_set_q(afx_rans* s, const double* q)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (S.halo) {  // q is the GLOBAL vector: pick this rank's cells (owned, rings, boundary ghosts)
            double* local = reinterpret_cast<double*>(S.h_stage);
#pragma omp parallel for
            for (int64_t l = 0; l < (int64_t)S.NT; ++l) std::memcpy(local + 4 * (size_t)l, q + 4 * (size_t)S.cell_l2g[l], 32);
            q = local;
        }
        S.from_ref_order4(q, S.q.p);
        S.sync_ghost_rows();
        CK(cudaStreamSynchronize(S.st));
        S.jac_valid = false;
    });
}

int afx_rans_get_q(afx_rans* s, double* q)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.halo) { S.to_ref_order4(S.q.p, q); return; }
        // partitioned: fill this rank's entries of the GLOBAL vector, leave the rest untouched
        double* local = reinterpret_cast<double*>(S.h_stage);
        S.to_ref_order4(S.q.p, local);
#pragma omp parallel for
        for (int64_t l = 0; l < (int64_t)S.NT; ++l) std::memcpy(q + 4 * (size_t)S.cell_l2g[l], local + 4 * (size_t)l, 32);
    });
}

int afx_rans_set_q_local(afx_rans* s, const double* q_local)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        S.from_ref_order4(q_local, S.q.p);
        S.sync_ghost_rows();
        CK(cudaStreamSynchronize(S.st));
        S.jac_valid = false;
    });
}

int afx_rans_get_q_local(afx_rans* s, double* q_local)
{
    if (!s) return null_handle();
    return guard([&] { s->s.use(); s->s.to_ref_order4(s->s.q.p, q_local); });
}

int afx_rans_get_field(afx_rans* s, int field, double* out)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        switch (field) {
            case AFX_F_Q: S.to_ref_order4(S.q.p, out); break;
            case AFX_F_QW: S.to_ref_order4(S.qW.p, out); break;
            case AFX_F_GX: S.to_ref_order4(S.gx.p, out); break;
            case AFX_F_GY: S.to_ref_order4(S.gy.p, out); break;
            case AFX_F_LIMITERS: S.to_ref_order4(S.lim.p, out); break;
            case AFX_F_RHS: S.to_ref_order4(S.rhs.p, out); break;
            case AFX_F_DT: {
                S.kt->permute1(S.dt.p, S.dt_ref.p, S.perm_c_old2new.p, S.NT, S.N, S.st);
                ++S.launches;
                CK(cudaMemcpyAsync(out, S.dt_ref.p, (size_t)S.NT * sizeof(double), cudaMemcpyDeviceToHost, S.st));
                CK(cudaStreamSynchronize(S.st));
                break;
            }
            default: throw afx::InvalidArg("unknown field id");
        }
    });
}

int afx_rans_boundary_variables(afx_rans* s, afx_bvars* out) { return (!s || !out) ? null_handle() : (s->s.boundary_variables(out) ? 1 : 0); }

int afx_rans_uniform_residual(afx_rans* s, double* norm)
{
    if (!s) return null_handle();
    return guard([&] { const double v = s->s.uniform_residual(); if (norm) *norm = v; });
}

int afx_rans_step_explicit(afx_rans* s, double relaxation, double* norm)
{
    if (!s) return null_handle();
    return guard([&] {
        double v = 0;
        s->s.run_explicit(relaxation, 1, &v);
        if (norm) *norm = v;
        s->s.check_norm(v);
    });
}

int afx_rans_run_explicit(afx_rans* s, double relaxation, int n_iter, double* norms)
{
    if (!s) return null_handle();
    return guard([&] {
        if (n_iter <= 0) return;
        std::vector<double> tmp;
        double* dst = norms;
        if (!dst) { tmp.resize((size_t)n_iter); dst = tmp.data(); }
        s->s.run_explicit(relaxation, n_iter, dst);
        s->s.check_norm(dst[n_iter - 1]);
    });
}

int afx_rans_phase_dt_gradients(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.bcs_set) throw afx::InvalidArg("set_bcs has not been called");
        S.push_params(S.relax_dev < 0 ? 1.0 : S.relax_dev);
        S.launch_dt_grad(true, true);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(S.st));
    });
}

int afx_rans_phase_limiters(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        S.launch_limiter(S.q.p);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(S.st));
    });
}

int afx_rans_phase_residual(afx_rans* s, double* norm)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.bcs_set) throw afx::InvalidArg("set_bcs has not been called");
        S.push_params(0.0);  // relax = 0: the stage "update" below leaves the state untouched
        S.launch_flux(S.q.p, false, afx::d4{0, 0, 0, 0});
        S.launch_gather<0, 1>(S.q.p, S.stage.p, S.qW.p, 0., false);
        CK(cudaGetLastError());
        const double v = S.fetch_last_norm();
        if (norm) *norm = v;
    });
}

int afx_rans_residual(afx_rans* s, double* norm)
{
    if (!s) return null_handle();
    return guard([&] { const double v = s->s.residual_rhs(); if (norm) *norm = v; });
}

int afx_rans_fill_jacobian(afx_rans* s)
{
    if (!s) return null_handle();
    return guard([&] { s->s.fill_jacobian(); });
}

int afx_rans_get_jacobian_blocks(afx_rans* s, double* diag, double* off01, double* off10)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.jac_valid) throw afx::InvalidArg("afx_rans_fill_jacobian has not been called for the current state");
        std::vector<double> hJ((size_t)S.E * 64), hD((size_t)S.NT * 16);
        CK(cudaMemcpy(hJ.data(), S.J.p, hJ.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hD.data(), S.D.p, hD.size() * sizeof(double), cudaMemcpyDeviceToHost));
        if (diag)
            for (uint32_t n = 0; n < S.NT; ++n) std::memcpy(diag + 16 * (size_t)S.c_new2old[n], hD.data() + 16 * (size_t)n, 16 * sizeof(double));
        for (uint32_t n = 0; n < S.E; ++n) {
            const size_t e = S.f_new2old[n];
            if (off01) std::memcpy(off01 + 16 * e, hJ.data() + 64 * (size_t)n + 16, 16 * sizeof(double));
            if (off10) {
                const bool two_sided = S.h_fcells[2 * (size_t)n + 1] < S.N;
                // one-sided faces never touch row c1 (solver.h:1051-1054)
                for (int k = 0; k < 16; ++k) off10[16 * e + k] = two_sided ? hJ[64 * (size_t)n + 32 + k] : 0.0;
            }
        }
    });
}

int afx_rans_compute(afx_rans* s)
{
    if (!s) return null_handle();
    int r = 0;
    const int rc = guard([&] { r = s->s.compute_preconditioner(); });
    return rc ? rc : (r == 0 ? AFX_OK : AFX_ERR_NUMERIC);
}

int afx_rans_step_implicit(afx_rans* s, double relaxation, double tol, int rhs_iterations, double* norm)
{
    if (!s) return null_handle();
    double v = -1;
    const int rc = guard([&] {
        v = s->s.step_implicit(relaxation, tol, rhs_iterations);
    });
    if (norm) *norm = v;
    if (rc) return rc;
    return v < 0 ? AFX_ERR_NUMERIC : AFX_OK;
}

int afx_rans_set_linear_solver(afx_rans* s, int restart, int max_iterations, double tolerance, int precond_sweeps)
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        if (restart < 1 || max_iterations < 1 || tolerance <= 0 || precond_sweeps < 1) throw afx::InvalidArg("bad linear solver settings");
        if (restart != S.gmres_restart) { S.Dinv.free(); S.precond_valid = false; }  // buffers are sized by the restart length
        if (restart != S.gmres_restart || tolerance != S.gmres_tol || precond_sweeps != S.precond_sweeps) S.invalidate_kry_graphs();  // baked into the nodes
        S.gmres_restart = restart; S.gmres_max_iter = max_iterations; S.gmres_tol = tolerance; S.precond_sweeps = precond_sweeps;
    });
}

int afx_rans_last_linear_iterations(afx_rans* s) { return s ? s->s.last_linear_iters : null_handle(); }

int afx_rans_wall_forces(afx_rans* s, int patch, double out[3])
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        afx_bvars far;
        S.boundary_variables(&far);
        // chord extent and moment centre of the patch, post.h:314-338
        double xmin = 0., xmax = 0., ym = 0.;
        uint32_t n_added = 0;
        for (uint32_t b = 0; b < S.G; ++b) {
            if (S.h_bnd_patch[b] != patch) continue;
            if (!n_added) { xmin = xmax = S.h_bcx[b]; ym = S.h_bcy[b]; }
            else { xmin = std::min(xmin, S.h_bcx[b]); xmax = std::max(xmax, S.h_bcx[b]); ym += S.h_bcy[b]; }
            ++n_added;
        }
        if (S.halo) {  // extents of the whole patch, not of this rank's piece
            const auto& H = *S.halo;
            if (patch < 0 || patch >= (int)H.patch_count.size()) throw afx::InvalidArg("unknown patch");
            xmin = H.patch_xmin[patch]; xmax = H.patch_xmax[patch]; ym = H.patch_ysum[patch]; n_added = H.patch_count[patch];
        }
        if (!n_added) throw afx::InvalidArg("patch has no boundary edges");
        ym /= (double)n_added;
        const double xm = (xmax - xmin) * 0.25 + xmin;
        S.kt->wall_forces(afx::WallArgs{S.bface.p, S.bpatch.p, S.G, patch, S.bcx.p, S.bcy.p, S.gas.gamma, far.p, far.mach, xmin, xmax, xm, ym,
                                        S.scratch.p, nullptr}, S.dm, S.q.p, S.st);
        ++S.launches;
        CK(cudaGetLastError());
        S.comm_allreduce(S.scratch.p, 3);
        CK(cudaMemcpyAsync(S.h_pinned + 16, S.scratch.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, S.st));
        CK(cudaStreamSynchronize(S.st));
        const double fx = S.h_pinned[16], fy = S.h_pinned[17], cm = S.h_pinned[18], aoa = far.angle;
        out[1] = fx * std::cos(aoa) + fy * std::sin(aoa);   // cd, post.h:383
        out[0] = -fx * std::sin(aoa) + fy * std::cos(aoa);  // cl, post.h:384
        out[2] = cm;
    });
}

int afx_rans_wall_cp(afx_rans* s, int patch, double* cp)
{
    if (!s) return null_handle();
    int count = 0;
    const int rc = guard([&] {
        auto& S = s->s;
        S.use();
        for (uint32_t b = 0; b < S.G; ++b) count += (S.h_bnd_patch[b] == patch);
        if (!cp || !count) return;
        afx_bvars far;
        S.boundary_variables(&far);
        afx::DBuf<double> d_cp;
        d_cp.alloc(S.G);
        d_cp.zero(S.st);
        S.kt->wall_forces(afx::WallArgs{S.bface.p, S.bpatch.p, S.G, patch, S.bcx.p, S.bcy.p, S.gas.gamma, far.p, far.mach, 0., 1., 0., 0.,
                                        S.scratch.p, d_cp.p}, S.dm, S.q.p, S.st);
        ++S.launches;
        std::vector<double> h(S.G);
        CK(cudaMemcpyAsync(h.data(), d_cp.p, S.G * sizeof(double), cudaMemcpyDeviceToHost, S.st));
        CK(cudaStreamSynchronize(S.st));
        int k = 0;
        for (uint32_t b = 0; b < S.G; ++b) if (S.h_bnd_patch[b] == patch) cp[k++] = h[b];
    });
    return rc ? rc : count;
}

struct afx_prolongation {
    afx_rans* coarse = nullptr;
    afx_rans* fine = nullptr;
    afx::DBuf<uint32_t> row_begin, col;
    afx::DBuf<double> w;
    cudaEvent_t ready = nullptr;
    ~afx_prolongation() { if (ready) cudaEventDestroy(ready); }
};

int afx_prolongation_create(afx_prolongation** out, afx_rans* coarse, afx_rans* fine, const uint32_t* row_begin, const uint32_t* col, const double* w)
{
    if (!out || !coarse || !fine || !row_begin || (!col && row_begin[fine->s.NT]) || (!w && row_begin[fine->s.NT])) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_prolongation* p = nullptr;
    const int rc = guard([&] {
        auto& F = fine->s; auto& Cs = coarse->s;
        if (F.device != Cs.device) throw afx::InvalidArg("both mesh levels must live on the same device");
        if (F.halo || Cs.halo) throw afx::InvalidArg("the device-resident prolongation is for un-partitioned levels");
        F.use();
        const size_t nnz = row_begin[F.NT];
        for (uint32_t r = 0; r < F.NT; ++r) if (row_begin[r] > row_begin[r + 1]) throw afx::InvalidArg("row_begin is not ascending");
        for (size_t k = 0; k < nnz; ++k) if (col[k] >= Cs.NT) throw afx::InvalidArg("prolongation column outside the coarse mesh");
        p = new afx_prolongation;
        p->coarse = coarse; p->fine = fine;
        p->row_begin.upload(std::vector<uint32_t>(row_begin, row_begin + F.NT + 1), F.st);
        p->col.upload(nnz ? std::vector<uint32_t>(col, col + nnz) : std::vector<uint32_t>(1, 0u), F.st);
        p->w.upload(nnz ? std::vector<double>(w, w + nnz) : std::vector<double>(1, 0.0), F.st);
        CK(cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming));
        CK(cudaStreamSynchronize(F.st));
    });
    if (rc) { delete p; return rc; }
    *out = p;
    return AFX_OK;
}
void afx_prolongation_free(afx_prolongation* p) { delete p; }

int afx_prolongation_apply(afx_prolongation* p)
{
    if (!p) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    return guard([&] {
        auto& F = p->fine->s; auto& Cs = p->coarse->s;
        F.use();
        CK(cudaEventRecord(p->ready, Cs.st));          // the coarse state as of now (its stream's order)
        CK(cudaStreamWaitEvent(F.st, p->ready, 0));
        // always the strict kernels: the reference's sums in the reference's order, whatever mode the stage kernels run in
        afx::strict::table().prolongate(F.NT, p->row_begin.p, p->col.p, p->w.p, F.perm_c_new2old.p, Cs.perm_c_old2new.p, Cs.q.p, F.q.p, F.st);
        ++F.launches;
        F.sync_ghost_rows();
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(F.st));
        F.jac_valid = false;
    });
}

namespace {
// multigrid<T>::run_solver (multigrid.h:182-293) on one level; returns AFX_OK / AFX_ERR_NUMERIC (the reference's "return 1") / an error
int run_level(afx_rans* s, const afx_sweep_settings* st, int* iters_out, double* err_out)
{
    int rc = AFX_OK;
    double err_0 = 0, err = 0, cfl = st->start_cfl;
    if ((rc = afx_rans_uniform_residual(s, &err_0)) != AFX_OK) return rc;
    int i = 0;
    do {
        afx_rans_set_cfl(s, cfl);
        if (st->implicit) {
            if ((rc = afx_rans_fill_jacobian(s)) != AFX_OK) return rc;
            rc = afx_rans_compute(s);
            if (rc == AFX_OK) rc = afx_rans_step_implicit(s, st->relaxation, err_0 * st->tolerance, st->rhs_iterations, &err);
        } else {
            rc = afx_rans_step_explicit(s, st->relaxation, &err);
        }
        // multigrid.h:227,285: "if (err < 0) return 1" comes before the iteration is counted (iters++), with err = -1 / err_0
        if (rc == AFX_ERR_NUMERIC) { if (iters_out) *iters_out = i; if (err_out) *err_out = -1.0 / err_0; return rc; }
        if (rc != AFX_OK) return rc;
        if (i == 0 && err > 2 * err_0) err_0 = err;
        err /= err_0;
        if (st->implicit) cfl = std::min(st->start_cfl + (i + 1) * st->slope_cfl, st->max_cfl);
        ++i;
    } while (err > st->tolerance && i < st->max_iterations);
    if (iters_out) *iters_out = i;
    if (err_out) *err_out = err;
    return AFX_OK;
}
}  // namespace

// Rans::run_airfoil (rans.h:78-106) over multigrid<T>::run(false) (multigrid.h:295-363): per angle the whole full-multigrid
// start-up from the coarsest level, every level warm-started -- level 0 from its own state of the previous angle, level i > 0
// from the prolongation of level i-1.  The states never leave the device.
int afx_rans_sweep_fmg(afx_rans* const* levels, afx_prolongation* const* prolongations, int n_levels, const afx_sweep_settings* st,
                       int farfield_patch, int wall_patch, const double* alphas_deg, int n_alpha, int reinit, double* cl, double* cd, double* cm,
                       int* iterations, double* residual)
{
    if (!levels || n_levels < 1 || !st || (n_alpha > 0 && !alphas_deg) || (n_levels > 1 && !prolongations)) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    for (int l = 0; l < n_levels; ++l) {
        if (!levels[l]) { afx::set_error("null level"); return AFX_ERR_INVALID; }
        auto& S = levels[l]->s;
        if (!S.bcs_set) { afx::set_error("set_bcs has not been called"); return AFX_ERR_INVALID; }
        if (farfield_patch < 0 || farfield_patch >= (int)S.patch_kinds.size() || S.patch_kinds[(size_t)farfield_patch] != AFX_BC_FARFIELD) {
            afx::set_error("farfield_patch is not a far-field patch of the last set_bcs");
            return AFX_ERR_INVALID;
        }
        if (l > 0 && (!prolongations[l - 1] || prolongations[l - 1]->coarse != levels[l - 1] || prolongations[l - 1]->fine != levels[l])) {
            afx::set_error("prolongations[l-1] must map levels[l-1] to levels[l]");
            return AFX_ERR_INVALID;
        }
    }
    if (n_alpha <= 0) return AFX_OK;
    int rc = AFX_OK;
    bool any_failed = false;
    auto apply_alpha = [&](afx_rans* s, double alpha_deg) {
        auto& S = s->s;
        std::vector<uint8_t> kinds(S.patch_kinds);
        std::vector<afx_bvars> vars(S.patch_vars);
        vars[(size_t)farfield_patch].angle = alpha_deg * 0.01745;  // rans.h:94
        return afx_rans_set_bcs(s, (int)kinds.size(), kinds.data(), vars.data());
    };
    // rans.h:86-88: the first angle is in place when the coarsest field is initialised
    if ((rc = apply_alpha(levels[0], alphas_deg[0])) != AFX_OK) return rc;
    if (reinit && (rc = afx_rans_init(levels[0])) != AFX_OK) return rc;
    for (int a = 0; a < n_alpha; ++a) {
        for (int l = 0; l < n_levels; ++l) if ((rc = apply_alpha(levels[l], alphas_deg[a])) != AFX_OK) return rc;  // rans.h:94-97
        if ((rc = afx_rans_refill_bcs(levels[0])) != AFX_OK) return rc;                                              // multigrid.h:303
        int it_total = 0, it = 0;
        double err = 0;
        afx_rans* last = levels[0];
        for (int l = 0; l < n_levels; ++l) {
            if (l > 0) {  // multigrid.h:306-311 / 339-344
                if ((rc = afx_rans_bcs_from_internal(levels[l - 1])) != AFX_OK) return rc;
                if ((rc = afx_prolongation_apply(prolongations[l - 1])) != AFX_OK) return rc;
                if ((rc = afx_rans_refill_bcs(levels[l])) != AFX_OK) return rc;
            }
            last = levels[l];
            rc = run_level(levels[l], st, &it, &err);
            it_total += it;
            // multigrid.h:313/355: a failed linear solve ends the run on this level and run() hands THIS level's solver back;
            // run_airfoil takes its wall profile and goes on to the next angle (rans.h:98-104) -- so does the sweep
            if (rc == AFX_ERR_NUMERIC) { any_failed = true; break; }
            if (rc != AFX_OK) return rc;
        }
        if (iterations) iterations[a] = it_total;
        if (residual) residual[a] = err;
        double f[3];
        if ((rc = afx_rans_wall_forces(last, wall_patch, f)) != AFX_OK) return rc;  // rans.h:99-102
        if (cl) cl[a] = f[0];
        if (cd) cd[a] = f[1];
        if (cm) cm[a] = f[2];
    }
    return any_failed ? AFX_ERR_NUMERIC : AFX_OK;  // every angle is filled in either way
}

int afx_rans_sweep(afx_rans* s, const afx_sweep_settings* st, int farfield_patch, int wall_patch, const double* alphas_deg, int n_alpha,
                   int reinit, double* cl, double* cd, double* cm, int* iterations, double* residual)
{
    if (!s) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    afx_rans* one[1] = {s};
    return afx_rans_sweep_fmg(one, nullptr, 1, st, farfield_patch, wall_patch, alphas_deg, n_alpha, reinit, cl, cd, cm, iterations, residual);
}

int afx_rans_profile_halo_ms(afx_rans* s, double out[2])
{
    if (!s || !out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    out[0] = s->s.prof_halo_ms[0]; out[1] = s->s.prof_halo_ms[1];
    return AFX_OK;
}
int afx_rans_last_device_ms(afx_rans* s, double* ms)
{
    if (!s || !ms) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *ms = s->s.last_ms;
    return AFX_OK;
}
int64_t afx_rans_launch_count(afx_rans* s) { return s ? s->s.launches : 0; }

int afx_rans_profile_explicit(afx_rans* s, double relaxation, int n_iter, double out_ms[6])
{
    if (!s) return null_handle();
    return guard([&] {
        auto& S = s->s;
        S.use();
        if (!S.bcs_set) throw afx::InvalidArg("set_bcs has not been called");
        S.push_params(relaxation);
        S.refresh_tile_k3a();
        for (int k = 0; k < 6; ++k) out_ms[k] = 0;
        const bool grads = S.visc_not_inviscid || S.second_order;
        const bool fused = S.fused_stage() || S.pipe_stage();  // one kernel per stage: its time goes to out_ms[5]
        const afx::d4* in[3] = {S.q.p, S.qkA.p, S.qkB.p};
        afx::d4* outp[3] = {S.qkA.p, S.qkB.p, S.q.p};
        const double alpha[3] = {0.25, 0.5, 1.};
        cudaEvent_t ev[14];
        for (auto& e : ev) CK(cudaEventCreate(&e));
        bool halo_split[3] = {false, false, false};
        S.prof_halo_ms[0] = S.prof_halo_ms[1] = 0;
        for (int it = 0; it < n_iter; ++it) {
            int e = 0;
            CK(cudaEventRecord(ev[e++], S.st));
            const bool lim0 = S.lim0_in_dt_grad();
            S.launch_dt_grad(grads, grads, lim0);
            CK(cudaEventRecord(ev[e++], S.st));
            for (int st = 0; st < 3; ++st) {
                if (fused) {
                    CK(cudaEventRecord(ev[e++], S.st));
                    CK(cudaEventRecord(ev[e++], S.st));
                    if (S.fused_stage()) S.launch_stage(st, in[st], outp[st], alpha[st]);
                    else S.launch_pipe(st, in[st], outp[st], alpha[st], S.second_order && !(st == 0 && lim0));
                    CK(cudaEventRecord(ev[e++], S.st));
                    CK(cudaEventRecord(ev[e++], S.st));
                    continue;
                }
                if (S.second_order && !(st == 0 && lim0)) S.launch_limiter(in[st]);
                CK(cudaEventRecord(ev[e++], S.st));
                S.launch_flux(in[st], false, afx::d4{0, 0, 0, 0});
                CK(cudaEventRecord(ev[e++], S.st));
                const bool split = S.halo && S.halo->p2p && !(S.halo_overlap && S.n_front > 0 && S.n_front < S.n_upd);
                S.prof_stage = split ? st : -1;
                if (st < 2) S.launch_gather<0, 0>(in[st], outp[st], S.qW.p, alpha[st], grads);
                else S.launch_gather<0, 1>(in[st], outp[st], S.qW.p, alpha[st], grads);
                S.prof_stage = -1;
                halo_split[st] = split;
                CK(cudaEventRecord(ev[e++], S.st));
                CK(cudaEventRecord(ev[e++], S.st));
            }
            S.ensure_halo();
            CK(cudaStreamSynchronize(S.st));
            float ms;
            CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); out_ms[0] += ms;
            for (int st = 0; st < 3; ++st) {
                if (fused) { CK(cudaEventElapsedTime(&ms, ev[3 + 4 * st], ev[4 + 4 * st])); out_ms[5] += ms; continue; }
                CK(cudaEventElapsedTime(&ms, ev[1 + 4 * st], ev[2 + 4 * st])); out_ms[1] += ms;
                CK(cudaEventElapsedTime(&ms, ev[2 + 4 * st], ev[3 + 4 * st])); out_ms[2] += ms;
                CK(cudaEventElapsedTime(&ms, ev[3 + 4 * st], ev[4 + 4 * st])); out_ms[3] += ms;
                CK(cudaEventElapsedTime(&ms, ev[4 + 4 * st], ev[5 + 4 * st])); out_ms[4] += ms;
                if (halo_split[st]) {  // peer-memory halo: flag hand-off + wait + scatter, taken out of the gather/update figure
                    CK(cudaEventElapsedTime(&ms, S.evp[8 + 2 * st], S.evp[9 + 2 * st])); out_ms[4] += ms; out_ms[3] -= ms;
                    if (!S.halo->early_signal) {
                        float a_ = 0, b_ = 0;
                        CK(cudaEventElapsedTime(&a_, S.evp[8 + 2 * st], S.evp[2 + st])); CK(cudaEventElapsedTime(&b_, S.evp[2 + st], S.evp[9 + 2 * st]));
                        S.prof_halo_ms[0] += a_ / n_iter; S.prof_halo_ms[1] += b_ / n_iter;
                    }
                }
            }
        }
        for (auto& e : ev) cudaEventDestroy(e);
        for (int k = 0; k < 6; ++k) out_ms[k] /= n_iter;
        S.jac_valid = false;
    });
}

}  // extern "C"
