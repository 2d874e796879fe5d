/*
 * afx_rans.h -- C ABI of libaeroflex_rans_b200.so: the B200 (sm_100a) drop-in
 * for AeroFLEX's src/rans residual / pseudo-time hot path.
 *
 * The reference has no FFI for this path; its seam is the C++ class
 * rans::solver (fill / compute / solve, solver.h:171-173) used as the template
 * argument of rans::multigrid<solverType> (multigrid.h:28-56).  The adapter
 * class rans::gpuSolver (aeroflex_b200/host/rans_b200/gpu_solver.h) mirrors
 * that class and forwards every call to the entry points below; each entry
 * point cites the reference member it replaces (paths relative to
 * src/rans/include/rans/ of the reference).
 *
 * Conventions
 *  - plain C types only; no exceptions cross; every int-returning function
 *    returns 0 on success or a negative afx_status, with afx_last_error()
 *    giving a message for the calling thread;
 *  - all cell / edge / boundary arrays are in the REFERENCE's order and
 *    layout (mesh.h:209-246; state = 4 doubles per cell, real cells first,
 *    then one ghost cell per boundary edge, solver.h:182-196).  Renumbering
 *    for coalescing is internal and invisible;
 *  - double precision throughout; indices are uint32_t (core.h:22);
 *  - one handle drives one GPU from one host thread at a time; distinct
 *    handles are independent.
 */
#ifndef AFX_RANS_H
#define AFX_RANS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFX_EDGE_NULL 0xFFFFFFFFu /* mesh.h:35 MESH_EDGE_NULL */

typedef enum {
    AFX_OK = 0,
    AFX_ERR_INVALID = -1,   /* bad argument / unknown name (std::invalid_argument, std::out_of_range) */
    AFX_ERR_CUDA = -2,      /* CUDA runtime failure or no usable device */
    AFX_ERR_NUMERIC = -3,   /* NaN/Inf residual, linear solver failure (the reference returns -1) */
    AFX_ERR_IO = -4,        /* mesh / config file problems (std::runtime_error) */
    AFX_ERR_COMM = -5       /* NCCL failure */
} afx_status;

/* edge flux kinds chosen by solver::set_bcs (solver.h:216-246) */
enum { AFX_BC_INTERNAL = 0, AFX_BC_FARFIELD = 1, AFX_BC_SLIPWALL = 2, AFX_BC_WALL = 3,
       /* the literal type "inlet-outlet": like every unknown type it keeps the internal flux against the ghost cell, but
        * solver::get_boundary_variables (solver.h:597-611) stops its search at such an edge and returns the defaults */
       AFX_BC_INLET_OUTLET = 4 };
/* Settings::viscosity_options (core.h:176) */
enum { AFX_VISC_INVISCID = 0, AFX_VISC_LAMINAR = 1, AFX_VISC_SA = 2 };
/* Settings::gradient_options (core.h:175), by meaning not by index */
enum { AFX_GRAD_GREEN_GAUSS = 0, AFX_GRAD_LEAST_SQUARES = 1 };
/* arithmetic mode of the kernels.  STRICT: the reference's expression order, no FMA contraction -> vectors are
 * bit-identical to the CPU reference built without -march (its default).  FAST (default): the same formulas
 * with reciprocals shared between divisions by one denominator and FMA contraction; every face flux agrees
 * with STRICT to a few ulp, residual histories to ~1e-13 relative (BASELINE tolerance: 1e-10). */
enum { AFX_MATH_STRICT = 0, AFX_MATH_FAST = 1 };
/* fields readable with afx_rans_get_field (solver.h:54-69) */
enum { AFX_F_Q = 0, AFX_F_QW = 1, AFX_F_GX = 2, AFX_F_GY = 3, AFX_F_LIMITERS = 4, AFX_F_DT = 5, AFX_F_RHS = 6 };

/* rans::gas, core.h:27-46 */
typedef struct afx_gas {
    double gamma, R, mu_L, Pr_L, cp;
} afx_gas;

/* rans::boundary_variables, core.h:61-84 */
typedef struct afx_bvars {
    double mach, angle, T, p;
} afx_bvars;

/* Borrowed view of a rans::mesh (mesh.h:209-246). */
typedef struct afx_mesh_desc {
    uint32_t n_cells;               /* nRealCells */
    uint32_t n_ghost;               /* boundaryEdges.size() */
    uint32_t n_edges;               /* edgesCells.cols() */
    const uint32_t* edges_cells;    /* [E][2]  edgesCells */
    const double* edges_nx;         /* [E] edgesNormalsX (out of cell 0) */
    const double* edges_ny;
    const double* edges_len;
    const double* edges_cx;         /* [E] edgesCentersX */
    const double* edges_cy;
    const double* cells_cx;         /* [N+G] cellsCentersX */
    const double* cells_cy;
    const double* cells_area;       /* [N+G] */
    const uint32_t* cells_edges;    /* [N][4] cellsEdges, AFX_EDGE_NULL padded */
    const uint8_t* cells_is_tri;    /* [N] */
    const uint32_t* boundary_edges; /* [G] boundaryEdges */
    const int32_t* boundary_patch;  /* [G] patch id of boundaryEdgesPhysicals[b] */
} afx_mesh_desc;

const char* afx_last_error(void);
const char* afx_version(void);
/* number of CUDA devices visible, or a negative afx_status */
int afx_device_count(void);

/* page-locked host memory for state transfers (afx_rans_set_q / get_q run at PCIe speed from it) */
void* afx_pinned_alloc(size_t bytes);
void afx_pinned_free(void* p);

/* ------------------------------------------------------------------ */
/* Mesh ingest (host only; replaces rans::mesh, mesh.h:250,834-884)     */
/* ------------------------------------------------------------------ */
typedef struct afx_mesh afx_mesh;

/* rans::mesh(filename): Gmsh MSH 4.1 ASCII, triangles + quads (mesh.h:457-738) */
int afx_mesh_read_msh(afx_mesh** out, const char* path);
/* the same construction from in-memory elements: cells [nc][4] (triangles
 * padded with node 0, mesh.h:715-717), boundary segments with a patch id each */
int afx_mesh_from_elements(afx_mesh** out, uint32_t n_nodes, const double* x, const double* y,
                           uint32_t n_cells, const uint32_t* cells, const uint8_t* is_tri,
                           uint32_t n_bnd, const uint32_t* b0, const uint32_t* b1, const int32_t* bpatch,
                           int n_patch, const char* const* patch_names);
/* Synthetic NACA0012 O-mesh: ni cells around x nj radial
 * layers, the inner n_quad_layers layers quads, the rest split into two
 * triangles each; patches "wall" and "farfield"; cells in (j,i) order. */
int afx_mesh_synth_omesh(afx_mesh** out, uint32_t ni, uint32_t nj, uint32_t n_quad_layers, double far_radius);
void afx_mesh_free(afx_mesh* m);
int afx_mesh_get_desc(const afx_mesh* m, afx_mesh_desc* out);
uint32_t afx_mesh_n_nodes(const afx_mesh* m);
int afx_mesh_n_patches(const afx_mesh* m);
const char* afx_mesh_patch_name(const afx_mesh* m, int patch);
int afx_mesh_patch_id(const afx_mesh* m, const char* name); /* -1 if absent */
/* raw inputs back out (nodes, connectivity, boundary segments); any pointer may be NULL */
int afx_mesh_get_elements(const afx_mesh* m, double* x, double* y, uint32_t* cells, uint32_t* b0, uint32_t* b1);
/* write MSH 4.1 ASCII that rans::mesh can read back (so the CPU reference can run on synthetic meshes) */
int afx_mesh_write_msh(const afx_mesh* m, const char* path);

/* ------------------------------------------------------------------ */
/* Solver (replaces rans::solver / explicitSolver / implicitSolver)     */
/* ------------------------------------------------------------------ */
typedef struct afx_rans afx_rans;

/* solver(mesh, gas, viscosity_model) + set_mesh_and_gas (solver.h:108-110,177-197).
 * device = CUDA ordinal.  All state vectors start at zero (the reference
 * leaves them uninitialised, see DESIGN.md "F9"). */
int afx_rans_create(afx_rans** out, const afx_mesh_desc* mesh, const afx_gas* gas, int viscosity_model, int device);
void afx_rans_destroy(afx_rans* s);

/* ---- multi-GPU: one process per GPU, one partition per process (no counterpart in the reference) ----
 * Host-only planning: cut the global mesh into `nranks` pieces along a Hilbert curve, each with a 2-cell halo. */
typedef struct afx_partition afx_partition;
int afx_partition_create(afx_partition** out, const afx_mesh_desc* global_mesh, int nranks, int rank);
void afx_partition_free(afx_partition* p);
/* the piece as a mesh of its own, reference layout: cells = owned | ring 1 | ring 2, then its boundary ghosts */
int afx_partition_get_desc(const afx_partition* p, afx_mesh_desc* out);
/* out = {n_owned, n_ring1, n_ring2, n_boundary_ghosts, n_local_edges, n_peers, rank, nranks} */
int afx_partition_info(const afx_partition* p, uint32_t out[8]);
const uint32_t* afx_partition_cell_l2g(const afx_partition* p); /* local cell (incl. ghosts) -> global cell */
const uint32_t* afx_partition_edge_l2g(const afx_partition* p); /* local edge -> global edge, ascending */
/* exchange plan with peer i: owned cells to send, halo cells to fill (local ids, matching order on both sides) */
int afx_partition_peer(const afx_partition* p, int i, int* peer_rank, const uint32_t** send, uint32_t* n_send,
                       const uint32_t** recv, uint32_t* n_recv);
/* NCCL bootstrap: rank 0 makes the 128-byte id, the host program broadcasts it (torch.distributed, MPI, ...) */
int afx_nccl_unique_id(char out[128]);
/* a solver on this rank's piece; halo cells are refreshed over NCCL after every stage, norms and forces are
 * all-reduced.  Collective: every rank must make the same calls.  set_q takes the GLOBAL vector, get_q fills
 * this rank's entries of the GLOBAL vector. */
int afx_rans_create_partitioned(afx_rans** out, const afx_partition* part, const afx_gas* gas, int viscosity_model,
                                int device, const char nccl_id[128]);

/* In-process group: the partitioned solvers of ONE host process (one host thread per handle; all ranks on one device, or
 * one device each).  It stands where NCCL stands: the halo is exchanged by device-to-device copies between two host
 * rendezvous (the "staged" halo, no CUDA graph), the norms and forces are summed on the host in rank order.  This is what
 * lets a box with a single GPU run every partition plan on hardware (NCCL refuses two ranks on one device), and what a
 * single-process driver of several GPUs uses.  afx_rans_p2p_export / _connect work on group members too (the blobs then
 * carry plain device pointers): the peer-memory push, its flag hand-off and the captured graph run exactly as between
 * processes.  Every collective call must be made by all ranks, each from its own host thread; a rank that never arrives
 * breaks the group after 120 s and the waiting ranks return AFX_ERR_COMM.  afx_group_abort breaks it at once. */
typedef struct afx_group afx_group;
int afx_group_create(afx_group** out, int nranks);
void afx_group_free(afx_group* g);   /* after the last member solver is destroyed */
void afx_group_abort(afx_group* g);
int afx_rans_create_partitioned_group(afx_rans** out, const afx_partition* part, const afx_gas* gas, int viscosity_model,
                                      int device, afx_group* group);

/* Halo over NVLink peer memory instead of NCCL (one node, one process per GPU, CUDA IPC): every rank exports a
 * blob (size returned in *size; pass blob = NULL to query), the host program all-gathers the blobs in rank order
 * and hands them to connect.  After that the update kernel itself stores the send layer into the peers' receive
 * buffers and a flag hand-off replaces the NCCL send/recv group; norms and forces still use NCCL all-reduce. */
int afx_rans_p2p_export(afx_rans* s, void* blob, size_t* size);
int afx_rans_p2p_connect(afx_rans* s, const void* blobs, size_t blob_size, int nranks);
/* A halo wait that sees no flag from a peer within AFX_HALO_TIMEOUT_MS (default 20000; 0 = wait for ever) gives up: the
 * run that contains it returns AFX_ERR_COMM and the solver's state is no longer valid.  The peer-memory path is used only
 * if EVERY rank's plan fits it (<= 8 peers); otherwise all ranks stay on the collective exchange.
 * 0 single GPU, 1 NCCL (or staged in-process) halo, 2 peer-memory halo */
int afx_rans_halo_mode(afx_rans* s);

/* solver::set_bcs (solver.h:200-247): kind and far-field variables per patch id */
int afx_rans_set_bcs(afx_rans* s, int n_patch, const uint8_t* patch_kind, const afx_bvars* patch_vars);
/* set_second_order / set_gradient_scheme / set_limiter_k (solver.h:135,149-152,162) */
int afx_rans_set_options(afx_rans* s, int second_order, int gradient_scheme, double limiter_k);
/* Which limiter function calc_limiters (solver.h:517-593) evaluates: Venkatakrishnan (the reference's default build, solver.h:578-584)
 * or Michalak (solver.h:557-576 + michalak_limiter, physics.h:581-592 -- what the reference compiles when RANS_MICHALAK_LIMITER is
 * defined; nothing in its build defines it).  Replaces that compile-time switch with a run-time one; the adapter headers map the
 * macro onto this call.  Michalak runs on the three-kernel stage (no limiter inside the dt/gradient kernel, no fused stage). */
enum { AFX_LIMITER_VENKATAKRISHNAN = 0, AFX_LIMITER_MICHALAK = 1 };
int afx_rans_set_limiter(afx_rans* s, int limiter);
/* arithmetic mode (AFX_MATH_*); the environment variable AFX_MATH=strict|fast sets the default at creation */
int afx_rans_set_math_mode(afx_rans* s, int mode);
int afx_rans_get_math_mode(afx_rans* s);
/* Stage kernels of the explicit iteration: 0 (default) = limiter / face flux / gather+update as three kernels;
 * 1 = ONE persistent kernel per Runge-Kutta stage on shared-memory tiles (limiter + MUSCL + flux + gather + update;
 * second-order non-laminar runs).  The tiles are built at creation only if the environment has AFX_FUSED=1 (the cells
 * are then numbered by recursive graph bisection); without them this call leaves the three-kernel stage in place.
 * Strict-mode states are bit-identical either way. */
int afx_rans_set_fused(afx_rans* s, int on);
/* out[0] = 1 if the fused stage kernel is in use, out[1] = tiles, out[2] = cells per tile, out[3] = dynamic shared
 * memory per CTA (bytes), out[4] = resident CTAs per SM, out[5] = largest local cell count, out[6] = largest local
 * face count, out[7] = total local cells of all tiles (own + ring 1 + state-only; / n_cells = staging overhead) */
int afx_rans_tile_info(afx_rans* s, uint64_t out[8]);
/* Pipelined stage kernel (default for inviscid / "spallart-allmaras" runs; AFX_PIPE=0 or on = 0 selects the three-kernel stage):
 * ONE persistent kernel per Runge-Kutta stage sweeps the mesh in chunks with limiter, face-flux and gather/update phases a
 * few chunks apart, so the limiters, the flux buffer and the re-read states and gradients change hands inside the L2 instead
 * of through HBM.  Same arithmetic, same bits as the three-kernel stage.
 * info: {active, cells per chunk, chunks, work items per stage, far faces, far cells, lagF << 32 | lagU, CTAs} */
int afx_rans_set_pipelined(afx_rans* s, int on);
int afx_rans_pipe_info(afx_rans* s, uint64_t out[8]);
/* Host-only (no device needed): renumber `mesh` as afx_rans_create would, cut it into tiles of at most `tile_cells`
 * cells and verify the plan against the connectivity.  n_tiles out; per_tile[4*t..] = own cells, ring-1 cells,
 * state-only cells, local faces of tile t (up to `cap` tiles); smem_bytes = dynamic shared memory k_stage would need. */
int afx_tiling_plan(const afx_mesh_desc* mesh, uint32_t tile_cells, const uint32_t* limits /* NULL or {max local cells, max own+ring1,
                    max faces, max ring ids}: tiles beyond are cut in two */, uint32_t* n_tiles, uint32_t* per_tile, uint32_t cap, uint64_t* smem_bytes);
/* the same for one rank's piece of a partitioned mesh: only the owned cells are tiled, the send layer by its own tiles */
int afx_tiling_plan_partition(const afx_partition* part, uint32_t tile_cells, const uint32_t* limits, uint32_t* n_tiles, uint32_t* per_tile,
                              uint32_t cap, uint64_t* smem_bytes);
/* solver::set_cfl (solver.h:250-252) */
int afx_rans_set_cfl(afx_rans* s, double cfl);
/* solver::init / refill_bcs / bcs_from_internal (solver.h:615-631, 259-287) */
int afx_rans_init(afx_rans* s);
int afx_rans_refill_bcs(afx_rans* s);
int afx_rans_bcs_from_internal(afx_rans* s);
/* solver::get_q as value transfer: q has 4*(N+G) doubles in reference order */
int afx_rans_set_q(afx_rans* s, const double* q);
int afx_rans_get_q(afx_rans* s, double* q);
/* the same for a partitioned handle in ITS OWN numbering: 4*(local cells + local ghosts) doubles in the order of
 * afx_partition_get_desc (owned | ring 1 | ring 2 | boundary ghosts) -- what a distributed host program holds.
 * On a single-GPU handle they are set_q / get_q. */
int afx_rans_set_q_local(afx_rans* s, const double* q_local);
int afx_rans_get_q_local(afx_rans* s, double* q_local);
/* one of AFX_F_*: 4*(N+G) doubles (N+G for AFX_F_DT) in reference order */
int afx_rans_get_field(afx_rans* s, int field, double* out);
/* solver::get_boundary_variables (solver.h:597-611); returns 1 if a far-field patch was found, else 0 (defaults) */
int afx_rans_boundary_variables(afx_rans* s, afx_bvars* out);

/* solver::get_uniform_residual (solver.h:636-690), accumulating from zero */
int afx_rans_uniform_residual(afx_rans* s, double* norm);
/* explicitSolver::solve (solver.h:802-828): local dt, 3 stages of
 * [wall ghosts, gradients, limiter, residual, update]; *norm = ||qW||_2 */
int afx_rans_step_explicit(afx_rans* s, double relaxation, double* norm);
/* n_iter explicit iterations back to back without host round trips;
 * norms[n_iter] (may be NULL) receives every iteration's norm at the end */
int afx_rans_run_explicit(afx_rans* s, double relaxation, int n_iter, double* norms);
/* the single phases, for parity tests: on_qk=0 works on q.  They mirror
 * calc_dt / set_walls_from_internal+calc_gradients / calc_limiters /
 * explicitSolver::calc_residual (solver.h:308,289,425,517,745) */
int afx_rans_phase_dt_gradients(afx_rans* s);
int afx_rans_phase_limiters(afx_rans* s);
int afx_rans_phase_residual(afx_rans* s, double* norm);
/* implicitSolver::fillRhoRHS (solver.h:1079-1152): RHS into AFX_F_RHS, *norm = ||RhoVector||_2 */
int afx_rans_residual(afx_rans* s, double* norm);
/* implicitSolver::fillRhoLHS (solver.h:979-1071): block Jacobian on the device */
int afx_rans_fill_jacobian(afx_rans* s);
/* the 4x4 blocks of that matrix: diag[(N+G)][16], off01[E][16] (row c0, col c1), off10[E][16]; row-major blocks */
int afx_rans_get_jacobian_blocks(afx_rans* s, double* diag, double* off01, double* off10);
/* implicitSolver::compute (solver.h:1160-1167): prepares the preconditioner of the last afx_rans_fill_jacobian.  The
 * reference factorises an Eigen::IncompleteLUT there; this library inverts the 4x4 diagonal blocks for a
 * block-Jacobi smoother.  AFX_ERR_NUMERIC where the reference returns -1. */
int afx_rans_compute(afx_rans* s);
/* implicitSolver::solve (solver.h:1170-1213): RHS, left-preconditioned restarted GMRES on the device with the
 * frozen Jacobian, q += relaxation * dq, up to 1 + rhs_iterations times with the reference's early exits;
 * *norm = final ||RhoVector||_2.  AFX_ERR_NUMERIC (and *norm = -1) where the reference returns -1.  The linear
 * solver is NOT Eigen's: per-iteration histories of the implicit path are "parity unpinned" (DESIGN.md).
 * Partitioned handles: every rank calls it; the halo rows of the Krylov vectors are fetched from their owners before
 * each matrix-vector product and Jacobi sweep, inner products are summed over the ranks. */
int afx_rans_step_implicit(afx_rans* s, double relaxation, double tol, int rhs_iterations, double* norm);
/* defaults mirror solver.h:906-910 (restart 30, 500 iterations, tolerance 1e-2); precond_sweeps block-Jacobi sweeps */
int afx_rans_set_linear_solver(afx_rans* s, int restart, int max_iterations, double tolerance, int precond_sweeps);
int afx_rans_last_linear_iterations(afx_rans* s);
/* get_wall_profile (post.h:301-387): out = {cl, cd, cm} of one patch */
int afx_rans_wall_forces(afx_rans* s, int patch, double out_cl_cd_cm[3]);
/* CpProfile::calc_cp (post.h:248-298): cp of the owner cell of every boundary edge of the patch, boundary order;
 * returns the count (cp may be NULL to query it) */
int afx_rans_wall_cp(afx_rans* s, int patch, double* cp);

/* device time of the kernels of the last step/run call, milliseconds (CUDA events on the solver's stream) */
/* Warm-started angle-of-attack sweep on one mesh level -- the per-angle body of Rans::run_airfoil (rans.h:86-104) with
 * multigrid<T>::run_solver (multigrid.h:182-293) inside the library: the far-field angle of `farfield_patch` is set to
 * alpha * 0.01745 (the reference's deg -> rad literal, rans.h:94) and set_bcs is applied; init() runs once before the
 * first angle if `reinit`; then per angle refill_bcs(), the pseudo-time loop (explicit: constant CFL start_cfl; implicit:
 * fill + compute + solve with the CFL ramp start_cfl + (i+1) slope_cfl <= max_cfl and rhs_iterations sub-iterations)
 * until residual / uniform-flow residual <= tolerance or max_iterations, and get_wall_profile of `wall_patch`
 * (post.h:301-387).  cl/cd/cm/iterations/residual have n_alpha entries (any may be NULL).  A chain of angles per GPU is
 * how the polar database of the VLM viscous correction is sharded (BASELINE config 5: no communication between chains).
 * Where the reference's run_solver returns 1 (a failed linear solve, multigrid.h:227,285) the sweep does what run_airfoil does:
 * it takes the forces of the level the run ended on and goes on to the next angle (residual = -1 / uniform-flow residual for that
 * angle, the failed iteration not counted); every angle is filled in and the call returns AFX_ERR_NUMERIC if any angle failed.
 * Partitioned handles: called by every rank with the same arguments (explicit and implicit). */
typedef struct afx_sweep_settings {
    int implicit;            /* 0: explicitSolver, 1: implicitSolver */
    double relaxation, start_cfl, slope_cfl, max_cfl, tolerance;
    int rhs_iterations, max_iterations;
} afx_sweep_settings;
int afx_rans_sweep(afx_rans* s, const afx_sweep_settings* settings, int farfield_patch, int wall_patch, const double* alphas_deg,
                   int n_alpha, int reinit, double* cl, double* cd, double* cm, int* iterations, double* residual);

/* FMG prolongation between two mesh levels ON THE DEVICE (replaces `solvers[i].get_q() = mappers[i-1] * solvers[i-1].get_q()`,
 * multigrid.h:308,341; the weights are those of multigrid::gen_mapper, multigrid.h:100-178).  CSR with one weight per
 * (fine cell, coarse cell) pair: row r = fine cell r in the fine mesh's reference order (ghost rows included, n_fine_cells +
 * n_fine_ghosts rows), col = coarse cell in the coarse mesh's reference order, columns ascending inside a row.  apply()
 * overwrites the fine solver's state with P q_coarse, summed in column order (the reference's bits); both solvers on one
 * device, neither partitioned. */
typedef struct afx_prolongation afx_prolongation;
int afx_prolongation_create(afx_prolongation** out, afx_rans* coarse, afx_rans* fine, const uint32_t* row_begin, const uint32_t* col,
                            const double* w);
void afx_prolongation_free(afx_prolongation* p);
int afx_prolongation_apply(afx_prolongation* p);
/* The whole of Rans::run_airfoil (rans.h:78-106): per angle, multigrid<T>::run(false) (multigrid.h:295-363) over `n_levels`
 * mesh levels -- level 0 warm-started from its own state of the previous angle, level l from prolongations[l-1] applied to
 * level l-1 after bcs_from_internal -- each level iterated by run_solver as in afx_rans_sweep; forces from the last level run.
 * iterations[a] = iterations summed over the levels.  States never leave the device.  levels[l]: set_bcs and set_options done. */
int afx_rans_sweep_fmg(afx_rans* const* levels, afx_prolongation* const* prolongations, int n_levels, const afx_sweep_settings* settings,
                       int farfield_patch, int wall_patch, const double* alphas_deg, int n_alpha, int reinit, double* cl, double* cd,
                       double* cm, int* iterations, double* residual);

int afx_rans_last_device_ms(afx_rans* s, double* ms);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t afx_rans_launch_count(afx_rans* s);
/* per-phase device milliseconds of one explicit iteration, timed with CUDA events between the phases:
 * out[0]=dt+gradients, out[1]=limiter (3 stages), out[2]=face flux (3 stages), out[3]=gather+update (3 stages),
 * out[4]=halo exchange (3 stages; 0 on one GPU; peer-memory halo: flag hand-off + wait + scatter, not counted in out[3]), out[5]=fused stage kernel (3 stages; then out[1..3] are 0) */
int afx_rans_profile_explicit(afx_rans* s, double relaxation, int n_iter, double out_ms[6]);
/* of the last afx_rans_profile_explicit on a peer-memory halo without the in-kernel hand-off (AFX_HALO_EARLY_SIGNAL=0):
 * device milliseconds per iteration of the signalling kernel and of the wait + scatter kernel (3 exchanges) */
int afx_rans_profile_halo_ms(afx_rans* s, double out[2]);

#ifdef __cplusplus
}
#endif
#endif /* AFX_RANS_H */
