/*
 * rans_oracle.h -- CPU restatement of AeroFLEX's src/rans hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the sm_100a CUDA
 * path in aeroflex_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may link or call it; the
 * product library never does.
 *
 * Parity status: PINNED for the explicit path, the implicit right-hand side,
 * the finite-difference Jacobian blocks, wall forces and mesh metrics -- the
 * restatement is checked entry-by-entry against the unmodified reference
 * headers compiled here (oracle/_ref, see oracle/Makefile and
 * oracle/ref_driver.cpp) and against the frozen outputs of that build under
 * tests/golden/.  UNPINNED for the implicit linear solve (Eigen GMRES + ILUT,
 * un-vendored third-party code).
 *
 * All citations are file:line under /root/reference/src/rans/include/rans/.
 * Plain C99, double precision, no FMA contraction assumed (build with
 * -ffp-contract=off for the parity build).
 */
#ifndef RANS_ORACLE_H
#define RANS_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EDGE_NULL 0xFFFFFFFFu /* mesh.h:35 */

/* edge flux kinds, solver.h:203-246 */
enum { ORC_INTERNAL = 0, ORC_FARFIELD = 1, ORC_SLIPWALL = 2, ORC_WALL = 3,
       ORC_INLET_OUTLET = 4 /* flux of an unknown type (internal), but get_boundary_variables stops here: solver.h:603-606 */ };
/* gradient schemes, solver.h:428,470 */
enum { ORC_GREEN_GAUSS = 0, ORC_LEAST_SQUARES = 1 };

typedef struct { /* core.h:27-46 */
    double gamma, R, mu_L, Pr_L, cp;
} orc_gas;

typedef struct { /* core.h:61-84 */
    double mach, angle, T, p;
} orc_bvars;

/* Geometry in the reference's own layout (mesh.h:209-246). Ghost cells are
 * appended after the N real cells, one per boundary edge, in boundary order. */
typedef struct {
    uint32_t N, G, E;
    uint32_t *edge_cells; /* [E][2]  edgesCells */
    double *enx, *eny, *elen, *ecx, *ecy; /* [E] */
    double *ccx, *ccy, *area; /* [N+G] */
    uint32_t *cell_edges; /* [N][4]  cellsEdges, ORC_EDGE_NULL padded */
    uint8_t *is_tri; /* [N+G] */
    uint32_t *bnd_edge; /* [G]  boundaryEdges */
    int32_t *bnd_patch; /* [G]  index of the physical name of boundary edge b */
} orc_mesh;

typedef struct {
    orc_mesh m;
    orc_gas g;
    /* set_bcs state (solver.h:200-247) */
    uint8_t *edge_kind; /* [E] */
    uint8_t *bnd_kind; /* [G] */
    orc_bvars *bnd_vars; /* [G] boundary_vars */
    int viscous_type; /* 0 inviscid flux, 1 laminar; 2 never reached */
    int visc_not_inviscid; /* viscosity_model != "inviscid" (solver.h:810,983,1083) */
    int second_order;
    int gradient_scheme;
    double limiter_k;
    double cfl;
    /* state (solver.h:54-69), all length 4(N+G) except dt */
    double *q, *qk, *qW, *gx, *gy, *lim, *qmin, *qmax, *rhs;
    double *dt; /* [N+G] */
    double *lsq; /* [N][4] row-major 2x2 (dT d)^-1, solver.h:402-422 */
    /* scratch of the threaded variant (orc_explicit_solve_omp): faces of a cell in ascending edge id, face fluxes */
    uint32_t *cf_sorted; /* [N][4] */
    double *fluxbuf; /* [E][4] */
    int limiter_kind; /* 0 Venkatakrishnan (default build), 1 Michalak (RANS_MICHALAK_LIMITER build, solver.h:557-576) */
} orc_solver;

/* ---- physics.h ---- */
double orc_pressure(const double q[4], double gamma);
void orc_flux_internal(const orc_gas *g, int viscous_type, double nx, double ny,
                       const double qL[4], const double qR[4],
                       const double gx[4], const double gy[4], double f[4]);
void orc_bc_vars(int kind, const orc_gas *g, double nx, double ny,
                 const double qL[4], const double qbc[4], double qR[4]);
void orc_flux(int kind, const orc_gas *g, int viscous_type, double nx, double ny,
              const double qL[4], const double qR[4],
              const double gx[4], const double gy[4], double f[4]);
void orc_fd_jacobian(int kind, const orc_gas *g, int viscous_type, double nx, double ny,
                     const double qL[4], const double qR[4],
                     const double gx[4], const double gy[4], double J[64]);
void orc_get_conservative(const orc_bvars *v, const orc_gas *g, double q[4]);

/* ---- mesh.h ---- */
/* Build the reference's edge/metric/ghost arrays from nodes, cells and
 * boundary segments (mesh.h:317-453, 744-787, 834-884). cells: [nc][4] with
 * triangles padded by node 0 (mesh.h:715-717). Returns 0, or -1 if a boundary
 * segment is not an edge of the mesh (mesh.h:757). Free with orc_mesh_free. */
int orc_mesh_build(orc_mesh *m, uint32_t nn, const double *x, const double *y,
                   uint32_t nc, const uint32_t *cells, const uint8_t *is_tri,
                   uint32_t nb, const uint32_t *b0, const uint32_t *b1,
                   const int32_t *bpatch);
void orc_mesh_free(orc_mesh *m);

/* ---- solver.h ---- */
int orc_solver_init(orc_solver *s, const orc_mesh *m_borrowed, const orc_gas *g,
                    int viscosity_model /*0 inviscid,1 laminar,2 spallart-allmaras*/);
void orc_solver_free(orc_solver *s);
/* kinds/vars per patch id, as bcs.at(name) would give (solver.h:216-230) */
void orc_set_bcs(orc_solver *s, int npatch, const uint8_t *patch_kind,
                 const orc_bvars *patch_vars);
void orc_set_gradient_scheme(orc_solver *s, int scheme);
void orc_init_field(orc_solver *s);
void orc_refill_bcs(orc_solver *s);
void orc_bcs_from_internal(orc_solver *s);
int orc_boundary_variables(const orc_solver *s, orc_bvars *out);
void orc_set_walls_from_internal(orc_solver *s, double *q_);
void orc_calc_dt(orc_solver *s);
void orc_calc_gradients(orc_solver *s);
void orc_calc_limiters(orc_solver *s, const double *q_);
void orc_average_gradients(const orc_solver *s, uint32_t c0, uint32_t c1,
                           double gradx[4], double grady[4]);
void orc_explicit_residual(orc_solver *s, const double *q_);
double orc_explicit_solve(orc_solver *s, double relaxation);
double orc_implicit_rhs(orc_solver *s);
double orc_uniform_residual(orc_solver *s);
/* fillRhoLHS (solver.h:979-1071) in block form: diag[(N+G)][16] and, per edge,
 * off01[E][16] (row c0, col c1) and off10[E][16] (row c1, col c0). */
void orc_implicit_lhs(orc_solver *s, double *diag, double *off01, double *off10);
double orc_norm(const double *v, size_t n);

/* ---- post.h:301-387 ---- */
int orc_wall_forces(const orc_solver *s, int patch, double out_cl_cd_cm[3]);

/* An 8-thread variant of orc_explicit_solve for timing only (cell-based
 * gathers instead of the reference's serial scatter loops; same arithmetic per
 * face, same per-cell accumulation order). */
double orc_explicit_solve_omp(orc_solver *s, double relaxation);

#ifdef __cplusplus
}
#endif
#endif
