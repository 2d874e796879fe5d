/*
 * rans_oracle.c -- CPU restatement of AeroFLEX's src/rans hot path.
 * TEST INFRASTRUCTURE ONLY (see rans_oracle.h for the rules and parity status).
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src/rans/include/rans/).  Expression association follows the
 * reference so that, without FMA contraction, results are bit-identical to the
 * reference headers compiled with a plain-container Eigen stand-in.
 */
#include "rans_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* physics.h                                                           */
/* ------------------------------------------------------------------ */

/* physics.h:48-53 (same expression inlined at solver.h:332, physics.h:183,472,474) */
double orc_pressure(const double q[4], double gamma)
{
    return (gamma - 1) * (q[3] - 0.5 / q[0] * (q[1] * q[1] + q[2] * q[2]));
}

/* physics.h:84-86 */
static double sabs_(double x) { return sqrt(x * x + 1e-4); }
/* physics.h:133-135 */
static double entropy_fix_(double l, double d) { return l > d ? l : (l * l + d * d) / (2 * d); }

/* physics.h:33-46 */
static void grad_p_(double gp[2], const double q[4], const double gx[4], const double gy[4], double gamma)
{
    gp[0] = 2. * gx[3];
    gp[0] -= gx[1] * q[1] / q[0] + q[1] * (gx[1] * q[0] - gx[0] * q[1]) / (q[0] * q[0]);
    gp[0] -= gx[2] * q[2] / q[0] + q[2] * (gx[2] * q[0] - gx[0] * q[2]) / (q[0] * q[0]);
    gp[0] *= 0.5 * (gamma - 1);
    gp[1] = 2. * gy[3];
    gp[1] -= gy[1] * q[1] / q[0] + q[1] * (gy[1] * q[0] - gy[0] * q[1]) / (q[0] * q[0]);
    gp[1] -= gy[2] * q[2] / q[0] + q[2] * (gy[2] * q[0] - gy[0] * q[2]) / (q[0] * q[0]);
    gp[1] *= 0.5 * (gamma - 1);
}

/* Roe flux with smoothed |.| and Harten fix, optional laminar term.
 * physics.h:160-264 */
void orc_flux_internal(const orc_gas *g, int viscous_type, double nx, double ny,
                       const double qL[4], const double qR[4],
                       const double gx[4], const double gy[4], double f[4])
{
    const double gam = g->gamma;
    const double V_L = (qL[1] * nx + qL[2] * ny) / qL[0];
    const double V_R = (qR[1] * nx + qR[2] * ny) / qR[0];
    const double pL = (gam - 1) * (qL[3] - 0.5 / qL[0] * (qL[1] * qL[1] + qL[2] * qL[2]));
    const double pR = (gam - 1) * (qR[3] - 0.5 / qR[0] * (qR[1] * qR[1] + qR[2] * qR[2]));

    /* central part, physics.h:186-189 */
    f[0] = (V_L * qL[0] + V_R * qR[0]) * 0.5;
    f[1] = (V_L * qL[1] + pL * nx + V_R * qR[1] + pR * nx) * 0.5;
    f[2] = (V_L * qL[2] + pL * ny + V_R * qR[2] + pR * ny) * 0.5;
    f[3] = (V_L * (qL[3] + pL) + V_R * (qR[3] + pR)) * 0.5;

    /* Roe averages, physics.h:196-211 */
    const double uL = qL[1] / qL[0], uR = qR[1] / qR[0];
    const double vL = qL[2] / qL[0], vR = qR[2] / qR[0];
    const double sL = sqrt(qL[0]), sR = sqrt(qR[0]);
    const double rho = sR * sL;
    const double u = (uL * sL + uR * sR) / (sL + sR);
    const double v = (vL * sL + vR * sR) / (sL + sR);
    const double h = ((qL[3] + pL) / qL[0] * sL + (qR[3] + pR) / qR[0] * sR) / (sL + sR);
    const double q2 = u * u + v * v;
    const double c = sqrt((gam - 1.) * (h - 0.5 * q2));
    const double V = u * nx + v * ny;
    const double VR = uR * nx + vR * ny;
    const double VL = uL * nx + vL * ny;

    /* physics.h:215-217 */
    const double l_cm = entropy_fix_(sabs_(V - c), 0.05 * c);
    const double l_c = entropy_fix_(sabs_(V), 0.05 * c);
    const double l_cp = entropy_fix_(sabs_(V + c), 0.05 * c);

    /* physics.h:220-223 */
    const double k1 = l_cm * ((pR - pL) - rho * c * (VR - VL)) / (2. * c * c);
    const double k2 = l_c * ((qR[0] - qL[0]) - (pR - pL) / (c * c));
    const double k3 = l_c * rho;
    const double k5 = l_cp * ((pR - pL) + rho * c * (VR - VL)) / (2 * c * c);

    /* physics.h:225-228 */
    f[0] -= 0.5 * (k1 + k2 + k5);
    f[1] -= 0.5 * (k1 * (u - c * nx) + k2 * u + k3 * (uR - uL - (VR - VL) * nx) + k5 * (u + c * nx));
    f[2] -= 0.5 * (k1 * (v - c * ny) + k2 * v + k3 * (vR - vL - (VR - VL) * ny) + k5 * (v + c * ny));
    f[3] -= 0.5 * (k1 * (h - c * V) + k2 * q2 * 0.5 + k3 * (u * (uR - uL) + v * (vR - vL) - V * (VR - VL)) + k5 * (h + c * V));

    if (viscous_type == 1) { /* physics.h:230-257 */
        double qc[4], gp[2], gT[2], gu[2], gv[2];
        for (int k = 0; k < 4; ++k) qc[k] = 0.5 * (qL[k] + qR[k]);
        const double p = orc_pressure(qc, gam);
        grad_p_(gp, qc, gx, gy, gam);
        gT[0] = (1. / g->R) * ((gp[0] * qc[0] - gx[0] * p) / (qc[0] * qc[0])); /* physics.h:62-67 */
        gT[1] = (1. / g->R) * ((gp[1] * qc[0] - gy[0] * p) / (qc[0] * qc[0]));
        gu[0] = (qc[0] * gx[1] - qc[1] * gx[0]) / (qc[0] * qc[0]); /* physics.h:73-74 */
        gu[1] = (qc[0] * gy[1] - qc[1] * gy[0]) / (qc[0] * qc[0]);
        gv[0] = (qc[0] * gx[2] - qc[2] * gx[0]) / (qc[0] * qc[0]); /* physics.h:80-81 */
        gv[1] = (qc[0] * gy[2] - qc[2] * gy[0]) / (qc[0] * qc[0]);
        const double mu = g->mu_L;
        const double kk = g->cp * g->mu_L / g->Pr_L; /* core.h:43-45 */
        const double div_v = gu[0] + gv[1];
        const double txx = 2. * mu * (gu[0] - div_v / 3.);
        const double tyy = 2. * mu * (gv[1] - div_v / 3.);
        const double txy = mu * (gu[1] + gv[0]);
        const double ph0 = qc[1] / qc[0] * txx + qc[2] / qc[0] * txy + kk * gT[0];
        const double ph1 = qc[1] / qc[0] * txy + qc[2] / qc[0] * tyy + kk * gT[1];
        f[1] -= nx * txx + ny * txy;
        f[2] -= nx * txy + ny * tyy;
        f[3] -= nx * ph0 + ny * ph1;
    }
    /* viscous_type == 2: empty in the reference (physics.h:259-261) */
}

/* flux::vars -- the state the face sees on its right (physics.h:267-276,
 * 311-339, 377-405, 446-530) */
void orc_bc_vars(int kind, const orc_gas *g, double nx, double ny,
                 const double qL[4], const double qbc[4], double qR[4])
{
    const double gam = g->gamma;
    if (kind == ORC_INTERNAL) { /* physics.h:275 */
        for (int k = 0; k < 4; ++k) qR[k] = (qL[k] + qbc[k]) * 0.5;
    } else if (kind == ORC_SLIPWALL) { /* physics.h:328-336 */
        const double rhoV = qL[1] * nx + qL[2] * ny;
        qR[0] = qL[0];
        qR[1] = qL[1] - 2. * rhoV * nx;
        qR[2] = qL[2] - 2. * rhoV * ny;
        qR[3] = qL[3];
    } else if (kind == ORC_WALL) { /* physics.h:399-402 */
        qR[0] = qL[0];
        qR[1] = -qL[1];
        qR[2] = -qL[2];
        qR[3] = qL[3];
    } else { /* farfield, physics.h:465-527 */
        const double rho = qL[0], rho_u = qL[1], rho_v = qL[2], rho_e = qL[3];
        const double bc_rho = qbc[0];
        const double bc_u = qbc[1] / bc_rho;
        const double bc_v = qbc[2] / bc_rho;
        const double bc_p = (gam - 1) * (qbc[3] - 0.5 / bc_rho * (qbc[1] * qbc[1] + qbc[2] * qbc[2]));
        const double p = (gam - 1) * (rho_e - 0.5 / rho * (rho_u * rho_u + rho_v * rho_v));
        const double c = sqrt(gam * p / rho);
        const double mach = sqrt(rho_u * rho_u + rho_v * rho_v) / (rho * c);
        const double io = rho_u * nx + rho_v * ny;
        if (mach > 1) {
            if (io < 0) {
                qR[0] = bc_rho;
                qR[1] = bc_rho * bc_u;
                qR[2] = bc_rho * bc_v;
                qR[3] = bc_p / (gam - 1) + 0.5 * bc_rho * (bc_u * bc_u + bc_v * bc_v);
            } else {
                qR[0] = rho; qR[1] = rho_u; qR[2] = rho_v; qR[3] = rho_e;
            }
        } else {
            const double pa = bc_p, rhoa = bc_rho, ua = bc_u, va = bc_v;
            const double pd = p, rhod = rho, ud = rho_u / rho, vd = rho_v / rho;
            const double rho0 = rho, c0 = c;
            if (io < 0) { /* subsonic inlet, physics.h:512-518 */
                const double pb = 0.5 * (pa + pd - rho0 * c0 * (nx * (ua - ud) + ny * (va - vd)));
                qR[0] = rhoa + (pb - pa) / (c0 * c0);
                qR[1] = qR[0] * (ua - nx * (pa - pb) / (rho0 * c0));
                qR[2] = qR[0] * (va - ny * (pa - pb) / (rho0 * c0));
                qR[3] = pb / (gam - 1) + 0.5 / qR[0] * (qR[1] * qR[1] + qR[2] * qR[2]);
            } else { /* subsonic outlet, physics.h:519-526; `va` is the reference's */
                const double pb = pa;
                qR[0] = rhod + (pb - pd) / (c0 * c0);
                qR[1] = qR[0] * (ud + nx * (pd - pb) / (rho0 * c0));
                qR[2] = qR[0] * (va + ny * (pd - pb) / (rho0 * c0));
                qR[3] = pb / (gam - 1) + 0.5 / qR[0] * (qR[1] * qR[1] + qR[2] * qR[2]);
            }
        }
    }
}

/* (*edges_flux_functions[e])(qL, qR, gx, gy): internal faces take the
 * gradients, boundary faces build their ghost state and call the Roe flux
 * with zero gradients (physics.h:296-298, 361-363, 426-431). */
void orc_flux(int kind, const orc_gas *g, int viscous_type, double nx, double ny,
              const double qL[4], const double qR[4],
              const double gx[4], const double gy[4], double f[4])
{
    if (kind == ORC_INTERNAL) {
        orc_flux_internal(g, viscous_type, nx, ny, qL, qR, gx, gy, f);
    } else {
        static const double zero[4] = {0, 0, 0, 0};
        double qb[4];
        orc_bc_vars(kind, g, nx, ny, qL, qR, qb);
        orc_flux_internal(g, viscous_type, nx, ny, qL, qb, zero, zero, f);
    }
}

/* Forward-difference 8x8 flux Jacobian, row-major J[r*8+c]. physics.h:533-578.
 * The perturbed entry is restored by subtraction, so later columns see the
 * rounding residue exactly as the reference does (physics.h:556-558,568-570). */
void orc_fd_jacobian(int kind, const orc_gas *g, int viscous_type, double nx, double ny,
                     const double qL_in[4], const double qR_in[4],
                     const double gx[4], const double gy[4], double J[64])
{
    double qL[4], qR[4], f[4], fp[4];
    memcpy(qL, qL_in, sizeof qL);
    memcpy(qR, qR_in, sizeof qR);
    orc_flux(kind, g, viscous_type, nx, ny, qL, qR, gx, gy, f);
    for (int i = 0; i < 4; ++i) {
        {
            const double h = fmax(1e-6, fabs(qL[i]) * 1e-6);
            qL[i] += h;
            orc_flux(kind, g, viscous_type, nx, ny, qL, qR, gx, gy, fp);
            qL[i] -= h;
            for (int r = 0; r < 4; ++r) {
                J[r * 8 + i] = (fp[r] - f[r]) / h;
                J[(r + 4) * 8 + i] = -J[r * 8 + i];
            }
        }
        {
            const double h = fmax(1e-6, fabs(qR[i]) * 1e-6);
            qR[i] += h;
            orc_flux(kind, g, viscous_type, nx, ny, qL, qR, gx, gy, fp);
            qR[i] -= h;
            for (int r = 0; r < 4; ++r) {
                J[r * 8 + i + 4] = (fp[r] - f[r]) / h;
                J[(r + 4) * 8 + i + 4] = -J[r * 8 + i + 4];
            }
        }
    }
}

/* core.h:73-83 */
void orc_get_conservative(const orc_bvars *b, const orc_gas *g, double q[4])
{
    const double c = sqrt(g->gamma * g->R * b->T);
    const double u = b->mach * c * cos(b->angle);
    const double v = b->mach * c * sin(b->angle);
    const double rho = b->p / (g->R * b->T);
    const double rhoE = b->p / (g->gamma - 1) + 0.5 * rho * (u * u + v * v);
    q[0] = rho; q[1] = rho * u; q[2] = rho * v; q[3] = rhoE;
}

/* ------------------------------------------------------------------ */
/* mesh.h                                                              */
/* ------------------------------------------------------------------ */

typedef struct { uint64_t *key; uint32_t *val; size_t cap; } edge_map;

static uint64_t pair_key_(uint32_t a, uint32_t b)
{
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a; /* mesh.h:291-292 */
    return ((uint64_t)lo << 32) | hi;
}
static size_t hash_(uint64_t k, size_t cap)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (size_t)(k & (cap - 1));
}
static uint32_t map_find_(const edge_map *m, uint64_t k)
{
    for (size_t i = hash_(k, m->cap);; i = (i + 1) & (m->cap - 1)) {
        if (m->val[i] == ORC_EDGE_NULL) return ORC_EDGE_NULL;
        if (m->key[i] == k) return m->val[i];
    }
}
static void map_put_(edge_map *m, uint64_t k, uint32_t v)
{
    size_t i = hash_(k, m->cap);
    while (m->val[i] != ORC_EDGE_NULL) i = (i + 1) & (m->cap - 1);
    m->key[i] = k; m->val[i] = v;
}

int orc_mesh_build(orc_mesh *m, uint32_t nn, const double *x, const double *y,
                   uint32_t nc, const uint32_t *cells, const uint8_t *is_tri,
                   uint32_t nb, const uint32_t *b0, const uint32_t *b1,
                   const int32_t *bpatch)
{
    (void)nn;
    memset(m, 0, sizeof *m);
    const size_t emax = (size_t)4 * nc + 1;
    edge_map em;
    em.cap = 16; while (em.cap < 2 * emax) em.cap <<= 1;
    em.key = (uint64_t *)malloc(em.cap * sizeof(uint64_t));
    em.val = (uint32_t *)malloc(em.cap * sizeof(uint32_t));
    memset(em.val, 0xFF, em.cap * sizeof(uint32_t));

    uint32_t *ecell = (uint32_t *)calloc(2 * emax, sizeof(uint32_t));
    uint32_t *enode = (uint32_t *)calloc(2 * emax, sizeof(uint32_t));
    uint32_t E = 0;
    /* add_cell_edges, mesh.h:317-341, called for cells in order (mesh.h:844-846) */
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t sz = is_tri[c] ? 3 : 4;
        for (uint32_t i = 0; i < sz; ++i) {
            const uint32_t j = (i < sz - 1) ? i + 1 : 0;
            const uint32_t a = cells[4 * c + i], b = cells[4 * c + j];
            const uint64_t k = pair_key_(a, b);
            if (map_find_(&em, k) == ORC_EDGE_NULL) {
                ecell[2 * E] = c; ecell[2 * E + 1] = 0;
                enode[2 * E] = a; enode[2 * E + 1] = b;
                map_put_(&em, k, E);
                ++E;
            }
        }
    }
    m->N = nc; m->G = nb; m->E = E;
    const size_t NT = (size_t)nc + nb;
    m->edge_cells = (uint32_t *)malloc(2 * (size_t)E * sizeof(uint32_t));
    memcpy(m->edge_cells, ecell, 2 * (size_t)E * sizeof(uint32_t));
    m->enx = (double *)calloc(E, sizeof(double)); m->eny = (double *)calloc(E, sizeof(double));
    m->elen = (double *)calloc(E, sizeof(double));
    m->ecx = (double *)calloc(E, sizeof(double)); m->ecy = (double *)calloc(E, sizeof(double));
    m->ccx = (double *)calloc(NT, sizeof(double)); m->ccy = (double *)calloc(NT, sizeof(double));
    m->area = (double *)calloc(NT, sizeof(double));
    m->cell_edges = (uint32_t *)malloc(4 * (size_t)nc * sizeof(uint32_t));
    m->is_tri = (uint8_t *)malloc(NT);
    m->bnd_edge = (uint32_t *)malloc((size_t)(nb ? nb : 1) * sizeof(uint32_t));
    m->bnd_patch = (int32_t *)malloc((size_t)(nb ? nb : 1) * sizeof(int32_t));
    memcpy(m->is_tri, is_tri, nc);

    /* convert_node_face_info, mesh.h:344-375 */
    for (uint32_t c = 0; c < nc; ++c) {
        uint32_t ce[4];
        ce[0] = map_find_(&em, pair_key_(cells[4 * c + 0], cells[4 * c + 1]));
        ce[1] = map_find_(&em, pair_key_(cells[4 * c + 1], cells[4 * c + 2]));
        if (is_tri[c]) {
            ce[2] = map_find_(&em, pair_key_(cells[4 * c + 2], cells[4 * c + 0]));
            ce[3] = ORC_EDGE_NULL;
        } else {
            ce[2] = map_find_(&em, pair_key_(cells[4 * c + 2], cells[4 * c + 3]));
            ce[3] = map_find_(&em, pair_key_(cells[4 * c + 3], cells[4 * c + 0]));
        }
        for (int k = 0; k < 4; ++k) {
            m->cell_edges[4 * c + k] = ce[k];
            if (ce[k] != ORC_EDGE_NULL && c != m->edge_cells[2 * ce[k]]) m->edge_cells[2 * ce[k] + 1] = c;
        }
    }
    /* compute_mesh, mesh.h:378-453 */
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t sz = is_tri[c] ? 3 : 4;
        double cx = 0., cy = 0.;
        for (uint32_t j = 0; j < sz; ++j) {
            cx += x[cells[4 * c + j]] / ((double)sz);
            cy += y[cells[4 * c + j]] / ((double)sz);
        }
        m->ccx[c] = cx; m->ccy[c] = cy;
    }
    for (uint32_t e = 0; e < E; ++e) {
        const uint32_t n0 = enode[2 * e], n1 = enode[2 * e + 1];
        m->ecx[e] = (x[n1] + x[n0]) * 0.5;
        m->ecy[e] = (y[n1] + y[n0]) * 0.5;
        const double dex = x[n1] - x[n0], dey = y[n1] - y[n0];
        const double l = sqrt(dex * dex + dey * dey);
        m->elen[e] = l;
        m->enx[e] = -dey / l;
        m->eny[e] = dex / l;
        const uint32_t c0 = m->edge_cells[2 * e];
        const double dot = m->enx[e] * (m->ecx[e] - m->ccx[c0]) + m->eny[e] * (m->ecy[e] - m->ccy[c0]);
        if (dot < 0) { m->enx[e] *= -1.; m->eny[e] *= -1.; }
    }
    for (uint32_t c = 0; c < nc; ++c) {
        const double x1 = x[cells[4 * c]], x2 = x[cells[4 * c + 1]], x3 = x[cells[4 * c + 2]];
        const double y1 = y[cells[4 * c]], y2 = y[cells[4 * c + 1]], y3 = y[cells[4 * c + 2]];
        if (is_tri[c]) {
            m->area[c] = 0.5 * fabs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2));
        } else {
            const double x4 = x[cells[4 * c + 3]], y4 = y[cells[4 * c + 3]];
            m->area[c] = 0.5 * fabs(x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))
                       + 0.5 * fabs(x1 * (y3 - y4) + x3 * (y4 - y1) + x4 * (y1 - y3));
        }
    }
    /* add_boundary_cells, mesh.h:744-787 */
    int rc = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t e = map_find_(&em, pair_key_(b0[b], b1[b]));
        if (e == ORC_EDGE_NULL) { rc = -1; break; }
        const uint32_t c = m->edge_cells[2 * e];
        const double dx = m->ecx[e] - m->ccx[c], dy = m->ecy[e] - m->ccy[c];
        const double dist = sqrt(dx * dx + dy * dy);
        m->area[nc + b] = m->area[c];
        m->ccx[nc + b] = m->ecx[e] + dist * m->enx[e];
        m->ccy[nc + b] = m->ecy[e] + dist * m->eny[e];
        m->is_tri[nc + b] = 1;
        m->edge_cells[2 * e + 1] = nc + b;
        m->bnd_edge[b] = e;
        m->bnd_patch[b] = bpatch ? bpatch[b] : 0;
    }
    free(em.key); free(em.val); free(ecell); free(enode);
    if (rc) orc_mesh_free(m);
    return rc;
}

void orc_mesh_free(orc_mesh *m)
{
    free(m->edge_cells); free(m->enx); free(m->eny); free(m->elen); free(m->ecx); free(m->ecy);
    free(m->ccx); free(m->ccy); free(m->area); free(m->cell_edges); free(m->is_tri);
    free(m->bnd_edge); free(m->bnd_patch);
    memset(m, 0, sizeof *m);
}

/* ------------------------------------------------------------------ */
/* solver.h                                                            */
/* ------------------------------------------------------------------ */

/* set_mesh_and_gas, solver.h:177-197.  The reference leaves the vectors
 * uninitialised; the oracle zero-fills them. */
int orc_solver_init(orc_solver *s, const orc_mesh *m, const orc_gas *g, int viscosity_model)
{
    memset(s, 0, sizeof *s);
    s->m = *m;
    s->g = *g;
    /* solver.h:203-208: only "laminar" reaches a non-zero flux type; the
     * settings string "spallart-allmaras" never matches "spalart-allmaras". */
    s->viscous_type = (viscosity_model == 1) ? 1 : 0;
    s->visc_not_inviscid = (viscosity_model != 0);
    s->second_order = 1;
    s->gradient_scheme = ORC_GREEN_GAUSS;
    s->limiter_k = 5.;
    s->limiter_kind = 0;
    s->cfl = 1;
    const size_t NT = (size_t)m->N + m->G, n4 = 4 * NT;
    double **v[] = {&s->q, &s->qk, &s->qW, &s->gx, &s->gy, &s->lim, &s->qmin, &s->qmax, &s->rhs};
    for (size_t i = 0; i < sizeof v / sizeof v[0]; ++i) {
        *v[i] = (double *)calloc(n4, sizeof(double));
        if (!*v[i]) return -1;
    }
    s->dt = (double *)calloc(NT, sizeof(double));
    s->edge_kind = (uint8_t *)calloc(m->E ? m->E : 1, 1);
    s->bnd_kind = (uint8_t *)calloc(m->G ? m->G : 1, 1);
    s->bnd_vars = (orc_bvars *)calloc(m->G ? m->G : 1, sizeof(orc_bvars));
    s->lsq = NULL;
    return 0;
}

void orc_solver_free(orc_solver *s)
{
    free(s->q); free(s->qk); free(s->qW); free(s->gx); free(s->gy); free(s->lim);
    free(s->qmin); free(s->qmax); free(s->rhs); free(s->dt);
    free(s->edge_kind); free(s->bnd_kind); free(s->bnd_vars); free(s->lsq); free(s->cf_sorted); free(s->fluxbuf);
    memset(s, 0, sizeof *s);
}

/* solver.h:200-247 */
void orc_set_bcs(orc_solver *s, int npatch, const uint8_t *patch_kind, const orc_bvars *patch_vars)
{
    const orc_mesh *m = &s->m;
    memset(s->edge_kind, 0, m->E);
    for (uint32_t b = 0; b < m->G; ++b) {
        const int p = m->bnd_patch[b];
        const uint8_t kind = (p >= 0 && p < npatch) ? patch_kind[p] : 0;
        orc_bvars dflt = {0.2, 0, 1, 1.}; /* core.h:62-65 */
        s->bnd_kind[b] = kind;
        s->bnd_vars[b] = (kind == ORC_FARFIELD) ? patch_vars[p] : dflt; /* solver.h:219-229 */
        s->edge_kind[m->bnd_edge[b]] = kind <= ORC_WALL ? kind : ORC_INTERNAL; /* "inlet-outlet" has no flux class: solver.h:237-245 */
    }
}

/* solver.h:402-422 */
static void lsq_matrices_(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    free(s->lsq);
    s->lsq = (double *)malloc(4 * (size_t)m->N * sizeof(double));
    for (uint32_t i = 0; i < m->N; ++i) {
        const uint32_t sz = m->is_tri[i] ? 3 : 4;
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
        for (uint32_t j = 0; j < sz; ++j) {
            const uint32_t e = m->cell_edges[4 * i + j];
            const uint32_t n = m->edge_cells[2 * e] == i ? m->edge_cells[2 * e + 1] : m->edge_cells[2 * e];
            const double d0 = m->ccx[n] - m->ccx[i], d1 = m->ccy[n] - m->ccy[i];
            a00 += d0 * d0; a01 += d0 * d1; a10 += d1 * d0; a11 += d1 * d1;
        }
        /* 2x2 inverse as 1/det * adjugate */
        const double det = a00 * a11 - a10 * a01;
        const double inv = 1. / det;
        s->lsq[4 * i + 0] = a11 * inv;
        s->lsq[4 * i + 1] = -a01 * inv;
        s->lsq[4 * i + 2] = -a10 * inv;
        s->lsq[4 * i + 3] = a00 * inv;
    }
}

/* solver.h:149-152 */
void orc_set_gradient_scheme(orc_solver *s, int scheme)
{
    s->gradient_scheme = scheme;
    if (scheme == ORC_LEAST_SQUARES) lsq_matrices_(s);
}

/* solver.h:597-611 */
int orc_boundary_variables(const orc_solver *s, orc_bvars *out)
{
    orc_bvars dflt = {0.2, 0, 1, 1.};
    *out = dflt;
    for (uint32_t b = 0; b < s->m.G; ++b) {
        if (s->bnd_kind[b] == ORC_FARFIELD) { *out = s->bnd_vars[b]; return 1; }
        if (s->bnd_kind[b] == ORC_INLET_OUTLET) { *out = s->bnd_vars[b]; return 0; } /* solver.h:603-606: the defaults */
    }
    return 0;
}

/* solver.h:615-631 */
void orc_init_field(orc_solver *s)
{
    orc_bvars v; double q0[4];
    orc_boundary_variables(s, &v);
    orc_get_conservative(&v, &s->g, q0);
    for (uint32_t i = 0; i < s->m.N; ++i) memcpy(s->q + 4 * (size_t)i, q0, sizeof q0);
}

/* solver.h:259-274 */
void orc_refill_bcs(orc_solver *s)
{
    for (uint32_t b = 0; b < s->m.G; ++b) {
        const uint32_t c1 = s->m.edge_cells[2 * s->m.bnd_edge[b] + 1];
        orc_get_conservative(&s->bnd_vars[b], &s->g, s->q + 4 * (size_t)c1);
    }
}

/* solver.h:276-287 */
void orc_bcs_from_internal(orc_solver *s)
{
    for (uint32_t b = 0; b < s->m.G; ++b) {
        const uint32_t e = s->m.bnd_edge[b];
        memcpy(s->q + 4 * (size_t)s->m.edge_cells[2 * e + 1], s->q + 4 * (size_t)s->m.edge_cells[2 * e], 32);
    }
}

/* solver.h:289-305 */
void orc_set_walls_from_internal(orc_solver *s, double *q_)
{
    for (uint32_t b = 0; b < s->m.G; ++b) {
        if (s->bnd_kind[b] == ORC_WALL || s->bnd_kind[b] == ORC_SLIPWALL) {
            const uint32_t e = s->m.bnd_edge[b];
            memcpy(q_ + 4 * (size_t)s->m.edge_cells[2 * e + 1], q_ + 4 * (size_t)s->m.edge_cells[2 * e], 32);
        }
    }
}

static int two_sided_(int kind) { return kind == ORC_INTERNAL; } /* physics.h:138,284,349,414 */

/* solver.h:308-356 */
void orc_calc_dt(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    const double *q = s->q;
    const double gam = s->g.gamma;
    const size_t NT = (size_t)m->N + m->G;
    for (size_t i = 0; i < NT; ++i) s->dt[i] = 0;
    for (uint32_t e = 0; e < m->E; ++e) {
        const uint32_t c0 = m->edge_cells[2 * e], c1 = m->edge_cells[2 * e + 1];
        const size_t k0 = 4 * (size_t)c0, k1 = 4 * (size_t)c1;
        const double nx = m->enx[e], ny = m->eny[e];
        double eig;
        if (two_sided_(s->edge_kind[e])) {
            const double V_L = (q[k0 + 1] * nx + q[k0 + 2] * ny) / q[k0];
            const double V_R = (q[k1 + 1] * nx + q[k1 + 2] * ny) / q[k1];
            const double p_L = (gam - 1) * (q[k0 + 3] - 0.5 / q[k0] * (q[k0 + 1] * q[k0 + 1] + q[k0 + 2] * q[k0 + 2]));
            const double p_R = (gam - 1) * (q[k1 + 3] - 0.5 / q[k1] * (q[k1 + 1] * q[k1 + 1] + q[k1 + 2] * q[k1 + 2]));
            const double eig_L = sqrt(p_L * gam / q[k0]) + fabs(V_L);
            const double eig_R = sqrt(p_R * gam / q[k1]) + fabs(V_R);
            eig = (eig_L < eig_R) ? eig_R : eig_L; /* std::max(eig_L, eig_R) */
        } else {
            const double V_L = (q[k0 + 1] * nx + q[k0 + 2] * ny) / q[k0];
            const double p_L = (gam - 1) * (q[k0 + 3] - 0.5 / q[k0] * (q[k0 + 1] * q[k0 + 1] + q[k0 + 2] * q[k0 + 2]));
            eig = sqrt(p_L * gam / q[k0]) + fabs(V_L);
        }
        s->dt[c0] += eig * m->elen[e];
        if (two_sided_(s->edge_kind[e])) s->dt[c1] += eig * m->elen[e];
    }
    for (size_t i = 0; i < NT; ++i) s->dt[i] = s->cfl * m->area[i] / s->dt[i];
}

/* solver.h:425-514.  Differentiates the member q whatever the caller passes
 * (SURVEY F5); face states come from flux::vars. */
void orc_calc_gradients(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    const double *q = s->q;
    const size_t NT = (size_t)m->N + m->G;
    if (s->gradient_scheme == ORC_GREEN_GAUSS) {
        memset(s->gx, 0, 4 * NT * sizeof(double));
        memset(s->gy, 0, 4 * NT * sizeof(double));
        for (uint32_t e = 0; e < m->E; ++e) {
            const uint32_t i = m->edge_cells[2 * e], j = m->edge_cells[2 * e + 1];
            const double dxif = m->ecx[e] - m->ccx[i], dyif = m->ecy[e] - m->ccy[i];
            const double dif = sqrt(dxif * dxif + dyif * dyif);
            const double dxij = m->ccx[i] - m->ccx[j], dyij = m->ccy[i] - m->ccy[j];
            const double dij = sqrt(dxij * dxij + dyij * dyij);
            const double w = dif / dij;
            const double *qL = q + 4 * (size_t)i;
            double qR[4];
            orc_bc_vars(s->edge_kind[e], &s->g, m->enx[e], m->eny[e], qL, q + 4 * (size_t)j, qR);
            for (int k = 0; k < 4; ++k) {
                const double fk = (qL[k] * (1.0 - w) + qR[k] * w) * m->elen[e];
                s->gx[4 * (size_t)i + k] += fk * m->enx[e];
                s->gy[4 * (size_t)i + k] += fk * m->eny[e];
                s->gx[4 * (size_t)j + k] -= fk * m->enx[e];
                s->gy[4 * (size_t)j + k] -= fk * m->eny[e];
            }
        }
        for (uint32_t i = 0; i < m->N; ++i)
            for (int k = 0; k < 4; ++k) {
                s->gx[4 * (size_t)i + k] /= m->area[i];
                s->gy[4 * (size_t)i + k] /= m->area[i];
            }
        for (size_t i = m->N; i < NT; ++i)
            for (int k = 0; k < 4; ++k) { s->gx[4 * i + k] *= 0; s->gy[4 * i + k] *= 0; }
    } else {
        for (uint32_t i = 0; i < m->N; ++i) {
            const uint32_t sz = m->is_tri[i] ? 3 : 4;
            double d[4][2], dq[4][4];
            const double *qL = q + 4 * (size_t)i;
            for (uint32_t j = 0; j < sz; ++j) {
                const uint32_t e = m->cell_edges[4 * i + j];
                const uint32_t n = m->edge_cells[2 * e] == i ? m->edge_cells[2 * e + 1] : m->edge_cells[2 * e];
                d[j][0] = m->ccx[n] - m->ccx[i];
                d[j][1] = m->ccy[n] - m->ccy[i];
                double qR[4];
                orc_bc_vars(s->edge_kind[e], &s->g, m->enx[e], m->eny[e], qL, q + 4 * (size_t)n, qR);
                for (int k = 0; k < 4; ++k) dq[j][k] = qL[k] - qR[k];
            }
            /* grad = M * dT * delta, evaluated (M*dT) first then times the column */
            const double *M = s->lsq + 4 * (size_t)i;
            double MdT[2][4];
            for (uint32_t j = 0; j < sz; ++j) {
                MdT[0][j] = M[0] * d[j][0] + M[1] * d[j][1];
                MdT[1][j] = M[2] * d[j][0] + M[3] * d[j][1];
            }
            for (int k = 0; k < 4; ++k) {
                double g0 = 0, g1 = 0;
                for (uint32_t j = 0; j < sz; ++j) { g0 += MdT[0][j] * dq[j][k]; g1 += MdT[1][j] * dq[j][k]; }
                s->gx[4 * (size_t)i + k] = g0;
                s->gy[4 * (size_t)i + k] = g1;
            }
        }
        for (size_t i = m->N; i < NT; ++i)
            for (int k = 0; k < 4; ++k) { s->gx[4 * i + k] *= 0; s->gy[4 * i + k] *= 0; }
    }
}

/* michalak_limiter, physics.h:581-592 (RANS_YT = 2.0; the constants as the reference writes them) */
static double orc_michalak_phi(double y)
{
    if (y >= 2.0) return 1.0;
    const double a = 1.0 / (2.0 * 2.0) - 2.0 / (2.0 * 2.0 * 2.0);
    const double b = -3.0 / 2.0 * a * 2.0 - 0.5 / 2.0;
    return a * y * y * y + b * y * y + y;
}
/* the limiter of one component at one face under RANS_MICHALAK_LIMITER, solver.h:553-576 */
static double orc_michalak_one(double dqg, double delta_max, double delta_min, double K3a)
{
    const double dMaxMin2 = (delta_max - delta_min) * (delta_max - delta_min);
    double lim = 1.0, sig;
    if (dMaxMin2 <= K3a) sig = 1.;
    else if (dMaxMin2 <= 2 * K3a) { const double y = (dMaxMin2 / K3a - 1.0); sig = 2.0 * y * y * y - 3.0 * y * y + 1.0; }
    else sig = 0.;
    if (sig < 1.0) {
        if (dqg > 1e-14) lim = orc_michalak_phi(delta_max / dqg);
        else if (dqg < -1e-14) lim = orc_michalak_phi(delta_min / dqg);
        else lim = 1.0;
    }
    return sig + (1.0 - sig) * lim;
}

/* calc_limiters, solver.h:517-593: Venkatakrishnan (the default build) or, limiter_kind = 1, the RANS_MICHALAK_LIMITER build */
void orc_calc_limiters(orc_solver *s, const double *q_)
{
    const orc_mesh *m = &s->m;
    const size_t n4 = 4 * ((size_t)m->N + m->G);
    for (size_t i = 0; i < n4; ++i) { s->lim[i] = 1; s->qmin[i] = q_[i]; s->qmax[i] = q_[i]; }
    for (uint32_t e = 0; e < m->E; ++e) {
        const size_t i = 4 * (size_t)m->edge_cells[2 * e], j = 4 * (size_t)m->edge_cells[2 * e + 1];
        for (int k = 0; k < 4; ++k) {
            s->qmin[i + k] = fmin(s->qmin[i + k], q_[j + k]);
            s->qmin[j + k] = fmin(s->qmin[j + k], q_[i + k]);
            s->qmax[i + k] = fmax(s->qmax[i + k], q_[j + k]);
            s->qmax[j + k] = fmax(s->qmax[j + k], q_[i + k]);
        }
    }
    for (uint32_t e = 0; e < m->E; ++e) {
        for (int side = 0; side < 2; ++side) {
            const uint32_t id = m->edge_cells[2 * e + side];
            if (id >= m->N) continue;
            const double dx = m->ecx[e] - m->ccx[id], dy = m->ecy[e] - m->ccy[id];
            const double sqrt_area = sqrt(m->area[id]);
            for (int k = 0; k < 4; ++k) {
                const size_t ik = 4 * (size_t)id + k;
                const double dqg = s->gx[ik] * dx + s->gy[ik] * dy;
                const double dmax = s->qmax[ik] - q_[ik];
                const double dmin = s->qmin[ik] - q_[ik];
                const double Ka = s->limiter_k * sqrt_area;
                const double K3a = Ka * Ka * Ka;
                double lim = 1.0;
                if (s->limiter_kind == 1)
                    lim = orc_michalak_one(dqg, dmax, dmin, K3a);
                else if (dqg > 1e-16)
                    lim = 1 / dqg * ((dmax * dmax + K3a) * dqg + 2 * dqg * dqg * dmax) / (dmax * dmax + 2 * dqg * dqg + dmax * dqg + K3a);
                else if (dqg < -1e-16)
                    lim = 1 / dqg * ((dmin * dmin + K3a) * dqg + 2 * dqg * dqg * dmin) / (dmin * dmin + 2 * dqg * dqg + dmin * dqg + K3a);
                s->lim[ik] = fmin(s->lim[ik], lim);
            }
        }
    }
}

/* solver.h:359-398 ("direction" from the SUM of centres, member q) */
void orc_average_gradients(const orc_solver *s, uint32_t c0, uint32_t c1, double gradx[4], double grady[4])
{
    const orc_mesh *m = &s->m;
    if (c0 == c1) {
        for (int i = 0; i < 4; ++i) { gradx[i] = s->gx[4 * (size_t)c0 + i]; grady[i] = s->gy[4 * (size_t)c0 + i]; }
        return;
    }
    double t0 = m->ccx[c0] + m->ccx[c1], t1 = m->ccy[c0] + m->ccy[c1];
    const double l = sqrt(t0 * t0 + t1 * t1);
    t0 /= l; t1 /= l;
    for (int i = 0; i < 4; ++i) {
        const double gdir = (s->q[4 * (size_t)c0 + i] - s->q[4 * (size_t)c1 + i]) / l;
        const double bx = (s->gx[4 * (size_t)c0 + i] + s->gx[4 * (size_t)c1 + i]) * 0.5;
        const double by = (s->gy[4 * (size_t)c0 + i] + s->gy[4 * (size_t)c1 + i]) * 0.5;
        const double dot = bx * t0 + by * t1;
        gradx[i] = bx - (dot - gdir) * t0;
        grady[i] = by - (dot - gdir) * t1;
    }
}

/* the face loop shared by explicitSolver::calc_residual (solver.h:751-792) and
 * implicitSolver::fillRhoRHS (solver.h:1097-1141) */
static void face_loop_(orc_solver *s, const double *q_, double *out)
{
    const orc_mesh *m = &s->m;
    for (uint32_t e = 0; e < m->E; ++e) {
        const uint32_t c0 = m->edge_cells[2 * e], c1 = m->edge_cells[2 * e + 1];
        const size_t k0 = 4 * (size_t)c0, k1 = 4 * (size_t)c1;
        double gradx[4], grady[4], f[4];
        orc_average_gradients(s, c0, c1, gradx, grady);
        if (s->second_order) {
            const double d0x = m->ecx[e] - m->ccx[c0], d0y = m->ecy[e] - m->ccy[c0];
            const double d1x = m->ecx[e] - m->ccx[c1], d1y = m->ecy[e] - m->ccy[c1];
            double qL[4], qR[4];
            for (int k = 0; k < 4; ++k) {
                qL[k] = q_[k0 + k] + (s->gx[k0 + k] * d0x + s->gy[k0 + k] * d0y) * s->lim[k0 + k];
                qR[k] = q_[k1 + k] + (s->gx[k1 + k] * d1x + s->gy[k1 + k] * d1y) * s->lim[k1 + k];
            }
            orc_flux(s->edge_kind[e], &s->g, s->viscous_type, m->enx[e], m->eny[e], qL, qR, gradx, grady, f);
        } else {
            orc_flux(s->edge_kind[e], &s->g, s->viscous_type, m->enx[e], m->eny[e], q_ + k0, q_ + k1, gradx, grady, f);
        }
        for (int i = 0; i < 4; ++i) {
            const double fl = f[i] * m->elen[e];
            out[k0 + i] -= fl;
            if (two_sided_(s->edge_kind[e])) out[k1 + i] += fl;
        }
    }
}

/* solver.h:745-799 */
void orc_explicit_residual(orc_solver *s, const double *q_)
{
    const orc_mesh *m = &s->m;
    memset(s->qW, 0, 4 * ((size_t)m->N + m->G) * sizeof(double));
    face_loop_(s, q_, s->qW);
    for (uint32_t i = 0; i < m->N; ++i)
        for (int j = 0; j < 4; ++j) s->qW[4 * (size_t)i + j] /= m->area[i];
}

double orc_norm(const double *v, size_t n)
{
    double a = 0;
    for (size_t i = 0; i < n; ++i) a += v[i] * v[i];
    return sqrt(a);
}

/* explicitSolver::solve, solver.h:802-828 */
double orc_explicit_solve(orc_solver *s, double relaxation)
{
    static const double alpha[3] = {0.25, 0.5, 1.}; /* solver.h:723 */
    const orc_mesh *m = &s->m;
    const size_t n4 = 4 * ((size_t)m->N + m->G);
    orc_calc_dt(s);
    memcpy(s->qk, s->q, n4 * sizeof(double));
    for (int st = 0; st < 3; ++st) {
        if (s->visc_not_inviscid | s->second_order) {
            orc_set_walls_from_internal(s, s->qk);
            orc_calc_gradients(s);
            if (s->second_order) orc_calc_limiters(s, s->qk);
        }
        orc_explicit_residual(s, s->qk);
        for (uint32_t i = 0; i < m->N; ++i)
            for (int j = 0; j < 4; ++j)
                s->qk[4 * (size_t)i + j] = s->q[4 * (size_t)i + j] + s->qW[4 * (size_t)i + j] * s->dt[i] * alpha[st] * relaxation;
    }
    memcpy(s->q, s->qk, n4 * sizeof(double));
    return orc_norm(s->qW, n4);
}

/* implicitSolver::fillRhoRHS, solver.h:1079-1152; returns RhoVector.norm() */
double orc_implicit_rhs(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    const size_t n4 = 4 * ((size_t)m->N + m->G);
    orc_calc_dt(s);
    if (s->second_order | s->visc_not_inviscid) {
        orc_set_walls_from_internal(s, s->q);
        orc_calc_gradients(s);
    }
    if (s->second_order) orc_calc_limiters(s, s->q);
    memset(s->rhs, 0, n4 * sizeof(double));
    face_loop_(s, s->q, s->rhs);
    for (size_t i = 4 * (size_t)m->N; i < n4; ++i) s->rhs[i] = 0;
    return orc_norm(s->rhs, n4);
}

/* solver.h:636-690.  The reference accumulates into whatever qW held
 * (SURVEY F9); the oracle starts from zero. */
double orc_uniform_residual(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    const size_t n4 = 4 * ((size_t)m->N + m->G);
    orc_bvars v; double qf[4];
    orc_boundary_variables(s, &v);
    orc_get_conservative(&v, &s->g, qf);
    memset(s->qW, 0, n4 * sizeof(double));
    for (uint32_t e = 0; e < m->E; ++e) {
        const uint32_t c0 = m->edge_cells[2 * e], c1 = m->edge_cells[2 * e + 1];
        double gradx[4], grady[4], f[4];
        orc_average_gradients(s, c0, c1, gradx, grady);
        orc_flux(s->edge_kind[e], &s->g, s->viscous_type, m->enx[e], m->eny[e], qf, qf, gradx, grady, f);
        for (int i = 0; i < 4; ++i) {
            const double fl = f[i] * m->elen[e];
            s->qW[4 * (size_t)c0 + i] -= fl;
            if (two_sided_(s->edge_kind[e])) s->qW[4 * (size_t)c1 + i] += fl;
        }
    }
    return orc_norm(s->qW, n4);
}

/* implicitSolver::fillRhoLHS, solver.h:979-1071, in 4x4 block form */
void orc_implicit_lhs(orc_solver *s, double *diag, double *off01, double *off10)
{
    const orc_mesh *m = &s->m;
    const size_t NT = (size_t)m->N + m->G;
    orc_calc_dt(s);
    if (s->visc_not_inviscid) {
        orc_set_walls_from_internal(s, s->q);
        orc_calc_gradients(s);
    }
    memset(diag, 0, 16 * NT * sizeof(double));
    memset(off01, 0, 16 * (size_t)m->E * sizeof(double));
    memset(off10, 0, 16 * (size_t)m->E * sizeof(double));
    for (size_t i = 0; i < NT; ++i)
        for (int j = 0; j < 4; ++j) diag[16 * i + 5 * j] = m->area[i] / s->dt[i];
    for (uint32_t e = 0; e < m->E; ++e) {
        const uint32_t c0 = m->edge_cells[2 * e], c1 = m->edge_cells[2 * e + 1];
        double gradx[4], grady[4], J[64];
        orc_average_gradients(s, c0, c1, gradx, grady);
        orc_fd_jacobian(s->edge_kind[e], &s->g, s->viscous_type, m->enx[e], m->eny[e],
                        s->q + 4 * (size_t)c0, s->q + 4 * (size_t)c1, gradx, grady, J);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                diag[16 * (size_t)c0 + 4 * i + j] += J[i * 8 + j] * m->elen[e];
                off01[16 * (size_t)e + 4 * i + j] += J[i * 8 + j + 4] * m->elen[e];
                if (two_sided_(s->edge_kind[e])) {
                    off10[16 * (size_t)e + 4 * i + j] += J[(i + 4) * 8 + j] * m->elen[e];
                    diag[16 * (size_t)c1 + 4 * i + j] += J[(i + 4) * 8 + j + 4] * m->elen[e];
                }
            }
    }
    for (size_t c = m->N; c < NT; ++c) /* ghost rows = identity, solver.h:1062-1070 */
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) diag[16 * c + 4 * i + j] = (i == j) ? 1 : 0;
}

/* get_wall_profile, post.h:301-387 -> {cl, cd, cm} */
int orc_wall_forces(const orc_solver *s, int patch, double out[3])
{
    const orc_mesh *m = &s->m;
    orc_bvars far;
    orc_boundary_variables(s, &far);
    const double p_inf = far.p, mach_inf = far.mach, gam = s->g.gamma;
    double xmin = 0., xmax = 0., y_moment = 0.;
    uint32_t n_added = 0;
    for (uint32_t b = 0; b < m->G; ++b) {
        if (m->bnd_patch[b] != patch) continue;
        const uint32_t e = m->bnd_edge[b];
        if (!n_added) { xmin = m->ecx[e]; xmax = m->ecx[e]; y_moment = m->ecy[e]; }
        else { xmin = fmin(xmin, m->ecx[e]); xmax = fmax(xmax, m->ecx[e]); y_moment += m->ecy[e]; }
        ++n_added;
    }
    if (!n_added) { out[0] = out[1] = out[2] = 0; return -1; }
    y_moment /= (double)n_added;
    const double x_moment = (xmax - xmin) * 0.25 + xmin;
    double cd = 0., cl = 0., cm = 0.;
    for (uint32_t b = 0; b < m->G; ++b) {
        if (m->bnd_patch[b] != patch) continue;
        const uint32_t e = m->bnd_edge[b];
        const double *qc = s->q + 4 * (size_t)m->edge_cells[2 * e];
        const double p = orc_pressure(qc, gam); /* core.h:145-147 */
        const double cp = 2. / (gam * mach_inf * mach_inf) * (p / p_inf - 1.);
        const double fxi = cp * m->enx[e] * m->elen[e] / (xmax - xmin);
        const double fyi = cp * m->eny[e] * m->elen[e] / (xmax - xmin);
        const double mi = (m->ecx[e] - x_moment) / (xmax - xmin) * fyi - (m->ecy[e] - y_moment) / (xmax - xmin) * fxi;
        cd += fxi; cl += fyi; cm -= mi;
    }
    const double fx = cd, fy = cl, aoa = far.angle;
    out[1] = fx * cos(aoa) + fy * sin(aoa);
    out[0] = -fx * sin(aoa) + fy * cos(aoa);
    out[2] = cm;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Threaded variant for TIMING the port on all host cores.  The reference's face loops are serial scatters
 * (solver.h:313,432,522,535,751); here every phase is a gather over cells (or a map over faces), with each
 * cell's faces visited in ascending edge id, so every sum has the reference's order and the result is
 * bit-identical to orc_explicit_solve.  Green-Gauss or least-squares, any flux kind.
 * ------------------------------------------------------------------ */
static void build_cf_(orc_solver *s)
{
    const orc_mesh *m = &s->m;
    s->cf_sorted = (uint32_t *)malloc(4 * (size_t)m->N * sizeof(uint32_t));
    s->fluxbuf = (double *)malloc(4 * (size_t)m->E * sizeof(double));
    for (uint32_t i = 0; i < m->N; ++i) {
        uint32_t e[4];
        const uint32_t sz = m->is_tri[i] ? 3 : 4;
        for (uint32_t k = 0; k < 4; ++k) e[k] = k < sz ? m->cell_edges[4 * (size_t)i + k] : ORC_EDGE_NULL;
        for (int a = 1; a < 4; ++a)
            for (int b = a; b > 0 && e[b] < e[b - 1]; --b) { const uint32_t t = e[b]; e[b] = e[b - 1]; e[b - 1] = t; }
        memcpy(s->cf_sorted + 4 * (size_t)i, e, sizeof e);
    }
}

static double spectral_(const double *q, double nx, double ny, double gam)
{
    const double V = (q[1] * nx + q[2] * ny) / q[0];
    const double p = (gam - 1) * (q[3] - 0.5 / q[0] * (q[1] * q[1] + q[2] * q[2]));
    return sqrt(p * gam / q[0]) + fabs(V);
}

double orc_explicit_solve_omp(orc_solver *s, double relaxation)
{
    static const double alpha[3] = {0.25, 0.5, 1.};
    const orc_mesh *m = &s->m;
    const size_t NT = (size_t)m->N + m->G, n4 = 4 * NT;
    const double gam = s->g.gamma;
    if (!s->cf_sorted) build_cf_(s);
    const double *q = s->q;
    /* calc_dt */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m->N; ++i) {
        double sum = 0;
        for (int k = 0; k < 4; ++k) {
            const uint32_t e = s->cf_sorted[4 * i + k];
            if (e == ORC_EDGE_NULL) break;
            const uint32_t c0 = m->edge_cells[2 * (size_t)e], c1 = m->edge_cells[2 * (size_t)e + 1];
            const double eL = spectral_(q + 4 * (size_t)c0, m->enx[e], m->eny[e], gam);
            double eig = eL;
            if (two_sided_(s->edge_kind[e])) { const double eR = spectral_(q + 4 * (size_t)c1, m->enx[e], m->eny[e], gam); eig = (eL < eR) ? eR : eL; }
            sum += eig * m->elen[e];
        }
        s->dt[i] = s->cfl * m->area[i] / sum;
    }
    memcpy(s->qk, s->q, n4 * sizeof(double));
    for (int st = 0; st < 3; ++st) {
        double *qk = s->qk;
        if (s->visc_not_inviscid | s->second_order) {
            orc_set_walls_from_internal(s, qk);
            if (s->gradient_scheme == ORC_GREEN_GAUSS) {
#pragma omp parallel for schedule(static)
                for (long i = 0; i < (long)m->N; ++i) {
                    double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0};
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t e = s->cf_sorted[4 * i + k];
                        if (e == ORC_EDGE_NULL) break;
                        const uint32_t c0 = m->edge_cells[2 * (size_t)e], c1 = m->edge_cells[2 * (size_t)e + 1];
                        const double dxif = m->ecx[e] - m->ccx[c0], dyif = m->ecy[e] - m->ccy[c0];
                        const double dxij = m->ccx[c0] - m->ccx[c1], dyij = m->ccy[c0] - m->ccy[c1];
                        const double w = sqrt(dxif * dxif + dyif * dyif) / sqrt(dxij * dxij + dyij * dyij);
                        const double *qL = q + 4 * (size_t)c0;
                        double qR[4];
                        orc_bc_vars(s->edge_kind[e], &s->g, m->enx[e], m->eny[e], qL, q + 4 * (size_t)c1, qR);
                        for (int c = 0; c < 4; ++c) {
                            const double fk = (qL[c] * (1.0 - w) + qR[c] * w) * m->elen[e];
                            if ((uint32_t)i == c0) { ax[c] += fk * m->enx[e]; ay[c] += fk * m->eny[e]; }
                            else { ax[c] -= fk * m->enx[e]; ay[c] -= fk * m->eny[e]; }
                        }
                    }
                    for (int c = 0; c < 4; ++c) { s->gx[4 * i + c] = ax[c] / m->area[i]; s->gy[4 * i + c] = ay[c] / m->area[i]; }
                }
                for (size_t i = 4 * (size_t)m->N; i < n4; ++i) { s->gx[i] = 0; s->gy[i] = 0; }
            } else {
                orc_calc_gradients(s);  /* least squares is already a loop over cells in the reference */
            }
            if (s->second_order) {
#pragma omp parallel for schedule(static)
                for (long i = 0; i < (long)m->N; ++i) {
                    double lo[4], hi[4], l[4] = {1, 1, 1, 1};
                    for (int c = 0; c < 4; ++c) lo[c] = hi[c] = qk[4 * i + c];
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t e = s->cf_sorted[4 * i + k];
                        if (e == ORC_EDGE_NULL) break;
                        const uint32_t c0 = m->edge_cells[2 * (size_t)e], c1 = m->edge_cells[2 * (size_t)e + 1];
                        const double *qn = qk + 4 * (size_t)((uint32_t)i == c0 ? c1 : c0);
                        for (int c = 0; c < 4; ++c) { lo[c] = fmin(lo[c], qn[c]); hi[c] = fmax(hi[c], qn[c]); }
                    }
                    const double Ka = s->limiter_k * sqrt(m->area[i]), K3a = Ka * Ka * Ka;
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t e = s->cf_sorted[4 * i + k];
                        if (e == ORC_EDGE_NULL) break;
                        const double dx = m->ecx[e] - m->ccx[i], dy = m->ecy[e] - m->ccy[i];
                        for (int c = 0; c < 4; ++c) {
                            const double dqg = s->gx[4 * i + c] * dx + s->gy[4 * i + c] * dy;
                            const double dmax = hi[c] - qk[4 * i + c], dmin = lo[c] - qk[4 * i + c];
                            double v = 1.0;
                            if (s->limiter_kind == 1) v = orc_michalak_one(dqg, dmax, dmin, K3a);
                            else if (dqg > 1e-16) v = 1 / dqg * ((dmax * dmax + K3a) * dqg + 2 * dqg * dqg * dmax) / (dmax * dmax + 2 * dqg * dqg + dmax * dqg + K3a);
                            else if (dqg < -1e-16) v = 1 / dqg * ((dmin * dmin + K3a) * dqg + 2 * dqg * dqg * dmin) / (dmin * dmin + 2 * dqg * dqg + dmin * dqg + K3a);
                            l[c] = fmin(l[c], v);
                        }
                    }
                    for (int c = 0; c < 4; ++c) s->lim[4 * i + c] = l[c];
                }
                for (size_t i = 4 * (size_t)m->N; i < n4; ++i) s->lim[i] = 1;
            }
        }
        /* face fluxes, once per face */
#pragma omp parallel for schedule(static)
        for (long e = 0; e < (long)m->E; ++e) {
            const uint32_t c0 = m->edge_cells[2 * e], c1 = m->edge_cells[2 * e + 1];
            const size_t k0 = 4 * (size_t)c0, k1 = 4 * (size_t)c1;
            double gradx[4], grady[4], f[4], qL[4], qR[4];
            orc_average_gradients(s, c0, c1, gradx, grady);
            for (int c = 0; c < 4; ++c) { qL[c] = qk[k0 + c]; qR[c] = qk[k1 + c]; }
            if (s->second_order) {
                const double d0x = m->ecx[e] - m->ccx[c0], d0y = m->ecy[e] - m->ccy[c0];
                const double d1x = m->ecx[e] - m->ccx[c1], d1y = m->ecy[e] - m->ccy[c1];
                for (int c = 0; c < 4; ++c) {
                    qL[c] = qk[k0 + c] + (s->gx[k0 + c] * d0x + s->gy[k0 + c] * d0y) * s->lim[k0 + c];
                    qR[c] = qk[k1 + c] + (s->gx[k1 + c] * d1x + s->gy[k1 + c] * d1y) * s->lim[k1 + c];
                }
            }
            orc_flux(s->edge_kind[e], &s->g, s->viscous_type, m->enx[e], m->eny[e], qL, qR, gradx, grady, f);
            for (int c = 0; c < 4; ++c) s->fluxbuf[4 * (size_t)e + c] = f[c] * m->elen[e];
        }
        /* gather, residual, stage update */
        memset(s->qW + 4 * (size_t)m->N, 0, 4 * (size_t)m->G * sizeof(double));
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)m->N; ++i) {
            double r[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k) {
                const uint32_t e = s->cf_sorted[4 * i + k];
                if (e == ORC_EDGE_NULL) break;
                const double *f = s->fluxbuf + 4 * (size_t)e;
                if (m->edge_cells[2 * (size_t)e] == (uint32_t)i) for (int c = 0; c < 4; ++c) r[c] -= f[c];
                else for (int c = 0; c < 4; ++c) r[c] += f[c];
            }
            for (int c = 0; c < 4; ++c) {
                s->qW[4 * i + c] = r[c] / m->area[i];
                qk[4 * i + c] = s->q[4 * i + c] + s->qW[4 * i + c] * s->dt[i] * alpha[st] * relaxation;
            }
        }
        for (uint32_t b = 0; b < m->G; ++b) {  /* ghost rows of qW for two-sided boundary faces */
            const uint32_t e = m->bnd_edge[b];
            if (two_sided_(s->edge_kind[e])) for (int c = 0; c < 4; ++c) s->qW[4 * ((size_t)m->N + b) + c] = s->fluxbuf[4 * (size_t)e + c];
        }
    }
    memcpy(s->q, s->qk, n4 * sizeof(double));
    return orc_norm(s->qW, n4);
}
