// rans_physics.cuh -- per-face physics of the AeroFLEX rans solver as sm_100a
// device functions.  Every function cites the reference lines it implements
// (src/rans/include/rans/physics.h unless noted).  Expression association
// follows the reference and the translation unit is compiled with -fmad=false,
// so in double precision (IEEE div/sqrt) the results are bit-identical to the
// CPU code built without FMA contraction.
#pragma once
#include "rans_types.h"

#ifndef AFX_FAST
#define AFX_FAST 0
#endif
#if AFX_FAST
#define AFX_NS fast
#else
#define AFX_NS strict
#endif

// Two arithmetic modes share this file (one translation unit each, rans_kernels_tu.cu):
//   strict (AFX_FAST=0, -fmad=false): the reference's expression order, every division where the reference
//          divides -> bit-identical to the CPU reference built without FMA contraction;
//   fast   (AFX_FAST=1, -fmad=true) : the same formulas with reciprocals shared between divisions by the same
//          denominator and FMA contraction allowed -> each face flux within a few ulp of strict.
namespace afx {
namespace AFX_NS {

__device__ __forceinline__ d4 mk4(double a, double b, double c, double d) { d4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

// min / max as one compare and select.  fmin()/fmax() add NaN quieting (compare + two selects + a fix-up per half
// word, ~7 instructions for a double on sm_100); the limiter takes 64 of them per cell.
__device__ __forceinline__ double dmin2(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double dmax2(double a, double b) { return a < b ? b : a; }

// physics.h:48-53
__device__ __forceinline__ double pressure(const d4& q, double gam)
{
    return (gam - 1) * (q.w - 0.5 / q.x * (q.y * q.y + q.z * q.z));
}

// physics.h:84-86
__device__ __forceinline__ double sabs(double x) { return sqrt(x * x + 1e-4); }
// physics.h:133-135
__device__ __forceinline__ double entropy_fix(double l, double d) { return l > d ? l : (l * l + d * d) / (2 * d); }

#if AFX_FAST
// Fast-mode reciprocal / square roots: the 20-bit hardware seed (MUFU.RCP64H / MUFU.RSQ64H) plus two Newton steps in
// FMA arithmetic, ~1-2 ulp for normal positive arguments, no special-case branch (the states here are densities,
// sound speeds and positive denominators).  The IEEE division / sqrt sequences of strict mode cost 3-4x as many
// instructions, most of them integer fix-up.
__device__ __forceinline__ double fast_rcp(double a)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0); x = fma(x, e, x);
    e = fma(-a, x, 1.0); x = fma(x, e, x);
    return x;
}
__device__ __forceinline__ double fast_rsqrt(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-(a * y), y, 1.0); y = fma(0.5 * y, e, y);
    e = fma(-(a * y), y, 1.0); y = fma(0.5 * y, e, y);
    return y;
}
__device__ __forceinline__ double fast_sqrt(double a)
{
    const double y = fast_rsqrt(a);
    const double g = a * y;
    return fma(0.5 * y, fma(-g, g, a), g);  // one more correction of the product
}
__device__ __forceinline__ double sabs_fast(double x) { return fast_sqrt(fma(x, x, 1e-4)); }

// Roe flux, physics.h:180-228: one reciprocal square root per density gives 1/rho and sqrt(rho), one for c^2 gives c
// and 1/c^2; one reciprocal for 1/(sL+sR) -- instead of the reference's 19 divisions and 6 square roots
__device__ __forceinline__ d4 roe_flux(const d4& qL, const d4& qR, double nx, double ny, double gam)
{
    const double gm1 = gam - 1;
    const double yL = fast_rsqrt(qL.x), yR = fast_rsqrt(qR.x);
    const double rL = yL * yL, rR = yR * yR;
    const double uL = qL.y * rL, vL = qL.z * rL, uR = qR.y * rR, vR = qR.z * rR;
    const double VL = uL * nx + vL * ny, VR = uR * nx + vR * ny;
    const double pL = gm1 * (qL.w - 0.5 * (qL.y * uL + qL.z * vL));
    const double pR = gm1 * (qR.w - 0.5 * (qR.y * uR + qR.z * vR));
    d4 f;
    f.x = (VL * qL.x + VR * qR.x) * 0.5;
    f.y = (VL * qL.y + VR * qR.y + (pL + pR) * nx) * 0.5;
    f.z = (VL * qL.z + VR * qR.z + (pL + pR) * ny) * 0.5;
    f.w = (VL * (qL.w + pL) + VR * (qR.w + pR)) * 0.5;

    const double sL = qL.x * yL, sR = qR.x * yR;
    const double rho = sR * sL;
    const double rs = fast_rcp(sL + sR);
    const double wL = sL * rs, wR = sR * rs;
    const double u = uL * wL + uR * wR;
    const double v = vL * wL + vR * wR;
    const double h = (qL.w + pL) * rL * wL + (qR.w + pR) * rR * wR;
    const double q2 = u * u + v * v;
    const double c2 = gm1 * (h - 0.5 * q2);
    const double yc = fast_rsqrt(c2);
    const double c = c2 * yc;
    const double rc2 = yc * yc;
    const double V = u * nx + v * ny;
    const double d = 0.05 * c, hd = 10.0 * yc;  // 1/(2d) = 10/c
    const double a_cm = sabs_fast(V - c), a_c = sabs_fast(V), a_cp = sabs_fast(V + c);
    const double l_cm = a_cm > d ? a_cm : (a_cm * a_cm + d * d) * hd;
    const double l_c = a_c > d ? a_c : (a_c * a_c + d * d) * hd;
    const double l_cp = a_cp > d ? a_cp : (a_cp * a_cp + d * d) * hd;
    const double dp = pR - pL, dV = VR - VL, rcdV = rho * c * dV;
    const double k1 = l_cm * (dp - rcdV) * (0.5 * rc2);
    const double k2 = l_c * ((qR.x - qL.x) - dp * rc2);
    const double k3 = l_c * rho;
    const double k5 = l_cp * (dp + rcdV) * (0.5 * rc2);
    const double du = uR - uL, dv = vR - vL;
    f.x -= 0.5 * (k1 + k2 + k5);
    f.y -= 0.5 * (k1 * (u - c * nx) + k2 * u + k3 * (du - dV * nx) + k5 * (u + c * nx));
    f.z -= 0.5 * (k1 * (v - c * ny) + k2 * v + k3 * (dv - dV * ny) + k5 * (v + c * ny));
    f.w -= 0.5 * (k1 * (h - c * V) + k2 * q2 * 0.5 + k3 * (u * du + v * dv - V * dV) + k5 * (h + c * V));
    return f;
}
#else
// Roe flux, physics.h:180-228
__device__ __forceinline__ d4 roe_flux(const d4& qL, const d4& qR, double nx, double ny, double gam)
{
    const double V_L = (qL.y * nx + qL.z * ny) / qL.x;
    const double V_R = (qR.y * nx + qR.z * ny) / qR.x;
    const double pL = (gam - 1) * (qL.w - 0.5 / qL.x * (qL.y * qL.y + qL.z * qL.z));
    const double pR = (gam - 1) * (qR.w - 0.5 / qR.x * (qR.y * qR.y + qR.z * qR.z));

    d4 f;
    f.x = (V_L * qL.x + V_R * qR.x) * 0.5;
    f.y = (V_L * qL.y + pL * nx + V_R * qR.y + pR * nx) * 0.5;
    f.z = (V_L * qL.z + pL * ny + V_R * qR.z + pR * ny) * 0.5;
    f.w = (V_L * (qL.w + pL) + V_R * (qR.w + pR)) * 0.5;

    const double uL = qL.y / qL.x, uR = qR.y / qR.x;
    const double vL = qL.z / qL.x, vR = qR.z / qR.x;
    const double sL = sqrt(qL.x), sR = sqrt(qR.x);
    const double rho = sR * sL;
    const double u = (uL * sL + uR * sR) / (sL + sR);
    const double v = (vL * sL + vR * sR) / (sL + sR);
    const double h = ((qL.w + pL) / qL.x * sL + (qR.w + pR) / qR.x * sR) / (sL + sR);
    const double q2 = u * u + v * v;
    const double c = sqrt((gam - 1.) * (h - 0.5 * q2));
    const double V = u * nx + v * ny;
    const double VR = uR * nx + vR * ny;
    const double VL = uL * nx + vL * ny;

    const double l_cm = entropy_fix(sabs(V - c), 0.05 * c);
    const double l_c = entropy_fix(sabs(V), 0.05 * c);
    const double l_cp = entropy_fix(sabs(V + c), 0.05 * c);

    const double k1 = l_cm * ((pR - pL) - rho * c * (VR - VL)) / (2. * c * c);
    const double k2 = l_c * ((qR.x - qL.x) - (pR - pL) / (c * c));
    const double k3 = l_c * rho;
    const double k5 = l_cp * ((pR - pL) + rho * c * (VR - VL)) / (2 * c * c);

    f.x -= 0.5 * (k1 + k2 + k5);
    f.y -= 0.5 * (k1 * (u - c * nx) + k2 * u + k3 * (uR - uL - (VR - VL) * nx) + k5 * (u + c * nx));
    f.z -= 0.5 * (k1 * (v - c * ny) + k2 * v + k3 * (vR - vL - (VR - VL) * ny) + k5 * (v + c * ny));
    f.w -= 0.5 * (k1 * (h - c * V) + k2 * q2 * 0.5 + k3 * (u * (uR - uL) + v * (vR - vL) - V * (VR - VL)) + k5 * (h + c * V));
    return f;
}
#endif

// Laminar viscous part, physics.h:230-257 with helpers :33-82. Subtracted from f in place.
__device__ __forceinline__ void laminar_flux(d4& f, const d4& qL, const d4& qR, const d4& gx, const d4& gy,
                                             double nx, double ny, const GasC& g)
{
    d4 qc;
    qc.x = 0.5 * (qL.x + qR.x); qc.y = 0.5 * (qL.y + qR.y); qc.z = 0.5 * (qL.z + qR.z); qc.w = 0.5 * (qL.w + qR.w);
    const double p = pressure(qc, g.gamma);
    double gp0 = 2. * gx.w;
    gp0 -= gx.y * qc.y / qc.x + qc.y * (gx.y * qc.x - gx.x * qc.y) / (qc.x * qc.x);
    gp0 -= gx.z * qc.z / qc.x + qc.z * (gx.z * qc.x - gx.x * qc.z) / (qc.x * qc.x);
    gp0 *= 0.5 * (g.gamma - 1);
    double gp1 = 2. * gy.w;
    gp1 -= gy.y * qc.y / qc.x + qc.y * (gy.y * qc.x - gy.x * qc.y) / (qc.x * qc.x);
    gp1 -= gy.z * qc.z / qc.x + qc.z * (gy.z * qc.x - gy.x * qc.z) / (qc.x * qc.x);
    gp1 *= 0.5 * (g.gamma - 1);
    const double gT0 = (1. / g.R) * ((gp0 * qc.x - gx.x * p) / (qc.x * qc.x));
    const double gT1 = (1. / g.R) * ((gp1 * qc.x - gy.x * p) / (qc.x * qc.x));
    const double gu0 = (qc.x * gx.y - qc.y * gx.x) / (qc.x * qc.x);
    const double gu1 = (qc.x * gy.y - qc.y * gy.x) / (qc.x * qc.x);
    const double gv0 = (qc.x * gx.z - qc.z * gx.x) / (qc.x * qc.x);
    const double gv1 = (qc.x * gy.z - qc.z * gy.x) / (qc.x * qc.x);
    const double mu = g.mu_L;
    const double kk = g.cp * g.mu_L / g.Pr_L;  // core.h:43-45
    const double div_v = gu0 + gv1;
    const double txx = 2. * mu * (gu0 - div_v / 3.);
    const double tyy = 2. * mu * (gv1 - div_v / 3.);
    const double txy = mu * (gu1 + gv0);
    const double ph0 = qc.y / qc.x * txx + qc.z / qc.x * txy + kk * gT0;
    const double ph1 = qc.y / qc.x * txy + qc.z / qc.x * tyy + kk * gT1;
    f.y -= nx * txx + ny * txy;
    f.z -= nx * txy + ny * tyy;
    f.w -= nx * ph0 + ny * ph1;
}

// flux::vars of the four flux classes: physics.h:267-276 (internal: average),
// 311-339 (slip wall), 377-405 (wall), 446-530 (far field)
__device__ __forceinline__ d4 bc_vars(int kind, const d4& qL, const d4& qbc, double nx, double ny, double gam)
{
    d4 r;
    if (kind == K_INTERNAL) {
        r.x = (qL.x + qbc.x) * 0.5; r.y = (qL.y + qbc.y) * 0.5; r.z = (qL.z + qbc.z) * 0.5; r.w = (qL.w + qbc.w) * 0.5;
    } else if (kind == K_SLIPWALL) {
        const double rhoV = qL.y * nx + qL.z * ny;
        r.x = qL.x; r.y = qL.y - 2. * rhoV * nx; r.z = qL.z - 2. * rhoV * ny; r.w = qL.w;
    } else if (kind == K_WALL) {
        r.x = qL.x; r.y = -qL.y; r.z = -qL.z; r.w = qL.w;
    } else {
        const double rho = qL.x, rho_u = qL.y, rho_v = qL.z, rho_e = qL.w;
        const double bc_rho = qbc.x;
        const double bc_u = qbc.y / bc_rho;
        const double bc_v = qbc.z / bc_rho;
        const double bc_p = (gam - 1) * (qbc.w - 0.5 / bc_rho * (qbc.y * qbc.y + qbc.z * qbc.z));
        const double p = (gam - 1) * (rho_e - 0.5 / rho * (rho_u * rho_u + rho_v * rho_v));
        const double c = sqrt(gam * p / rho);
        const double mach = sqrt(rho_u * rho_u + rho_v * rho_v) / (rho * c);
        const double io = rho_u * nx + rho_v * ny;
        if (mach > 1) {
            if (io < 0) {
                r.x = bc_rho; r.y = bc_rho * bc_u; r.z = bc_rho * bc_v;
                r.w = bc_p / (gam - 1) + 0.5 * bc_rho * (bc_u * bc_u + bc_v * bc_v);
            } else {
                r = qL;
            }
        } else {
            const double pa = bc_p, rhoa = bc_rho, ua = bc_u, va = bc_v;
            const double pd = p, rhod = rho, ud = rho_u / rho, vd = rho_v / rho;
            const double rho0 = rho, c0 = c;
            if (io < 0) {
                const double pb = 0.5 * (pa + pd - rho0 * c0 * (nx * (ua - ud) + ny * (va - vd)));
                r.x = rhoa + (pb - pa) / (c0 * c0);
                r.y = r.x * (ua - nx * (pa - pb) / (rho0 * c0));
                r.z = r.x * (va - ny * (pa - pb) / (rho0 * c0));
                r.w = pb / (gam - 1) + 0.5 / r.x * (r.y * r.y + r.z * r.z);
            } else {  // outlet; `va` as in the reference (physics.h:524)
                const double pb = pa;
                r.x = rhod + (pb - pd) / (c0 * c0);
                r.y = r.x * (ud + nx * (pd - pb) / (rho0 * c0));
                r.z = r.x * (va + ny * (pd - pb) / (rho0 * c0));
                r.w = pb / (gam - 1) + 0.5 / r.x * (r.y * r.y + r.z * r.z);
            }
        }
    }
    return r;
}

// (*edges_flux_functions[e])(qL, qR, gx, gy): internal faces use the face
// gradient (laminar only); boundary faces build the ghost state and call the
// Roe flux with zero gradients (physics.h:296-298,361-363,426-431), for which
// the laminar term vanishes identically.
template <int VISC>
__device__ __forceinline__ d4 face_flux(int kind, const d4& qL, const d4& qR, const d4& gfx, const d4& gfy,
                                        double nx, double ny, const GasC& g)
{
    if (kind == K_INTERNAL) {
        d4 f = roe_flux(qL, qR, nx, ny, g.gamma);
        if (VISC == 1) laminar_flux(f, qL, qR, gfx, gfy, nx, ny, g);
        return f;
    }
    const d4 qb = bc_vars(kind, qL, qR, nx, ny, g.gamma);
    d4 f = roe_flux(qL, qb, nx, ny, g.gamma);
    if (VISC == 1) {
        // zero gradients: every stress/heat-flux term is an exact +0/-0; the
        // reference still evaluates f -= 0, which leaves f unchanged bit for bit
        // except -0 -> +0 is impossible here (f - (+0) keeps f).  Nothing to do.
    }
    return f;
}

}  // namespace AFX_NS
}  // namespace afx
