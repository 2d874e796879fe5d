// ref_driver.cpp -- C entry points into the UNMODIFIED AeroFLEX reference
// headers (/root/reference/src/rans/include/rans/*.h), compiled where they
// lie against the Eigen-API stand-in in oracle/eigen_shim.  TEST
// INFRASTRUCTURE: builds oracle/_ref/libafx_ref.so (see oracle/Makefile), used
// to pin oracle/rans_oracle.c, to generate tests/golden/ (oracle/make_golden.py)
// and as the "reference" CPU arm of bench.py.  Never linked by the product.
//
// No reference source is copied: the headers are #included from the read-only
// reference tree at build time.  Built with -fno-access-control so that the
// individual phases (calc_dt, calc_gradients, calc_limiters, calc_residual,
// fillRhoRHS, fillRhoLHS) can be called and their vectors read one at a time.
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <filesystem>
#include <unistd.h>

#include <rans/multigrid.h>
#include <rans/post.h>

namespace {

struct RefSolver {
    std::unique_ptr<rans::explicitSolver> ex;
    std::unique_ptr<rans::implicitSolver> im;
    std::map<std::string, rans::boundary_condition> bcs;
    rans::solver& s() { return ex ? static_cast<rans::solver&>(*ex) : static_cast<rans::solver&>(*im); }
};

thread_local std::string g_err;

template <class F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown exception"; return -1; }
}

std::unique_ptr<rans::flux> make_flux(int kind, rans::gas& g, double& nx, double& ny, int viscous_type) {
    switch (kind) {  // same mapping as solver.h:237-245
        case 0: return std::make_unique<rans::internal_flux>(g, nx, ny, viscous_type);
        case 1: return std::make_unique<rans::farfield_flux>(g, nx, ny, viscous_type);
        case 2: return std::make_unique<rans::slip_wall_flux>(g, nx, ny, viscous_type);
        default: return std::make_unique<rans::wall_flux>(g, nx, ny, viscous_type);
    }
}

Eigen::VectorXd vec4(const double* p) { Eigen::VectorXd v(4); for (int i = 0; i < 4; ++i) v(i) = p[i]; return v; }

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---------------- mesh (mesh.h) ----------------
void* ref_mesh_load(const char* path) {
    rans::mesh* m = nullptr;
    if (guarded([&] { m = new rans::mesh(std::string(path)); })) return nullptr;
    return m;
}
void ref_mesh_free(void* h) { delete static_cast<rans::mesh*>(h); }

// sizes: nodes, real cells, ghost cells, edges, number of physical names
void ref_mesh_sizes(void* h, uint32_t out[5]) {
    auto& m = *static_cast<rans::mesh*>(h);
    out[0] = (uint32_t)m.nodesX.size();
    out[1] = m.nRealCells;
    out[2] = (uint32_t)m.boundaryEdges.size();
    out[3] = m.edgesCells.cols();
    out[4] = (uint32_t)m.physicalNames.size();
}

// raw inputs of the mesh builder: nodes, cell connectivity, boundary segments
void ref_mesh_inputs(void* h, double* x, double* y, uint32_t* cells, uint8_t* is_tri, uint32_t* b0, uint32_t* b1) {
    auto& m = *static_cast<rans::mesh*>(h);
    std::copy(m.nodesX.begin(), m.nodesX.end(), x);
    std::copy(m.nodesY.begin(), m.nodesY.end(), y);
    for (uint32_t c = 0; c < m.nRealCells; ++c) {
        for (int k = 0; k < 4; ++k) cells[4 * c + k] = m.cellsNodes(c, k);
        is_tri[c] = m.cellsIsTriangle[c];
    }
    std::copy(m.boundaryEdges0.begin(), m.boundaryEdges0.end(), b0);
    std::copy(m.boundaryEdges1.begin(), m.boundaryEdges1.end(), b1);
}

// derived arrays exactly as the reference built them
void ref_mesh_arrays(void* h, uint32_t* edge_cells, uint32_t* edge_nodes, double* enx, double* eny, double* elen,
                     double* ecx, double* ecy, double* ccx, double* ccy, double* area, uint32_t* cell_edges,
                     uint32_t* bnd_edge) {
    auto& m = *static_cast<rans::mesh*>(h);
    const uint32_t E = m.edgesCells.cols(), N = m.nRealCells;
    for (uint32_t e = 0; e < E; ++e) {
        edge_cells[2 * e] = m.edgesCells(e, 0); edge_cells[2 * e + 1] = m.edgesCells(e, 1);
        edge_nodes[2 * e] = m.edgesNodes(e, 0); edge_nodes[2 * e + 1] = m.edgesNodes(e, 1);
    }
    std::copy(m.edgesNormalsX.begin(), m.edgesNormalsX.end(), enx);
    std::copy(m.edgesNormalsY.begin(), m.edgesNormalsY.end(), eny);
    std::copy(m.edgesLengths.begin(), m.edgesLengths.end(), elen);
    std::copy(m.edgesCentersX.begin(), m.edgesCentersX.end(), ecx);
    std::copy(m.edgesCentersY.begin(), m.edgesCentersY.end(), ecy);
    std::copy(m.cellsCentersX.begin(), m.cellsCentersX.end(), ccx);
    std::copy(m.cellsCentersY.begin(), m.cellsCentersY.end(), ccy);
    std::copy(m.cellsAreas.begin(), m.cellsAreas.end(), area);
    for (uint32_t c = 0; c < N; ++c) for (int k = 0; k < 4; ++k) cell_edges[4 * c + k] = m.cellsEdges(c, k);
    std::copy(m.boundaryEdges.begin(), m.boundaryEdges.end(), bnd_edge);
}

// physical name of boundary edge b (copied into buf)
void ref_mesh_boundary_name(void* h, uint32_t b, char* buf, int n) {
    auto& m = *static_cast<rans::mesh*>(h);
    std::strncpy(buf, m.boundaryEdgesPhysicals[b].c_str(), (size_t)n - 1);
    buf[n - 1] = 0;
}

// ---------------- solver (solver.h) ----------------
void* ref_solver_new(void* mesh_h, int implicit, const char* viscosity, double gamma, double R) {
    auto* rs = new RefSolver;
    rans::gas g; g.gamma = gamma; g.R = R;
    auto& m = *static_cast<rans::mesh*>(mesh_h);
    if (guarded([&] {
            if (implicit) rs->im = std::make_unique<rans::implicitSolver>(m, g, std::string(viscosity));
            else rs->ex = std::make_unique<rans::explicitSolver>(m, g, std::string(viscosity));
        })) { delete rs; return nullptr; }
    return rs;
}
void ref_solver_free(void* h) { delete static_cast<RefSolver*>(h); }

void ref_solver_set_gas(void* h, double mu_L, double Pr_L, double cp) {
    auto& g = static_cast<RefSolver*>(h)->s().g; g.mu_L = mu_L; g.Pr_L = Pr_L; g.cp = cp;
}

void ref_solver_add_bc(void* h, const char* name, const char* type, double mach, double angle, double T, double p) {
    auto* rs = static_cast<RefSolver*>(h);
    rans::boundary_condition bc; bc.bc_type = type;
    bc.vars_far.mach = mach; bc.vars_far.angle = angle; bc.vars_far.T = T; bc.vars_far.p = p;
    rs->bcs[name] = bc;
}
int ref_solver_apply_bcs(void* h) {
    auto* rs = static_cast<RefSolver*>(h);
    return guarded([&] { rs->s().set_bcs(rs->bcs); });
}
int ref_solver_set_options(void* h, int second_order, const char* gradient_scheme, double limiter_k, double cfl) {
    auto* rs = static_cast<RefSolver*>(h);
    return guarded([&] {
        rs->s().set_second_order(second_order != 0);
        rs->s().set_gradient_scheme(gradient_scheme);
        rs->s().set_limiter_k(limiter_k);
        rs->s().set_cfl(cfl);
    });
}
void ref_solver_set_cfl(void* h, double cfl) { static_cast<RefSolver*>(h)->s().set_cfl(cfl); }
void ref_solver_init(void* h) { static_cast<RefSolver*>(h)->s().init(); }
void ref_solver_refill_bcs(void* h) { static_cast<RefSolver*>(h)->s().refill_bcs(); }
void ref_solver_bcs_from_internal(void* h) { static_cast<RefSolver*>(h)->s().bcs_from_internal(); }
double ref_solver_uniform_residual(void* h) { return static_cast<RefSolver*>(h)->s().get_uniform_residual(); }

// name: q qk qW gx gy limiters dt rhs
static Eigen::VectorXd* pick(RefSolver* rs, const char* name) {
    const std::string n(name);
    rans::solver& s = rs->s();
    if (n == "q") return &s.q;
    if (n == "qW") return &s.qW;
    if (n == "gx") return &s.gx;
    if (n == "gy") return &s.gy;
    if (n == "limiters") return &s.limiters;
    if (n == "dt") return &s.dt;
    if (n == "qk" && rs->ex) return &rs->ex->qk;
    if (n == "rhs" && rs->im) return &rs->im->RhoVector;
    return nullptr;
}
long ref_solver_get(void* h, const char* name, double* out) {
    auto* v = pick(static_cast<RefSolver*>(h), name);
    if (!v) return -1;
    if (out) std::copy(v->data(), v->data() + v->size(), out);
    return (long)v->size();
}
long ref_solver_set(void* h, const char* name, const double* in) {
    auto* v = pick(static_cast<RefSolver*>(h), name);
    if (!v) return -1;
    std::copy(in, in + v->size(), v->data());
    return (long)v->size();
}

// single phases; which = 0 operates on q, 1 on qk (explicit only)
void ref_solver_calc_dt(void* h) { static_cast<RefSolver*>(h)->s().calc_dt(); }
void ref_solver_walls(void* h, int which) {
    auto* rs = static_cast<RefSolver*>(h);
    rs->s().set_walls_from_internal(which && rs->ex ? rs->ex->qk : rs->s().q);
}
void ref_solver_calc_gradients(void* h, int which) {
    auto* rs = static_cast<RefSolver*>(h);
    rs->s().calc_gradients(which && rs->ex ? rs->ex->qk : rs->s().q);
}
void ref_solver_calc_limiters(void* h, int which) {
    auto* rs = static_cast<RefSolver*>(h);
    rs->s().calc_limiters(which && rs->ex ? rs->ex->qk : rs->s().q);
}
int ref_solver_calc_residual(void* h, int which) {  // explicitSolver::calc_residual -> qW
    auto* rs = static_cast<RefSolver*>(h);
    if (!rs->ex) return -1;
    rs->ex->calc_residual(which ? rs->ex->qk : rs->ex->q);
    return 0;
}
double ref_solver_explicit_solve(void* h, double relaxation) {
    auto* rs = static_cast<RefSolver*>(h);
    return rs->ex ? rs->ex->solve(relaxation) : -2.0;
}
double ref_solver_implicit_rhs(void* h) {  // fillRhoRHS -> RhoVector, returns its norm
    auto* rs = static_cast<RefSolver*>(h);
    if (!rs->im) return -2.0;
    rs->im->fillRhoRHS();
    return rs->im->RhoVector.norm();
}
// fillRhoLHS, then the 4x4 blocks: diag[(N+G)][16], off01[E][16] (row c0, col c1), off10[E][16]
int ref_solver_implicit_lhs(void* h, double* diag, double* off01, double* off10) {
    auto* rs = static_cast<RefSolver*>(h);
    if (!rs->im) return -1;
    return guarded([&] {
        auto& s = *rs->im;
        s.fillRhoLHS();
        const uint32_t NT = (uint32_t)s.m.cellsAreas.size(), E = s.m.edgesCells.cols();
        for (uint32_t c = 0; c < NT; ++c)
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) diag[16 * c + 4 * i + j] = s.RhoMatrix.coeffRef(4 * c + i, 4 * c + j);
        for (uint32_t e = 0; e < E; ++e) {
            const uint32_t c0 = s.m.edgesCells(e, 0), c1 = s.m.edgesCells(e, 1);
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
                off01[16 * e + 4 * i + j] = s.RhoMatrix.coeffRef(4 * c0 + i, 4 * c1 + j);
                off10[16 * e + 4 * i + j] = s.RhoMatrix.coeffRef(4 * c1 + i, 4 * c0 + j);
            }
        }
    });
}
// one implicit outer iteration as multigrid<implicitSolver>::run_solver does it
double ref_solver_implicit_step(void* h, double relaxation, double tol, int rhs_iterations) {
    auto* rs = static_cast<RefSolver*>(h);
    if (!rs->im) return -2.0;
    double err = -1;
    if (guarded([&] { rs->im->fill(); if (rs->im->compute() == 0) err = rs->im->solve(relaxation, tol, rhs_iterations); })) return -3.0;
    return err;
}

int ref_solver_wall_forces(void* h, const char* patch, double out_cl_cd_cm[3]) {
    auto* rs = static_cast<RefSolver*>(h);
    return guarded([&] {
        rans::wallProfile wp = rans::get_wall_profile(rs->s(), patch);
        out_cl_cd_cm[0] = wp.cl; out_cl_cd_cm[1] = wp.cd; out_cl_cd_cm[2] = wp.cm;
    });
}

// ---------------- single-face physics (physics.h) ----------------
static rans::gas mk_gas(const double* gas5) {
    rans::gas g; g.gamma = gas5[0]; g.R = gas5[1]; g.mu_L = gas5[2]; g.Pr_L = gas5[3]; g.cp = gas5[4]; return g;
}
void ref_flux(int kind, const double* gas5, int viscous_type, double nx, double ny, const double* qL, const double* qR,
              const double* gx, const double* gy, double* f) {
    rans::gas g = mk_gas(gas5);
    auto fl = make_flux(kind, g, nx, ny, viscous_type);
    Eigen::VectorXd r = (*fl)(vec4(qL), vec4(qR), vec4(gx), vec4(gy));
    for (int i = 0; i < 4; ++i) f[i] = r(i);
}
void ref_vars(int kind, const double* gas5, double nx, double ny, const double* qL, const double* qbc, double* qR) {
    rans::gas g = mk_gas(gas5);
    auto fl = make_flux(kind, g, nx, ny, 0);
    Eigen::VectorXd r = fl->vars(vec4(qL), vec4(qbc));
    for (int i = 0; i < 4; ++i) qR[i] = r(i);
}
void ref_fd_jacobian(int kind, const double* gas5, int viscous_type, double nx, double ny, const double* qL,
                     const double* qR, const double* gx, const double* gy, double* J /*8x8 row-major*/) {
    rans::gas g = mk_gas(gas5);
    auto fl = make_flux(kind, g, nx, ny, viscous_type);
    Eigen::VectorXd l = vec4(qL), r = vec4(qR);
    Eigen::MatrixXd Jm = rans::calc_convective_jacobian(*fl, l, r, vec4(gx), vec4(gy));
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) J[8 * i + j] = Jm(i, j);
}
void ref_get_conservative(double mach, double angle, double T, double p, const double* gas5, double* q) {
    rans::gas g = mk_gas(gas5);
    rans::boundary_variables v(mach, angle, T, p);
    auto c = v.get_conservative(g);
    q[0] = c.rho; q[1] = c.rhou; q[2] = c.rhov; q[3] = c.rhoe;
}

// ---------------- the multigrid driver on given mesh files (multigrid.h, rans.h:78-106) ----------------
// Runs the reference's own FMG + alpha sweep (the body of Rans::run_airfoil with
// the mesh paths passed in instead of the hard-coded ../../../../examples path
// and without the VTU save).  Returns the number of alphas done, <0 on error.
int ref_run_sweep(int n_mesh, const char** mesh_paths, int implicit, const char* viscosity, const char* gradient,
                  int second_order, double relaxation, double start_cfl, double slope_cfl, double max_cfl,
                  double tolerance, int rhs_iterations, int max_iterations, double limiter_k, double gamma, double R,
                  double mach, double T, double p, const char* wall_type, int n_alpha, const double* alphas_deg,
                  double* cl, double* cd, double* cm, int* iters_out, double* seconds_out, int quiet) {
    int done = 0;
    const int rc = guarded([&] {
        std::streambuf* old = nullptr;
        std::ostringstream sink;
        if (quiet) old = std::cout.rdbuf(sink.rdbuf());
        struct Restore { std::streambuf* o; ~Restore() { if (o) std::cout.rdbuf(o); } } restore{old};

        GUIHandler gui;
        rans::Settings st;
        st.g.gamma = gamma; st.g.R = R;
        st.bcs["farfield"].bc_type = "farfield";
        st.bcs["farfield"].vars_far = rans::boundary_variables(mach, 0., T, p);
        st.bcs["wall"].bc_type = wall_type;
        st.set_solver_type(implicit ? "implicit" : "explicit");
        st.set_gradient_scheme(gradient);
        st.set_viscosity_model(viscosity);
        st.second_order = second_order != 0; st.relaxation = relaxation;
        st.start_cfl = start_cfl; st.slope_cfl = slope_cfl; st.max_cfl = max_cfl;
        st.tolerance = tolerance; st.rhs_iterations = rhs_iterations; st.max_iterations = max_iterations;
        st.limiter_k = limiter_k;
        std::vector<double> residuals = {1.0};
        std::atomic<int> iters{0};
        rans::CpProfile profile;
        std::vector<rans::mesh> ms;
        for (int i = 0; i < n_mesh; ++i) ms.push_back(rans::mesh(std::string(mesh_paths[i])));

        auto body = [&](auto tag) {
            using T_ = typename decltype(tag)::type;
            st.bcs["farfield"].vars_far.angle = alphas_deg[0] * 0.01745;
            rans::multigrid<T_> multi(ms, st, gui, residuals, iters, profile);
            multi.solvers[0].init();
            const auto t0 = std::chrono::steady_clock::now();
            for (int a = 0; a < n_alpha; ++a) {
                iters = 0;  // the reference never resets this inside a sweep; avoid its overflow
                st.bcs["farfield"].vars_far.angle = alphas_deg[a] * 0.01745;
                for (auto& s : multi.solvers) s.set_bcs(st.bcs);
                rans::solver& s = multi.run(false);
                rans::wallProfile wp = rans::get_wall_profile(s, "wall");
                cl[a] = wp.cl; cd[a] = wp.cd; cm[a] = wp.cm;
                if (iters_out) iters_out[a] = iters;
                ++done;
            }
            if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        };
        struct TagI { using type = rans::implicitSolver; };
        struct TagE { using type = rans::explicitSolver; };
        if (implicit) body(TagI{}); else body(TagE{});
    });
    return rc ? rc : done;
}

// multigrid<explicitSolver>::gen_mapper(0) applied to q_coarse (multigrid.h:100-178, 312): q_fine = mapper * q_coarse
int ref_prolongate(const char* coarse_path, const char* fine_path, const double* q_coarse, double* q_fine) {
    return guarded([&] {
        std::ostringstream sink;
        std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
        struct Restore { std::streambuf* o; ~Restore() { std::cout.rdbuf(o); } } restore{old};
        GUIHandler gui;
        rans::Settings st;
        st.bcs["farfield"].bc_type = "farfield";
        st.bcs["wall"].bc_type = "slip-wall";
        st.set_solver_type("explicit");
        std::vector<double> residuals = {1.0};
        std::atomic<int> iters{0};
        rans::CpProfile profile;
        std::vector<rans::mesh> ms = {rans::mesh(std::string(coarse_path)), rans::mesh(std::string(fine_path))};
        rans::multigrid<rans::explicitSolver> multi(ms, st, gui, residuals, iters, profile);
        Eigen::VectorXd& qc = multi.solvers[0].get_q();
        std::copy(q_coarse, q_coarse + qc.size(), qc.data());
        Eigen::VectorXd qf = multi.mappers[0] * qc;
        std::copy(qf.data(), qf.data() + qf.size(), q_fine);
    });
}

}  // extern "C"
