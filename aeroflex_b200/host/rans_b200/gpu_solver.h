// gpu_solver.h -- rans::solver / explicitSolver / implicitSolver with the reference's public surface
// (src/rans/include/rans/solver.h:104-175, 721-742, 852-970), every call forwarded to libaeroflex_rans_b200.so.
// multigrid<solverType>, get_wall_profile and Rans (this directory) use it exactly as the reference uses its
// CPU classes.  State lives on the GPU; get_q() hands out a host mirror that is refreshed on request and pushed
// back with set_q()/push_q() -- the one place where the drop-in differs (the reference returns a live reference).
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mesh.h"

namespace rans {

// CUDA device new solvers are created on (one process per GPU: set it once from LOCAL_RANK / --device)
inline int& default_device() { static int d = 0; return d; }

class solver {
protected:
    struct Deleter { void operator()(afx_rans* s) const { afx_rans_destroy(s); } };
    gas g;
    mesh m;
    std::shared_ptr<afx_rans> h_;
    std::vector<double> q_host_;
    bool second_order = true;
    double cfl = 1;
    uint print_interval = 1;
    std::string viscosity_model;
    std::string gradient_scheme = "green-gauss";
    double limiter_k = 5.;
    std::string airfoil_name;
    int device_ = 0;

    static void check(int rc) {
        if (rc == AFX_OK) return;
        if (rc == AFX_ERR_INVALID) throw std::out_of_range(afx_last_error());
        throw std::runtime_error(afx_last_error());
    }
    static int visc_id(const std::string& v) { return v == "inviscid" ? AFX_VISC_INVISCID : (v == "laminar" ? AFX_VISC_LAMINAR : AFX_VISC_SA); }
    void create() {
        afx_rans* s = nullptr;
        const afx_gas cg = g.c_abi();
        check(afx_rans_create(&s, &m.desc(), &cg, visc_id(viscosity_model), device_));
        h_.reset(s, Deleter());
#ifdef RANS_MICHALAK_LIMITER  // the reference's compile-time switch (solver.h:557) selects the library's run-time one
        check(afx_rans_set_limiter(s, AFX_LIMITER_MICHALAK));
#endif
        q_host_.assign(4 * (m.cellsAreas.size()), 0.0);
    }
    void push_options() {
        check(afx_rans_set_options(h_.get(), second_order ? 1 : 0,
                                   gradient_scheme == "least-squares" ? AFX_GRAD_LEAST_SQUARES : AFX_GRAD_GREEN_GAUSS, limiter_k));
    }

public:
    std::map<std::string, boundary_condition> bcs;

    solver() {}
    solver(const mesh& m_in, const gas& g_in, std::string viscosity_model_, int device = -1)
        : g(g_in), m(m_in), viscosity_model(std::move(viscosity_model_)), device_(device < 0 ? default_device() : device) { create(); }
    // like the reference (solver.h:112-116): a copy is a fresh solver on the same mesh with the same bcs, not a state copy
    solver(const solver& s) : solver(s.get_cmesh(), s.get_gas(), s.get_viscosity_model(), s.device_) {
        bcs = s.get_bcs(); print_interval = s.get_print_interval(); second_order = s.second_order;
        push_options();  // the fresh device solver starts from its defaults: hand it the copied second_order (as solver.h:112-116 copies it)
    }
    solver& operator=(const solver& rhs) {
        if (this == &rhs) return *this;
        g = rhs.g; m = rhs.m; viscosity_model = rhs.viscosity_model; device_ = rhs.device_;
        create();
        bcs = rhs.bcs; print_interval = rhs.print_interval; second_order = rhs.second_order;
        push_options();
        return *this;
    }
    virtual ~solver() {}

    afx_rans* handle() const { return h_.get(); }

    boundary_variables get_boundary_variables() {  // solver.h:597-611
        afx_bvars v;
        afx_rans_boundary_variables(h_.get(), &v);
        return boundary_variables(v.mach, v.angle, v.T, v.p);
    }

    void set_bcs(std::map<std::string, boundary_condition> bcs_in) {  // solver.h:200-247
        bcs = bcs_in;
        const int np = m.n_patches();
        std::vector<uint8_t> kind((size_t)std::max(np, 1), 0);
        std::vector<afx_bvars> vars((size_t)std::max(np, 1), afx_bvars{0.2, 0., 1., 1.});
        std::vector<char> used((size_t)std::max(np, 1), 0);
        for (const auto& name : m.boundaryEdgesPhysicals) used[(size_t)m.patch_id(name)] = 1;
        for (int p = 0; p < np; ++p) {
            if (!used[(size_t)p]) continue;
            const boundary_condition& bc = bcs.at(m.patch_name(p));  // std::out_of_range like the reference
            kind[(size_t)p] = bc.bc_type == "farfield" ? AFX_BC_FARFIELD : bc.bc_type == "slip-wall" ? AFX_BC_SLIPWALL : bc.bc_type == "wall" ? AFX_BC_WALL : bc.bc_type == "inlet-outlet" ? AFX_BC_INLET_OUTLET : AFX_BC_INTERNAL;
            vars[(size_t)p] = afx_bvars{bc.vars_far.mach, bc.vars_far.angle, bc.vars_far.T, bc.vars_far.p};
        }
        check(afx_rans_set_bcs(h_.get(), np, kind.data(), vars.data()));
    }
    void set_cfl(const double& cfl_in) { cfl = cfl_in; check(afx_rans_set_cfl(h_.get(), cfl)); }
    void set_limiter_k(const double& x) { limiter_k = x; push_options(); }
    double get_limiter_k() const { return limiter_k; }
    void set_airfoil_name(const std::string& x) { airfoil_name = x; }
    std::string get_airfoil_name() const { return airfoil_name; }
    void init() { check(afx_rans_init(h_.get())); }
    double get_uniform_residual() { double v = 0; check(afx_rans_uniform_residual(h_.get(), &v)); return v; }

    // host mirror of the state vector, 4 doubles per cell, real cells then ghosts (solver.h:255-257)
    std::vector<double>& get_q() { check(afx_rans_get_q(h_.get(), q_host_.data())); return q_host_; }
    solution get_solution() { return solution(get_q(), g); }  // solver.h:693-697, on the refreshed host mirror
    void set_q(const std::vector<double>& q) {
        if (q.size() != q_host_.size()) throw std::invalid_argument("state vector has the wrong length");
        check(afx_rans_set_q(h_.get(), q.data()));
    }
    void push_q() { check(afx_rans_set_q(h_.get(), q_host_.data())); }

    gas& get_gas() { return g; }
    const gas& get_gas() const { return g; }
    void set_gradient_scheme(std::string grads) { gradient_scheme = grads; push_options(); }
    std::string get_gradient_scheme() const { return gradient_scheme; }
    std::string get_viscosity_model() const { return viscosity_model; }
    void refill_bcs() { check(afx_rans_refill_bcs(h_.get())); }
    void bcs_from_internal() { check(afx_rans_bcs_from_internal(h_.get())); }
    void set_second_order(const bool x = true) { second_order = x; push_options(); }
    bool get_second_order(const bool = true) const { return second_order; }
    mesh& get_mesh() { return m; }
    const mesh& get_cmesh() const { return m; }
    uint get_print_interval() const { return print_interval; }
    std::map<std::string, boundary_condition> get_bcs() const { return bcs; }
    // arithmetic mode of the kernels: "strict" (bit-identical to the CPU reference) or "fast"
    void set_math_mode(const std::string& mode) { check(afx_rans_set_math_mode(h_.get(), mode == "strict" ? AFX_MATH_STRICT : AFX_MATH_FAST)); }

    virtual void fill() {}
    virtual int compute() { return -1; }
    virtual double solve(const double = 1, const double = 0, const int = 5) { return -1; }
};

class explicitSolver : public solver {  // solver.h:721-742
public:
    explicitSolver(const mesh& m_in, const gas& g_in, std::string viscosity_model_, int device = -1) : solver(m_in, g_in, viscosity_model_, device) {}
    explicitSolver(const explicitSolver& s) : solver(s) {}  // the reference's implicit copy: solver's copy constructor (fresh state, same bcs / order)
    void fill() override {}
    int compute() override { return 0; }
    double solve(const double relaxation = 1, const double = 0, const int = 5) override {  // solver.h:802-828
        double v = -1;
        const int rc = afx_rans_step_explicit(h_.get(), relaxation, &v);
        if (rc == AFX_ERR_NUMERIC) return -1;
        check(rc);
        return v;
    }
    // n iterations back to back on the device; returns the residual norms
    std::vector<double> run(int n, double relaxation) {
        std::vector<double> norms((size_t)std::max(n, 0));
        const int rc = afx_rans_run_explicit(h_.get(), relaxation, n, norms.data());
        if (rc != AFX_ERR_NUMERIC) check(rc);
        return norms;
    }
};

class implicitSolver : public solver {  // solver.h:852-970
public:
    implicitSolver(const mesh& m_in, const gas& g_in, std::string viscosity_model_, int device = -1) : solver(m_in, g_in, viscosity_model_, device) {}
    implicitSolver(const implicitSolver& s) : solver(s) {}
    void fill() override { check(afx_rans_fill_jacobian(h_.get())); }  // fillRhoLHS, solver.h:973-976
    int compute() override {                                            // solver.h:1160-1167
        const int rc = afx_rans_compute(h_.get());
        if (rc == AFX_ERR_NUMERIC) return -1;
        check(rc);
        return 0;
    }
    double solve(const double relaxation = 1, const double tol = 0, const int rhs_iterations = 5) override {  // solver.h:1170-1213
        double v = -1;
        const int rc = afx_rans_step_implicit(h_.get(), relaxation, tol, rhs_iterations, &v);
        if (rc == AFX_ERR_NUMERIC) return -1;
        check(rc);
        return v;
    }
};

}  // namespace rans
