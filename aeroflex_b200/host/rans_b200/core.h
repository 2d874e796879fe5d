// core.h -- gas, boundary variables, boundary conditions and Settings with the conf.ini contract of the reference
// (src/rans/include/rans/core.h:27-46, 61-84, 112-115, 166-308).  Same names, defaults, section/key names and
// error behaviour; independent code.
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/afx_rans.h"
#include "config.h"

using uint = unsigned int;

namespace rans {

struct gas {  // core.h:27-46
    double R = 0.71428571428, mu_L = 1e-5, Pr_L = 0.72, Pr_T = 0.9, cp = 1., gamma = 1.4;
    gas(double R_ = 0.71428571428, double mu_L_ = 1e-5, double Pr_L_ = 0.72, double Pr_T_ = 0.9, double gamma_ = 1.4)
        : R(R_), mu_L(mu_L_), Pr_L(Pr_L_), Pr_T(Pr_T_), gamma(gamma_) {}
    double k() const { return cp * mu_L / Pr_L; }
    afx_gas c_abi() const { return afx_gas{gamma, R, mu_L, Pr_L, cp}; }
};

struct conservative_variables {
    double rho, rhou, rhov, rhoe;
};

struct boundary_variables {  // core.h:61-84
    double mach = 0.2, angle = 0, T = 1, p = 1.;
    boundary_variables() {}
    boundary_variables(double machi, double anglei = 0, double Ti = 1, double pi = 1.) : mach(machi), angle(anglei), T(Ti), p(pi) {}
    conservative_variables get_conservative(const gas& g) const {
        const double c = std::sqrt(g.gamma * g.R * T);
        const double u = mach * c * std::cos(angle), v = mach * c * std::sin(angle);
        const double rho = p / (g.R * T);
        return {rho, rho * u, rho * v, p / (g.gamma - 1) + 0.5 * rho * (u * u + v * v)};
    }
};

struct boundary_condition {  // core.h:112-115
    std::string bc_type;
    boundary_variables vars_far;
};

class solution {  // core.h:119-164: primitive variables of cell i from the conservative state
    const std::vector<double>& q;
    const gas& g;
public:
    solution(const std::vector<double>& q_, const gas& g_) : q(q_), g(g_) {}
    double gamma() const { return g.gamma; }
    double rho(int i) const { return q[4 * (size_t)i]; }
    double rhou(int i) const { return q[4 * (size_t)i + 1]; }
    double rhov(int i) const { return q[4 * (size_t)i + 2]; }
    double rhoe(int i) const { return q[4 * (size_t)i + 3]; }
    double p(int i) const { return (g.gamma - 1) * (rhoe(i) - 0.5 / rho(i) * (rhou(i) * rhou(i) + rhov(i) * rhov(i))); }
    double u(int i) const { return rhou(i) / rho(i); }
    double v(int i) const { return rhov(i) / rho(i); }
    double T(int i) const { return p(i) / (g.R * rho(i)); }
    double c(int i) const { return std::sqrt(g.gamma * p(i) / rho(i)); }
    double mach(int i) const { return std::sqrt(rhou(i) * rhou(i) + rhov(i) * rhov(i)) / (rho(i) * c(i)); }
};

struct Settings {  // core.h:166-232
    gas g;
    std::vector<std::string> meshes;
    std::map<std::string, boundary_condition> bcs;
    std::vector<std::string> solver_options = {"explicit", "implicit"};
    std::vector<std::string> gradient_options = {"least-squares", "green-gauss"};
    std::vector<std::string> viscosity_options = {"inviscid", "laminar", "spallart-allmaras"};
    int solver = 1, gradient = 1, viscosity = 0;
    bool second_order = true;
    double start_cfl = 40., slope_cfl = 50., max_cfl = 100., relaxation = 0.9, tolerance = 1e-4;
    int rhs_iterations = 5, max_iterations = 300;
    double alpha_start = 1.0, alpha_end = 7.0, alpha_step = 3.0;
    double limiter_k = 5.;
    std::string airfoil_name = "wall";
    std::string outfilename;
    int read_failure = 1;

    std::string solver_type() { return solver_options.at(solver); }
    void set_solver_type(const std::string& t) { if (t == "explicit") solver = 0; else if (t == "implicit") solver = 1; }
    std::string gradient_scheme() { return gradient_options.at(gradient); }
    void set_gradient_scheme(const std::string& t) { if (t == "least-squares") gradient = 0; else if (t == "green-gauss") gradient = 1; }
    std::string viscosity_model() { return viscosity_options.at(viscosity); }
    void set_viscosity_model(const std::string& t) {
        if (t == "inviscid") viscosity = 0; else if (t == "laminar") viscosity = 1; else if (t == "spallart-allmaras") viscosity = 2;
    }

    // core.h:238-272
    void import_config_file(tiny::config& io) {
        if (io.how_many("rans-bc") != 2) throw std::runtime_error("[RANS] Invalid number of boundary conditions");
        g.gamma = io.get<double>("rans-gas", "gamma");
        g.R = io.get<double>("rans-gas", "R");
        for (int i = 0; i < io.how_many("rans-bc"); i++) {
            const std::string type = io.get_i<std::string>("rans-bc", "type", i), name = io.get_i<std::string>("rans-bc", "name", i);
            bcs[name].bc_type = type;
            if (type == "farfield") {
                bcs[name].vars_far.T = io.get_i<double>("rans-bc", "T", i);
                bcs[name].vars_far.mach = io.get_i<double>("rans-bc", "mach", i);
                bcs[name].vars_far.angle = io.get_i<double>("rans-bc", "angle", i);
                bcs[name].vars_far.p = io.get_i<double>("rans-bc", "p", i);
            }
        }
        set_solver_type(io.get<std::string>("rans-solver", "solver"));
        set_gradient_scheme(io.get<std::string>("rans-solver", "gradient"));
        set_viscosity_model(io.get<std::string>("rans-solver", "viscosity"));
        second_order = io.get<bool>("rans-solver", "second_order");
        relaxation = io.get<double>("rans-solver", "relaxation");
        start_cfl = io.get<double>("rans-solver", "start_cfl");
        slope_cfl = io.get<double>("rans-solver", "slope_cfl");
        max_cfl = io.get<double>("rans-solver", "max_cfl");
        tolerance = io.get<double>("rans-solver", "tolerance");
        rhs_iterations = io.get<int>("rans-solver", "rhs_iterations");
        max_iterations = io.get<int>("rans-solver", "max_iterations");
        limiter_k = io.get<double>("rans-solver", "limiter_k");
        alpha_start = io.get<double>("rans-alphas", "alpha_start");
        alpha_end = io.get<double>("rans-alphas", "alpha_end");
        alpha_step = io.get<double>("rans-alphas", "alpha_step");
    }

    // core.h:274-308 (the reference writes `relaxation` through bool_to_string; kept as a number here so that
    // an exported file imports back to the same settings)
    void export_config_file(tiny::config& io) {
        io.config["rans-gas"]["gamma"] = std::to_string(g.gamma);
        io.config["rans-gas"]["R"] = std::to_string(g.R);
        io.config_vec["rans-bc"] = {};
        for (auto& [name, bc] : bcs)
            io.config_vec["rans-bc"].push_back({{"type", bc.bc_type}, {"name", name}, {"T", std::to_string(bc.vars_far.T)},
                                                {"mach", std::to_string(bc.vars_far.mach)}, {"angle", std::to_string(bc.vars_far.angle)},
                                                {"p", std::to_string(bc.vars_far.p)}});
        auto& s = io.config["rans-solver"];
        s["solver"] = solver_type(); s["gradient"] = gradient_scheme(); s["viscosity"] = viscosity_model();
        s["second_order"] = second_order ? "true" : "false";
        s["relaxation"] = std::to_string(relaxation);
        s["start_cfl"] = std::to_string(start_cfl); s["slope_cfl"] = std::to_string(slope_cfl); s["max_cfl"] = std::to_string(max_cfl);
        s["tolerance"] = std::to_string(tolerance); s["rhs_iterations"] = std::to_string(rhs_iterations);
        s["max_iterations"] = std::to_string(max_iterations); s["limiter_k"] = std::to_string(limiter_k);
        io.config["rans-alphas"]["alpha_start"] = std::to_string(alpha_start);
        io.config["rans-alphas"]["alpha_end"] = std::to_string(alpha_end);
        io.config["rans-alphas"]["alpha_step"] = std::to_string(alpha_step);
    }
};

}  // namespace rans
