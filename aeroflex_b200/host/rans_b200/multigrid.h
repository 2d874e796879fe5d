// multigrid.h -- full-multigrid start-up and the outer iteration loops of the reference
// (src/rans/include/rans/multigrid.h:28-363) over GPU solvers.  Same template interface, same CFL ramp, residual
// normalisation, stop/pause polling and residual history; the prolongation weights are the reference's (same cells,
// same sums in the same order), searched with OpenMP instead of one thread.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <iostream>
#include <memory>
#include <thread>

#include "common.h"
#include "post.h"

namespace rans {

struct ProlongationRow { uint32_t begin, end; };
struct Prolongation {  // rows of the sparse (4m x 4n) mapper, one weight per (fine cell, coarse cell)
    std::vector<ProlongationRow> rows;
    std::vector<uint32_t> col;
    std::vector<double> w;
    std::vector<double> apply(const std::vector<double>& qc) const {
        std::vector<double> qf(4 * rows.size(), 0.0);
#pragma omp parallel for
        for (long i = 0; i < (long)rows.size(); ++i)
            for (int k = 0; k < 4; ++k) {
                double s = 0;
                for (uint32_t p = rows[(size_t)i].begin; p < rows[(size_t)i].end; ++p) s += w[p] * qc[4 * (size_t)col[p] + k];
                qf[4 * (size_t)i + k] = s;
            }
        return qf;
    }
};

// inverse-distance prolongation, multigrid.h:100-178: coarse cell j contributes to fine cell i when
// d2 < 2 A_j, weight 1/max(0.1 sqrt(A_j), d), rows normalised by their sum (accumulated over ascending j).
// The reference tests all m x n pairs; make_prolongation finds the same pairs through a k-d tree over the coarse
// centres whose nodes carry their bounding box and the largest 2 A_j below them: a node is skipped when the
// point is at least that far from the box, which -- rounding being monotonic -- never drops a pair the reference keeps.
// Rows are sorted by j and summed in that order, so weights and q_fine are bit-identical to the reference's.
inline Prolongation make_prolongation_bruteforce(const mesh& coarse, const mesh& fine) {
    const long m = (long)fine.cellsAreas.size(), n = (long)coarse.cellsAreas.size();
    std::vector<std::vector<std::pair<uint32_t, double>>> rows((size_t)m);
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < m; ++i) {
        const double xi = fine.cellsCentersX[(size_t)i], yi = fine.cellsCentersY[(size_t)i];
        double scale = 0;
        auto& r = rows[(size_t)i];
        for (long j = 0; j < n; ++j) {
            const double xj = coarse.cellsCentersX[(size_t)j], yj = coarse.cellsCentersY[(size_t)j];
            const double d2 = (xi - xj) * (xi - xj) + (yi - yj) * (yi - yj);
            const double r2 = coarse.cellsAreas[(size_t)j];
            if (d2 < 2 * r2) {
                const double si = 1 / std::max(0.1 * std::sqrt(r2), std::sqrt(d2));
                scale += si;
                r.emplace_back((uint32_t)j, si);
            }
        }
        for (auto& e : r) e.second = e.second / scale;
    }
    Prolongation P;
    P.rows.resize((size_t)m);
    for (long i = 0; i < m; ++i) {
        P.rows[(size_t)i].begin = (uint32_t)P.col.size();
        for (auto& e : rows[(size_t)i]) { P.col.push_back(e.first); P.w.push_back(e.second); }
        P.rows[(size_t)i].end = (uint32_t)P.col.size();
    }
    return P;
}

struct CoarseTree {  // k-d tree over the coarse cell centres
    struct Node { double x0, x1, y0, y1, r2max; uint32_t lo, hi; int left, right; };
    std::vector<uint32_t> idx;
    std::vector<Node> nodes;
    const mesh& c;
    explicit CoarseTree(const mesh& coarse) : c(coarse) {
        idx.resize(c.cellsAreas.size());
        for (size_t k = 0; k < idx.size(); ++k) idx[k] = (uint32_t)k;
        nodes.reserve(idx.size() / 4 + 16);
        if (!idx.empty()) build(0, (uint32_t)idx.size());
    }
    int build(uint32_t lo, uint32_t hi) {
        Node nd{1e300, -1e300, 1e300, -1e300, 0., lo, hi, -1, -1};
        for (uint32_t k = lo; k < hi; ++k) {
            const uint32_t j = idx[k];
            nd.x0 = std::min(nd.x0, c.cellsCentersX[j]); nd.x1 = std::max(nd.x1, c.cellsCentersX[j]);
            nd.y0 = std::min(nd.y0, c.cellsCentersY[j]); nd.y1 = std::max(nd.y1, c.cellsCentersY[j]);
            nd.r2max = std::max(nd.r2max, 2 * c.cellsAreas[j]);
        }
        const int me = (int)nodes.size();
        nodes.push_back(nd);
        if (hi - lo > 16) {
            const bool by_x = (nd.x1 - nd.x0) >= (nd.y1 - nd.y0);
            const uint32_t mid = lo + (hi - lo) / 2;
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](uint32_t a, uint32_t b) {
                return by_x ? c.cellsCentersX[a] < c.cellsCentersX[b] : c.cellsCentersY[a] < c.cellsCentersY[b];
            });
            const int l = build(lo, mid), r = build(mid, hi);
            nodes[(size_t)me].left = l; nodes[(size_t)me].right = r;
        }
        return me;
    }
    // every coarse cell j with |x - x_j|^2 < 2 A_j, unordered
    void query(double x, double y, std::vector<uint32_t>& out) const {
        if (nodes.empty()) return;
        int stack[64], top = 0;
        stack[top++] = 0;
        while (top) {
            const Node& nd = nodes[(size_t)stack[--top]];
            const double dx = x < nd.x0 ? nd.x0 - x : (x > nd.x1 ? x - nd.x1 : 0.), dy = y < nd.y0 ? nd.y0 - y : (y > nd.y1 ? y - nd.y1 : 0.);
            if (dx * dx + dy * dy >= nd.r2max) continue;
            if (nd.left < 0) {
                for (uint32_t k = nd.lo; k < nd.hi; ++k) {
                    const uint32_t j = idx[k];
                    const double xj = c.cellsCentersX[j], yj = c.cellsCentersY[j];
                    const double d2 = (x - xj) * (x - xj) + (y - yj) * (y - yj);
                    if (d2 < 2 * c.cellsAreas[j]) out.push_back(j);
                }
            } else { stack[top++] = nd.left; stack[top++] = nd.right; }
        }
    }
};

inline Prolongation make_prolongation(const mesh& coarse, const mesh& fine) {
    const long m = (long)fine.cellsAreas.size();
    const CoarseTree tree(coarse);
    std::vector<std::vector<std::pair<uint32_t, double>>> rows((size_t)m);
#pragma omp parallel
    {
        std::vector<uint32_t> js;
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < m; ++i) {
            const double xi = fine.cellsCentersX[(size_t)i], yi = fine.cellsCentersY[(size_t)i];
            js.clear();
            tree.query(xi, yi, js);
            std::sort(js.begin(), js.end());  // the reference's j loop is ascending: same summation order
            double scale = 0;
            auto& r = rows[(size_t)i];
            r.reserve(js.size());
            for (const uint32_t j : js) {
                const double xj = coarse.cellsCentersX[j], yj = coarse.cellsCentersY[j];
                const double d2 = (xi - xj) * (xi - xj) + (yi - yj) * (yi - yj);
                const double r2 = coarse.cellsAreas[j];
                const double si = 1 / std::max(0.1 * std::sqrt(r2), std::sqrt(d2));
                scale += si;
                r.emplace_back(j, si);
            }
            for (auto& e : r) e.second = e.second / scale;
        }
    }
    Prolongation P;
    P.rows.resize((size_t)m);
    for (long i = 0; i < m; ++i) {
        P.rows[(size_t)i].begin = (uint32_t)P.col.size();
        for (auto& e : rows[(size_t)i]) { P.col.push_back(e.first); P.w.push_back(e.second); }
        P.rows[(size_t)i].end = (uint32_t)P.col.size();
    }
    return P;
}

template <class solverType>
class multigrid {
public:
    GUIHandler& gui;
    Settings& settings;
    std::vector<double>& residuals;
    std::atomic<int>& iters;
    CpProfile& profile;
    std::vector<solverType> solvers;
    std::vector<Prolongation> mappers;
    // the same mappers resident on the GPU: level i -> i+1 without the state leaving the device (afx_prolongation_*)
    struct ProlDeleter { void operator()(afx_prolongation* p) const { afx_prolongation_free(p); } };
    std::vector<std::unique_ptr<afx_prolongation, ProlDeleter>> device_mappers;
    bool verbose = true;

    multigrid(std::vector<mesh> ms, Settings& settings, GUIHandler& gui, std::vector<double>& residuals, std::atomic<int>& iters,
              CpProfile& profile)
        : gui(gui), settings(settings), residuals(residuals), iters(iters), profile(profile) {  // multigrid.h:59-97
        solvers.reserve(ms.size());
        for (auto& mi : ms) {
            mi.compute_wall_dist(settings.bcs);
            solvers.push_back(solverType(mi, settings.g, settings.viscosity_model()));
            solvers.back().set_bcs(settings.bcs);
            solvers.back().set_second_order(settings.second_order);
            solvers.back().set_gradient_scheme(settings.gradient_scheme());
            solvers.back().set_limiter_k(settings.limiter_k);
        }
        gui.msg.push("[RANS] Multigrid : Precompute " + std::to_string(ms.size() - 1) + " matrices");
        if (ms.size() > 1) {
            mappers.resize(ms.size() - 1);
            for (uint i = 0; i < ms.size() - 1; ++i) {
                mappers[i] = gen_mapper(i);
                std::vector<uint32_t> row_begin(mappers[i].rows.size() + 1, 0u);
                for (size_t r = 0; r < mappers[i].rows.size(); ++r) row_begin[r + 1] = mappers[i].rows[r].end;
                afx_prolongation* dp = nullptr;
                if (afx_prolongation_create(&dp, solvers[i].handle(), solvers[i + 1].handle(), row_begin.data(), mappers[i].col.data(),
                                            mappers[i].w.data()) != AFX_OK)
                    throw std::runtime_error(afx_last_error());
                device_mappers.emplace_back(dp);
                gui.msg.push("[RANS] Matrix " + std::to_string(i + 1) + " done");
            }
        }
    }

    Prolongation gen_mapper(const uint level) { return make_prolongation(solvers[level].get_mesh(), solvers[level + 1].get_mesh()); }

    int run_solver(solverType& s);
    solverType& run(const bool reinit = true);
};

// multigrid.h:182-237
template <>
inline int multigrid<explicitSolver>::run_solver(explicitSolver& s) {
    const double cfl = settings.start_cfl;
    int i = 0;
    double err = 0;
    double err_0 = s.get_uniform_residual();
    do {
        while (gui.signal.pause) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        s.set_cfl(cfl);
        s.fill();
        const int ok = s.compute();
        if (ok == 0) {
            err = s.solve(settings.relaxation);
            if ((i == 0) & (err > 2 * err_0)) err_0 = err;
        } else err = -1;
        if (err > 0) err /= err_0;
        if (verbose && i % 100 == 0) std::cout << "Iteration " << i << " Residual = " << err << std::endl;
        if (err < 0) return 1;
        if (i % 10 == 0) {  // every 10th residual, bounded (the reference overruns this vector in long sweeps)
            if ((size_t)iters < residuals.size()) residuals[(size_t)iters] = err;
            iters++;
        }
        i++;
    } while ((err > settings.tolerance) && (i < settings.max_iterations) && !gui.signal.stop);
    return 0;
}

// multigrid.h:239-293
template <>
inline int multigrid<implicitSolver>::run_solver(implicitSolver& s) {
    double cfl = settings.start_cfl;
    int i = 0;
    double err = 0;
    double err_0 = s.get_uniform_residual();
    do {
        while (gui.signal.pause) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        s.set_cfl(cfl);
        s.fill();
        const int ok = s.compute();
        if (ok == 0) {
            err = s.solve(settings.relaxation, err_0 * settings.tolerance, settings.rhs_iterations);
            if ((i == 0) & (err > 2 * err_0)) err_0 = err;
        } else err = -1;
        err /= err_0;
        cfl = std::min(settings.start_cfl + (i + 1) * settings.slope_cfl, settings.max_cfl);
        if (verbose) std::cout << "Iteration " << i << " Residual = " << err << std::endl;
        i++;
        if (err < 0) return 1;
        profile.calc_cp(s, settings.airfoil_name);
        if ((size_t)iters < residuals.size()) residuals[(size_t)iters] = err;
        iters++;
    } while ((err > settings.tolerance) && (i < settings.max_iterations) && !gui.signal.stop);
    return 0;
}

// multigrid.h:295-363 (both specialisations share this body; the explicit one sizes the history by max_iters/10)
template <class solverType>
inline solverType& multigrid<solverType>::run(const bool reinit) {
    const int max_iters = settings.max_iterations;
    const size_t per_level = std::is_same<solverType, explicitSolver>::value ? (size_t)(max_iters / 10 + 1) : (size_t)max_iters;
    if (residuals.size() < (size_t)iters + solvers.size() * per_level) residuals.resize((size_t)iters + solvers.size() * per_level);
    if (reinit) solvers[0].init();
    solvers[0].refill_bcs();
    for (uint i = 0; i < solvers.size(); ++i) {
        profile.calc_chord_coords(solvers[i], settings.airfoil_name);
        if (verbose) std::cout << "\nMultigrid : Stage " << i + 1 << "/" << solvers.size() << "\n" << std::endl;
        if (i > 0) {  // map the last solution to the current grid
            solvers[i - 1].bcs_from_internal();
            if (i - 1 < device_mappers.size()) {  // q_fine = mapper * q_coarse on the device, the reference's sums in the reference's order
                if (afx_prolongation_apply(device_mappers[i - 1].get()) != AFX_OK) throw std::runtime_error(afx_last_error());
            } else solvers[i].set_q(mappers[i - 1].apply(solvers[i - 1].get_q()));
            solvers[i].refill_bcs();
        }
        const int state = run_solver(solvers[i]);
        if (state || gui.signal.stop) return solvers[i];
    }
    return solvers[solvers.size() - 1];
}

}  // namespace rans
