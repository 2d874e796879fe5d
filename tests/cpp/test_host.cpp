// CPU-side checks of the C++ host adapter (no GPU needed): conf.ini parsing and the FMG prolongation.
//   test_host conf <conf.ini> <roundtrip.ini>         -> prints the parsed rans settings
//   test_host prolong <coarse.msh> <fine.msh> <q.bin> <out.bin>
//   test_host walldist <mesh.msh> [wall bc_type]      -> tree search vs all-pairs scan
//   test_host prolong_check <coarse.msh> <fine.msh>   -> tree search vs all-pairs search
#include <cstdio>
#include <fstream>
#include <iostream>

#include "rans_b200/multigrid.h"

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "";
    try {
        if (mode == "conf" && argc > 2) {
            tiny::config io;
            io.read(argv[2]);
            rans::Settings s;
            s.import_config_file(io);
            std::printf("solver=%s viscosity=%s gradient=%s second_order=%d relaxation=%.17g start_cfl=%.17g slope_cfl=%.17g max_cfl=%.17g "
                        "tolerance=%.17g rhs_iterations=%d max_iterations=%d limiter_k=%.17g alpha=%.17g:%.17g:%.17g gamma=%.17g R=%.17g\n",
                        s.solver_type().c_str(), s.viscosity_model().c_str(), s.gradient_scheme().c_str(), (int)s.second_order, s.relaxation,
                        s.start_cfl, s.slope_cfl, s.max_cfl, s.tolerance, s.rhs_iterations, s.max_iterations, s.limiter_k, s.alpha_start,
                        s.alpha_end, s.alpha_step, s.g.gamma, s.g.R);
            for (auto& [name, bc] : s.bcs)
                std::printf("bc %s type=%s mach=%.17g angle=%.17g T=%.17g p=%.17g\n", name.c_str(), bc.bc_type.c_str(), bc.vars_far.mach,
                            bc.vars_far.angle, bc.vars_far.T, bc.vars_far.p);
            // export -> import round trip
            tiny::config out;
            out.sections = {"rans-gas", "rans-bc", "rans-solver", "rans-alphas"};
            s.export_config_file(out);
            const std::string rt = argc > 3 ? argv[3] : "/tmp/afx_conf_roundtrip.ini";
            out.write(rt);
            tiny::config back;
            back.read(rt);
            rans::Settings s2;
            s2.import_config_file(back);
            std::printf("roundtrip %d\n", (int)(s2.solver == s.solver && s2.max_iterations == s.max_iterations && s2.bcs.size() == s.bcs.size() &&
                                                 s2.bcs["wall"].bc_type == s.bcs["wall"].bc_type && s2.second_order == s.second_order));
            return 0;
        }
        if (mode == "prolong" && argc > 5) {
            rans::mesh coarse(argv[2]), fine(argv[3]);
            std::vector<double> qc(4 * coarse.cellsAreas.size());
            std::ifstream(argv[4], std::ios::binary).read(reinterpret_cast<char*>(qc.data()), (std::streamsize)(qc.size() * sizeof(double)));
            const rans::Prolongation P = rans::make_prolongation(coarse, fine);
            const std::vector<double> qf = P.apply(qc);
            std::ofstream(argv[5], std::ios::binary).write(reinterpret_cast<const char*>(qf.data()), (std::streamsize)(qf.size() * sizeof(double)));
            std::printf("prolong rows=%zu nnz=%zu\n", P.rows.size(), P.w.size());
            return 0;
        }
        if (mode == "prolong_check" && argc > 3) {  // k-d tree search against the reference's all-pairs loop: identical rows
            rans::mesh coarse(argv[2]), fine(argv[3]);
            const rans::Prolongation A = rans::make_prolongation(coarse, fine), B = rans::make_prolongation_bruteforce(coarse, fine);
            bool same = A.col == B.col && A.w == B.w && A.rows.size() == B.rows.size();
            for (size_t i = 0; same && i < A.rows.size(); ++i) same = A.rows[i].begin == B.rows[i].begin && A.rows[i].end == B.rows[i].end;
            std::printf("prolong_check rows=%zu nnz=%zu identical=%d\n", A.rows.size(), A.w.size(), (int)same);
            return same ? 0 : 1;
        }
        if (mode == "walldist" && argc > 2) {  // k-d tree wall distance against the reference's all-pairs scan (mesh.h:794-830)
            rans::mesh m(argv[2]);
            std::map<std::string, rans::boundary_condition> bcs;
            bcs["wall"].bc_type = argc > 3 ? argv[3] : "wall";
            bcs["farfield"].bc_type = "farfield";
            m.compute_wall_dist(bcs);
            bool same = true;
            double dmax = 0;
            for (size_t i = 0; i < m.cellsAreas.size(); ++i) {
                double mind = 1;
                bool first = true;
                for (size_t j = 0; j < m.boundaryEdges.size(); ++j) {
                    const std::string& t = bcs.at(m.boundaryEdgesPhysicals[j]).bc_type;
                    if (t == "wall" || t == "slip-wall") {
                        const double dx = m.cellsCentersX[i] - m.edgesCentersX[m.boundaryEdges[j]], dy = m.cellsCentersY[i] - m.edgesCentersY[m.boundaryEdges[j]];
                        const double d = std::sqrt(dx * dx + dy * dy);
                        mind = first ? d : std::min(mind, d);
                        first = false;
                    }
                }
                same = same && mind == m.wall_dist[i];
                dmax = std::max(dmax, m.wall_dist[i]);
            }
            std::printf("walldist cells=%zu nodes=%zu max=%.6g identical=%d tri0=%d n0=%u,%u,%u,%u\n", m.cellsAreas.size(), m.nodesX.size(), dmax, (int)same,
                        (int)m.cellsIsTriangle[0], m.cellsNodes(0, 0), m.cellsNodes(0, 1), m.cellsNodes(0, 2), m.cellsNodes(0, 3));
            return same ? 0 : 1;
        }
    } catch (std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "usage: test_host conf <ini> | prolong <coarse.msh> <fine.msh> <q.bin> <out.bin>\n");
    return 2;
}
