// rans.h -- the rans::Rans facade (reference: src/rans/include/rans/rans.h:25-123): settings, residual history,
// alpha list and solve_airfoil(), which fills the polar database the VLM viscous correction reads.
#pragma once
#include <string>

#include "multigrid.h"

namespace rans {

class Rans {
public:
    std::vector<double> residuals = {1.0};
    std::atomic<int> iters = 0;
    CpProfile profile;
    Settings settings;
    std::vector<mesh> ms;
    std::vector<double> alphas;
    bool mesh_loaded = false;
    GUIHandler& gui;
    // where <airfoil>_coarse.msh / <airfoil>_mid.msh live; the reference hard-codes this relative path (rans.h:84-85)
    std::string mesh_dir = "../../../../examples/rans/";
    bool verbose = true;
    // the reference writes <airfoil>_<alpha>.vtu into the working directory after every angle (rans.h:103); off by default
    // here (a 64-angle sweep is 64 files), on with save_vtu -- same file names, under vtu_dir
    bool save_vtu = false;
    std::string vtu_dir = "";

    explicit Rans(GUIHandler& gui) : gui(gui) {}

    void compute_alphas() {  // rans.h:53-57
        for (double i = settings.alpha_start; i <= settings.alpha_end; i += settings.alpha_step) alphas.push_back(i);
    }
    void input() {  // rans.h:59-66
        if (!mesh_loaded) {
            for (const auto& mesh_name : settings.meshes) ms.push_back(mesh(mesh_name));
            mesh_loaded = true;
        }
    }

    template <class T>
    void run_airfoil(const std::string& airfoil, database::airfoil& db) {  // rans.h:78-106
        gui.msg.push("[RANS] Solving airfoil: " + airfoil);
        ms.clear();
        ms.push_back(mesh(mesh_dir + airfoil + "_coarse.msh"));
        ms.push_back(mesh(mesh_dir + airfoil + "_mid.msh"));
        settings.bcs["farfield"].vars_far.angle = db.alpha.at(0) * 0.01745;
        multigrid<T> multi(ms, settings, gui, residuals, iters, profile);
        multi.verbose = verbose;
        multi.solvers[0].init();
        for (auto& alpha : db.alpha) {
            gui.msg.push("[RANS] Solving for alpha = " + std::to_string(alpha) + " deg.");
            settings.bcs["farfield"].vars_far.angle = alpha * 0.01745;  // the reference's deg -> rad literal (rans.h:94)
            for (auto& s : multi.solvers) s.set_bcs(settings.bcs);
            rans::solver& s = multi.run(false);
            wallProfile wp = get_wall_profile(s, "wall");
            db.cl.push_back(wp.cl);
            db.cd.push_back(wp.cd);
            db.cmy.push_back(wp.cm);
            if (save_vtu) save(vtu_dir + airfoil + "_" + std::to_string(alpha) + ".vtu", s);  // rans.h:103
            if (gui.signal.stop) break;
        }
    }

    void solve_airfoil(const std::string& airfoil, database::airfoil& db) {  // rans.h:116-123
        iters = 0;
        if (settings.solver_type() == "implicit") run_airfoil<implicitSolver>(airfoil, db);
        else if (settings.solver_type() == "explicit") run_airfoil<explicitSolver>(airfoil, db);
    }

    template <class T>
    void run() {  // rans.h:68-76: multigrid over settings.meshes, then the VTU file
        input();
        multigrid<T> multi(ms, settings, gui, residuals, iters, profile);
        multi.verbose = verbose;
        rans::solver& s = multi.run(true);
        save(settings.outfilename, s);
        std::cout << "Saved results to file " << settings.outfilename << "\n" << std::endl;
    }
    void solve() {  // rans.h:108-114
        if (settings.solver_type() == "implicit") run<implicitSolver>();
        else if (settings.solver_type() == "explicit") run<explicitSolver>();
    }

    template <class T>
    rans::solver* run_meshes(std::unique_ptr<multigrid<T>>& keep) {  // rans.h:68-76 without the VTU writer
        keep.reset(new multigrid<T>(ms, settings, gui, residuals, iters, profile));
        keep->verbose = verbose;
        return &keep->run(true);
    }
};

}  // namespace rans
