// post.h -- wall profile and force coefficients (reference: src/rans/include/rans/post.h:41-55, 182-387).
// The integrals run on the device (afx_rans_wall_forces: warp-shuffle reduction); x/y come from the mesh.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "gpu_solver.h"

namespace rans {

struct wallProfile {  // post.h:41-55
    std::vector<double> x, y, cp;
    double cd = 0., cl = 0., cm = 0.;
    void reserve(const uint& n) { x.reserve(n); y.reserve(n); cp.reserve(n); }
};

inline wallProfile get_wall_profile(solver& solv, std::string patch_name) {  // post.h:301-387
    wallProfile wall;
    mesh& m = solv.get_mesh();
    const int patch = m.patch_id(patch_name);
    for (uint b = 0; b < m.boundaryEdges.size(); ++b)
        if (m.boundaryEdgesPhysicals[b] == patch_name) {
            wall.x.push_back(m.edgesCentersX[m.boundaryEdges[b]]);
            wall.y.push_back(m.edgesCentersY[m.boundaryEdges[b]]);
        }
    if (patch < 0 || wall.x.empty()) return wall;  // the reference divides by zero here; we return zeros
    wall.cp.resize(wall.x.size());
    if (afx_rans_wall_cp(solv.handle(), patch, wall.cp.data()) < 0) throw std::runtime_error(afx_last_error());
    double f[3];
    if (afx_rans_wall_forces(solv.handle(), patch, f)) throw std::runtime_error(afx_last_error());
    wall.cl = f[0]; wall.cd = f[1]; wall.cm = f[2];
    return wall;
}

class CpProfile {  // post.h:182-299, the part the solver loop touches
public:
    double x_moment = 0, y_moment = 0;
    std::mutex m_mutex;
    std::vector<double> x, y, cp;
    bool filled = false;

    void calc_chord_coords(solver& s, std::string& af) {
        std::scoped_lock lock(m_mutex);
        mesh& m = s.get_mesh();
        filled = false;
        x.clear(); y.clear(); cp.clear();
        for (uint b = 0; b < m.boundaryEdges.size(); ++b)
            if (m.boundaryEdgesPhysicals[b] == af) {
                x.push_back(m.edgesCentersX[m.boundaryEdges[b]]);
                y.push_back(m.edgesCentersY[m.boundaryEdges[b]]);
                cp.push_back(0.0);
            }
    }
    void calc_cp(solver& s, std::string& af) {
        std::scoped_lock lock(m_mutex);
        const int patch = s.get_mesh().patch_id(af);
        if (patch >= 0 && !cp.empty()) afx_rans_wall_cp(s.handle(), patch, cp.data());
        filled = true;
    }
};

}  // namespace rans
