// post.h -- wall profile and force coefficients (reference: src/rans/include/rans/post.h:41-55, 182-387).
// The integrals run on the device (afx_rans_wall_forces: warp-shuffle reduction); x/y come from the mesh.
#pragma once
#include <fstream>
#include <iomanip>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "gpu_solver.h"

namespace rans {

inline std::string double2string(const double& x, const int precision) {  // post.h:32-36
    std::stringstream stream;
    stream << std::fixed << std::setprecision(precision) << x;
    return stream.str();
}

// VTU writer, post.h:58-180: same XML layout, array names, order and number formatting (points with std::to_string,
// cell data with 16 fixed decimals), so files diff clean against the reference's for equal states.  The state is
// fetched from the GPU once; the wall distance is the mesh's (compute_wall_dist).
inline void save(const std::string filename, solver& solv) {
    solution q = solv.get_solution();
    mesh& m = solv.get_mesh();
    std::ofstream out(filename);
    out << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"BigEndian\">\n";
    out << "  <UnstructuredGrid>\n";
    out << "    <Piece NumberOfPoints=\"" << m.nodesX.size() << "\" NumberOfCells=\"" << m.nRealCells << "\">\n";
    out << "      <Points>\n";
    out << "        <DataArray type=\"Float32\" NumberOfComponents=\"3\" Format=\"ascii\">\n";
    for (size_t i = 0; i < m.nodesX.size(); ++i) out << "          " << std::to_string(m.nodesX[i]) << " " << std::to_string(m.nodesY[i]) << " 0.0\n";
    out << "        </DataArray>\n";
    out << "      </Points>\n";
    out << "      <Cells>\n";
    out << "        <DataArray type=\"Int32\" Name=\"connectivity\" Format=\"ascii\">\n";
    out << "          ";
    for (uint fi = 0; fi < m.nRealCells; ++fi) {
        const uint nn = m.cellsIsTriangle[fi] ? 3u : 4u;
        for (uint k = 0; k < nn; ++k) out << m.cellsNodes(fi, k) << "  ";
        out << "\n          ";
    }
    out << "        </DataArray>\n";
    out << "        <DataArray type=\"Int32\" Name=\"offsets\" Format=\"ascii\">\n";
    out << "          ";
    uint current_offset = 0;
    for (uint i = 0; i < m.nRealCells; ++i) { current_offset += m.cellsIsTriangle[i] ? 3 : 4; out << current_offset << "  "; }
    out << "\n";
    out << "        </DataArray>\n";
    out << "        <DataArray type=\"Int32\" Name=\"types\" Format=\"ascii\">\n";
    out << "          ";
    for (uint i = 0; i < m.nRealCells; ++i) out << (m.cellsIsTriangle[i] ? "5  " : "9  ");
    out << "\n";
    out << "        </DataArray>\n";
    out << "      </Cells>\n";
    out << "      <CellData Scalars=\"scalars\">\n";
    auto scalar = [&](const char* name, auto&& f) {
        out << "        <DataArray type=\"Float32\" Name=\"" << name << "\" Format=\"ascii\">\n";
        for (uint i = 0; i < m.nRealCells; ++i) out << "          " << double2string(f((int)i), 16) << "\n";
        out << "        </DataArray>\n";
    };
    scalar("Wall Distance", [&](int i) { return m.wall_dist[(size_t)i]; });
    scalar("Mach", [&](int i) { return q.mach(i); });
    scalar("Density", [&](int i) { return q.rho(i); });
    scalar("Pressure", [&](int i) { return q.p(i); });
    scalar("Temperature", [&](int i) { return q.T(i); });
    out << "        <DataArray type=\"Float32\" Name=\"Velocity\" NumberOfComponents=\"3\" Format=\"ascii\">\n";
    for (uint i = 0; i < m.nRealCells; ++i) out << "          " << double2string(q.u((int)i), 16) << " " << double2string(q.v((int)i), 16) << " 0.0\n";
    out << "        </DataArray>\n";
    out << "      </CellData>\n";
    out << "    </Piece>\n";
    out << "  </UnstructuredGrid>\n";
    out << "</VTKFile>";
}

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
