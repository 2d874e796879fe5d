// mesh_host.h -- host mesh container behind the afx_mesh handle of the C ABI.
// Holds the arrays of rans::mesh (reference mesh.h:209-246) in reference order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/afx_rans.h"

namespace afx {

struct HostMesh {
    // inputs (file order)
    std::vector<double> x, y;          // nodesX / nodesY
    std::vector<uint32_t> cells;       // [N][4] cellsNodes, triangles padded with node 0
    std::vector<uint8_t> is_tri;       // [N+G] cellsIsTriangle (ghosts are "triangles", mesh.h:775)
    std::vector<uint32_t> b0, b1;      // boundaryEdges0/1
    std::vector<int32_t> bpatch;       // patch id of every boundary segment
    std::vector<std::string> patch_names;
    // derived (build())
    uint32_t N = 0, G = 0, E = 0;
    std::vector<uint32_t> edge_cells, edge_nodes;  // [E][2]
    std::vector<double> enx, eny, elen, ecx, ecy;  // [E]
    std::vector<double> ccx, ccy, area;            // [N+G]
    std::vector<uint32_t> cell_edges;              // [N][4]
    std::vector<uint32_t> bnd_edge;                // [G]

    void build();
    void read_msh(const std::string& path);
    void write_msh(const std::string& path) const;
    void synth_omesh(uint32_t ni, uint32_t nj, uint32_t n_quad_layers, double far_radius);
    afx_mesh_desc desc() const;
};

}  // namespace afx

struct afx_mesh {
    afx::HostMesh m;
};
