// mesh.h -- rans::mesh on top of the afx_mesh_* entry points (reference: src/rans/include/rans/mesh.h:198-273).
// Same public array names in the same order and orientation, so code written against the reference's mesh
// (post-processing, prolongation, the solver adapter) reads it unchanged.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "core.h"

namespace rans {

const uint MESH_EDGE_NULL = 4294967295u;

template <uint N>
class meshArray {  // mesh.h:126-141: (row, col) view of a flat uint vector
    const uint32_t* p_ = nullptr;
    uint n_ = 0;
public:
    meshArray() {}
    meshArray(const uint32_t* p, uint n) : p_(p), n_(n) {}
    const uint32_t& operator()(uint i, uint j) const { return p_[(size_t)i * N + j]; }
    uint cols() const { return n_; }
    uint rows() const { return N; }
};

class mesh {
    struct Deleter { void operator()(afx_mesh* m) const { afx_mesh_free(m); } };
    std::shared_ptr<afx_mesh> h_;
    afx_mesh_desc d_{};
    void bind() {
        if (afx_mesh_get_desc(h_.get(), &d_)) throw std::runtime_error(afx_last_error());
        nRealCells = d_.n_cells;
        const size_t NT = (size_t)d_.n_cells + d_.n_ghost, E = d_.n_edges;
        edgesCells = meshArray<2>(d_.edges_cells, d_.n_edges);
        cellsEdges = meshArray<4>(d_.cells_edges, d_.n_cells);
        edgesLengths.assign(d_.edges_len, d_.edges_len + E);
        edgesNormalsX.assign(d_.edges_nx, d_.edges_nx + E); edgesNormalsY.assign(d_.edges_ny, d_.edges_ny + E);
        edgesCentersX.assign(d_.edges_cx, d_.edges_cx + E); edgesCentersY.assign(d_.edges_cy, d_.edges_cy + E);
        cellsAreas.assign(d_.cells_area, d_.cells_area + NT);
        cellsCentersX.assign(d_.cells_cx, d_.cells_cx + NT); cellsCentersY.assign(d_.cells_cy, d_.cells_cy + NT);
        cellsIsTriangle.assign(NT, true);
        for (uint i = 0; i < d_.n_cells; ++i) cellsIsTriangle[i] = d_.cells_is_tri[i] != 0;
        boundaryEdges.assign(d_.boundary_edges, d_.boundary_edges + d_.n_ghost);
        boundaryEdgesPhysicals.clear();
        for (uint b = 0; b < d_.n_ghost; ++b) boundaryEdgesPhysicals.push_back(afx_mesh_patch_name(h_.get(), d_.boundary_patch[b]));
    }
public:
    std::string filename;
    meshArray<2> edgesCells;
    std::vector<double> edgesLengths, edgesNormalsX, edgesNormalsY, edgesCentersX, edgesCentersY;
    std::vector<uint> boundaryEdges;
    std::vector<std::string> boundaryEdgesPhysicals;
    meshArray<4> cellsEdges;
    std::vector<bool> cellsIsTriangle;
    std::vector<double> cellsAreas, cellsCentersX, cellsCentersY;
    uint nRealCells = 0;

    mesh() {}
    explicit mesh(std::string filename_) { read_file(filename_); }  // mesh.h:250

    void read_file(std::string filename_in) {  // mesh.h:834-884; std::invalid_argument / std::runtime_error on bad input
        filename = filename_in;
        afx_mesh* m = nullptr;
        const int rc = afx_mesh_read_msh(&m, filename.c_str());
        if (rc == AFX_ERR_INVALID) throw std::invalid_argument(afx_last_error());
        if (rc) throw std::runtime_error(afx_last_error());
        h_.reset(m, Deleter());
        bind();
    }
    static mesh synthetic_omesh(uint ni, uint nj, uint n_quad_layers, double far_radius = 150.0) {
        afx_mesh* m = nullptr;
        if (afx_mesh_synth_omesh(&m, ni, nj, n_quad_layers, far_radius)) throw std::runtime_error(afx_last_error());
        mesh r; r.h_.reset(m, Deleter()); r.bind();
        return r;
    }
    // the wall distance of the reference (mesh.h:794-830) feeds only its empty SA branch and the VTU writer: not kept
    void compute_wall_dist(const std::map<std::string, boundary_condition>&) {}

    const afx_mesh_desc& desc() const { return d_; }
    afx_mesh* handle() const { return h_.get(); }
    int patch_id(const std::string& name) const { return h_ ? afx_mesh_patch_id(h_.get(), name.c_str()) : -1; }
    int n_patches() const { return h_ ? afx_mesh_n_patches(h_.get()) : 0; }
    std::string patch_name(int p) const { const char* s = afx_mesh_patch_name(h_.get(), p); return s ? s : ""; }
};

}  // namespace rans
