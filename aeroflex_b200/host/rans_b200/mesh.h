// mesh.h -- rans::mesh on top of the afx_mesh_* entry points (reference: src/rans/include/rans/mesh.h:198-273).
// Same public array names in the same order and orientation, so code written against the reference's mesh
// (post-processing, prolongation, the solver adapter) reads it unchanged.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
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
        // nodes and connectivity (mesh.h:209-216), for the VTU writer
        const uint32_t nn = afx_mesh_n_nodes(h_.get());
        nodesX.assign(nn, 0.); nodesY.assign(nn, 0.);
        cells_nodes_ = std::make_shared<std::vector<uint32_t>>((size_t)4 * d_.n_cells, 0u);  // shared: copies of the mesh keep the view valid
        if (afx_mesh_get_elements(h_.get(), nodesX.data(), nodesY.data(), cells_nodes_->data(), nullptr, nullptr)) throw std::runtime_error(afx_last_error());
        cellsNodes = meshArray<4>(cells_nodes_->data(), d_.n_cells);
        wall_dist.assign(NT, 1.0);
    }
    std::shared_ptr<std::vector<uint32_t>> cells_nodes_;
public:
    std::string filename;
    meshArray<2> edgesCells;
    std::vector<double> edgesLengths, edgesNormalsX, edgesNormalsY, edgesCentersX, edgesCentersY;
    std::vector<uint> boundaryEdges;
    std::vector<std::string> boundaryEdgesPhysicals;
    meshArray<4> cellsEdges;
    std::vector<bool> cellsIsTriangle;
    std::vector<double> cellsAreas, cellsCentersX, cellsCentersY;
    std::vector<double> nodesX, nodesY;
    meshArray<4> cellsNodes;      // triangles: the 4th entry repeats node 0 (mesh.h:715-722)
    std::vector<double> wall_dist;
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
    // Wall distance, mesh.h:794-830: for every cell (ghosts included) the smallest distance from its centre to the
    // CENTRE of a boundary edge whose bc_type is "wall" or "slip-wall"; 1 if there is none.  The reference scans all
    // N x G pairs; here the wall-edge centres go into a k-d tree.  sqrt is monotonic and correctly rounded, so the
    // square root of the smallest squared distance is the reference's smallest distance bit for bit.
    void compute_wall_dist(const std::map<std::string, boundary_condition>& bcs) {
        std::vector<double> wx, wy;
        for (size_t j = 0; j < boundaryEdges.size(); ++j) {
            const std::string& t = bcs.at(boundaryEdgesPhysicals[j]).bc_type;
            if (t == "wall" || t == "slip-wall") { wx.push_back(edgesCentersX[boundaryEdges[j]]); wy.push_back(edgesCentersY[boundaryEdges[j]]); }
        }
        wall_dist.assign(cellsAreas.size(), 1.0);
        if (wx.empty()) return;
        struct Node { double x0, x1, y0, y1; uint32_t lo, hi; int left, right; };
        std::vector<uint32_t> idx(wx.size());
        for (size_t k = 0; k < idx.size(); ++k) idx[k] = (uint32_t)k;
        std::vector<Node> nodes;
        struct Builder {
            std::vector<uint32_t>& idx; std::vector<Node>& nodes; const std::vector<double>& wx; const std::vector<double>& wy;
            int build(uint32_t lo, uint32_t hi) {
                Node nd{1e300, -1e300, 1e300, -1e300, lo, hi, -1, -1};
                for (uint32_t k = lo; k < hi; ++k) {
                    nd.x0 = std::min(nd.x0, wx[idx[k]]); nd.x1 = std::max(nd.x1, wx[idx[k]]);
                    nd.y0 = std::min(nd.y0, wy[idx[k]]); nd.y1 = std::max(nd.y1, wy[idx[k]]);
                }
                const int me = (int)nodes.size();
                nodes.push_back(nd);
                if (hi - lo > 8) {
                    const bool by_x = (nd.x1 - nd.x0) >= (nd.y1 - nd.y0);
                    const uint32_t mid = lo + (hi - lo) / 2;
                    std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                                     [&](uint32_t a, uint32_t b) { return by_x ? wx[a] < wx[b] : wy[a] < wy[b]; });
                    const int l = build(lo, mid), r = build(mid, hi);
                    nodes[(size_t)me].left = l; nodes[(size_t)me].right = r;
                }
                return me;
            }
        } builder{idx, nodes, wx, wy};
        builder.build(0, (uint32_t)idx.size());
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)cellsAreas.size(); ++i) {
            const double cx = cellsCentersX[(size_t)i], cy = cellsCentersY[(size_t)i];
            double best = 1e300;
            int stack[64], top = 0;
            stack[top++] = 0;
            while (top) {
                const Node& nd = nodes[(size_t)stack[--top]];
                const double bx = cx < nd.x0 ? nd.x0 - cx : (cx > nd.x1 ? cx - nd.x1 : 0.), by = cy < nd.y0 ? nd.y0 - cy : (cy > nd.y1 ? cy - nd.y1 : 0.);
                if (bx * bx + by * by > best) continue;  // monotonic rounding: never prunes the minimiser
                if (nd.left < 0) {
                    for (uint32_t k = nd.lo; k < nd.hi; ++k) {
                        const double dx = cx - wx[idx[k]], dy = cy - wy[idx[k]];
                        best = std::min(best, dx * dx + dy * dy);
                    }
                } else { stack[top++] = nd.left; stack[top++] = nd.right; }
            }
            wall_dist[(size_t)i] = std::sqrt(best);
        }
    }

    const afx_mesh_desc& desc() const { return d_; }
    afx_mesh* handle() const { return h_.get(); }
    int patch_id(const std::string& name) const { return h_ ? afx_mesh_patch_id(h_.get(), name.c_str()) : -1; }
    int n_patches() const { return h_ ? afx_mesh_n_patches(h_.get()) : 0; }
    std::string patch_name(int p) const { const char* s = afx_mesh_patch_name(h_.get(), p); return s ? s : ""; }
};

}  // namespace rans
