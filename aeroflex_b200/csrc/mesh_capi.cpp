// mesh_capi.cpp -- afx_mesh_* entry points of the C ABI (host only, no CUDA).
#include <cstring>
#include <stdexcept>
#include <string>

#include "mesh_host.h"

namespace afx {
void set_error(const std::string& s);

afx_mesh_desc HostMesh::desc() const
{
    afx_mesh_desc d{};
    d.n_cells = N; d.n_ghost = G; d.n_edges = E;
    d.edges_cells = edge_cells.data();
    d.edges_nx = enx.data(); d.edges_ny = eny.data(); d.edges_len = elen.data();
    d.edges_cx = ecx.data(); d.edges_cy = ecy.data();
    d.cells_cx = ccx.data(); d.cells_cy = ccy.data(); d.cells_area = area.data();
    d.cells_edges = cell_edges.data(); d.cells_is_tri = is_tri.data();
    d.boundary_edges = bnd_edge.data(); d.boundary_patch = bpatch.data();
    return d;
}
}  // namespace afx

namespace {
template <class F>
int guard_mesh(afx_mesh** out, F&& f)
{
    if (!out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_mesh* m = new afx_mesh;
    try {
        f(m->m);
        *out = m;
        return AFX_OK;
    } catch (const std::invalid_argument& e) { afx::set_error(e.what()); delete m; return AFX_ERR_INVALID; }
    catch (const std::out_of_range& e) { afx::set_error(e.what()); delete m; return AFX_ERR_INVALID; }
    catch (const std::exception& e) { afx::set_error(e.what()); delete m; return AFX_ERR_IO; }
    catch (...) { afx::set_error("unknown error"); delete m; return AFX_ERR_IO; }
}
}  // namespace

extern "C" {

int afx_mesh_read_msh(afx_mesh** out, const char* path)
{
    if (!path) { afx::set_error("null path"); return AFX_ERR_INVALID; }
    return guard_mesh(out, [&](afx::HostMesh& m) { m.read_msh(path); });
}

int afx_mesh_from_elements(afx_mesh** out, uint32_t n_nodes, const double* x, const double* y, uint32_t n_cells,
                           const uint32_t* cells, const uint8_t* is_tri, uint32_t n_bnd, const uint32_t* b0,
                           const uint32_t* b1, const int32_t* bpatch, int n_patch, const char* const* patch_names)
{
    return guard_mesh(out, [&](afx::HostMesh& m) {
        if (!x || !y || !cells || !is_tri) throw std::invalid_argument("null mesh arrays");
        m.x.assign(x, x + n_nodes); m.y.assign(y, y + n_nodes);
        m.cells.assign(cells, cells + 4 * (size_t)n_cells);
        m.is_tri.assign(is_tri, is_tri + n_cells);
        for (size_t c = 0; c < n_cells; ++c)
            for (int k = 0; k < (m.is_tri[c] ? 3 : 4); ++k)
                if (m.cells[4 * c + k] >= n_nodes) throw std::invalid_argument("cell refers to a missing node");
        if (n_bnd) {
            if (!b0 || !b1) throw std::invalid_argument("null boundary arrays");
            m.b0.assign(b0, b0 + n_bnd); m.b1.assign(b1, b1 + n_bnd);
            if (bpatch) m.bpatch.assign(bpatch, bpatch + n_bnd); else m.bpatch.assign(n_bnd, 0);
        }
        for (int p = 0; p < n_patch; ++p) m.patch_names.push_back(patch_names && patch_names[p] ? patch_names[p] : ("patch" + std::to_string(p)));
        m.build();
    });
}

int afx_mesh_synth_omesh(afx_mesh** out, uint32_t ni, uint32_t nj, uint32_t n_quad_layers, double far_radius)
{
    return guard_mesh(out, [&](afx::HostMesh& m) { m.synth_omesh(ni, nj, n_quad_layers, far_radius); });
}

void afx_mesh_free(afx_mesh* m) { delete m; }

int afx_mesh_get_desc(const afx_mesh* m, afx_mesh_desc* out)
{
    if (!m || !out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = m->m.desc();
    return AFX_OK;
}

uint32_t afx_mesh_n_nodes(const afx_mesh* m) { return m ? (uint32_t)m->m.x.size() : 0u; }
int afx_mesh_n_patches(const afx_mesh* m) { return m ? (int)m->m.patch_names.size() : 0; }
const char* afx_mesh_patch_name(const afx_mesh* m, int p)
{
    return (m && p >= 0 && p < (int)m->m.patch_names.size()) ? m->m.patch_names[(size_t)p].c_str() : nullptr;
}
int afx_mesh_patch_id(const afx_mesh* m, const char* name)
{
    if (!m || !name) return -1;
    for (size_t p = 0; p < m->m.patch_names.size(); ++p)
        if (m->m.patch_names[p] == name) return (int)p;
    return -1;
}

int afx_mesh_get_elements(const afx_mesh* m, double* x, double* y, uint32_t* cells, uint32_t* b0, uint32_t* b1)
{
    const auto& h = m->m;
    if (x) std::memcpy(x, h.x.data(), h.x.size() * sizeof(double));
    if (y) std::memcpy(y, h.y.data(), h.y.size() * sizeof(double));
    if (cells) std::memcpy(cells, h.cells.data(), h.cells.size() * sizeof(uint32_t));
    if (b0) std::memcpy(b0, h.b0.data(), h.b0.size() * sizeof(uint32_t));
    if (b1) std::memcpy(b1, h.b1.data(), h.b1.size() * sizeof(uint32_t));
    return AFX_OK;
}

int afx_mesh_write_msh(const afx_mesh* m, const char* path)
{
    try { m->m.write_msh(path); return AFX_OK; }
    catch (const std::exception& e) { afx::set_error(e.what()); return AFX_ERR_IO; }
}

}  // extern "C"
