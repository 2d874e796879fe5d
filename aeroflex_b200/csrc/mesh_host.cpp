// mesh_host.cpp -- host-side mesh ingest for libaeroflex_rans_b200.so.
//
// Produces the arrays of rans::mesh (reference: src/rans/include/rans/mesh.h)
// in the reference's order and orientation, so that everything above the C ABI
// (edge ids, boundary order, ghost-cell numbering) means the same thing as in
// AeroFLEX.  The construction is array/sort based (no std::map), so it scales
// to the 16M-64M cell benchmark meshes:
//   half-edges (cell, side) are sorted by their node pair; the first
//   half-edge of a pair in (cell, side) order creates the edge, and edge ids
//   are the ranks of the creators -- which is exactly the numbering produced
//   by the reference's scan "append an edge the first time its node pair is
//   seen" (mesh.h:317-341, 844-846).
#include "mesh_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#if defined(_OPENMP)
#include <parallel/algorithm>
#define AFX_SORT __gnu_parallel::sort
#else
#define AFX_SORT std::sort
#endif

namespace afx {

namespace {

struct HalfEdge {
    uint64_t key;  // (min node << 32) | max node
    uint32_t idx;  // 4*cell + side
    bool operator<(const HalfEdge& o) const { return key != o.key ? key < o.key : idx < o.idx; }
};

inline uint64_t pair_key(uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; }

}  // namespace

// Connectivity + metrics + ghost cells.  mesh.h:317-453, 744-787, 834-884.
void HostMesh::build()
{
    const uint32_t nc = (uint32_t)(cells.size() / 4);
    N = nc;
    G = (uint32_t)b0.size();
    if (is_tri.size() < nc) throw std::invalid_argument("is_tri shorter than the cell list");

    // 1. half-edges in (cell, side) order, sides (0,1),(1,2),..,(last,0)  [mesh.h:320-323]
    std::vector<HalfEdge> he;
    he.reserve((size_t)4 * nc);
    for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t sz = is_tri[c] ? 3u : 4u;
        for (uint32_t s = 0; s < sz; ++s) {
            const uint32_t a = cells[4 * (size_t)c + s], b = cells[4 * (size_t)c + (s + 1 < sz ? s + 1 : 0)];
            he.push_back({pair_key(a, b), 4 * c + s});
        }
    }
    AFX_SORT(he.begin(), he.end());

    // 2. the smallest (cell, side) of every node pair creates the edge
    std::vector<uint8_t> creates((size_t)4 * nc, 0);
    for (size_t i = 0; i < he.size(); ++i)
        if (i == 0 || he[i].key != he[i - 1].key) creates[he[i].idx] = 1;
    std::vector<uint32_t> rank((size_t)4 * nc, 0);
    uint32_t ne = 0;
    for (size_t i = 0; i < creates.size(); ++i) { rank[i] = ne; ne += creates[i]; }
    E = ne;

    edge_cells.assign((size_t)2 * E, 0);
    edge_nodes.assign((size_t)2 * E, 0);
    cell_edges.assign((size_t)4 * nc, AFX_EDGE_NULL);
    // 3. hand the id to every half-edge of the group; later cells overwrite
    //    column 1 in cell order, as convert_node_face_info does (mesh.h:366-373)
    for (size_t i = 0; i < he.size();) {
        size_t j = i;
        while (j < he.size() && he[j].key == he[i].key) ++j;
        const uint32_t cr = he[i].idx, e = rank[cr];
        const uint32_t c = cr >> 2, s = cr & 3, sz = is_tri[c] ? 3u : 4u;
        edge_cells[2 * (size_t)e] = c;
        edge_nodes[2 * (size_t)e] = cells[4 * (size_t)c + s];
        edge_nodes[2 * (size_t)e + 1] = cells[4 * (size_t)c + (s + 1 < sz ? s + 1 : 0)];
        for (size_t k = i; k < j; ++k) {
            cell_edges[he[k].idx] = e;
            const uint32_t ck = he[k].idx >> 2;
            if (ck != c) edge_cells[2 * (size_t)e + 1] = ck;  // ascending idx => last cell wins
        }
        i = j;
    }

    const size_t NT = (size_t)N + G;
    ccx.assign(NT, 0.); ccy.assign(NT, 0.); area.assign(NT, 0.);
    enx.assign(E, 0.); eny.assign(E, 0.); elen.assign(E, 0.); ecx.assign(E, 0.); ecy.assign(E, 0.);
    is_tri.resize(NT, 1);

    // 4. metrics, compute_mesh (mesh.h:378-453)
#pragma omp parallel for
    for (int64_t c = 0; c < (int64_t)nc; ++c) {
        const uint32_t sz = is_tri[c] ? 3u : 4u;
        double sx = 0., sy = 0.;
        for (uint32_t j = 0; j < sz; ++j) {  // mean of the nodes, each term divided first (mesh.h:391-392)
            sx += x[cells[4 * (size_t)c + j]] / ((double)sz);
            sy += y[cells[4 * (size_t)c + j]] / ((double)sz);
        }
        ccx[c] = sx; ccy[c] = sy;
        const uint32_t* n = &cells[4 * (size_t)c];
        const double t1 = 0.5 * std::fabs(x[n[0]] * (y[n[1]] - y[n[2]]) + x[n[1]] * (y[n[2]] - y[n[0]]) + x[n[2]] * (y[n[0]] - y[n[1]]));
        if (is_tri[c]) area[c] = t1;  // mesh.h:435-437
        else area[c] = t1 + 0.5 * std::fabs(x[n[0]] * (y[n[2]] - y[n[3]]) + x[n[2]] * (y[n[3]] - y[n[0]]) + x[n[3]] * (y[n[0]] - y[n[2]]));  // mesh.h:442-446
    }
#pragma omp parallel for
    for (int64_t e = 0; e < (int64_t)E; ++e) {
        const uint32_t n0 = edge_nodes[2 * (size_t)e], n1 = edge_nodes[2 * (size_t)e + 1];
        ecx[e] = (x[n1] + x[n0]) * 0.5;  // mesh.h:403-404
        ecy[e] = (y[n1] + y[n0]) * 0.5;
        const double dex = x[n1] - x[n0], dey = y[n1] - y[n0];
        const double l = std::sqrt(dex * dex + dey * dey);
        elen[e] = l;
        double nx = -dey / l, ny = dex / l;  // mesh.h:410-411
        const uint32_t c0 = edge_cells[2 * (size_t)e];
        if (nx * (ecx[e] - ccx[c0]) + ny * (ecy[e] - ccy[c0]) < 0) { nx *= -1.; ny *= -1.; }  // out of cell 0, mesh.h:414-420
        enx[e] = nx; eny[e] = ny;
    }

    // 5. one ghost cell per boundary segment, in boundary order (mesh.h:744-787)
    bnd_edge.assign(G, 0);
    if (G) {
        std::vector<HalfEdge> bk(G);
        for (uint32_t b = 0; b < G; ++b) bk[b] = {pair_key(b0[b], b1[b]), b};
        AFX_SORT(bk.begin(), bk.end());
        // he is sorted by key: merge-join
        size_t i = 0;
        for (uint32_t k = 0; k < G; ++k) {
            while (i < he.size() && he[i].key < bk[k].key) ++i;
            if (i == he.size() || he[i].key != bk[k].key)
                throw std::invalid_argument("invalid edge ref (" + std::to_string((uint32_t)(bk[k].key >> 32)) + ", " +
                                            std::to_string((uint32_t)bk[k].key) + ")");  // mesh.h:757
            bnd_edge[bk[k].idx] = cell_edges[he[i].idx];
        }
        for (uint32_t b = 0; b < G; ++b) {
            const uint32_t e = bnd_edge[b], c = edge_cells[2 * (size_t)e];
            const double dx = ecx[e] - ccx[c], dy = ecy[e] - ccy[c];
            const double dist = std::sqrt(dx * dx + dy * dy);
            area[N + b] = area[c];
            ccx[N + b] = ecx[e] + dist * enx[e];
            ccy[N + b] = ecy[e] + dist * eny[e];
            is_tri[N + b] = 1;
            edge_cells[2 * (size_t)e + 1] = N + b;
        }
    }
}

// ---------------------------------------------------------------------------
// Gmsh MSH 4.1 ASCII (mesh.h:457-738).  Token based, section by section.
// ---------------------------------------------------------------------------
namespace {

std::vector<std::string> tokens(const std::string& line)
{
    std::vector<std::string> t;
    std::istringstream is(line);
    std::string w;
    while (is >> w) t.push_back(w);
    return t;
}

bool next_line(std::istream& in, std::string& line)
{
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.find_first_not_of(" \t") != std::string::npos) return true;
    }
    return false;
}

}  // namespace

void HostMesh::read_msh(const std::string& path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open mesh file " + path);
    std::map<long, std::string> phys_name;  // key = tag + 1000*dim (mesh.h:486)
    std::vector<long> phys_order;           // 1-D names in file order -> patch ids
    std::map<long, long> curve_phys;        // curve entity tag -> physical key (mesh.h:504)
    std::map<long, uint32_t> node_of_tag;   // file tag -> sequential id (mesh.h:567)
    std::map<long, int> patch_of_key;
    std::string line;
    bool have_format = false;
    while (next_line(in, line)) {
        if (line[0] != '$') continue;
        const std::string sec = line.substr(1);
        if (sec == "MeshFormat") {
            next_line(in, line);
            auto t = tokens(line);
            if (t.empty() || t[0].substr(0, 2) != "4." || (t.size() > 1 && t[1] != "0"))
                throw std::runtime_error("only Gmsh MSH 4.x ASCII is supported: " + path);
            have_format = true;
        } else if (sec == "PhysicalNames") {
            next_line(in, line);
            const long n = std::stol(line);
            for (long i = 0; i < n; ++i) {
                next_line(in, line);
                auto t = tokens(line);
                if (t.size() < 3) throw std::runtime_error("bad $PhysicalNames line");
                std::string nm = line.substr(line.find('"') == std::string::npos ? 0 : line.find('"'));
                nm.erase(std::remove(nm.begin(), nm.end(), '"'), nm.end());
                if (line.find('"') == std::string::npos) nm = t[2];
                const long dim = std::stol(t[0]), key = std::stol(t[1]) + 1000 * dim;
                phys_name[key] = nm;
                if (dim == 1) phys_order.push_back(key);
            }
        } else if (sec == "Entities") {
            next_line(in, line);
            auto t = tokens(line);
            if (t.size() < 4) throw std::runtime_error("bad $Entities header");
            const long np = std::stol(t[0]), ncv = std::stol(t[1]), ns = std::stol(t[2]), nv = std::stol(t[3]);
            for (long i = 0; i < np; ++i) next_line(in, line);
            for (long i = 0; i < ncv; ++i) {
                next_line(in, line);
                auto c = tokens(line);
                if (c.size() < 8) throw std::runtime_error("bad curve entity line");
                const long tag = std::stol(c[0]), nph = std::stol(c[7]);
                if (nph >= 1 && c.size() >= 9) curve_phys[tag] = std::stol(c[8]) + 1000;  // first physical tag
            }
            for (long i = 0; i < ns + nv; ++i) next_line(in, line);
        } else if (sec == "Nodes") {
            next_line(in, line);
            auto t = tokens(line);
            const long nblocks = std::stol(t.at(0));
            for (long b = 0; b < nblocks; ++b) {
                next_line(in, line);
                auto h = tokens(line);
                const long nb = std::stol(h.at(3));
                std::vector<long> tags((size_t)nb);
                for (long i = 0; i < nb; ++i) { next_line(in, line); tags[(size_t)i] = std::stol(line); }
                for (long i = 0; i < nb; ++i) {
                    next_line(in, line);
                    auto c = tokens(line);
                    node_of_tag[tags[(size_t)i]] = (uint32_t)x.size();
                    x.push_back(std::stod(c.at(0)));
                    y.push_back(std::stod(c.at(1)));
                }
            }
        } else if (sec == "Elements") {
            next_line(in, line);
            auto t = tokens(line);
            const long nblocks = std::stol(t.at(0));
            for (long b = 0; b < nblocks; ++b) {
                next_line(in, line);
                auto h = tokens(line);
                const long dim = std::stol(h.at(0)), ent = std::stol(h.at(1)), nb = std::stol(h.at(3));
                int patch = -1;
                if (dim == 1) {
                    auto it = curve_phys.find(ent);
                    if (it == curve_phys.end()) throw std::out_of_range("curve entity " + std::to_string(ent) + " has no physical tag");  // mesh.h:623 .at()
                    auto nm = phys_name.find(it->second);
                    if (nm == phys_name.end()) throw std::out_of_range("physical tag without a name");  // mesh.h:784 .at()
                    auto pk = patch_of_key.find(it->second);
                    if (pk == patch_of_key.end()) {
                        patch = (int)patch_names.size();
                        patch_of_key[it->second] = patch;
                        patch_names.push_back(nm->second);
                    } else patch = pk->second;
                }
                for (long i = 0; i < nb; ++i) {
                    next_line(in, line);
                    auto c = tokens(line);
                    const size_t nn = c.size() - 1;
                    if (dim == 1 && nn == 2) {
                        b0.push_back(node_of_tag.at(std::stol(c[1])));
                        b1.push_back(node_of_tag.at(std::stol(c[2])));
                        bpatch.push_back(patch);
                    } else if (dim == 2 && (nn == 3 || nn == 4)) {
                        for (size_t k = 0; k < nn; ++k) cells.push_back(node_of_tag.at(std::stol(c[k + 1])));
                        if (nn == 3) cells.push_back(0);  // mesh.h:715-717
                        is_tri.push_back(nn == 3);
                    }
                }
            }
        }
    }
    if (!have_format) throw std::runtime_error("not a Gmsh MSH file: " + path);
    // patches that exist by name but have no elements still get an id
    for (long key : phys_order)
        if (!patch_of_key.count(key)) { patch_of_key[key] = (int)patch_names.size(); patch_names.push_back(phys_name[key]); }
    build();
}

void HostMesh::write_msh(const std::string& path) const
{
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    const int np = (int)patch_names.size();
    std::fprintf(f, "$MeshFormat\n4.1 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n", np + 1);
    for (int p = 0; p < np; ++p) std::fprintf(f, "1 %d \"%s\"\n", p + 1, patch_names[(size_t)p].c_str());
    std::fprintf(f, "2 %d \"internal\"\n$EndPhysicalNames\n", np + 1);
    std::fprintf(f, "$Entities\n0 %d 1 0\n", np);
    for (int p = 0; p < np; ++p) std::fprintf(f, "%d 0 0 0 0 0 0 1 %d 0 \n", p + 1, p + 1);
    std::fprintf(f, "1 0 0 0 0 0 0 1 %d 0 \n$EndEntities\n", np + 1);
    const size_t nn = x.size();
    std::fprintf(f, "$Nodes\n1 %zu 1 %zu\n2 1 0 %zu\n", nn, nn, nn);
    for (size_t i = 0; i < nn; ++i) std::fprintf(f, "%zu\n", i + 1);
    for (size_t i = 0; i < nn; ++i) std::fprintf(f, "%.17g %.17g 0\n", x[i], y[i]);
    std::fprintf(f, "$EndNodes\n");
    // element blocks: runs of equal patch (1-D) and equal shape (2-D), file order preserved
    struct Run { int dim, tag, type; size_t lo, hi; };
    std::vector<Run> runs;
    for (size_t b = 0; b < b0.size();) {
        size_t e = b;
        while (e < b0.size() && bpatch[e] == bpatch[b]) ++e;
        runs.push_back({1, bpatch[b] + 1, 1, b, e});
        b = e;
    }
    for (size_t c = 0; c < N;) {
        size_t e = c;
        while (e < N && is_tri[e] == is_tri[c]) ++e;
        runs.push_back({2, 1, is_tri[c] ? 2 : 3, c, e});
        c = e;
    }
    const size_t nel = b0.size() + N;
    std::fprintf(f, "$Elements\n%zu %zu 1 %zu\n", runs.size(), nel, nel);
    size_t tag = 1;
    for (const Run& r : runs) {
        std::fprintf(f, "%d %d %d %zu\n", r.dim, r.tag, r.type, r.hi - r.lo);
        for (size_t i = r.lo; i < r.hi; ++i, ++tag) {
            if (r.dim == 1) std::fprintf(f, "%zu %u %u \n", tag, b0[i] + 1, b1[i] + 1);
            else if (r.type == 2) std::fprintf(f, "%zu %u %u %u \n", tag, cells[4 * i] + 1, cells[4 * i + 1] + 1, cells[4 * i + 2] + 1);
            else std::fprintf(f, "%zu %u %u %u %u \n", tag, cells[4 * i] + 1, cells[4 * i + 1] + 1, cells[4 * i + 2] + 1, cells[4 * i + 3] + 1);
        }
    }
    std::fprintf(f, "$EndElements\n");
    std::fclose(f);
}

// ---------------------------------------------------------------------------
// Synthetic NACA0012 O-mesh.  Deterministic.
// ---------------------------------------------------------------------------
void HostMesh::synth_omesh(uint32_t ni, uint32_t nj, uint32_t n_quad_layers, double far_radius)
{
    if (ni < 8 || nj < 2) throw std::invalid_argument("synthetic O-mesh needs ni >= 8 and nj >= 2");
    if (n_quad_layers > nj) n_quad_layers = nj;
    const double pi = 3.14159265358979323846;
    // geometric stretching ratio r with first-layer height h0 tied to the
    // surface spacing: h0 * (r^nj - 1)/(r - 1) = far_radius, h0 = (2/ni)/12
    const double h0 = 2.0 / ni / 12.0;
    double lo = 1.0 + 1e-9, hi = 2.0;
    for (int it = 0; it < 200; ++it) {
        const double r = 0.5 * (lo + hi);
        const double s = h0 * (std::pow(r, (double)nj) - 1.0) / (r - 1.0);
        (s > far_radius ? hi : lo) = r;
    }
    const double ratio = 0.5 * (lo + hi);
    std::vector<double> s(nj + 1);  // normalised radial coordinate, s[0]=0 .. s[nj]=1
    {
        double acc = 0., h = h0;
        s[0] = 0.;
        for (uint32_t j = 1; j <= nj; ++j) { acc += h; h *= ratio; s[j] = acc; }
        for (uint32_t j = 0; j <= nj; ++j) s[j] /= acc;
        s[nj] = 1.0;
    }
    x.resize((size_t)ni * (nj + 1)); y.resize((size_t)ni * (nj + 1));
    // surface points, counter-clockwise from the trailing edge over the upper side
    std::vector<double> xs(ni), ys(ni), nxs(ni), nys(ni);
    for (uint32_t i = 0; i < ni; ++i) {
        const double th = 2.0 * pi * i / ni;
        const double xc = 0.5 * (1.0 + std::cos(th));
        const double yt = 0.6 * (0.2969 * std::sqrt(xc) - 0.1260 * xc - 0.3516 * xc * xc + 0.2843 * xc * xc * xc - 0.1036 * xc * xc * xc * xc);
        xs[i] = xc;
        ys[i] = (i == 0) ? 0.0 : ((th <= pi) ? yt : -yt);
    }
    for (uint32_t i = 0; i < ni; ++i) {  // outward normals from central differences
        const uint32_t ip = (i + 1) % ni, im = (i + ni - 1) % ni;
        const double tx = xs[ip] - xs[im], ty = ys[ip] - ys[im], tl = std::sqrt(tx * tx + ty * ty);
        nxs[i] = ty / tl; nys[i] = -tx / tl;
    }
    const double d0 = 0.05, d1 = 2.0;  // chords: normal extrusion below d0, radial rays beyond d1
    for (uint32_t i = 0; i < ni; ++i) {
        const double th = 2.0 * pi * i / ni;
        const double xo = 0.5 + far_radius * std::cos(th), yo = far_radius * std::sin(th);
        for (uint32_t j = 0; j <= nj; ++j) {
            const double t = s[j], d = far_radius * t;
            double b = (d - d0) / (d1 - d0);
            b = b < 0 ? 0 : (b > 1 ? 1 : b);
            b = b * b * (3.0 - 2.0 * b);
            const double ax = xs[i] + nxs[i] * d, ay = ys[i] + nys[i] * d;          // along the normal
            const double bx = (1.0 - t) * xs[i] + t * xo, by = (1.0 - t) * ys[i] + t * yo;  // towards the far circle
            x[(size_t)j * ni + i] = (1.0 - b) * ax + b * bx;
            y[(size_t)j * ni + i] = (1.0 - b) * ay + b * by;
        }
    }
    cells.clear(); is_tri.clear(); b0.clear(); b1.clear(); bpatch.clear();
    patch_names = {"wall", "farfield"};
    const size_t ncell = (size_t)ni * n_quad_layers + (size_t)2 * ni * (nj - n_quad_layers);
    cells.reserve(4 * ncell); is_tri.reserve(ncell);
    for (uint32_t j = 0; j < nj; ++j)
        for (uint32_t i = 0; i < ni; ++i) {
            const uint32_t ip = (i + 1 == ni) ? 0 : i + 1;
            const uint32_t a = j * ni + i, b = j * ni + ip, c = (j + 1) * ni + ip, d = (j + 1) * ni + i;
            if (j < n_quad_layers) {
                cells.insert(cells.end(), {a, d, c, b});
                is_tri.push_back(0);
            } else if ((i + j) & 1) {
                cells.insert(cells.end(), {a, d, c, 0u}); is_tri.push_back(1);
                cells.insert(cells.end(), {a, c, b, 0u}); is_tri.push_back(1);
            } else {
                cells.insert(cells.end(), {a, d, b, 0u}); is_tri.push_back(1);
                cells.insert(cells.end(), {d, c, b, 0u}); is_tri.push_back(1);
            }
        }
    for (uint32_t i = 0; i < ni; ++i) { b0.push_back(i); b1.push_back(i + 1 == ni ? 0 : i + 1); bpatch.push_back(0); }
    for (uint32_t i = 0; i < ni; ++i) { b0.push_back(nj * ni + i); b1.push_back(nj * ni + (i + 1 == ni ? 0 : i + 1)); bpatch.push_back(1); }
    // every cell must be counter-clockwise with positive area
    for (size_t c = 0; c < is_tri.size(); ++c) {
        const uint32_t* n = &cells[4 * c];
        const uint32_t sz = is_tri[c] ? 3u : 4u;
        double a2 = 0;
        for (uint32_t k = 0; k < sz; ++k) { const uint32_t p = n[k], q = n[(k + 1) % sz]; a2 += x[p] * y[q] - x[q] * y[p]; }
        if (!(a2 > 0)) throw std::runtime_error("synthetic O-mesh produced an inverted cell; change ni/nj");
    }
    build();
}

}  // namespace afx
