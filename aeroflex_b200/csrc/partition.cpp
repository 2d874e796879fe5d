// partition.cpp -- see partition.h.  Host only.
#include "partition.h"

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <stdexcept>
#include <string>

#include "ordering.h"

namespace afx {

void set_error(const std::string& s);

namespace {

// the (up to 4) real neighbours of a real cell through its edges
template <class F>
inline void for_neighbours(const afx_mesh_desc& g, uint32_t c, F&& f)
{
    const uint32_t sz = g.cells_is_tri[c] ? 3u : 4u;
    for (uint32_t k = 0; k < sz; ++k) {
        const uint32_t e = g.cells_edges[4 * (size_t)c + k];
        if (e == AFX_EDGE_NULL) continue;
        const uint32_t a = g.edges_cells[2 * (size_t)e], b = g.edges_cells[2 * (size_t)e + 1];
        const uint32_t n = (a == c) ? b : a;
        if (n < g.n_cells) f(n);
    }
}

}  // namespace

void Partition::build(const afx_mesh_desc& g, int nranks_, int rank_)
{
    if (nranks_ < 1 || rank_ < 0 || rank_ >= nranks_) throw std::invalid_argument("bad rank / number of ranks");
    rank = rank_; nranks = nranks_;
    const uint32_t N = g.n_cells, G = g.n_ghost, E = g.n_edges;
    if ((uint32_t)nranks > N) throw std::invalid_argument("more ranks than cells");
    n_global_cells = N; n_global_ghost = G; n_global_edges = E;

    // 1. global order and ownership: rank r owns cells order[cut[r] .. cut[r+1])
    //    default: Hilbert curve over rank coordinates, cut evenly;
    //    AFX_PARTITION=graph: recursive graph bisection of the face-neighbour graph into nranks pieces (METIS-style graph
    //    partition; pieces in breadth-first order).  Every rank computes the same deterministic result.
    std::vector<uint32_t> all(N);
    std::iota(all.begin(), all.end(), 0u);
    std::vector<uint32_t> order;
    std::vector<uint32_t> cut(nranks + 1);
    const char* pm = getenv("AFX_PARTITION");
    if (pm && std::string(pm) == "graph") {
        std::vector<uint32_t> nbr((size_t)4 * N, 0xFFFFFFFFu);
        for (uint32_t c = 0; c < N; ++c) {
            uint32_t k = 0;
            for_neighbours(g, c, [&](uint32_t n) { nbr[4 * (size_t)c + k++] = n; });
        }
        std::vector<uint32_t> sizes;
        order = graph_partition_order(N, nbr.data(), std::move(all), (uint32_t)nranks, sizes);
        if ((int)sizes.size() != nranks) throw std::logic_error("partition: graph bisection returned a wrong number of pieces");
        cut[0] = 0;
        for (int r = 0; r < nranks; ++r) cut[r + 1] = cut[r] + sizes[r];
    } else if (pm && std::string(pm) != "hilbert" && pm[0]) {
        throw std::invalid_argument("AFX_PARTITION must be hilbert or graph");
    } else {
        order = hilbert_order(g.cells_cx, g.cells_cy, all);
        // Cut by WORK, not by cell count: the face kernel's time goes with the faces, and a quadrilateral carries 4/2 of them
        // against a triangle's 3/2.  With the per-cell and per-face times measured on one B200 at 16M cells (0.231 ns per cell
        // and iteration in the cell kernels, 0.0876 ns per face in the face kernel) a quad costs 1.12 triangles; on the 16M
        // O-mesh an all-quad piece of equal cell count ran 9 % longer than the average piece, and every stage ends with a
        // halo hand-off that waits for the slowest neighbour.  AFX_PARTITION_QUAD_WEIGHT overrides (percent; 100 = by cells).
        uint64_t wq = 112;
        if (const char* e = getenv("AFX_PARTITION_QUAD_WEIGHT")) wq = (uint64_t)std::max(1, atoi(e));
        std::vector<uint64_t> pre(N + 1, 0);
        for (uint32_t k = 0; k < N; ++k) pre[k + 1] = pre[k] + (g.cells_is_tri[order[k]] ? 100u : wq);
        cut[0] = 0; cut[nranks] = N;
        for (int r = 1; r < nranks; ++r) {
            const uint64_t want = pre[N] / (uint64_t)nranks * (uint64_t)r + pre[N] % (uint64_t)nranks * (uint64_t)r / (uint64_t)nranks;
            uint32_t c = (uint32_t)(std::lower_bound(pre.begin(), pre.end(), want) - pre.begin());
            c = std::max<uint32_t>(c, cut[r - 1] + 1);                       // every rank owns at least one cell
            cut[r] = std::min<uint32_t>(c, N - (uint32_t)(nranks - r));
        }
    }
    std::vector<uint32_t> pos(N);
    for (uint32_t k = 0; k < N; ++k) pos[order[k]] = k;
    auto owner_of = [&](uint32_t c) { return (int)(std::upper_bound(cut.begin(), cut.end(), pos[c]) - cut.begin()) - 1; };

    // 2. classes: 0 owned, 1 ring 1, 2 ring 2, 255 absent
    std::vector<uint8_t> cls(N, 255);
    std::vector<uint32_t> own(order.begin() + cut[rank], order.begin() + cut[rank + 1]);
    for (uint32_t c : own) cls[c] = 0;
    std::vector<uint32_t> r1, r2;
    for (uint32_t c : own) for_neighbours(g, c, [&](uint32_t n) { if (cls[n] == 255) { cls[n] = 1; r1.push_back(n); } });
    for (uint32_t c : r1) for_neighbours(g, c, [&](uint32_t n) { if (cls[n] == 255) { cls[n] = 2; r2.push_back(n); } });
    auto by_owner_then_curve = [&](uint32_t a, uint32_t b) { return pos[a] < pos[b]; };  // owner is monotone in pos
    std::sort(r1.begin(), r1.end(), by_owner_then_curve);
    std::sort(r2.begin(), r2.end(), by_owner_then_curve);
    n_own = (uint32_t)own.size(); n_r1 = (uint32_t)r1.size(); n_r2 = (uint32_t)r2.size();
    const uint32_t NR = n_real();

    std::vector<uint32_t> g2l(N + (size_t)G, AFX_EDGE_NULL);
    cell_l2g.clear();
    cell_l2g.reserve(NR);
    for (const auto* v : {&own, &r1, &r2}) for (uint32_t c : *v) { g2l[c] = (uint32_t)cell_l2g.size(); cell_l2g.push_back(c); }

    // 3. local edges = edges of owned and ring-1 cells, ascending global id
    edge_l2g.clear();
    for (uint32_t l = 0; l < n_own + n_r1; ++l) {
        const uint32_t c = cell_l2g[l];
        const uint32_t sz = g.cells_is_tri[c] ? 3u : 4u;
        for (uint32_t k = 0; k < sz; ++k) edge_l2g.push_back(g.cells_edges[4 * (size_t)c + k]);
    }
    std::sort(edge_l2g.begin(), edge_l2g.end());
    edge_l2g.erase(std::unique(edge_l2g.begin(), edge_l2g.end()), edge_l2g.end());
    const uint32_t EL = (uint32_t)edge_l2g.size();
    auto edge_local = [&](uint32_t e) { return (uint32_t)(std::lower_bound(edge_l2g.begin(), edge_l2g.end(), e) - edge_l2g.begin()); };

    // 4. boundary ghosts of the local edges, ascending global boundary index
    std::vector<uint8_t> edge_is_local_bnd;
    bnd_l2g.clear();
    for (uint32_t b = 0; b < G; ++b) {
        const uint32_t e = g.boundary_edges[b];
        const uint32_t le = edge_local(e);
        if (le < EL && edge_l2g[le] == e) bnd_l2g.push_back(b);
    }
    n_bc = (uint32_t)bnd_l2g.size();
    for (uint32_t k = 0; k < n_bc; ++k) {
        const uint32_t gc = g.edges_cells[2 * (size_t)g.boundary_edges[bnd_l2g[k]] + 1];  // global ghost cell id
        g2l[gc] = NR + k;
        cell_l2g.push_back(gc);
    }
    const uint32_t NT = NR + n_bc;

    // 5. local arrays
    edge_cells.resize(2 * (size_t)EL); enx.resize(EL); eny.resize(EL); elen.resize(EL); ecx.resize(EL); ecy.resize(EL);
    for (uint32_t l = 0; l < EL; ++l) {
        const uint32_t e = edge_l2g[l];
        const uint32_t a = g2l[g.edges_cells[2 * (size_t)e]], b = g2l[g.edges_cells[2 * (size_t)e + 1]];
        if (a == AFX_EDGE_NULL || b == AFX_EDGE_NULL) throw std::logic_error("partition: edge with a cell outside the halo");
        edge_cells[2 * (size_t)l] = a; edge_cells[2 * (size_t)l + 1] = b;
        enx[l] = g.edges_nx[e]; eny[l] = g.edges_ny[e]; elen[l] = g.edges_len[e]; ecx[l] = g.edges_cx[e]; ecy[l] = g.edges_cy[e];
    }
    ccx.resize(NT); ccy.resize(NT); area.resize(NT); is_tri.assign(NT, 1);
    for (uint32_t l = 0; l < NT; ++l) {
        const uint32_t c = cell_l2g[l];
        ccx[l] = g.cells_cx[c]; ccy[l] = g.cells_cy[c]; area[l] = g.cells_area[c];
        if (l < NR) is_tri[l] = g.cells_is_tri[c];
    }
    cell_edges.assign(4 * (size_t)NR, AFX_EDGE_NULL);
    for (uint32_t l = 0; l < n_own + n_r1; ++l) {
        const uint32_t c = cell_l2g[l];
        const uint32_t sz = g.cells_is_tri[c] ? 3u : 4u;
        for (uint32_t k = 0; k < sz; ++k) cell_edges[4 * (size_t)l + k] = edge_local(g.cells_edges[4 * (size_t)c + k]);
    }
    bnd_edge.resize(n_bc); bnd_patch.resize(n_bc);
    for (uint32_t k = 0; k < n_bc; ++k) {
        bnd_edge[k] = edge_local(g.boundary_edges[bnd_l2g[k]]);
        bnd_patch[k] = g.boundary_patch[bnd_l2g[k]];
    }

    // 6. exchange plan.  Receive side: ring cells grouped by owner, curve order inside a group.
    //    Send side: an owned cell goes to every other rank that owns a cell within two hops of it --
    //    the same set the receiver derives, in the same (curve) order.
    peers.clear();
    std::vector<int> peer_slot(nranks, -1);
    auto slot_of = [&](int r) {
        if (peer_slot[r] < 0) { peer_slot[r] = (int)peers.size(); peers.push_back(Peer{r, {}, {}}); }
        return peer_slot[r];
    };
    std::vector<uint32_t> halo(r1);
    halo.insert(halo.end(), r2.begin(), r2.end());
    std::sort(halo.begin(), halo.end(), by_owner_then_curve);
    for (uint32_t c : halo) peers[slot_of(owner_of(c))].recv.push_back(g2l[c]);
    std::vector<int> seen;
    for (uint32_t l = 0; l < n_own; ++l) {  // local order of owned cells IS curve order
        const uint32_t c = cell_l2g[l];
        seen.clear();
        auto visit = [&](uint32_t n) {
            if (cls[n] == 0) return;
            const int r = owner_of(n);
            if (std::find(seen.begin(), seen.end(), r) == seen.end()) seen.push_back(r);
        };
        for_neighbours(g, c, [&](uint32_t n1) { visit(n1); for_neighbours(g, n1, [&](uint32_t n2) { visit(n2); }); });
        for (int r : seen) peers[slot_of(r)].send.push_back(l);
    }
    std::sort(peers.begin(), peers.end(), [](const Peer& a, const Peer& b) { return a.rank < b.rank; });

    // 7. global patch extents (post.h:314-338)
    int npatch = 0;
    for (uint32_t b = 0; b < G; ++b) npatch = std::max(npatch, g.boundary_patch[b] + 1);
    patch_xmin.assign(npatch, 0.); patch_xmax.assign(npatch, 0.); patch_ysum.assign(npatch, 0.); patch_count.assign(npatch, 0);
    patch_order.clear();
    for (uint32_t b = 0; b < G; ++b) {
        const int p = g.boundary_patch[b];
        if (p < 0) continue;
        const uint32_t e = g.boundary_edges[b];
        if (!patch_count[p]) patch_order.push_back(p);
        if (!patch_count[p]) { patch_xmin[p] = patch_xmax[p] = g.edges_cx[e]; patch_ysum[p] = g.edges_cy[e]; }
        else { patch_xmin[p] = std::min(patch_xmin[p], g.edges_cx[e]); patch_xmax[p] = std::max(patch_xmax[p], g.edges_cx[e]); patch_ysum[p] += g.edges_cy[e]; }
        ++patch_count[p];
    }
}

afx_mesh_desc Partition::desc() const
{
    afx_mesh_desc d{};
    d.n_cells = n_real(); d.n_ghost = n_bc; d.n_edges = (uint32_t)edge_l2g.size();
    d.edges_cells = edge_cells.data();
    d.edges_nx = enx.data(); d.edges_ny = eny.data(); d.edges_len = elen.data(); d.edges_cx = ecx.data(); d.edges_cy = ecy.data();
    d.cells_cx = ccx.data(); d.cells_cy = ccy.data(); d.cells_area = area.data();
    d.cells_edges = cell_edges.data(); d.cells_is_tri = is_tri.data();
    d.boundary_edges = bnd_edge.data(); d.boundary_patch = bnd_patch.data();
    return d;
}

}  // namespace afx

extern "C" {

int afx_partition_create(afx_partition** out, const afx_mesh_desc* global, int nranks, int rank)
{
    if (!out || !global) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = nullptr;
    afx_partition* p = new afx_partition;
    try {
        p->p.build(*global, nranks, rank);
        *out = p;
        return AFX_OK;
    } catch (const std::exception& e) { afx::set_error(e.what()); delete p; return AFX_ERR_INVALID; }
}

void afx_partition_free(afx_partition* p) { delete p; }

int afx_partition_get_desc(const afx_partition* p, afx_mesh_desc* out)
{
    if (!p || !out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    *out = p->p.desc();
    return AFX_OK;
}

/* out = {n_own, n_ring1, n_ring2, n_boundary_ghosts, n_local_edges, n_peers, rank, nranks} */
int afx_partition_info(const afx_partition* p, uint32_t out[8])
{
    if (!p || !out) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    const auto& q = p->p;
    out[0] = q.n_own; out[1] = q.n_r1; out[2] = q.n_r2; out[3] = q.n_bc; out[4] = (uint32_t)q.edge_l2g.size();
    out[5] = (uint32_t)q.peers.size(); out[6] = (uint32_t)q.rank; out[7] = (uint32_t)q.nranks;
    return AFX_OK;
}

const uint32_t* afx_partition_cell_l2g(const afx_partition* p) { return p ? p->p.cell_l2g.data() : nullptr; }
const uint32_t* afx_partition_edge_l2g(const afx_partition* p) { return p ? p->p.edge_l2g.data() : nullptr; }

int afx_partition_peer(const afx_partition* p, int i, int* peer_rank, const uint32_t** send, uint32_t* n_send,
                       const uint32_t** recv, uint32_t* n_recv)
{
    if (!p) { afx::set_error("null argument"); return AFX_ERR_INVALID; }
    if (i < 0 || i >= (int)p->p.peers.size()) { afx::set_error("peer index out of range"); return AFX_ERR_INVALID; }
    const auto& q = p->p.peers[(size_t)i];
    if (peer_rank) *peer_rank = q.rank;
    if (send) *send = q.send.data();
    if (n_send) *n_send = (uint32_t)q.send.size();
    if (recv) *recv = q.recv.data();
    if (n_recv) *n_recv = (uint32_t)q.recv.size();
    return AFX_OK;
}

}  // extern "C"
