// tiling.cpp -- see tiling.h
#include "tiling.h"

#include <omp.h>

#include <algorithm>
#include <stdexcept>
#include <string>

namespace afx {
namespace {

constexpr uint32_t CF_NONE_ = 0xFFFFFFFFu, CF_SIDE_ = 0x80000000u, CF_ID_ = 0x3FFFFFFFu;

// small open-addressing map uint32 -> uint32, cleared in O(1) by bumping a stamp
struct SmallMap {
    std::vector<uint32_t> key, val, stamp;
    uint32_t mask = 0, cur = 0, used = 0;
    explicit SmallMap(uint32_t cap_pow2) : key(cap_pow2), val(cap_pow2), stamp(cap_pow2, 0), mask(cap_pow2 - 1) {}
    void clear() { ++cur; used = 0; }
    bool crowded() const { return used > (mask >> 1); }
    static uint32_t h(uint32_t k) { k *= 0x9E3779B1u; return k ^ (k >> 15); }
    // returns the slot of k; `found` tells whether it was present
    uint32_t* find(uint32_t k)
    {
        for (uint32_t i = h(k) & mask;; i = (i + 1) & mask) {
            if (stamp[i] != cur) return nullptr;
            if (key[i] == k) return &val[i];
        }
    }
    void put(uint32_t k, uint32_t v)
    {
        for (uint32_t i = h(k) & mask;; i = (i + 1) & mask) {
            if (stamp[i] != cur) { stamp[i] = cur; key[i] = k; val[i] = v; ++used; return; }
            if (key[i] == k) { val[i] = v; return; }
        }
    }
};

struct OneTile {
    TileHead hd{};
    std::vector<uint32_t> halo;
    std::vector<TileCell> ctab;
    std::vector<uint32_t> ftab;
};

}  // namespace

namespace {

// one tile: collect ring 1, the state-only cells and the local faces of cells [c0, c0+nc); false if a limit is exceeded
bool make_tile(uint32_t N, uint32_t n_grad, const uint32_t* cf, const uint32_t* cnb, uint32_t c0, uint32_t nc, const TileLimits& lim,
               SmallMap& cmap, SmallMap& fmap, std::vector<uint32_t>& loc, OneTile& T, const char*& err)
{
    cmap.clear(); fmap.clear(); loc.clear();
    T.halo.clear(); T.ctab.clear(); T.ftab.clear();
    for (uint32_t l = 0; l < nc; ++l) { cmap.put(c0 + l, l); loc.push_back(c0 + l); }
    // faces of the own cells, ring-1 cells; ghosts next to own cells are numbered after ring 1
    std::vector<uint32_t> late;
    for (uint32_t l = 0; l < nc; ++l) {
        const uint32_t c = c0 + l;
        for (int s = 0; s < 4; ++s) {
            const uint32_t v = cf[(size_t)s * N + c];
            if (v == CF_NONE_) continue;
            const uint32_t f = v & CF_ID_;
            if (!fmap.find(f)) { fmap.put(f, (uint32_t)T.ftab.size()); T.ftab.push_back(f); }
            const uint32_t j = cnb[(size_t)s * N + c];
            if (cmap.find(j)) continue;
            if (j < N) {
                if (j >= n_grad) { err = "tiling: a neighbour of an advanced cell has no gradient"; return true; }
                cmap.put(j, (uint32_t)loc.size()); loc.push_back(j);
            } else {
                cmap.put(j, 0xFFFFFFFEu);  // placeholder, numbered after ring 1
                late.push_back(j);
            }
        }
    }
    const uint32_t h1 = (uint32_t)loc.size() - nc;
    for (uint32_t j : late) { cmap.put(j, (uint32_t)loc.size()); loc.push_back(j); }
    // what the ring-1 limiters read
    for (uint32_t l = nc; l < nc + h1; ++l) {
        const uint32_t c = loc[l];
        for (int s = 0; s < 4; ++s) {
            const uint32_t j = cnb[(size_t)s * N + c];
            if (j == CF_NONE_ || cmap.find(j)) continue;
            if (cmap.crowded()) { if (nc > 1) return false; err = "tiling: tile neighbourhood too large"; return true; }
            cmap.put(j, (uint32_t)loc.size()); loc.push_back(j);
        }
    }
    const uint32_t nloc = (uint32_t)loc.size(), nf = (uint32_t)T.ftab.size(), nh = nloc - nc;
    const bool over = (lim.max_loc && nloc > lim.max_loc) || (lim.max_n1 && ((nc + h1 + 1u) & ~1u) > lim.max_n1) || (lim.max_nf && nf > lim.max_nf) ||
                      (lim.max_halo && ((nh + 3u) & ~3u) > lim.max_halo) || nloc >= TL_NONE || nf >= TL_SIDE;
    if (over) {
        if (nc > 1) return false;
        err = "tiling: a single cell exceeds the tile limits";
        return true;
    }
    T.hd.cell0 = c0; T.hd.nc = nc; T.hd.h1 = h1; T.hd.h2 = nloc - nc - h1; T.hd.nf = nf;
    T.halo.assign(loc.begin() + nc, loc.end());
    T.ctab.resize(nc + h1);
    for (uint32_t l = 0; l < nc + h1; ++l) {
        const uint32_t c = loc[l];
        TileCell tc;
        for (int s = 0; s < 4; ++s) {
            tc.nb[s] = (uint16_t)l; tc.fs[s] = TL_NONE;  // an empty slot points at the cell itself: min/max over it changes nothing
            const uint32_t v = cf[(size_t)s * N + c];
            if (v == CF_NONE_) continue;
            const uint32_t j = cnb[(size_t)s * N + c];
            const uint32_t* lj = cmap.find(j);
            if (!lj) { err = "tiling: neighbour missing from the tile"; return true; }
            tc.nb[s] = (uint16_t)*lj;
            if (const uint32_t* lf = fmap.find(v & CF_ID_)) tc.fs[s] = (uint16_t)(*lf | ((v & CF_SIDE_) ? TL_SIDE : 0));
        }
        T.ctab[l] = tc;
    }
    return true;
}

}  // namespace

TilePlan build_tiles(uint32_t N, uint32_t n_grad, const uint32_t* cf, const uint32_t* cnb, const std::vector<uint32_t>& sizes,
                     uint32_t n_front_tiles, uint32_t tile_cells, const TileLimits& lim)
{
    if (tile_cells < 1 || tile_cells > 4096) throw std::invalid_argument("tile size out of range");
    TilePlan P;
    P.tile_cells = tile_cells;
    // the given runs, cut to tile_cells; each becomes one tile or, where a limit is exceeded, a few (split in halves)
    struct Run { uint32_t start, count; bool front; };
    std::vector<Run> runs;
    {
        uint32_t c = 0;
        for (size_t r = 0; r < sizes.size(); ++r) {
            for (uint32_t o = 0; o < sizes[r]; o += tile_cells) runs.push_back(Run{c + o, std::min(tile_cells, sizes[r] - o), r < n_front_tiles});
            c += sizes[r];
        }
    }
    const size_t nr = runs.size();
    std::vector<std::vector<OneTile>> made(nr);
    uint32_t cap = 1024;
    while (cap < 16 * tile_cells) cap <<= 1;

    std::string err;  // exceptions must not leave the parallel region
#pragma omp parallel
    {
        SmallMap cmap(cap), fmap(cap);
        std::vector<uint32_t> loc;  // local -> global cell
#pragma omp for schedule(dynamic, 16)
        for (int64_t r = 0; r < (int64_t)nr; ++r) {
            std::vector<std::pair<uint32_t, uint32_t>> todo{{runs[r].start, runs[r].count}};  // LIFO keeps the pieces in cell order
            while (!todo.empty()) {
                const auto piece = todo.back();
                todo.pop_back();
                OneTile T;
                const char* e = nullptr;
                if (make_tile(N, n_grad, cf, cnb, piece.first, piece.second, lim, cmap, fmap, loc, T, e)) {
                    if (e) {
#pragma omp critical(afx_tiling_err)
                        if (err.empty()) err = e;
                        break;
                    }
                    made[r].push_back(std::move(T));
                } else {
                    const uint32_t half = piece.second / 2;
                    todo.emplace_back(piece.first + half, piece.second - half);
                    todo.emplace_back(piece.first, half);
                }
            }
        }
    }
    if (!err.empty()) throw std::invalid_argument(err);
    std::vector<OneTile> tiles;
    for (size_t r = 0; r < nr; ++r) {
        for (auto& t : made[r]) tiles.push_back(std::move(t));
        if (runs[r].front) P.n_front_tiles = (uint32_t)tiles.size();
    }
    const size_t nt = tiles.size();
    P.head.resize(nt);
    size_t oh = 0, oc = 0, of = 0;
    for (size_t t = 0; t < nt; ++t) {
        TileHead& h = tiles[t].hd;
        const uint32_t n1p = (h.nc + h.h1 + 1u) & ~1u, nhp = (h.h1 + h.h2 + 3u) & ~3u;
        h.off_halo = (uint32_t)oh; h.off_cell = (uint32_t)oc; h.off_face = (uint32_t)of;
        oh += nhp; oc += n1p; of += h.nf;
        if (oh > 0xFFFFFFF0ull || 4 * oc > 0xFFFFFFF0ull || of > 0xFFFFFFF0ull) throw std::invalid_argument("tiling: tables exceed 32-bit offsets");
        P.head[t] = h;
        P.max_loc = std::max(P.max_loc, h.nc + h.h1 + h.h2);
        P.max_n1 = std::max(P.max_n1, n1p);
        P.max_nf = std::max(P.max_nf, h.nf);
        P.max_nc = std::max(P.max_nc, (h.nc + 1u) & ~1u);
        P.max_halo = std::max(P.max_halo, nhp);
        P.local_cells += h.nc + h.h1 + h.h2;
    }
    TileCell empty;
    for (int s = 0; s < 4; ++s) { empty.nb[s] = 0; empty.fs[s] = TL_NONE; }
    P.halo.assign(oh, 0); P.ctab.assign(oc, empty); P.face.resize(of);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)nt; ++t) {
        const TileHead& h = P.head[t];
        std::copy(tiles[t].halo.begin(), tiles[t].halo.end(), P.halo.begin() + h.off_halo);
        const uint32_t nh = h.h1 + h.h2, nhp = (nh + 3u) & ~3u;
        for (uint32_t k = nh; k < nhp; ++k) P.halo[h.off_halo + k] = nh ? tiles[t].halo[nh - 1] : h.cell0;  // padding reads a valid cell
        std::copy(tiles[t].ctab.begin(), tiles[t].ctab.end(), P.ctab.begin() + h.off_cell);
        std::copy(tiles[t].ftab.begin(), tiles[t].ftab.end(), P.face.begin() + h.off_face);
    }
    return P;
}


std::string check_tiles(const TilePlan& P, uint32_t N, uint32_t n_upd, const uint32_t* cf, const uint32_t* cnb)
{
    std::vector<uint8_t> owned(n_upd, 0);
    std::string err;
    auto fail = [&](size_t t, const char* what) { if (err.empty()) err = "tile " + std::to_string(t) + ": " + what; };
    for (size_t t = 0; t < P.head.size() && err.empty(); ++t) {
        const TileHead& h = P.head[t];
        const uint32_t n1 = h.nc + h.h1, nloc = n1 + h.h2;
        if ((h.off_cell & 1u) || (h.off_halo & 3u)) fail(t, "table offsets are not aligned for bulk copies");
        auto gid = [&](uint32_t l) { return l < h.nc ? h.cell0 + l : P.halo[h.off_halo + l - h.nc]; };
        for (uint32_t l = 0; l < h.nc; ++l) {
            if (h.cell0 + l >= n_upd || owned[h.cell0 + l]++) { fail(t, "own cell outside the advanced range or owned twice"); break; }
        }
        std::vector<uint8_t> face_seen(h.nf, 0);
        for (uint32_t l = 0; l < n1 && err.empty(); ++l) {
            const uint32_t c = gid(l);
            if (c >= N) { fail(t, "a ghost cell is listed as own / ring 1"); break; }
            const TileCell& tc = P.ctab[h.off_cell + l];
            bool touches_tile = false;
            for (int s = 0; s < 4; ++s) {
                const uint32_t v = cf[(size_t)s * N + c], j = cnb[(size_t)s * N + c];
                if (v == CF_NONE_) { if (tc.nb[s] != l || tc.fs[s] != TL_NONE) fail(t, "empty slot does not point at its own cell"); continue; }
                if (tc.nb[s] == TL_NONE || tc.nb[s] >= nloc || gid(tc.nb[s]) != j) { fail(t, "local neighbour index is wrong"); break; }
                const bool nb_own = tc.nb[s] < h.nc;
                const bool want_face = l < h.nc || nb_own;  // faces with an end in the tile
                if (want_face != (tc.fs[s] != TL_NONE)) { fail(t, "face list does not match the tile's faces"); break; }
                if (!want_face) continue;
                touches_tile = true;
                const uint32_t lf = tc.fs[s] & 0x7FFFu;
                if (lf >= h.nf || P.face[h.off_face + lf] != (v & CF_ID_)) { fail(t, "local face index is wrong"); break; }
                if (((tc.fs[s] & TL_SIDE) != 0) != ((v & CF_SIDE_) != 0)) { fail(t, "side bit is wrong"); break; }
                if (l < h.nc && tc.nb[s] >= h.nc && tc.nb[s] < n1 && j >= N) { fail(t, "a ghost is numbered as ring 1"); break; }
                if (l < h.nc && tc.nb[s] >= n1 && j < N) { fail(t, "a real neighbour of an own cell is not ring 1"); break; }
                face_seen[lf] = 1;
            }
            if (l >= h.nc && !touches_tile) fail(t, "ring-1 cell shares no face with the tile");
        }
        for (uint32_t lf = 0; lf < h.nf && err.empty(); ++lf) if (!face_seen[lf]) fail(t, "local face without an own cell");
    }
    if (err.empty())
        for (uint32_t c = 0; c < n_upd; ++c) if (!owned[c]) { err = "cell " + std::to_string(c) + " is in no tile"; break; }
    return err;
}

}  // namespace afx
