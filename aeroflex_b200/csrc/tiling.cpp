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

TilePlan build_tiles(uint32_t N, uint32_t n_upd, uint32_t n_front, uint32_t n_grad, const uint32_t* cf, const uint32_t* cnb,
                     uint32_t tile_cells)
{
    if (tile_cells < 32 || tile_cells > 4096) throw std::invalid_argument("tile size out of range");
    TilePlan P;
    P.tile_cells = tile_cells;
    // tile boundaries: [0, n_front) and [n_front, n_upd) are tiled separately
    std::vector<uint32_t> start, count;
    auto cut = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t c = lo; c < hi; c += tile_cells) { start.push_back(c); count.push_back(std::min(tile_cells, hi - c)); }
    };
    if (n_front > 0 && n_front < n_upd) { cut(0, n_front); P.n_front_tiles = (uint32_t)start.size(); cut(n_front, n_upd); }
    else cut(0, n_upd);
    const size_t nt = start.size();
    std::vector<OneTile> tiles(nt);
    uint32_t cap = 1024;
    while (cap < 16 * tile_cells) cap <<= 1;

    std::string err;  // exceptions must not leave the parallel region
#pragma omp parallel
    {
        SmallMap cmap(cap), fmap(cap);
        std::vector<uint32_t> loc;  // local -> global cell
#pragma omp for schedule(dynamic, 16)
        for (int64_t t = 0; t < (int64_t)nt; ++t) {
            OneTile& T = tiles[t];
            const uint32_t c0 = start[t], nc = count[t];
            auto fail = [&](const char* what) {
#pragma omp critical(afx_tiling_err)
                if (err.empty()) err = what;
            };
            cmap.clear(); fmap.clear(); loc.clear();
            for (uint32_t l = 0; l < nc; ++l) { cmap.put(c0 + l, l); loc.push_back(c0 + l); }
            // faces of the own cells, ring-1 cells; ghosts next to own cells are collected after ring 1
            std::vector<uint32_t> late;  // state-only cells seen so far (ghosts of own cells)
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
                        if (j >= n_grad) { fail("tiling: a neighbour of an advanced cell has no gradient"); continue; }
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
                    if (cmap.crowded()) { fail("tiling: tile neighbourhood too large (cells are not ordered compactly)"); break; }
                    cmap.put(j, (uint32_t)loc.size()); loc.push_back(j);
                }
            }
            const uint32_t nloc = (uint32_t)loc.size();
            if (nloc >= TL_NONE || T.ftab.size() >= TL_SIDE) { fail("tiling: tile too large for 16-bit local indices"); continue; }
            T.hd.cell0 = c0; T.hd.nc = nc; T.hd.h1 = h1; T.hd.h2 = nloc - nc - h1; T.hd.nf = (uint32_t)T.ftab.size();
            T.halo.assign(loc.begin() + nc, loc.end());
            T.ctab.resize(nc + h1);
            for (uint32_t l = 0; l < nc + h1; ++l) {
                const uint32_t c = loc[l];
                TileCell tc;
                for (int s = 0; s < 4; ++s) {
                    tc.nb[s] = TL_NONE; tc.fs[s] = TL_NONE;
                    const uint32_t v = cf[(size_t)s * N + c];
                    if (v == CF_NONE_) continue;
                    const uint32_t j = cnb[(size_t)s * N + c];
                    const uint32_t* lj = cmap.find(j);
                    if (!lj) { fail("tiling: neighbour missing from the tile"); continue; }
                    tc.nb[s] = (uint16_t)*lj;
                    if (const uint32_t* lf = fmap.find(v & CF_ID_)) tc.fs[s] = (uint16_t)(*lf | ((v & CF_SIDE_) ? TL_SIDE : 0));
                }
                T.ctab[l] = tc;
            }
        }
    }
    if (!err.empty()) throw std::invalid_argument(err);
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
    for (int s = 0; s < 4; ++s) { empty.nb[s] = TL_NONE; empty.fs[s] = TL_NONE; }
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

}  // namespace afx
