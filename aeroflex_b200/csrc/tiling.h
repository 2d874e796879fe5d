// tiling.h -- shared-memory tiles of the renumbered mesh (host side of the fused stage kernel).
//
// The advanced cells [0, n_upd) are cut, in their (Hilbert) order, into tiles of
// at most `tile_cells` consecutive cells.  One CTA of k_stage owns one tile and
// stages in shared memory everything one Runge-Kutta stage of that tile needs:
//   local cells   [0, nc)              the tile's own cells
//                 [nc, nc+h1)          ring 1: REAL cells sharing a face with an own cell
//                                      (their limiter is recomputed by this tile)
//                 [nc+h1, nc+h1+h2)    everything else a ring-1 limiter or an own face
//                                      reads: ring-2 cells and ghost cells (state only)
//   local faces   [0, nf)              every face with an end in the tile, in order of
//                                      first appearance (own cells ascending, slots ascending)
// Per-cell tables carry LOCAL 16-bit indices so the kernel never touches the
// global connectivity; per-cell sums keep the slot order of cf[], i.e. the
// reference's ascending edge order, so results do not depend on the tiling.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace afx {

constexpr uint16_t TL_NONE = 0xFFFFu;
constexpr uint16_t TL_SIDE = 0x8000u;  // slot entry: this cell is cell1 of the local face

struct TileHead {            // 32 bytes, read by every thread of the CTA
    uint32_t cell0, nc;      // own cells [cell0, cell0+nc)
    uint32_t h1, h2;         // ring-1 cells, state-only cells
    uint32_t nf;             // local faces
    uint32_t off_halo;       // first entry of this tile in halo[] (multiple of 4; the tile's list is padded to a multiple of 4)
    uint32_t off_cell;       // first record of the per-cell tables (even; nc+h1 records padded to an even count)
    uint32_t off_face;       // first record of the per-face tables (nf records)
};

struct TileCell {            // 16 bytes per own / ring-1 cell
    uint16_t nb[4];          // local index of the cell across slot s (the cell's own index where the slot is empty)
    uint16_t fs[4];          // local face of slot s | TL_SIDE, TL_NONE if the face has no end in the tile
};

struct TilePlan {
    uint32_t tile_cells = 0;
    uint32_t n_front_tiles = 0;        // tiles covering the send layer [0, n_front) come first
    uint32_t max_loc = 0, max_n1 = 0, max_nf = 0, max_nc = 0, max_halo = 0;  // shared memory is sized by these (padded counts)
    std::vector<TileHead> head;
    std::vector<uint32_t> halo;        // global ids of the ring-1 and state-only cells, tile after tile (padded with the last id)
    std::vector<TileCell> ctab;        // padded records are empty cells
    std::vector<uint32_t> face;        // global (renumbered) id of every local face
    uint64_t local_cells = 0;          // sum of nc+h1+h2 over the tiles
};

struct TileLimits {  // a tile whose staging would exceed one of these is cut in two (0 = no limit)
    uint32_t max_loc = 0, max_n1 = 0, max_nf = 0, max_halo = 0;
};

// cf / cnb: [4][N] slot-major tables of the solver (valid for cells < n_grad); N real cells.
// `sizes`: consecutive runs of cells starting at cell 0 (ordering.h: graph_tile_order), the first `n_front_tiles` of
// them cover the send layer of a partitioned run.  Runs longer than tile_cells are cut into pieces of tile_cells.
TilePlan build_tiles(uint32_t N, uint32_t n_grad, const uint32_t* cf, const uint32_t* cnb, const std::vector<uint32_t>& sizes,
                     uint32_t n_front_tiles, uint32_t tile_cells, const TileLimits& lim = TileLimits());


// Verifies a plan against the connectivity it was built from: every advanced cell is owned by exactly one tile, the
// local neighbour / face indices of every own and ring-1 cell point at the right global cell / face with the right
// side bit, every face of an own cell is a local face and ring-1 cells list exactly the faces they share with the
// tile.  Returns an empty string or the first inconsistency.
std::string check_tiles(const TilePlan& P, uint32_t N, uint32_t n_upd, const uint32_t* cf, const uint32_t* cnb);

}  // namespace afx
