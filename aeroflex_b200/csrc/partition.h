// partition.h -- cutting one rans::mesh into per-GPU pieces with a 2-layer halo (host).
//
// Cells are ordered along a Hilbert curve and cut into `nranks` contiguous
// chunks (compact pieces, short interfaces; the same curve gives each GPU its
// coalescing-friendly numbering) -- or, with AFX_PARTITION=graph, by recursive
// graph bisection of the face-neighbour graph (METIS-style graph partition,
// ordering.h).  Rank r keeps, in this local order,
//   [ owned | ring 1 | ring 2 | boundary ghosts ]
// ring k = real cells of other ranks at graph distance k from an owned cell.
// With both rings' STATES received after every stage update, a rank can
// recompute gradients and limiters of ring-1 cells itself, so one exchange of
// 32 bytes per halo cell per Runge-Kutta stage is the only communication
// (SURVEY.md 8e).  Local edges keep the ascending global edge order, so
// per-cell sums run in the reference's order and a partitioned run is
// bit-identical to the single-GPU run in strict mode.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/afx_rans.h"

namespace afx {

struct Partition {
    int rank = 0, nranks = 1;
    uint32_t n_own = 0, n_r1 = 0, n_r2 = 0, n_bc = 0;
    uint32_t n_global_cells = 0, n_global_ghost = 0, n_global_edges = 0;
    std::vector<uint32_t> cell_l2g;  // local cell (incl. ghosts) -> global cell id in the reference numbering
    std::vector<uint32_t> edge_l2g;  // local edge -> global edge id, ascending
    std::vector<uint32_t> bnd_l2g;   // local boundary index -> global boundary index, ascending
    // local mesh in the reference's layout
    std::vector<uint32_t> edge_cells, cell_edges, bnd_edge;
    std::vector<double> enx, eny, elen, ecx, ecy, ccx, ccy, area;
    std::vector<uint8_t> is_tri;
    std::vector<int32_t> bnd_patch;
    // halo exchange plan: per peer, owned cells to send and halo cells to fill, both in global curve order
    struct Peer {
        int rank;
        std::vector<uint32_t> send;  // local ids of owned cells
        std::vector<uint32_t> recv;  // local ids of ring cells
    };
    std::vector<Peer> peers;
    // global extents of every patch, for force coefficients (post.h:314-338)
    std::vector<double> patch_xmin, patch_xmax, patch_ysum;
    std::vector<uint32_t> patch_count;
    // patch ids in the order of their first edge in the GLOBAL boundary list: solver::get_boundary_variables
    // (solver.h:597-611) takes the far-field state of the first far-field edge of the whole mesh, which a rank may not hold
    std::vector<int32_t> patch_order;

    void build(const afx_mesh_desc& g, int nranks, int rank);
    afx_mesh_desc desc() const;
    uint32_t n_real() const { return n_own + n_r1 + n_r2; }
};

}  // namespace afx

struct afx_partition {
    afx::Partition p;
};
