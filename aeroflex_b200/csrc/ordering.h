// ordering.h -- space-filling-curve ordering of cells (host).
// Used for the coalescing-friendly internal numbering of one GPU's cells and
// for cutting the mesh into contiguous, compact partitions (one per GPU).
#pragma once
#include <cstdint>
#include <vector>

namespace afx {

// Hilbert-curve order of the points (x[i], y[i]), i in idx, computed on RANK
// coordinates (the curve sees the x-rank and y-rank of every point, so strongly
// graded meshes -- 1e-5 chord cells at the wall, 10 chord cells at the far
// field -- are spread evenly over the curve).  Stable for ties.  Returns the
// permutation: result[k] = k-th point along the curve.
std::vector<uint32_t> hilbert_order(const double* x, const double* y, const std::vector<uint32_t>& idx);

// Recursive graph bisection of the cell graph restricted to `cells`: the set is cut in two along a breadth-first level
// front started from a pseudo-peripheral cell, and so on until every piece has at most `tile_cells` cells.  Returns the
// cells piece after piece (each piece in breadth-first order) and appends the piece sizes to `tile_sizes`: consecutive
// runs of the new numbering are compact patches of the GRAPH (short perimeter in faces), whatever the cells' shapes --
// which is what the shared-memory tiles of the fused stage kernel need (rank-space curves give ragged tiles on
// strongly stretched meshes).  nb[4*c+k] = k-th neighbour cell of c or 0xFFFFFFFF; n_total = size of the label space.
std::vector<uint32_t> graph_tile_order(uint32_t n_total, const uint32_t* nb, std::vector<uint32_t> cells, uint32_t tile_cells,
                                       std::vector<uint32_t>& tile_sizes);

// The same bisection cut into exactly k pieces of (almost) equal size: a METIS-style graph partition of the cell graph
// (recursive graph-growing bisection, no multilevel refinement).  piece_sizes[p] consecutive cells of the result form piece p.
std::vector<uint32_t> graph_partition_order(uint32_t n_total, const uint32_t* nb, std::vector<uint32_t> cells, uint32_t k,
                                            std::vector<uint32_t>& piece_sizes);

}  // namespace afx
