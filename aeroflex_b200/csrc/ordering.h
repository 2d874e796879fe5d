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

}  // namespace afx
