#include "ordering.h"

#include <algorithm>
#include <numeric>
#if defined(_OPENMP)
#include <parallel/algorithm>
#define AFX_SSORT __gnu_parallel::stable_sort
#else
#define AFX_SSORT std::stable_sort
#endif

namespace afx {

static uint64_t hilbert_d(uint32_t x, uint32_t y, int order)
{
    // classic xy -> d conversion on a 2^order x 2^order grid
    const uint32_t n1 = (order >= 32 ? 0xFFFFFFFFu : ((1u << order) - 1u));
    uint64_t d = 0;
    for (uint32_t s = 1u << (order - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1 : 0, ry = (y & s) ? 1 : 0;
        d += (uint64_t)s * s * ((3 * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) { x = n1 - x; y = n1 - y; }
            const uint32_t t = x; x = y; y = t;
        }
    }
    return d;
}

std::vector<uint32_t> hilbert_order(const double* x, const double* y, const std::vector<uint32_t>& idx)
{
    const uint32_t n = (uint32_t)idx.size();
    std::vector<uint32_t> out(idx);
    if (n < 2) return out;
    std::vector<uint32_t> pos(n), rx(n), ry(n);
    std::iota(pos.begin(), pos.end(), 0u);
    AFX_SSORT(pos.begin(), pos.end(), [&](uint32_t a, uint32_t b) { return x[idx[a]] < x[idx[b]]; });
    for (uint32_t r = 0; r < n; ++r) rx[pos[r]] = r;
    std::iota(pos.begin(), pos.end(), 0u);
    AFX_SSORT(pos.begin(), pos.end(), [&](uint32_t a, uint32_t b) { return y[idx[a]] < y[idx[b]]; });
    for (uint32_t r = 0; r < n; ++r) ry[pos[r]] = r;
    int order = 1;
    while ((1u << order) < n && order < 31) ++order;
    std::vector<uint64_t> key(n);
#pragma omp parallel for
    for (int64_t i = 0; i < (int64_t)n; ++i) key[i] = hilbert_d(rx[i], ry[i], order);
    std::iota(pos.begin(), pos.end(), 0u);
    AFX_SSORT(pos.begin(), pos.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (uint32_t k = 0; k < n; ++k) out[k] = idx[pos[k]];
    return out;
}

namespace {

struct Bisector {
    const uint32_t* nb;
    std::vector<uint32_t>& cells;
    std::vector<uint32_t> label;              // subset / visit marks, one fresh id per use
    uint32_t next_id = 1;
    uint32_t T;
    std::vector<std::pair<uint32_t, uint32_t>> leaves;  // (start, size)

    uint32_t fresh()
    {
        uint32_t v;
#pragma omp atomic capture
        v = next_id++;
        return v;
    }
    // breadth-first order of cells[lo,hi) (all labelled `in`) from `start`, relabelled `out`; components not reached
    // from `start` follow, each from its first cell in array order.  Returns the last cell of the FIRST component.
    uint32_t bfs(uint32_t lo, uint32_t hi, uint32_t start, uint32_t in, uint32_t out, std::vector<uint32_t>& ord)
    {
        ord.clear();
        uint32_t last_first = start, scan = lo;
        bool first = true;
        for (uint32_t seed = start;;) {
            size_t head = ord.size();
            label[seed] = out; ord.push_back(seed);
            while (head < ord.size()) {
                const uint32_t c = ord[head++];
                for (int k = 0; k < 4; ++k) {
                    const uint32_t j = nb[4 * (size_t)c + k];
                    if (j != 0xFFFFFFFFu && label[j] == in) { label[j] = out; ord.push_back(j); }
                }
            }
            if (first) { last_first = ord.back(); first = false; }
            if (ord.size() == hi - lo) break;
            while (label[cells[scan]] != in) ++scan;
            seed = cells[scan];
        }
        return last_first;
    }
    void split(uint32_t lo, uint32_t hi, uint32_t k)
    {
        const uint32_t n = hi - lo;
        std::vector<uint32_t> ord;
        ord.reserve(n);
        const uint32_t a = fresh(), b = fresh(), c = fresh();
        for (uint32_t i = lo; i < hi; ++i) label[cells[i]] = a;
        uint32_t far = bfs(lo, hi, cells[lo], a, b, ord);
        if (k > 1) {  // a second sweep from the far end: a better pseudo-peripheral start, levels along the long axis
            const uint32_t d = fresh();
            far = bfs(lo, hi, far, b, d, ord);
            bfs(lo, hi, far, d, c, ord);
        }
        std::copy(ord.begin(), ord.end(), cells.begin() + lo);
        if (k == 1) {
#pragma omp critical(afx_bisect_leaves)
            leaves.emplace_back(lo, n);
            return;
        }
        const uint32_t kl = k / 2;
        const uint32_t nl = (uint32_t)((uint64_t)n * kl / k);
        std::vector<uint32_t>().swap(ord);
#pragma omp task default(shared) if (n > 65536)
        split(lo, lo + nl, kl);
#pragma omp task default(shared) if (n > 65536)
        split(lo + nl, hi, k - kl);
#pragma omp taskwait
    }
};

}  // namespace

std::vector<uint32_t> graph_tile_order(uint32_t n_total, const uint32_t* nb, std::vector<uint32_t> cells, uint32_t tile_cells,
                                       std::vector<uint32_t>& tile_sizes)
{
    if (cells.empty()) return cells;
    Bisector B{nb, cells, std::vector<uint32_t>(n_total, 0u), 1u, tile_cells, {}};
    const uint32_t n = (uint32_t)cells.size();
    const uint32_t k = (n + tile_cells - 1) / tile_cells;
#pragma omp parallel
#pragma omp single
    B.split(0, n, k);
    std::sort(B.leaves.begin(), B.leaves.end());
    for (const auto& l : B.leaves) tile_sizes.push_back(l.second);
    return cells;
}

// k pieces of (almost) equal size by the same recursive level-structure bisection: piece p = cells [start_p, start_p + size_p)
// of the returned order.  This is the graph-growing bisection METIS uses for its initial partitions, applied recursively to
// the cell graph itself (no coarsening / refinement passes): pieces are connected fronts of the face-neighbour graph, whatever
// the cell shapes and sizes.
std::vector<uint32_t> graph_partition_order(uint32_t n_total, const uint32_t* nb, std::vector<uint32_t> cells, uint32_t k,
                                            std::vector<uint32_t>& piece_sizes)
{
    piece_sizes.clear();
    if (cells.empty() || k == 0) return cells;
    Bisector B{nb, cells, std::vector<uint32_t>(n_total, 0u), 1u, 0u, {}};
    const uint32_t n = (uint32_t)cells.size();
#pragma omp parallel
#pragma omp single
    B.split(0, n, std::min(k, n));
    std::sort(B.leaves.begin(), B.leaves.end());
    for (const auto& l : B.leaves) piece_sizes.push_back(l.second);
    return cells;
}

}  // namespace afx
