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

}  // namespace afx
