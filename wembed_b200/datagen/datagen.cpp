// Synthetic-workload helpers for bench.py and the tests (host side, not on the product path): the O(n) parts of
// wembed_b200/datasets.py that numpy does slowly at n = 1e7.  Results are identical to the numpy code paths
// (same comparisons in double, same edge order), which tests/test_cabi_host.py checks.
//
//   wbd_pairs_within   all index pairs (i < j) with ||p_i - p_j||^2 < r^2, lexicographic order
//                      (the rule of the reference's GeometricGraphSampler.cpp:10-51, found with a cell grid instead of its O(n^2) loop)
//   wbd_csr_canonical  CSR of an edge list that is already unique, sorted and has src < dst (Graph.cpp:87-150 invariants)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
std::vector<int32_t> g_edges;   // result of the last wbd_pairs_within call, copied out by wbd_take_edges
}

extern "C" {

int64_t wbd_pairs_within(int64_t n, const double* pts /* [n][2] */, double radius) {
    g_edges.clear();
    if (n <= 0) return 0;
    std::vector<int64_t> cx(n), cy(n);
    int64_t minx = INT64_MAX, miny = INT64_MAX, maxx = INT64_MIN, maxy = INT64_MIN;
    for (int64_t i = 0; i < n; ++i) {
        cx[i] = (int64_t)std::floor(pts[2 * i] / radius);
        cy[i] = (int64_t)std::floor(pts[2 * i + 1] / radius);
        minx = std::min(minx, cx[i]); maxx = std::max(maxx, cx[i]);
        miny = std::min(miny, cy[i]); maxy = std::max(maxy, cy[i]);
    }
    const int64_t nx = maxx - minx + 3, ny = maxy - miny + 3;   // one empty ring of cells around the data
    std::vector<int64_t> start(nx * ny + 1, 0);
    std::vector<int64_t> cell(n);
    for (int64_t i = 0; i < n; ++i) { cell[i] = (cx[i] - minx + 1) * ny + (cy[i] - miny + 1); ++start[cell[i] + 1]; }
    for (int64_t c = 0; c < nx * ny; ++c) start[c + 1] += start[c];
    std::vector<int32_t> member(n);
    {
        std::vector<int64_t> at(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; ++i) member[at[cell[i]]++] = (int32_t)i;     // ascending index inside every cell
    }
    const double r2 = radius * radius;
    const int threads = 16;
    std::vector<std::vector<int32_t>> part(threads);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int t = 0; t < threads; ++t) {
        std::vector<int32_t>& out = part[t];
        std::vector<int32_t> nb;
        const int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
        for (int64_t i = lo; i < hi; ++i) {
            nb.clear();
            const double xi = pts[2 * i], yi = pts[2 * i + 1];
            for (int64_t dx = -1; dx <= 1; ++dx)
                for (int64_t dy = -1; dy <= 1; ++dy) {
                    const int64_t c = cell[i] + dx * ny + dy;
                    for (int64_t k = start[c]; k < start[c + 1]; ++k) {
                        const int32_t j = member[k];
                        if (j <= i) continue;
                        const double ex = xi - pts[2 * (int64_t)j], ey = yi - pts[2 * (int64_t)j + 1];
                        if (ex * ex + ey * ey < r2) nb.push_back(j);
                    }
                }
            std::sort(nb.begin(), nb.end());
            for (int32_t j : nb) { out.push_back((int32_t)i); out.push_back(j); }
        }
    }
    size_t total = 0;
    for (auto& p : part) total += p.size();
    g_edges.reserve(total);
    for (auto& p : part) { g_edges.insert(g_edges.end(), p.begin(), p.end()); std::vector<int32_t>().swap(p); }
    return (int64_t)(g_edges.size() / 2);
}

void wbd_take_edges(int32_t* out /* [m][2] */) {
    std::memcpy(out, g_edges.data(), g_edges.size() * sizeof(int32_t));
    std::vector<int32_t>().swap(g_edges);
}

// edges: unique, sorted by (src, dst), src < dst.  rowPtr[n + 1], col[2 m].  Returns 0, or -1 if the list is not canonical.
int wbd_csr_canonical(int64_t n, int64_t m, const int32_t* edges, int32_t* rowPtr, int32_t* col) {
    std::vector<int64_t> deg(n + 1, 0);
    for (int64_t e = 0; e < m; ++e) {
        const int32_t s = edges[2 * e], d = edges[2 * e + 1];
        if (s < 0 || d >= n || s >= d) return -1;
        if (e > 0 && (edges[2 * e - 2] > s || (edges[2 * e - 2] == s && edges[2 * e - 1] >= d))) return -1;
        ++deg[s]; ++deg[d];
    }
    std::vector<int64_t> at(n + 1, 0);
    for (int64_t v = 0; v < n; ++v) at[v + 1] = at[v] + deg[v];
    if (at[n] > INT32_MAX) return -1;
    for (int64_t v = 0; v <= n; ++v) rowPtr[v] = (int32_t)at[v];
    // rows ascending: first the smaller neighbours (arrive in ascending src order), then the larger ones (ascending dst order)
    for (int64_t e = 0; e < m; ++e) col[at[edges[2 * e + 1]]++] = edges[2 * e];
    for (int64_t e = 0; e < m; ++e) col[at[edges[2 * e]]++] = edges[2 * e + 1];
    return 0;
}

}  // extern "C"
