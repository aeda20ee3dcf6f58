// Synthetic-workload helpers for bench.py and the tests (host side, not on the product path): the O(n) parts of
// wembed_b200/datasets.py that numpy does slowly at n = 1e7.  Results are identical to the numpy code paths
// (same comparisons in double, same edge order), which tests/test_cabi_host.py checks.
//
//   wbd_pairs_within   all index pairs (i < j) with ||p_i - p_j||^2 < r^2, lexicographic order
//                      (the rule of the reference's GeometricGraphSampler.cpp:10-51, found with a cell grid instead of its O(n^2) loop)
//   wbd_girg_pairs     threshold-GIRG edges: (i, j) with ||p_i - p_j||^2 < c^2 w_i w_j / W, lexicographic order (datasets.heavy_tailed_graph)
//   wbd_csr_canonical  CSR of an edge list that is already unique, sorted and has src < dst (Graph.cpp:87-150 invariants)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>

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

// Threshold GIRG-like edge rule of datasets.heavy_tailed_graph: vertices are binned into weight layers [2^k, 2^(k+1)); for every
// layer pair (ka <= kb) the vertices of layer kb are put on a grid whose cell is the largest distance that pair of layers can
// connect over, and every vertex of layer ka scans its 3 x 3 cells.  The predicate is evaluated exactly as the numpy code does
// (first factor = the vertex of the lower layer, or of the lower index inside one layer).  count_only: no edges are kept.
int64_t wbd_girg_pairs(int64_t n, const double* pts /* [n][2] */, const double* w, double c, double W, int count_only) {
    g_edges.clear();
    if (n <= 0) return 0;
    std::vector<int> cls(n);
    int maxCls = 0;
    for (int64_t i = 0; i < n; ++i) { cls[i] = (int)std::floor(std::log2(w[i])); maxCls = std::max(maxCls, cls[i]); }
    std::vector<std::vector<int32_t>> layer(maxCls + 1);
    for (int64_t i = 0; i < n; ++i) layer[cls[i]].push_back((int32_t)i);
    const int threads = 16;
    std::vector<std::vector<int64_t>> part(threads);
    std::vector<int64_t> counts(threads, 0);
    const double cc = c * c;
    for (int ka = 0; ka <= maxCls; ++ka) {
        const std::vector<int32_t>& A = layer[ka];
        if (A.empty()) continue;
        for (int kb = ka; kb <= maxCls; ++kb) {
            const std::vector<int32_t>& B = layer[kb];
            if (B.empty()) continue;
            const double rmax = std::min(1.5, c * std::sqrt(std::ldexp(1.0, ka + 1) * std::ldexp(1.0, kb + 1) / W));
            // grid over the unit square (positions are uniform in [0, 1)^2), cells of side >= rmax
            const int64_t g = std::max<int64_t>(1, std::min<int64_t>(4096, (int64_t)std::floor(1.0 / rmax)));
            const double inv = (double)g;
            std::vector<int64_t> start(g * g + 1, 0);
            std::vector<int32_t> cellOf(B.size());
            for (size_t k = 0; k < B.size(); ++k) {
                const int64_t gx = std::min<int64_t>(g - 1, (int64_t)(pts[2 * (int64_t)B[k]] * inv)), gy = std::min<int64_t>(g - 1, (int64_t)(pts[2 * (int64_t)B[k] + 1] * inv));
                cellOf[k] = (int32_t)(gx * g + gy);
                ++start[cellOf[k] + 1];
            }
            for (int64_t q = 0; q < g * g; ++q) start[q + 1] += start[q];
            std::vector<int32_t> member(B.size());
            {
                std::vector<int64_t> at(start.begin(), start.end() - 1);
                for (size_t k = 0; k < B.size(); ++k) member[at[cellOf[k]]++] = B[k];
            }
            const bool same = ka == kb;
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
            for (int64_t ai = 0; ai < (int64_t)A.size(); ++ai) {
                const int t = omp_get_thread_num();
                const int64_t i = A[ai];
                const double xi = pts[2 * i], yi = pts[2 * i + 1], wi = w[i];
                const int64_t gx = std::min<int64_t>(g - 1, (int64_t)(xi * inv)), gy = std::min<int64_t>(g - 1, (int64_t)(yi * inv));
                for (int64_t x = std::max<int64_t>(0, gx - 1); x <= std::min<int64_t>(g - 1, gx + 1); ++x)
                    for (int64_t y = std::max<int64_t>(0, gy - 1); y <= std::min<int64_t>(g - 1, gy + 1); ++y)
                        for (int64_t k = start[x * g + y]; k < start[x * g + y + 1]; ++k) {
                            const int64_t j = member[k];
                            if (same && j <= i) continue;
                            const double ex = xi - pts[2 * j], ey = yi - pts[2 * j + 1];
                            if (ex * ex + ey * ey < cc * wi * w[j] / W) {
                                ++counts[t];
                                if (!count_only) part[t].push_back(i < j ? (i << 32) | j : (j << 32) | i);
                            }
                        }
            }
        }
    }
    int64_t total = 0;
    for (int64_t k : counts) total += k;
    if (count_only) return total;
    std::vector<int64_t> keys;
    keys.reserve((size_t)total);
    for (auto& p : part) { keys.insert(keys.end(), p.begin(), p.end()); std::vector<int64_t>().swap(p); }
    std::sort(keys.begin(), keys.end());
    g_edges.resize(keys.size() * 2);
    for (size_t e = 0; e < keys.size(); ++e) { g_edges[2 * e] = (int32_t)(keys[e] >> 32); g_edges[2 * e + 1] = (int32_t)(keys[e] & 0xffffffff); }
    return total;
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
