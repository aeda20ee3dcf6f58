#include "graph.hpp"

#include <algorithm>
#include <limits>
#include <stdexcept>
#include <thread>

namespace wembed {
namespace impl {

EmbeddingGraph::EmbeddingGraph(const std::vector<std::pair<int, int>>& edges) { build(0, edges); }
EmbeddingGraph::EmbeddingGraph(int numVertices, const std::vector<std::pair<int, int>>& edges) { build(numVertices, edges); }

void EmbeddingGraph::build(int numVertices, const std::vector<std::pair<int, int>>& edges) {
    // Graph::constructFromEdges semantics (Graph.cpp:87-150) without its std::map<int, std::set<int>>: symmetrise, order every row
    // ascending, drop duplicates and ALL self loops (the reference skips only the first one it meets and then overruns its edge
    // array, Graph.cpp:124-128).  Counting sort by source vertex, then each (short) row is sorted and deduplicated in place:
    // O(m) passes over the edge list instead of a 2m-key comparison sort.
    int maxId = -1;
    for (const auto& [a, b] : edges) {
        if (a < 0 || b < 0) throw std::invalid_argument("wembed::Graph: negative vertex id in the edge list");
        maxId = std::max(maxId, std::max(a, b));
    }
    const int n = std::max(numVertices, maxId + 1);   // Graph.cpp:101: number of nodes = largest id + 1
    std::vector<std::int64_t> start(static_cast<std::size_t>(n) + 1, 0);
    for (const auto& [a, b] : edges) {
        if (a == b) continue;
        start[static_cast<std::size_t>(a) + 1]++;
        start[static_cast<std::size_t>(b) + 1]++;
    }
    for (int v = 0; v < n; ++v) start[v + 1] += start[v];
    if (start[n] > std::numeric_limits<std::int32_t>::max()) throw std::invalid_argument("wembed::Graph: more than 2^31 - 1 directed edges");
    std::vector<std::int32_t> raw(static_cast<std::size_t>(start[n]));
    {
        std::vector<std::int64_t> cursor(start.begin(), start.end() - 1);
        for (const auto& [a, b] : edges) {
            if (a == b) continue;
            raw[static_cast<std::size_t>(cursor[a]++)] = b;
            raw[static_cast<std::size_t>(cursor[b]++)] = a;
        }
    }
    rowPtr_.assign(static_cast<std::size_t>(n) + 1, 0);
    // rows are independent: sort + unique each, remember the deduplicated length
    const unsigned workers = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto sortRows = [&](int v0, int v1) {
        for (int v = v0; v < v1; ++v) {
            auto first = raw.begin() + start[v], last = raw.begin() + start[v + 1];
            std::sort(first, last);
            rowPtr_[static_cast<std::size_t>(v) + 1] = static_cast<std::int32_t>(std::unique(first, last) - first);
        }
    };
    if (n < 65536 || workers == 1) {
        sortRows(0, n);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < workers; ++t)
            pool.emplace_back(sortRows, static_cast<int>(static_cast<std::int64_t>(n) * t / workers), static_cast<int>(static_cast<std::int64_t>(n) * (t + 1) / workers));
        for (auto& th : pool) th.join();
    }
    for (int v = 0; v < n; ++v) rowPtr_[static_cast<std::size_t>(v) + 1] += rowPtr_[v];
    col_.resize(static_cast<std::size_t>(rowPtr_[n]));
    for (int v = 0; v < n; ++v)
        std::copy(raw.begin() + start[v], raw.begin() + start[v] + (rowPtr_[v + 1] - rowPtr_[v]), col_.begin() + rowPtr_[v]);
}

std::vector<std::int32_t> EmbeddingGraph::getEdges(std::int32_t v) const {
    std::vector<std::int32_t> out(getNumNeighbors(v));
    for (std::size_t i = 0; i < out.size(); ++i) out[i] = rowPtr_[v] + static_cast<std::int32_t>(i);
    return out;
}

std::vector<std::int32_t> EmbeddingGraph::getNeighbors(std::int32_t v) const {
    return std::vector<std::int32_t>(col_.begin() + rowPtr_[v], col_.begin() + rowPtr_[v + 1]);
}

bool EmbeddingGraph::areNeighbors(std::int32_t v, std::int32_t u) const {
    if (getNumNeighbors(v) > getNumNeighbors(u)) std::swap(v, u);
    return std::binary_search(col_.begin() + rowPtr_[v], col_.begin() + rowPtr_[v + 1], u);
}

std::string EmbeddingGraph::toString() const {
    std::string out = "Graph AdjList:\n";
    for (std::int32_t v = 0; v < getNumVertices(); ++v) {
        out += std::to_string(v) + ": ";
        for (std::int32_t e = rowPtr_[v]; e < rowPtr_[v + 1]; ++e) out += std::to_string(col_[e]) + " ";
        out += "\n";
    }
    return out;
}

}  // namespace impl
}  // namespace wembed
