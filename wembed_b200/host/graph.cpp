#include "graph.hpp"

#include <algorithm>

namespace wembed {
namespace impl {

EmbeddingGraph::EmbeddingGraph(const std::vector<std::pair<int, int>>& edges) { build(0, edges); }
EmbeddingGraph::EmbeddingGraph(int numVertices, const std::vector<std::pair<int, int>>& edges) { build(numVertices, edges); }

void EmbeddingGraph::build(int numVertices, const std::vector<std::pair<int, int>>& edges) {
    // symmetrise + dedupe + order rows by sorting 64-bit (src, dst) keys; all self loops are dropped
    // (the reference skips only the first one it meets and then overruns its edge array, Graph.cpp:124-128)
    std::vector<std::uint64_t> keys;
    keys.reserve(edges.size() * 2);
    int maxId = -1;
    for (const auto& [a, b] : edges) {
        maxId = std::max(maxId, std::max(a, b));
        if (a == b) continue;
        keys.push_back((static_cast<std::uint64_t>(static_cast<std::uint32_t>(a)) << 32) | static_cast<std::uint32_t>(b));
        keys.push_back((static_cast<std::uint64_t>(static_cast<std::uint32_t>(b)) << 32) | static_cast<std::uint32_t>(a));
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const int n = std::max(numVertices, maxId + 1);   // Graph.cpp:101: number of nodes = largest id + 1
    rowPtr_.assign(static_cast<std::size_t>(n) + 1, 0);
    col_.resize(keys.size());
    for (std::size_t i = 0; i < keys.size(); ++i) {
        rowPtr_[(keys[i] >> 32) + 1]++;
        col_[i] = static_cast<std::int32_t>(keys[i] & 0xffffffffu);
    }
    for (int v = 0; v < n; ++v) rowPtr_[v + 1] += rowPtr_[v];
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
