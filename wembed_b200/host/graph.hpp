// Host-side CSR graph with the invariants of the reference's Graph (src/graphLib/src/graph/Graph.cpp:87-150):
// symmetric, deduplicated, neighbours ascending, n = largest id + 1.  Built by sorting packed (src, dst) keys
// instead of the reference's std::map<int, std::set<int>> (3.4 s at n = 1e5).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace wembed {
namespace impl {

class EmbeddingGraph {
   public:
    EmbeddingGraph() : rowPtr_(1, 0) {}
    explicit EmbeddingGraph(const std::vector<std::pair<int, int>>& edges);   // Graph::constructFromEdges (Graph.cpp:139-150)
    EmbeddingGraph(int numVertices, const std::vector<std::pair<int, int>>& edges);   // Graph(map): vertices without edges kept (Graph.cpp:87-137)

    int32_t getNumVertices() const { return static_cast<int32_t>(rowPtr_.size()) - 1; }
    int32_t getNumEdges() const { return static_cast<int32_t>(col_.size() / 2); }
    int getNumNeighbors(int32_t v) const { return rowPtr_[v + 1] - rowPtr_[v]; }
    std::vector<int32_t> getEdges(int32_t v) const;
    std::vector<int32_t> getNeighbors(int32_t v) const;
    int32_t getEdgeTarget(int32_t e) const { return col_[e]; }
    bool areNeighbors(int32_t v, int32_t u) const;   // Graph.cpp:67-83 (binary search: rows are sorted)
    std::string toString() const;                    // Graph.cpp:165-180

    const std::vector<int32_t>& rowPtr() const { return rowPtr_; }
    const std::vector<int32_t>& col() const { return col_; }

   private:
    void build(int numVertices, const std::vector<std::pair<int, int>>& edges);
    std::vector<int32_t> rowPtr_, col_;
};

}  // namespace impl
}  // namespace wembed
