// Multilevel driver, host side: graph coarsening by size-capped label propagation and the
// layer-by-layer embedding loop.  Restates, with sort-based graph construction instead of std::map<std::set>:
//   LabelPropagation   src/embeddingLib/src/partition/LabelPropagation.cpp:13-239
//   coarsenGraph       src/graphLib/src/graph/GraphAlgorithms.cpp:107-143
//   GraphHierarchy     src/embeddingLib/src/partition/GraphHierarchy.cpp:5-58 (graphs + node parent pointers)
//   LayeredEmbedder    src/embeddingLib/src/embedder/LayeredEmbedder.cpp:5-95
// Every level is embedded by the device embedder; only the scalar / integer coarsening logic runs on the host, as it does in
// the reference (it is inherently sequential: nodes are processed one by one in degree order).
#pragma once
#include <memory>
#include <vector>

#include "embedder.hpp"
#include "graph.hpp"

namespace wembed {
namespace impl {

// PartitionerOptions (src/embeddingLib/include/partition/Partitioner.hpp:9-16)
struct CoarseningOptions {
    int maxIterations = 20;
    int maxClusterSize = 6;
    int finalGraphSize = 10;
};

using ParentPointerTree = std::vector<std::vector<int32_t>>;

// GraphAlgo::coarsenGraph: contracts every cluster to one vertex; edgeMap[e] = CSR slot of the coarse edge that fine slot e
// falls on, or -1 for an edge inside a cluster.
std::pair<EmbeddingGraph, std::vector<int32_t>> coarsenGraph(const EmbeddingGraph& g, const std::vector<int32_t>& clusterId);

// LabelPropagation::coarsenAllLayers: parent pointers of every layer, ending with "all remaining vertices -> 0" and {-1}.
ParentPointerTree coarsenAllLayers(const EmbeddingGraph& g, const std::vector<double>& edgeWeights, const CoarseningOptions& o = {});

struct Hierarchy {
    std::vector<EmbeddingGraph> graphs;     // graphs[0] = the input graph, graphs.back() = a single vertex
    ParentPointerTree parent;               // parent[l][v] = vertex of layer l+1 that contains v
    Hierarchy(const EmbeddingGraph& g, const CoarseningOptions& o = {});
    int numLayers() const { return static_cast<int>(graphs.size()); }
};

class LayeredDeviceEmbedder final : public EmbedderInterface {
   public:
    LayeredDeviceEmbedder(const EmbeddingGraph& graph, const Options& options);

    void calculateStep() override;
    bool isFinished() override { return currentLayer_ == 0 && current_->isFinished(); }
    void calculateEmbedding() override;
    EmbeddingGraph getCurrentGraph() override { return hierarchy_.graphs[currentLayer_]; }
    std::vector<std::vector<double>> getCoordinates() override { return current_->getCoordinates(); }
    std::vector<double> getWeights() override { return current_->getWeights(); }
    std::vector<PhaseTiming> getTimings() override;
    void setCoordinates(const std::vector<std::vector<double>>&) override;   // no effect, like the reference (LayeredEmbedder.cpp:26-30)
    void setWeights(const std::vector<double>&) override;                    // no effect (:32-36)
    int getNumVertices() const override { return current_->getNumVertices(); }
    int getEmbeddingDimension() const override { return current_->getEmbeddingDimension(); }
    void copyCoordinatesTo(double* out) const override { current_->copyCoordinatesTo(out); }
    EmbeddingLoss getLoss() const override { return current_->getLoss(); }
    double getCurrentLearningRate() const override { return current_->getCurrentLearningRate(); }
    double getLastRelDisplacement() const override { return current_->getLastRelDisplacement(); }
    double getLastRelLossImprovement() const override { return current_->getLastRelLossImprovement(); }

    int currentLayer() const { return currentLayer_; }
    long long iterations() const { return iterations_; }

   private:
    void expandPositions();
    Options opts_;
    Hierarchy hierarchy_;
    int currentLayer_;
    long long iterations_ = 0;
    double expandSeconds_ = 0.0, totalSeconds_ = 0.0;
    std::vector<PhaseTiming> finishedLayers_;
    std::unique_ptr<DeviceEmbedder> current_;
};

}  // namespace impl
}  // namespace wembed
