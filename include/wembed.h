/*
 * wembed.h - public C++ API of the B200-native WEmbed.
 *
 * Source compatible with the reference's include/wembed.h (Vraier/wembed, include/wembed.h:116-223):
 * the same namespace, types, enumerators, Options fields with the same defaults, free functions and
 * member functions, so code written against the reference recompiles against this header unchanged.
 * The implementation behind it is different: Embedder drives a device-resident embedding through the
 * C ABI in wembed_b200.h (there is no CPU code path for calculateStep).
 */
#ifndef WEMBED_PUBLIC_API_H
#define WEMBED_PUBLIC_API_H

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace wembed {

#ifndef _WEMBED_IS_IMPL
namespace impl {
class EmbeddingGraph;     // CSR graph, see wembed_b200/host/graph.hpp
class EmbedderInterface;  // see wembed_b200/host/embedder.hpp
}  // namespace impl
#endif

using NodeId = std::int32_t;
using EdgeId = std::int32_t;

class Embedder;

/* ---- enumerations (values as in the reference, wembed.h:24-42) ------------------------------------------ */

enum SpatialIndex : std::int32_t { IndexSNN = 1, IndexSprk = 2 };   /* both select the device index */
enum OptimizerType : std::int32_t { OptimizerSimple = 0, OptimizerAdam = 1 };
enum LRSchedule : std::int32_t { LRExponentialCooling = 0, LRLossAdaptive = 1 };
enum StopCriterion : std::int32_t { StopDisplacement = 0, StopLoss = 1 };

/* ---- plain data ------------------------------------------------------------------------------------------- */

struct Edge {
    NodeId src;
    NodeId dst;
};

/* One row of the phase timing tree; depth 0 = top level, value in seconds. */
struct TimingResult {
    std::uint64_t depth;
    std::string displayName;
    double value;
};

struct Loss {
    double attractive;
    double repulsive;
    double total;
};

/* Field names, order and defaults follow wembed.h:65-114 of the reference. */
struct Options {
    /* embedding */
    std::int32_t embeddingDimension = 4;
    bool useUnitWeights = false;          /* false: degree weights (WeightType::Degree), true: all weights 1 */
    double dimensionHint = -1.0;          /* > 0: weights are raised to embeddingDimension / dimensionHint */
    bool layeredEmbedding = false;        /* multilevel driver; not part of the device hot path (see createEmbedder) */

    /* forces */
    SpatialIndex indexType = IndexSprk;
    double attractionScale = 1.0;
    double repulsionScale = 1.0;
    double centreScale = 0.0;
    double edgeLength = 1.0;
    double expansionStretch = 1.0;

    /* gradient descent */
    OptimizerType optimizerType = OptimizerAdam;
    std::int32_t maxIterations = 10000;
    double simpleOptMaxDisplacement = 1.0;

    /* learning-rate schedule */
    LRSchedule lrSchedule = LRExponentialCooling;
    double learningRate = 10.0;
    std::int32_t warmupSteps = 20;
    double lrCoolingFactor = 0.995;
    double lrDecayFactor = 0.5;
    double lrDecayThreshold = 1e-2;
    std::int32_t lrAdaptPatience = 20;
    double lrGrowthFactor = 1.0;
    double lrGrowthThreshold = 1e-1;

    /* stopping */
    StopCriterion stopCriterion = StopLoss;
    double stopDisplacementTol = 3e-4;
    std::int32_t stopDisplacementPatience = 5;
    double lossSmoothingFactor = 0.3;
    std::int32_t lossRateWindow = 30;
    double stopLossTol = 1e-3;
    std::int32_t stopLossPatience = 50;
};

/* ---- Graph: move-only handle to a static undirected graph in CSR form ----------------------------------- */

class Graph {
   public:
    explicit Graph(std::unique_ptr<impl::EmbeddingGraph>&& graph);
    ~Graph();
    Graph(const Graph&) = delete;
    Graph& operator=(const Graph&) = delete;
    Graph(Graph&& other);
    Graph& operator=(Graph&& other);

    NodeId getNumVertices() const;
    EdgeId getNumEdges() const;                       /* undirected edges */

    std::vector<EdgeId> getEdges(NodeId v) const;     /* CSR slots of v */
    std::vector<NodeId> getNeighbors(NodeId v) const; /* ascending */
    int getNumNeighbors(NodeId v) const;
    NodeId getEdgeTarget(EdgeId e) const;
    bool areNeighbors(NodeId v, NodeId u) const;
    std::vector<Edge> getEdgeList() const;            /* every undirected edge once, src < dst */
    std::string toString() const;

    /* Additions of this build (not in the reference): read-only views of the CSR the embedder uploads, valid while the Graph lives. */
    const std::int32_t* csrOffsets() const;           /* getNumVertices() + 1 entries */
    const std::int32_t* csrTargets() const;           /* 2 * getNumEdges() entries, every row ascending */

   private:
    friend Embedder createEmbedder(const Graph& g, const Options& options);
    std::unique_ptr<impl::EmbeddingGraph> _graph;
};

/* ---- Embedder --------------------------------------------------------------------------------------------- */

class Embedder {
   public:
    explicit Embedder(std::unique_ptr<impl::EmbedderInterface>&& embedder);
    ~Embedder();
    Embedder(const Embedder&) = delete;
    Embedder& operator=(const Embedder&) = delete;
    Embedder(Embedder&& other);
    Embedder& operator=(Embedder&& other);

    void calculateStep();
    bool isFinished() const;
    void calculateEmbedding();

    std::int32_t getNumVertices() const;
    std::int32_t getEmbeddingDimension() const;
    void copyCoordinatesTo(double* out) const;        /* n * d doubles, row-major */

    Graph getCurrentGraph() const;
    std::vector<std::vector<double>> getCoordinates() const;
    std::vector<double> getWeights() const;
    void setCoordinates(const std::vector<std::vector<double>>& coordinates);
    void setWeights(const std::vector<double>& weights);

    std::vector<TimingResult> getTimings() const;
    Loss getLoss() const;
    double getCurrentLearningRate() const;
    double getLastRelDisplacement() const;
    double getLastRelLossImprovement() const;

    void writeCoordinates(const std::string& filePath, bool writeWeights = true) const;

   private:
    std::unique_ptr<impl::EmbedderInterface> _embedder;
};

/* ---- free functions --------------------------------------------------------------------------------------- */

Embedder createEmbedder(const Graph& g, const Options& options);
Graph graphFromEdges(const std::vector<Edge>& edges);
Graph graphFromEdgeListFile(const std::string& filePath, const std::string& comment = "#", const std::string& delimiter = " ");
std::vector<std::vector<double>> readCoordinatesFromFile(const std::string& filePath, const std::string& comment = "%",
                                                         const std::string& delimiter = ",");
std::string timingsToString(const std::vector<TimingResult>& timings);
void setSeed(int seed);

}  // namespace wembed

#endif /* WEMBED_PUBLIC_API_H */
