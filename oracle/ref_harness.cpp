// TEST INFRASTRUCTURE - C harness around the reference's own WembedEmbedder.
//
// Compiled (with -fno-access-control, so it can read the embedder's private state)
// together with the reference's unmodified sources, in place from /root/reference,
// into oracle/_ref/libwembed_ref.so by oracle/Makefile.  Nothing here re-implements
// reference logic: every call forwards to the reference
// (src/embeddingLib/src/embedder/WembedEmbedder.cpp).
#include <omp.h>

#include <cstring>
#include <memory>
#include <vector>

#include "Graph.hpp"
#include "GraphIO.hpp"
#include "GraphHierarchy.hpp"
#include "LabelPropagation.hpp"
#include "LayeredEmbedder.hpp"
#include "Rand.hpp"
#include "WembedEmbedder.hpp"
#include "oracle_api.h"

namespace {
struct RefHandle {
    Graph graph;
    EmbedderOptions opts;
    std::unique_ptr<WembedEmbedder> emb;
};

EmbedderOptions translate(const orc_options& o) {
    EmbedderOptions e;
    e.embeddingDimension = o.embeddingDimension;
    e.dimensionHint = o.dimensionHint;
    e.weightType = o.weightType == 0 ? WeightType::Unit : WeightType::Degree;
    e.indexType = IndexType::SNN;  // sprk is not buildable here (oracle/shims/sprk/sprk.h)
    e.doublingFactor = o.doublingFactor;
    e.attractionScale = o.attractionScale;
    e.repulsionScale = o.repulsionScale;
    e.centreScale = o.centreScale;
    e.edgeLength = o.edgeLength;
    e.optimizerType = o.optimizerType == 0 ? OptimizerType::Simple : OptimizerType::Adam;
    e.maxIterations = o.maxIterations;
    e.simpleOptMaxDisplacement = o.simpleOptMaxDisplacement;
    e.lrScheduleType = o.lrScheduleType == 0 ? LRScheduleType::ExponentialCooling : LRScheduleType::LossAdaptive;
    e.learningRate = o.learningRate;
    e.warmupSteps = o.warmupSteps;
    e.lrCoolingFactor = o.lrCoolingFactor;
    e.lrDecayFactor = o.lrDecayFactor;
    e.lrDecayThreshold = o.lrDecayThreshold;
    e.lrAdaptPatience = o.lrAdaptPatience;
    e.lrGrowthFactor = o.lrGrowthFactor;
    e.lrGrowthThreshold = o.lrGrowthThreshold;
    e.stopCriterion = o.stopCriterion == 0 ? StopCriterionType::Displacement : StopCriterionType::Loss;
    e.stopDisplacementTol = o.stopDisplacementTol;
    e.stopDisplacementPatience = o.stopDisplacementPatience;
    e.lossSmoothingFactor = o.lossSmoothingFactor;
    e.lossRateWindow = o.lossRateWindow;
    e.stopLossTol = o.stopLossTol;
    e.stopLossPatience = o.stopLossPatience;
    return e;
}
}  // namespace

extern "C" {

void ref_options_default(orc_options* o) {
    const EmbedderOptions e;
    std::memset(o, 0, sizeof(*o));
    o->embeddingDimension = e.embeddingDimension;
    o->weightType = static_cast<int>(e.weightType);
    o->optimizerType = static_cast<int>(e.optimizerType);
    o->maxIterations = e.maxIterations;
    o->lrScheduleType = static_cast<int>(e.lrScheduleType);
    o->warmupSteps = e.warmupSteps;
    o->lrAdaptPatience = e.lrAdaptPatience;
    o->stopCriterion = static_cast<int>(e.stopCriterion);
    o->stopDisplacementPatience = e.stopDisplacementPatience;
    o->lossRateWindow = e.lossRateWindow;
    o->stopLossPatience = e.stopLossPatience;
    o->numThreads = 0;
    o->dimensionHint = e.dimensionHint;
    o->attractionScale = e.attractionScale;
    o->repulsionScale = e.repulsionScale;
    o->centreScale = e.centreScale;
    o->edgeLength = e.edgeLength;
    o->doublingFactor = e.doublingFactor;
    o->simpleOptMaxDisplacement = e.simpleOptMaxDisplacement;
    o->learningRate = e.learningRate;
    o->lrCoolingFactor = e.lrCoolingFactor;
    o->lrDecayFactor = e.lrDecayFactor;
    o->lrDecayThreshold = e.lrDecayThreshold;
    o->lrGrowthFactor = e.lrGrowthFactor;
    o->lrGrowthThreshold = e.lrGrowthThreshold;
    o->stopDisplacementTol = e.stopDisplacementTol;
    o->lossSmoothingFactor = e.lossSmoothingFactor;
    o->stopLossTol = e.stopLossTol;
}

void* ref_create(int32_t n_hint, int64_t m, const int32_t* src, const int32_t* dst, const orc_options* o, int32_t seed,
                 int32_t init_state) {
    (void)n_hint;  // the reference derives n from the largest id (Graph.cpp:101)
    if (o->numThreads > 0) omp_set_num_threads(o->numThreads);
    Rand::setSeed(seed);
    std::vector<std::pair<int, int>> edges;
    edges.reserve(m);
    for (int64_t i = 0; i < m; ++i) edges.emplace_back(src[i], dst[i]);
    auto* h = new RefHandle{Graph(edges), translate(*o), nullptr};
    h->emb = std::make_unique<WembedEmbedder>(h->graph, h->opts, std::make_shared<util::Timer>(), init_state != 0);
    return h;
}

void ref_destroy(void* p) { delete static_cast<RefHandle*>(p); }

int32_t ref_num_vertices(void* p) { return static_cast<RefHandle*>(p)->graph.getNumVertices(); }
int64_t ref_num_directed_edges(void* p) { return 2 * static_cast<int64_t>(static_cast<RefHandle*>(p)->graph.getNumEdges()); }

void ref_csr(void* p, int32_t* row_ptr, int32_t* col) {
    const Graph& g = static_cast<RefHandle*>(p)->graph;
    const int n = g.getNumVertices();
    for (int v = 0; v <= n; ++v) row_ptr[v] = g.nodes[v].firstEdge;
    for (std::size_t e = 0; e < g.edges.size(); ++e) col[e] = g.edges[e].neighbour;
}

int32_t ref_are_neighbors(void* p, int32_t v, int32_t u) { return static_cast<RefHandle*>(p)->graph.areNeighbors(v, u) ? 1 : 0; }

void ref_set_coordinates(void* p, const double* c) {
    auto* h = static_cast<RefHandle*>(p);
    const int n = h->graph.getNumVertices(), d = h->opts.embeddingDimension;
    std::vector<std::vector<double>> rows(n, std::vector<double>(d));
    for (int v = 0; v < n; ++v)
        for (int k = 0; k < d; ++k) rows[v][k] = c[static_cast<std::size_t>(v) * d + k];
    h->emb->setCoordinates(rows);
}

void ref_set_weights(void* p, const double* w) {
    auto* h = static_cast<RefHandle*>(p);
    h->emb->setWeights(std::vector<double>(w, w + h->graph.getNumVertices()));
}

void ref_get_coordinates(void* p, double* c) { static_cast<RefHandle*>(p)->emb->copyCoordinatesTo(c); }

void ref_get_weights(void* p, double* w) {
    const auto ws = static_cast<RefHandle*>(p)->emb->getWeights();
    std::copy(ws.begin(), ws.end(), w);
}

void ref_get_forces(void* p, double* f) { static_cast<RefHandle*>(p)->emb->state.force.copyToFlat(f); }

void ref_step(void* p) { static_cast<RefHandle*>(p)->emb->calculateStep(); }

int32_t ref_is_finished(void* p) { return static_cast<RefHandle*>(p)->emb->isFinished() ? 1 : 0; }

int64_t ref_run(void* p) {
    auto* h = static_cast<RefHandle*>(p);
    h->emb->calculateEmbedding();
    return static_cast<int64_t>(h->emb->state.currentIteration);
}

void ref_get_stats(void* p, double* s) {
    auto* e = static_cast<RefHandle*>(p)->emb.get();
    s[ORC_LOSS_ATTRACT] = e->state.lastAttractLoss;
    s[ORC_LOSS_REPEL] = e->state.lastRepelLoss;
    s[ORC_LR] = e->state.lastLearningRate;
    s[ORC_REL_DISP] = e->state.lastRelDisplacement;
    s[ORC_REL_LOSS_IMPROVEMENT] = e->state.lastRelLossImprovement;
    s[ORC_ITERATION] = static_cast<double>(e->state.currentIteration);
    s[ORC_NUM_REP_PAIRS] = static_cast<double>(e->numRepForceCalculations);
    s[ORC_RESERVED] = 0.0;
}

int64_t ref_candidates(void* p, int32_t v, int32_t* out, int64_t cap) {
    auto* e = static_cast<RefHandle*>(p)->emb.get();
    static int lastBuiltFor = -1;
    (void)lastBuiltFor;
    e->updateIndex();
    VecBuffer<2> buf(e->opts.embeddingDimension);
    const std::vector<NodeId> c = e->getRepellingCandidatesForNode(v, buf);
    for (std::size_t i = 0; i < c.size() && static_cast<int64_t>(i) < cap; ++i) out[i] = c[i];
    return static_cast<int64_t>(c.size());
}


// ---- multilevel driver: reference-only helpers used to generate golden fixtures ------------------------

// LabelPropagation::coarsenAllLayers with the defaults of wembed::createEmbedder (src/wembed.cpp:229-233): returns the number
// of layers; layer_sizes[l] = vertices of layer l, parents = the parent pointers of all layers concatenated.
int32_t ref_coarsen(int64_t m, const int32_t* src, const int32_t* dst, int32_t* layer_sizes, int32_t max_layers, int32_t* parents,
                    int64_t cap) {
    std::vector<std::pair<int, int>> edges;
    for (int64_t i = 0; i < m; ++i) edges.emplace_back(src[i], dst[i]);
    Graph g(edges);
    std::vector<double> edgeWeights(g.getNumEdges() * 2, 1.0);
    LabelPropagation coarsener(PartitionerOptions{}, g, edgeWeights);
    const ParentPointerTree tree = coarsener.coarsenAllLayers();
    int64_t at = 0;
    for (std::size_t l = 0; l < tree.size() && static_cast<int32_t>(l) < max_layers; ++l) {
        layer_sizes[l] = static_cast<int32_t>(tree[l].size());
        for (NodeId p : tree[l])
            if (at < cap) parents[at++] = p;
    }
    return static_cast<int32_t>(tree.size());
}

// The whole layered embedding (LayeredEmbedder::calculateEmbedding): final coordinates / weights of layer 0; returns iterations.
int64_t ref_layered_run(int64_t m, const int32_t* src, const int32_t* dst, const orc_options* o, int32_t seed, double* coords,
                        double* weights, double* stats8) {
    if (o->numThreads > 0) omp_set_num_threads(o->numThreads);
    Rand::setSeed(seed);
    std::vector<std::pair<int, int>> edges;
    for (int64_t i = 0; i < m; ++i) edges.emplace_back(src[i], dst[i]);
    Graph g(edges);
    std::vector<double> edgeWeights(g.getNumEdges() * 2, 1.0);
    LabelPropagation coarsener(PartitionerOptions{}, g, edgeWeights);
    LayeredEmbedder emb(g, coarsener, translate(*o));
    emb.calculateEmbedding();
    emb.copyCoordinatesTo(coords);
    const auto w = emb.getWeights();
    std::copy(w.begin(), w.end(), weights);
    const EmbeddingLoss l = emb.getLoss();
    stats8[ORC_LOSS_ATTRACT] = l.attractive;
    stats8[ORC_LOSS_REPEL] = l.repulsive;
    stats8[ORC_LR] = emb.getCurrentLearningRate();
    stats8[ORC_REL_DISP] = emb.getLastRelDisplacement();
    stats8[ORC_ITERATION] = static_cast<double>(emb.currentIteration);
    return emb.currentIteration;
}

// The reference's own edge-list reader (GraphIO::readEdgeList, GraphIO.cpp:10-51) + Graph construction (Graph.cpp:87-150): the graph
// it builds is kept until ref_graph_csr copies it out.  Returns the number of vertices; *directed = CSR entries.
static Graph g_fileGraph;
int32_t ref_read_edge_list(const char* path, const char* comment, const char* delimiter, int64_t* directed) {
    g_fileGraph = GraphIO::readEdgeList(path, comment, delimiter);
    *directed = 2 * static_cast<int64_t>(g_fileGraph.getNumEdges());
    return g_fileGraph.getNumVertices();
}
void ref_graph_csr(int32_t* row_ptr, int32_t* col) {
    const int n = g_fileGraph.getNumVertices();
    for (int v = 0; v <= n; ++v) row_ptr[v] = g_fileGraph.nodes[v].firstEdge;
    for (std::size_t e = 0; e < g_fileGraph.edges.size(); ++e) col[e] = g_fileGraph.edges[e].neighbour;
}

}  // extern "C"
