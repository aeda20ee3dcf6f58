#include "hierarchy.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <numeric>

namespace wembed {
namespace impl {

namespace {

// compactClusterIds (LabelPropagation.cpp:204-221): renumber in order of first appearance, e.g. [1,3,1,6,6,5,5] -> [0,1,0,2,2,3,3]
std::vector<int32_t> compactIds(const std::vector<int32_t>& ids) {
    const int n = static_cast<int>(ids.size());
    std::vector<int32_t> map(n, -1), out(n);
    int next = 0;
    for (int v = 0; v < n; ++v) {
        if (map[ids[v]] == -1) map[ids[v]] = next++;
        out[v] = map[ids[v]];
    }
    return out;
}

// labelPropagation (LabelPropagation.cpp:58-112): nodes in ascending-degree order (std::sort, as the reference), each moves to
// the neighbouring cluster it has most edge weight to, provided the cluster stays within maxClusterSize.
std::vector<int32_t> labelPropagation(const EmbeddingGraph& g, const std::vector<double>& edgeW, const CoarseningOptions& o) {
    const int n = g.getNumVertices();
    const auto& rp = g.rowPtr();
    const auto& col = g.col();
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](const int32_t& a, const int32_t& b) { return g.getNumNeighbors(a) < g.getNumNeighbors(b); });
    std::vector<int32_t> cluster(n);
    std::iota(cluster.begin(), cluster.end(), 0);
    std::vector<double> edgeSum(n, 0.0);
    std::vector<int> clusterSize(n, 0);   // the reference starts the sizes at 0 as well (:69), so the cap counts moves, not members
    for (int it = 0; it < o.maxIterations; ++it) {
        for (int i = 0; i < n; ++i) {
            const int v = order[i];
            for (int e = rp[v]; e < rp[v + 1]; ++e) edgeSum[cluster[col[e]]] += edgeW[e];
            int best = cluster[v];
            const int original = cluster[v];
            double bestW = 0.0;
            for (int e = rp[v]; e < rp[v + 1]; ++e) {
                const int c = cluster[col[e]];
                if (edgeSum[c] > bestW && ((clusterSize[c] + 1) <= o.maxClusterSize || c == original)) {
                    bestW = edgeSum[c];
                    best = c;
                }
                edgeSum[c] = 0.0;
            }
            clusterSize[best] += 1;
            clusterSize[original] -= 1;
            cluster[v] = best;
        }
    }
    return compactIds(cluster);
}

// aggressivePropagation (LabelPropagation.cpp:114-180): vertices that did not merge in the previous round join the
// neighbour they have most weight to, regardless of the size cap; isolated vertices are paired up.
std::vector<int32_t> aggressivePropagation(const EmbeddingGraph& g, const std::vector<double>& edgeW, const std::vector<int32_t>& prevParents) {
    const int n = g.getNumVertices();
    const auto& rp = g.rowPtr();
    const auto& col = g.col();
    std::vector<int> numChildren(n, 0);
    std::vector<int32_t> cluster(n, -1);
    std::vector<double> edgeSum(n, 0.0);
    std::vector<int32_t> isolated;
    for (int32_t p : prevParents) numChildren[p] += 1;
    for (int v = 0; v < n; ++v) {
        if (numChildren[v] > 1) { cluster[v] = v; continue; }
        if (rp[v + 1] > rp[v]) {
            for (int e = rp[v]; e < rp[v + 1]; ++e) edgeSum[col[e]] += edgeW[e];
            int best = -1;
            double bestW = -1.0;
            for (int e = rp[v]; e < rp[v + 1]; ++e) {
                const int t = col[e];
                if (edgeSum[t] > bestW) { bestW = edgeSum[t]; best = t; }
                edgeSum[t] = 0.0;
            }
            cluster[v] = best;
        } else {
            isolated.push_back(v);
        }
    }
    for (std::size_t i = 0; i < isolated.size(); ++i) cluster[isolated[i]] = (i % 2 == 1) ? isolated[i - 1] : isolated[i];
    return compactIds(cluster);
}

// calculateNewEdgeWeights (LabelPropagation.cpp:223-239)
std::vector<double> coarseEdgeWeights(const std::vector<double>& fine, const std::vector<int32_t>& edgeMap, std::size_t numCoarse) {
    std::vector<double> out(numCoarse, 0.0);
    for (std::size_t e = 0; e < fine.size(); ++e)
        if (edgeMap[e] != -1) out[edgeMap[e]] += fine[e];
    return out;
}

}  // namespace

std::pair<EmbeddingGraph, std::vector<int32_t>> coarsenGraph(const EmbeddingGraph& g, const std::vector<int32_t>& clusterId) {
    const int n = g.getNumVertices();
    const auto& rp = g.rowPtr();
    const auto& col = g.col();
    int numClusters = 0;
    for (int32_t c : clusterId) numClusters = std::max(numClusters, c + 1);
    std::vector<std::pair<int, int>> coarse;
    for (int v = 0; v < n; ++v)
        for (int e = rp[v]; e < rp[v + 1]; ++e)
            if (clusterId[v] != clusterId[col[e]]) coarse.emplace_back(clusterId[v], clusterId[col[e]]);
    EmbeddingGraph result(numClusters, coarse);   // every cluster is a vertex, also those without outside edges (GraphAlgorithms.cpp:113-116)
    std::vector<int32_t> edgeMap(col.size());
    const auto& crp = result.rowPtr();
    const auto& ccol = result.col();
    for (int v = 0; v < n; ++v) {
        for (int e = rp[v]; e < rp[v + 1]; ++e) {
            const int a = clusterId[v], b = clusterId[col[e]];
            if (a == b) { edgeMap[e] = -1; continue; }
            edgeMap[e] = static_cast<int32_t>(std::lower_bound(ccol.begin() + crp[a], ccol.begin() + crp[a + 1], b) - ccol.begin());
        }
    }
    return {std::move(result), std::move(edgeMap)};
}

ParentPointerTree coarsenAllLayers(const EmbeddingGraph& g, const std::vector<double>& edgeWeights, const CoarseningOptions& o) {
    ParentPointerTree parents;
    EmbeddingGraph current = g;
    std::vector<double> weights = edgeWeights;
    double shrink = 0.0;   // always a normal label propagation first (LabelPropagation.cpp:23)
    while (current.getNumVertices() > o.finalGraphSize && current.getNumEdges() > 0) {
        std::vector<int32_t> mapping = shrink < 0.5 ? labelPropagation(current, weights, o) : aggressivePropagation(current, weights, parents.back());
        auto coarse = coarsenGraph(current, mapping);
        weights = coarseEdgeWeights(weights, coarse.second, coarse.first.col().size());
        shrink = static_cast<double>(coarse.first.getNumVertices()) / static_cast<double>(current.getNumVertices());
        parents.push_back(std::move(mapping));
        current = std::move(coarse.first);
    }
    parents.emplace_back(current.getNumVertices(), 0);   // everything that is left goes into one vertex
    parents.emplace_back(1, -1);                         // end of the hierarchy
    return parents;
}

Hierarchy::Hierarchy(const EmbeddingGraph& g, const CoarseningOptions& o) {
    parent = coarsenAllLayers(g, std::vector<double>(g.col().size(), 1.0), o);   // unit edge weights (src/wembed.cpp:231)
    EmbeddingGraph current = g;
    for (std::size_t l = 0; l < parent.size(); ++l) {
        graphs.push_back(current);
        if (l + 1 < parent.size()) current = coarsenGraph(current, parent[l]).first;
    }
}

LayeredDeviceEmbedder::LayeredDeviceEmbedder(const EmbeddingGraph& graph, const Options& options)
    : opts_(options), hierarchy_(graph), currentLayer_(hierarchy_.numLayers() - 1),
      current_(std::make_unique<DeviceEmbedder>(hierarchy_.graphs[currentLayer_], options)) {}

void LayeredDeviceEmbedder::calculateStep() {
    ++iterations_;
    if (current_->isFinished() && currentLayer_ > 0) expandPositions();   // LayeredEmbedder.cpp:5-11
    current_->calculateStep();
}

void LayeredDeviceEmbedder::calculateEmbedding() {
    const auto t0 = std::chrono::steady_clock::now();
    iterations_ = 0;
    while (!isFinished()) calculateStep();
    totalSeconds_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void LayeredDeviceEmbedder::setCoordinates(const std::vector<std::vector<double>>&) {
    std::cout << "[WARNING] Setting coordinates for layered embedder has no effect" << std::endl;
}
void LayeredDeviceEmbedder::setWeights(const std::vector<double>&) {
    std::cout << "[WARNING] Setting weights for layered embedder has no effect" << std::endl;
}

std::vector<PhaseTiming> LayeredDeviceEmbedder::getTimings() {
    std::vector<PhaseTiming> out;
    out.push_back({0, "Embedding", totalSeconds_});
    out.push_back({1, "Expanding Positions", expandSeconds_});
    for (PhaseTiming t : finishedLayers_) out.push_back(t);
    for (PhaseTiming t : current_->getTimings()) { t.depth += 1; out.push_back(t); }
    return out;
}

// LayeredEmbedder::expandPositions (LayeredEmbedder.cpp:46-95): children start at their parent's position scaled by
// (newN / oldN)^(1/d) * expansionStretch.  The reference adds a random unit vector scaled by pow(totalContainedNodes, 1/d), but
// NodeInformation::totalContainedNodes is never assigned (GraphHierarchy.hpp:15 initialises it to 0 and nothing writes it), so
// the offset is exactly zero and siblings start coincident - the tie-break of the force kernels separates them.  The unit
// vectors are still drawn from the global generator so its stream advances like the reference's.
void LayeredDeviceEmbedder::expandPositions() {
    const auto t0 = std::chrono::steady_clock::now();
    const int d = opts_.embeddingDimension;
    const EmbeddingGraph& fine = hierarchy_.graphs[currentLayer_ - 1];
    const int newN = fine.getNumVertices(), oldN = hierarchy_.graphs[currentLayer_].getNumVertices();
    std::vector<double> old(static_cast<std::size_t>(oldN) * d);
    current_->copyCoordinatesTo(old.data());
    std::vector<double> w = opts_.useUnitWeights ? std::vector<double>(newN, 1.0)
                                                 : DeviceEmbedder::rescaleWeights(opts_.dimensionHint, d, DeviceEmbedder::degreeWeights(fine));
    const double stretch = std::pow(static_cast<double>(newN) / static_cast<double>(oldN), 1.0 / d) * opts_.expansionStretch;
    std::vector<std::vector<double>> x(newN, std::vector<double>(d));
    for (int v = 0; v < newN; ++v) {
        const int p = hierarchy_.parent[currentLayer_ - 1][v];
        for (int k = 0; k < d; ++k) {
            std::normal_distribution<double> dist(0.0, 1.0);   // setToRandomUnitVector (DVec.hpp:412-424): stream parity only
            (void)dist(GlobalRandom::generator());
        }
        for (int k = 0; k < d; ++k) x[v][k] = stretch * old[static_cast<std::size_t>(p) * d + k];
    }
    for (PhaseTiming t : current_->getTimings()) { t.depth += 1; finishedLayers_.push_back(t); }
    --currentLayer_;
    current_ = std::make_unique<DeviceEmbedder>(fine, opts_, /*initializeState=*/false);
    current_->setCoordinates(x);
    current_->setWeights(w);
    expandSeconds_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace impl
}  // namespace wembed

// Test hook (CPU only, no device needed): the parent pointers of every layer for a graph given as an edge list.
extern "C" int wbh_coarsen(long long m, const int* src, const int* dst, int* layerSizes, int maxLayers, int* parents, long long cap) {
    std::vector<std::pair<int, int>> edges;
    for (long long i = 0; i < m; ++i) edges.emplace_back(src[i], dst[i]);
    const wembed::impl::EmbeddingGraph g(edges);
    const auto tree = wembed::impl::coarsenAllLayers(g, std::vector<double>(g.col().size(), 1.0));
    long long at = 0;
    for (std::size_t l = 0; l < tree.size() && static_cast<int>(l) < maxLayers; ++l) {
        layerSizes[l] = static_cast<int>(tree[l].size());
        for (int p : tree[l])
            if (at < cap) parents[at++] = p;
    }
    return static_cast<int>(tree.size());
}
