// Implementation of include/wembed.h over the device embedder (the role of the reference's src/wembed.cpp).
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>

#include "embedder.hpp"
#include "graph.hpp"
#include "hierarchy.hpp"
#include "wembed.h"

namespace wembed {

namespace {
// util::splitIntoTokens (src/utilLib/src/StringManipulation.cpp:44-60): split at every delimiter; a trailing
// delimiter yields a trailing empty token.
std::vector<std::string> tokens(std::string line, const std::string& delimiter) {
    std::vector<std::string> out;
    while (!line.empty()) {
        const std::size_t at = line.find(delimiter);
        if (at == std::string::npos) {
            out.push_back(line);
            break;
        }
        out.push_back(line.substr(0, at));
        line = line.substr(at + delimiter.size());
        if (line.empty()) out.push_back(line);
    }
    return out;
}
}  // namespace

// ---- Graph -------------------------------------------------------------------------------------------------------
Graph::Graph(std::unique_ptr<impl::EmbeddingGraph>&& graph) : _graph(std::move(graph)) {}
Graph::~Graph() = default;
Graph::Graph(Graph&& other) = default;
Graph& Graph::operator=(Graph&& other) = default;
NodeId Graph::getNumVertices() const { return _graph->getNumVertices(); }
EdgeId Graph::getNumEdges() const { return _graph->getNumEdges(); }
std::vector<EdgeId> Graph::getEdges(NodeId v) const { return _graph->getEdges(v); }
std::vector<NodeId> Graph::getNeighbors(NodeId v) const { return _graph->getNeighbors(v); }
int Graph::getNumNeighbors(NodeId v) const { return _graph->getNumNeighbors(v); }
NodeId Graph::getEdgeTarget(EdgeId e) const { return _graph->getEdgeTarget(e); }
bool Graph::areNeighbors(NodeId v, NodeId u) const { return _graph->areNeighbors(v, u); }
std::string Graph::toString() const { return _graph->toString(); }
const std::int32_t* Graph::csrOffsets() const { return _graph->rowPtr().data(); }
const std::int32_t* Graph::csrTargets() const { return _graph->col().data(); }

std::vector<Edge> Graph::getEdgeList() const {
    std::vector<Edge> out;
    out.reserve(_graph->getNumEdges());
    for (NodeId v = 0; v < _graph->getNumVertices(); ++v)
        for (NodeId u : _graph->getNeighbors(v))
            if (v < u) out.push_back({v, u});
    return out;
}

// ---- Embedder ----------------------------------------------------------------------------------------------------
Embedder::Embedder(std::unique_ptr<impl::EmbedderInterface>&& embedder) : _embedder(std::move(embedder)) {}
Embedder::~Embedder() = default;
Embedder::Embedder(Embedder&& other) = default;
Embedder& Embedder::operator=(Embedder&& other) = default;
void Embedder::calculateStep() { _embedder->calculateStep(); }
bool Embedder::isFinished() const { return _embedder->isFinished(); }
void Embedder::calculateEmbedding() { _embedder->calculateEmbedding(); }
std::int32_t Embedder::getNumVertices() const { return _embedder->getNumVertices(); }
std::int32_t Embedder::getEmbeddingDimension() const { return _embedder->getEmbeddingDimension(); }
void Embedder::copyCoordinatesTo(double* out) const { _embedder->copyCoordinatesTo(out); }
Graph Embedder::getCurrentGraph() const { return Graph(std::make_unique<impl::EmbeddingGraph>(_embedder->getCurrentGraph())); }
std::vector<std::vector<double>> Embedder::getCoordinates() const { return _embedder->getCoordinates(); }
std::vector<double> Embedder::getWeights() const { return _embedder->getWeights(); }
void Embedder::setCoordinates(const std::vector<std::vector<double>>& coordinates) { _embedder->setCoordinates(coordinates); }
void Embedder::setWeights(const std::vector<double>& weights) { _embedder->setWeights(weights); }
double Embedder::getCurrentLearningRate() const { return _embedder->getCurrentLearningRate(); }
double Embedder::getLastRelDisplacement() const { return _embedder->getLastRelDisplacement(); }
double Embedder::getLastRelLossImprovement() const { return _embedder->getLastRelLossImprovement(); }

std::vector<TimingResult> Embedder::getTimings() const {
    std::vector<TimingResult> out;
    for (const auto& t : _embedder->getTimings()) out.push_back({static_cast<std::uint64_t>(t.depth), t.displayName, t.seconds});
    return out;
}

Loss Embedder::getLoss() const {
    const impl::EmbeddingLoss l = _embedder->getLoss();
    return {l.attractive, l.repulsive, l.total};
}

// EmbeddingIO::writeCoordinates (src/embeddingLib/src/embeddingIO/EmbeddingIO.cpp:194-222): "id,x1,...,xd[,w]", 16 significant digits
void Embedder::writeCoordinates(const std::string& filePath, bool writeWeights) const {
    const auto x = _embedder->getCoordinates();
    const auto w = writeWeights ? _embedder->getWeights() : std::vector<double>();
    std::ofstream out(filePath);
    out << std::setprecision(std::numeric_limits<double>::digits10 + 1);
    for (std::size_t i = 0; i < x.size(); ++i) {
        out << i;
        for (double e : x[i]) out << "," << e;
        if (writeWeights) out << "," << w[i];
        out << "\n";
    }
}

// ---- free functions ------------------------------------------------------------------------------------------------
Embedder createEmbedder(const Graph& g, const Options& options) {
    if (options.layeredEmbedding)   // multilevel driver: coarsen with label propagation, embed layer by layer (wembed.cpp:229-233)
        return Embedder(std::make_unique<impl::LayeredDeviceEmbedder>(*g._graph, options));
    return Embedder(std::make_unique<impl::DeviceEmbedder>(*g._graph, options));
}

Graph graphFromEdges(const std::vector<Edge>& edges) {
    std::vector<std::pair<int, int>> pairs;
    pairs.reserve(edges.size());
    for (const Edge& e : edges) pairs.emplace_back(e.src, e.dst);
    return Graph(std::make_unique<impl::EmbeddingGraph>(pairs));
}

// GraphIO::readEdgeList (src/graphLib/src/graphIO/GraphIO.cpp:10-51): comment-prefixed lines skipped, the first two tokens are ids.
Graph graphFromEdgeListFile(const std::string& filePath, const std::string& comment, const std::string& delimiter) {
    std::ifstream in(filePath);
    if (!in.good()) throw std::runtime_error("Could not find file: " + filePath);   // the reference aborts the process here
    std::vector<std::pair<int, int>> pairs;
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind(comment, 0) == 0) continue;
        const auto t = tokens(line, delimiter);
        if (t.size() < 2) continue;
        try {
            pairs.emplace_back(std::stoi(t[0]), std::stoi(t[1]));
        } catch (const std::exception&) {
            continue;
        }
    }
    return Graph(std::make_unique<impl::EmbeddingGraph>(pairs));
}

// EmbeddingIO::readCoordinatesFromFile (EmbeddingIO.cpp:110-157): "id,c1,...,ck" rows, ids consecutive from 0
std::vector<std::vector<double>> readCoordinatesFromFile(const std::string& filePath, const std::string& comment, const std::string& delimiter) {
    std::ifstream in(filePath);
    if (!in.good()) throw std::runtime_error("Error while reading file: " + filePath);
    std::map<int, std::vector<double>> rows;
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind(comment, 0) == 0) continue;
        const auto t = tokens(line, delimiter);
        if (t.empty()) continue;
        std::vector<double> c(t.size() - 1);
        for (std::size_t i = 1; i < t.size(); ++i) c[i - 1] = std::stod(t[i]);
        rows[std::stoi(t[0])] = std::move(c);
    }
    std::vector<std::vector<double>> out;
    out.reserve(rows.size());
    for (auto& kv : rows) out.push_back(std::move(kv.second));
    return out;
}

// util::timingsToStringRepresentation (src/utilLib/src/Timings.cpp:64-78)
std::string timingsToString(const std::vector<TimingResult>& timings) {
    std::ostringstream out;
    for (const TimingResult& t : timings) {
        for (std::uint64_t i = 0; i < t.depth; ++i) out << "   ";
        std::ostringstream number;
        number << std::setprecision(4) << t.value << "s";
        out << "+- " << std::left << std::setw(15) << number.str() << t.displayName << std::endl;
    }
    return out.str();
}

void setSeed(int seed) { impl::GlobalRandom::setSeed(seed); }

}  // namespace wembed
