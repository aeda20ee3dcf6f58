// pybind11 module "wembed": the same Python surface as the reference's python/bindings.cpp:11-133
// (enums exported by value; Edge, TimingResult, Loss, Options, Graph, Embedder; the six free functions; __version__).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include "wembed.h"

namespace py = pybind11;
using namespace wembed;

#define WB_OPTION_FIELDS(X)                                                                                               \
    X(embeddingDimension) X(useUnitWeights) X(dimensionHint) X(layeredEmbedding) X(indexType) X(attractionScale)         \
    X(repulsionScale) X(centreScale) X(edgeLength) X(expansionStretch) X(optimizerType) X(maxIterations)                 \
    X(simpleOptMaxDisplacement) X(lrSchedule) X(learningRate) X(warmupSteps) X(lrCoolingFactor) X(lrDecayFactor)         \
    X(lrDecayThreshold) X(lrAdaptPatience) X(lrGrowthFactor) X(lrGrowthThreshold) X(stopCriterion) X(stopDisplacementTol) \
    X(stopDisplacementPatience) X(lossSmoothingFactor) X(lossRateWindow) X(stopLossTol) X(stopLossPatience)

PYBIND11_MODULE(wembed, m) {
    m.doc() = "WEmbed: weighted low-dimensional graph embeddings (B200-native build)";

    py::enum_<SpatialIndex>(m, "SpatialIndex").value("IndexSNN", IndexSNN).value("IndexSprk", IndexSprk).export_values();
    py::enum_<OptimizerType>(m, "OptimizerType").value("OptimizerSimple", OptimizerSimple).value("OptimizerAdam", OptimizerAdam).export_values();
    py::enum_<LRSchedule>(m, "LRSchedule").value("LRExponentialCooling", LRExponentialCooling).value("LRLossAdaptive", LRLossAdaptive).export_values();
    py::enum_<StopCriterion>(m, "StopCriterion").value("StopDisplacement", StopDisplacement).value("StopLoss", StopLoss).export_values();

    py::class_<Edge>(m, "Edge")
        .def(py::init<NodeId, NodeId>(), py::arg("src"), py::arg("dst"))
        .def_readwrite("src", &Edge::src)
        .def_readwrite("dst", &Edge::dst)
        .def("__repr__", [](const Edge& e) { return "Edge(" + std::to_string(e.src) + ", " + std::to_string(e.dst) + ")"; });

    py::class_<TimingResult>(m, "TimingResult")
        .def_readonly("depth", &TimingResult::depth)
        .def_readonly("display_name", &TimingResult::displayName)
        .def_readonly("value", &TimingResult::value);

    py::class_<Loss>(m, "Loss")
        .def_readonly("attractive", &Loss::attractive)
        .def_readonly("repulsive", &Loss::repulsive)
        .def_readonly("total", &Loss::total)
        .def("__repr__", [](const Loss& l) {
            return "Loss(attractive=" + std::to_string(l.attractive) + ", repulsive=" + std::to_string(l.repulsive) +
                   ", total=" + std::to_string(l.total) + ")";
        });

    py::class_<Options> options(m, "Options");
    options.def(py::init<>());
#define X(field) options.def_readwrite(#field, &Options::field);
    WB_OPTION_FIELDS(X)
#undef X

    py::class_<Graph>(m, "Graph")
        .def("getNumVertices", &Graph::getNumVertices)
        .def("getNumEdges", &Graph::getNumEdges)
        .def("getEdges", &Graph::getEdges)
        .def("getNeighbors", &Graph::getNeighbors)
        .def("getNumNeighbors", &Graph::getNumNeighbors)
        .def("getEdgeTarget", &Graph::getEdgeTarget)
        .def("areNeighbors", &Graph::areNeighbors)
        .def("getEdgeList", &Graph::getEdgeList)
        // addition of this build: the CSR as two numpy arrays (copies), what the embedder uploads to the device
        .def("csr", [](const Graph& g) {
            const py::ssize_t n = g.getNumVertices(), m2 = 2 * (py::ssize_t)g.getNumEdges();
            py::array_t<std::int32_t> offsets(n + 1), targets(m2);
            std::memcpy(offsets.mutable_data(), g.csrOffsets(), sizeof(std::int32_t) * (n + 1));
            if (m2) std::memcpy(targets.mutable_data(), g.csrTargets(), sizeof(std::int32_t) * m2);
            return py::make_tuple(offsets, targets);
        })
        .def("__repr__", &Graph::toString);

    py::class_<Embedder>(m, "Embedder")
        .def("calculateStep", &Embedder::calculateStep)
        .def("isFinished", &Embedder::isFinished)
        .def("calculateEmbedding", &Embedder::calculateEmbedding, py::call_guard<py::gil_scoped_release>())
        .def("getNumVertices", &Embedder::getNumVertices)
        .def("getEmbeddingDimension", &Embedder::getEmbeddingDimension)
        .def("getCurrentGraph", &Embedder::getCurrentGraph)
        .def("getCoordinates", &Embedder::getCoordinates)
        .def("getWeights", &Embedder::getWeights)
        .def("setCoordinates", &Embedder::setCoordinates)
        .def("setWeights", &Embedder::setWeights)
        .def("getTimings", &Embedder::getTimings)
        .def("getLoss", &Embedder::getLoss)
        .def("getCurrentLearningRate", &Embedder::getCurrentLearningRate)
        .def("getLastRelDisplacement", &Embedder::getLastRelDisplacement)
        .def("getLastRelLossImprovement", &Embedder::getLastRelLossImprovement)
        .def("writeCoordinates", &Embedder::writeCoordinates, py::arg("filePath"), py::arg("writeWeights") = true);

    m.def("createEmbedder", &createEmbedder, py::arg("graph"), py::arg("options"));
    m.def("graphFromEdges", &graphFromEdges, py::arg("edges"));
    // addition of this build: the same constructor fed from an (m, 2) integer array instead of a list of Edge objects
    m.def("graphFromEdgeArray", [](py::array_t<std::int32_t, py::array::c_style | py::array::forcecast> a) {
        if (a.ndim() != 2 || a.shape(1) != 2) throw std::invalid_argument("graphFromEdgeArray: expected an array of shape (m, 2)");
        std::vector<Edge> edges((std::size_t)a.shape(0));
        const std::int32_t* p = a.data();
        for (std::size_t i = 0; i < edges.size(); ++i) edges[i] = Edge{p[2 * i], p[2 * i + 1]};
        py::gil_scoped_release release;
        return graphFromEdges(edges);
    }, py::arg("edges"));
    m.def("graphFromEdgeListFile", &graphFromEdgeListFile, py::arg("filePath"), py::arg("comment") = "#", py::arg("delimiter") = " ");
    m.def("readCoordinatesFromFile", &readCoordinatesFromFile, py::arg("filePath"), py::arg("comment") = "%", py::arg("delimiter") = ",");
    m.def("timingsToString", &timingsToString, py::arg("timings"));
    m.def("setSeed", &setSeed, py::arg("seed"));
    m.attr("__version__") = "b200-dev";
}
