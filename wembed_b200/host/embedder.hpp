// Host side of the drop-in boundary: the reference's EmbedderInterface surface
// (src/embeddingLib/include/embedder/EmbedderInterface.hpp:24-175) and the scalar logic that stays on the host -
// learning-rate schedules, loss / displacement monitors, the phase timer, the global generator - layered over the
// device step of include/wembed_b200.h.
#pragma once
#include <chrono>
#include <cstdint>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "graph.hpp"
#include "wembed.h"
#include "wembed_b200.h"

namespace wembed {
namespace impl {

struct EmbeddingLoss {
    double attractive, repulsive, total;
};

struct PhaseTiming {
    std::size_t depth;
    std::string displayName;
    double seconds;
};

// util::Timer (src/utilLib/src/Timings.cpp:9-62): a tree of accumulated seconds keyed by phase; here the leaves of a
// step are fed with CUDA-event times instead of wall clock.
class PhaseTimer {
   public:
    void start(const std::string& key, const std::string& displayName);
    void stop(const std::string& key);
    void add(const std::string& parent, const std::string& key, const std::string& displayName, double seconds);
    std::vector<PhaseTiming> results() const;
    std::string runningKey() const { return running_.empty() ? std::string() : running_.back().first; }

   private:
    struct Entry {
        std::string parent, key, displayName;
        double seconds;
    };
    std::size_t slot(const std::string& parent, const std::string& key, const std::string& displayName);
    void collect(std::size_t depth, const std::string& key, std::vector<PhaseTiming>& out) const;
    std::vector<std::pair<std::string, std::chrono::steady_clock::time_point>> running_;
    std::vector<Entry> entries_;
    std::unordered_map<std::string, std::size_t> index_;
};

// ConvergenceMonitor (src/embeddingLib/src/embedder/ConvergenceMonitor.cpp:6-42)
class LossMonitor {
   public:
    LossMonitor(double relTol, int patience, double smoothing, int window);
    void observe(double loss);
    bool converged() const { return stagnant_ >= patience_; }
    double rate() const { return rate_; }

   private:
    double relTol_, smoothing_;
    int patience_;
    std::vector<double> ring_;
    int head_ = 0, count_ = 0, observed_ = 0, stagnant_ = 0;
    double smoothed_ = 0.0;
    double rate_ = std::numeric_limits<double>::infinity();
};

// DisplacementMonitor (src/embeddingLib/src/embedder/DisplacementMonitor.cpp:5-14)
class MoveMonitor {
   public:
    MoveMonitor(double relTol, int patience) : relTol_(relTol), patience_(patience) {}
    void observe(double rel) { settled_ = rel < relTol_ ? settled_ + 1 : 0; }
    bool converged() const { return settled_ >= patience_; }

   private:
    double relTol_;
    int patience_, settled_ = 0;
};

// LRScheduler and its two schedules (src/embeddingLib/src/gradientOptimizer/LRScheduler.cpp:7-39)
class LearningRate {
   public:
    explicit LearningRate(const Options& o) : o_(o), current_(o.learningRate) {}
    double next(int iteration, const LossMonitor& monitor);   // call exactly once per step (LossAdaptive has state)

   private:
    Options o_;
    double current_;
    int growth_ = 0, decay_ = 0;
};

// Rand (src/utilLib/src/Rand.cpp:6-25): process-wide mt19937, seeded from random_device until setSeed is called.
struct GlobalRandom {
    static std::mt19937& generator();
    static std::uint32_t seed();
    static void setSeed(int seed);
};

class EmbedderInterface {
   public:
    virtual ~EmbedderInterface() = default;
    virtual void calculateStep() = 0;
    virtual bool isFinished() = 0;
    virtual void calculateEmbedding() = 0;
    virtual EmbeddingGraph getCurrentGraph() = 0;
    virtual std::vector<std::vector<double>> getCoordinates() = 0;
    virtual std::vector<double> getWeights() = 0;
    virtual std::vector<PhaseTiming> getTimings() = 0;
    virtual void setCoordinates(const std::vector<std::vector<double>>& coordinates) = 0;
    virtual void setWeights(const std::vector<double>& weights) = 0;
    virtual int getNumVertices() const = 0;
    virtual int getEmbeddingDimension() const = 0;
    virtual void copyCoordinatesTo(double* out) const = 0;
    virtual EmbeddingLoss getLoss() const = 0;
    virtual double getCurrentLearningRate() const = 0;
    virtual double getLastRelDisplacement() const = 0;
    virtual double getLastRelLossImprovement() const = 0;
};

// The device embedder: WembedEmbedder's role (src/embeddingLib/include/embedder/WembedEmbedder.hpp:16-144).
class DeviceEmbedder final : public EmbedderInterface {
   public:
    DeviceEmbedder(const EmbeddingGraph& graph, const Options& options, bool initializeState = true);
    ~DeviceEmbedder() override;
    DeviceEmbedder(const DeviceEmbedder&) = delete;
    DeviceEmbedder& operator=(const DeviceEmbedder&) = delete;

    void calculateStep() override;
    bool isFinished() override;
    void calculateEmbedding() override;
    EmbeddingGraph getCurrentGraph() override { return graph_; }
    std::vector<std::vector<double>> getCoordinates() override;
    std::vector<double> getWeights() override;
    std::vector<PhaseTiming> getTimings() override { return timer_.results(); }
    void setCoordinates(const std::vector<std::vector<double>>& coordinates) override;
    void setWeights(const std::vector<double>& weights) override;
    int getNumVertices() const override { return graph_.getNumVertices(); }
    int getEmbeddingDimension() const override { return opts_.embeddingDimension; }
    void copyCoordinatesTo(double* out) const override;
    EmbeddingLoss getLoss() const override { return {lossAttract_, lossRepel_, lossAttract_ + lossRepel_}; }
    double getCurrentLearningRate() const override { return lastLearningRate_; }
    double getLastRelDisplacement() const override { return lastRelDisplacement_; }
    double getLastRelLossImprovement() const override { return lastRelLossImprovement_; }

    static std::vector<double> degreeWeights(const EmbeddingGraph& g);                                 // WembedEmbedder.cpp:381-388
    static std::vector<double> rescaleWeights(double dimensionHint, double dimension, std::vector<double> w);  // :359-379

   private:
    void check(int status, const char* what) const;
    EmbeddingGraph graph_;
    Options opts_;
    wb_embedder* handle_ = nullptr;
    std::int64_t iteration_ = 0;
    double lossAttract_ = 0.0, lossRepel_ = 0.0, lastLearningRate_ = 0.0, lastRelDisplacement_ = 0.0, lastRelLossImprovement_ = 0.0;
    LossMonitor lossMonitor_;
    MoveMonitor moveMonitor_;
    LearningRate schedule_;
    PhaseTimer timer_;
};

}  // namespace impl
}  // namespace wembed
