#include "embedder.hpp"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>

namespace wembed {
namespace impl {

// ---- PhaseTimer ------------------------------------------------------------------------------------------------
std::size_t PhaseTimer::slot(const std::string& parent, const std::string& key, const std::string& displayName) {
    auto it = index_.find(key);
    if (it != index_.end()) return it->second;
    index_[key] = entries_.size();
    entries_.push_back({parent, key, displayName, 0.0});
    return entries_.size() - 1;
}

void PhaseTimer::start(const std::string& key, const std::string& displayName) {
    slot(running_.empty() ? std::string() : running_.back().first, key, displayName);
    running_.emplace_back(key, std::chrono::steady_clock::now());
}

void PhaseTimer::stop(const std::string& key) {
    if (running_.empty() || running_.back().first != key) return;
    entries_[index_[key]].seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - running_.back().second).count();
    running_.pop_back();
}

void PhaseTimer::add(const std::string& parent, const std::string& key, const std::string& displayName, double seconds) {
    entries_[slot(parent, key, displayName)].seconds += seconds;
}

void PhaseTimer::collect(std::size_t depth, const std::string& key, std::vector<PhaseTiming>& out) const {
    for (const Entry& e : entries_) {
        if (e.parent == key) {
            out.push_back({depth, e.displayName, e.seconds});
            collect(depth + 1, e.key, out);
        }
    }
}

std::vector<PhaseTiming> PhaseTimer::results() const {
    std::vector<PhaseTiming> out;
    collect(0, "", out);
    return out;
}

// ---- monitors and schedule ---------------------------------------------------------------------------------------
LossMonitor::LossMonitor(double relTol, int patience, double smoothing, int window)
    : relTol_(relTol), smoothing_(smoothing), patience_(patience), ring_(static_cast<std::size_t>(std::max(window, 1)) + 1, 0.0) {}

void LossMonitor::observe(double loss) {
    smoothed_ = observed_ == 0 ? loss : smoothing_ * loss + (1.0 - smoothing_) * smoothed_;
    ++observed_;
    ring_[head_] = smoothed_;
    head_ = (head_ + 1) % static_cast<int>(ring_.size());
    count_ = std::min(count_ + 1, static_cast<int>(ring_.size()));
    if (count_ >= static_cast<int>(ring_.size())) {
        const double start = ring_[head_];   // the oldest retained sample: Lbar(t - window)
        rate_ = (start - smoothed_) / std::max(std::abs(start), 1e-12);
    } else {
        rate_ = std::numeric_limits<double>::infinity();
    }
    stagnant_ = rate_ < relTol_ ? stagnant_ + 1 : 0;
}

double LearningRate::next(int iteration, const LossMonitor& monitor) {
    double lr;
    if (o_.lrSchedule == LRExponentialCooling) {
        lr = o_.learningRate * std::pow(o_.lrCoolingFactor, static_cast<double>(iteration));
    } else {
        const double r = monitor.rate();
        if (r > o_.lrGrowthThreshold) {
            decay_ = 0;
            if (++growth_ >= o_.lrAdaptPatience) { current_ *= o_.lrGrowthFactor; growth_ = 0; }
        } else if (r < o_.lrDecayThreshold) {
            growth_ = 0;
            if (++decay_ >= o_.lrAdaptPatience) { current_ *= o_.lrDecayFactor; decay_ = 0; }
        } else {
            growth_ = decay_ = 0;
        }
        lr = current_;
    }
    if (iteration < o_.warmupSteps) return lr * static_cast<double>(iteration) / static_cast<double>(o_.warmupSteps);
    return lr;
}

// ---- global generator --------------------------------------------------------------------------------------------
namespace {
struct RandState {
    std::uint32_t seed;
    std::mt19937 gen;
    RandState() : seed(std::random_device{}()), gen(seed) {}
};
RandState& randState() {
    static RandState s;
    return s;
}
}  // namespace
std::mt19937& GlobalRandom::generator() { return randState().gen; }
std::uint32_t GlobalRandom::seed() { return randState().seed; }
void GlobalRandom::setSeed(int seed) {
    randState().seed = static_cast<std::uint32_t>(seed);
    randState().gen = std::mt19937(static_cast<std::uint32_t>(seed));
}

// ---- DeviceEmbedder ----------------------------------------------------------------------------------------------
void DeviceEmbedder::check(int status, const char* what) const {
    if (status != WB_OK) throw std::runtime_error(std::string(what) + ": " + wb_last_error());
}

std::vector<double> DeviceEmbedder::degreeWeights(const EmbeddingGraph& g) {
    std::vector<double> w(g.getNumVertices());
    for (int v = 0; v < g.getNumVertices(); ++v) w[v] = g.getNumNeighbors(v) > 0 ? g.getNumNeighbors(v) : 1;
    return w;
}

std::vector<double> DeviceEmbedder::rescaleWeights(double dimensionHint, double dimension, std::vector<double> w) {
    if (dimensionHint > 0)
        for (double& e : w) e = std::pow(e, dimension / dimensionHint);
    double sum = 0.0;
    for (double e : w) sum += e;
    const double n = static_cast<double>(w.size());
    for (double& e : w) e = e * (n / sum);
    return w;
}

DeviceEmbedder::DeviceEmbedder(const EmbeddingGraph& graph, const Options& options, bool initializeState)
    : graph_(graph), opts_(options), lastLearningRate_(options.learningRate),
      lossMonitor_(options.stopLossTol, options.stopLossPatience, options.lossSmoothingFactor, options.lossRateWindow),
      moveMonitor_(options.stopDisplacementTol, options.stopDisplacementPatience), schedule_(options) {
    wb_options o;
    wb_options_default(&o);
    o.embedding_dimension = options.embeddingDimension;
    o.optimizer = options.optimizerType == OptimizerSimple ? WB_OPT_SIMPLE : WB_OPT_ADAM;
    o.attraction_scale = options.attractionScale;
    o.repulsion_scale = options.repulsionScale;
    o.centre_scale = options.centreScale;
    o.edge_length = options.edgeLength;
    o.simple_max_displacement = options.simpleOptMaxDisplacement;
    o.seed = GlobalRandom::seed();   // Rand::localGenerator keys its streams by the base seed (Rand.cpp:29-35)
    static const std::int32_t none[1] = {0};
    const std::int32_t* col = graph_.col().empty() ? none : graph_.col().data();
    check(wb_create(&handle_, graph_.getNumVertices(), graph_.rowPtr().data(), col, &o), "wb_create");
    wb_enable_timing(handle_, 1);
    if (!initializeState) return;   // the caller assigns coordinates and weights (WembedEmbedder.hpp:109-111)

    // EmbedderInterface::constructRandomCoordinates (EmbedderInterface.hpp:61-65): cube side pow((float)n, 1/d),
    // Rand::randomCoordinates draw order = vertex-major (Rand.cpp:101-109)
    const int n = graph_.getNumVertices(), d = opts_.embeddingDimension;
    const double side = std::pow(static_cast<float>(n), 1.0 / d);
    std::vector<double> x(static_cast<std::size_t>(n) * d);
    for (double& e : x) {
        std::uniform_real_distribution<double> dist(0.0, side);
        e = dist(GlobalRandom::generator());
    }
    check(wb_set_coordinates(handle_, x.data()), "wb_set_coordinates");
    if (opts_.useUnitWeights) {
        setWeights(std::vector<double>(n, 1.0));
    } else {
        setWeights(rescaleWeights(opts_.dimensionHint, d, degreeWeights(graph_)));
    }
}

DeviceEmbedder::~DeviceEmbedder() { wb_destroy(handle_); }

void DeviceEmbedder::calculateStep() {
    ++iteration_;
    wb_set_iteration(handle_, iteration_ - 1);
    const double lr = graph_.getNumVertices() <= 1 ? lastLearningRate_ : schedule_.next(static_cast<int>(iteration_), lossMonitor_);
    wb_step_stats st;
    check(wb_step(handle_, lr, &st), "wb_step");
    lossAttract_ = st.loss_attract;
    lossRepel_ = st.loss_repel;
    if (graph_.getNumVertices() <= 1) return;   // WembedEmbedder.cpp:19-21
    lastLearningRate_ = lr;
    lastRelDisplacement_ = st.rel_displacement;
    moveMonitor_.observe(st.rel_displacement);
    lossMonitor_.observe(st.loss_attract + st.loss_repel);
    lastRelLossImprovement_ = lossMonitor_.rate();
    double ms[6];
    if (wb_get_phase_times(handle_, ms) == WB_OK) {
        // the reference's timer keys (WembedEmbedder.cpp:28-58), fed with device times
        const std::string parent = timer_.runningKey();   // "embedding_all" inside calculateEmbedding, top level otherwise
        timer_.add(parent, "index", "Construct spacial index", ms[0] * 1e-3);
        timer_.add(parent, "attracting_forces", "Compute Attracting Forces + Applying Forces (fused)", ms[1] * 1e-3);
        timer_.add(parent, "repelling_forces", "Compute Repelling Forces", ms[2] * 1e-3);
        timer_.add(parent, "gravity", "Move graph towards centre", ms[4] * 1e-3);
    }
}

bool DeviceEmbedder::isFinished() {
    if (iteration_ >= opts_.maxIterations) return true;
    if (graph_.getNumVertices() <= 1) return true;
    return opts_.stopCriterion == StopDisplacement ? moveMonitor_.converged() : lossMonitor_.converged();
}

void DeviceEmbedder::calculateEmbedding() {
    timer_.start("embedding_all", "Embedding");
    iteration_ = 0;   // WembedEmbedder.cpp:80 - the optimizer state and the monitors are not reset
    while (!isFinished()) calculateStep();
    timer_.stop("embedding_all");
}

std::vector<std::vector<double>> DeviceEmbedder::getCoordinates() {
    const int n = graph_.getNumVertices(), d = opts_.embeddingDimension;
    std::vector<double> flat(static_cast<std::size_t>(n) * d);
    copyCoordinatesTo(flat.data());
    std::vector<std::vector<double>> out(n, std::vector<double>(d));
    for (int v = 0; v < n; ++v) std::copy(flat.begin() + static_cast<std::size_t>(v) * d, flat.begin() + static_cast<std::size_t>(v + 1) * d, out[v].begin());
    return out;
}

void DeviceEmbedder::copyCoordinatesTo(double* out) const { check(wb_get_coordinates(handle_, out), "wb_get_coordinates"); }

std::vector<double> DeviceEmbedder::getWeights() {
    std::vector<double> w(graph_.getNumVertices());
    check(wb_get_weights(handle_, w.data()), "wb_get_weights");
    return w;
}

void DeviceEmbedder::setCoordinates(const std::vector<std::vector<double>>& coordinates) {
    const int n = graph_.getNumVertices(), d = opts_.embeddingDimension;
    if (static_cast<int>(coordinates.size()) != n) throw std::invalid_argument("setCoordinates: one row per vertex expected");
    const int coordDim = coordinates.empty() ? 0 : static_cast<int>(coordinates[0].size());
    if (coordDim != d)   // the reference only warns and copies min(d, coordDim) columns (WembedEmbedder.cpp:108-118)
        std::cout << "[WARNING] Dimension of coordinates (" << coordDim << ") does not match embedding dimension (" << d << ")" << std::endl;
    std::vector<double> flat(static_cast<std::size_t>(n) * d);
    if (coordDim < d) copyCoordinatesTo(flat.data());
    for (int v = 0; v < n; ++v)
        for (int k = 0; k < std::min(d, coordDim); ++k) flat[static_cast<std::size_t>(v) * d + k] = coordinates[v][k];
    check(wb_set_coordinates(handle_, flat.data()), "wb_set_coordinates");
}

void DeviceEmbedder::setWeights(const std::vector<double>& weights) {
    if (static_cast<int>(weights.size()) != graph_.getNumVertices()) throw std::invalid_argument("setWeights: one weight per vertex expected");
    check(wb_set_weights(handle_, weights.data()), "wb_set_weights");
}

}  // namespace impl
}  // namespace wembed

// Test hook (CPU only): the host scalar logic of one embedding run - learning-rate schedule, loss monitor, displacement monitor,
// stop decision - replayed over a given sequence of per-step losses and relative displacements, exactly in the order
// DeviceEmbedder::calculateStep uses them.  o = {lrSchedule, learningRate, warmupSteps, lrCoolingFactor, lrDecayFactor,
// lrDecayThreshold, lrAdaptPatience, lrGrowthFactor, lrGrowthThreshold, stopCriterion, stopDisplacementTol,
// stopDisplacementPatience, lossSmoothingFactor, lossRateWindow, stopLossTol, stopLossPatience, maxIterations}.
// Returns the first iteration after which isFinished() holds (0 if never within `steps`).
extern "C" int wbh_host_logic_trace(const double* o, int steps, const double* loss, const double* relDisp, double* outLr, double* outRate) {
    using namespace wembed;
    Options opt;
    opt.lrSchedule = o[0] == 0.0 ? LRExponentialCooling : LRLossAdaptive;
    opt.learningRate = o[1]; opt.warmupSteps = (int)o[2]; opt.lrCoolingFactor = o[3]; opt.lrDecayFactor = o[4];
    opt.lrDecayThreshold = o[5]; opt.lrAdaptPatience = (int)o[6]; opt.lrGrowthFactor = o[7]; opt.lrGrowthThreshold = o[8];
    opt.stopCriterion = o[9] == 0.0 ? StopDisplacement : StopLoss;
    opt.stopDisplacementTol = o[10]; opt.stopDisplacementPatience = (int)o[11]; opt.lossSmoothingFactor = o[12];
    opt.lossRateWindow = (int)o[13]; opt.stopLossTol = o[14]; opt.stopLossPatience = (int)o[15]; opt.maxIterations = (int)o[16];
    impl::LossMonitor lossMon(opt.stopLossTol, opt.stopLossPatience, opt.lossSmoothingFactor, opt.lossRateWindow);
    impl::MoveMonitor moveMon(opt.stopDisplacementTol, opt.stopDisplacementPatience);
    impl::LearningRate schedule(opt);
    int finishedAt = 0;
    for (int it = 1; it <= steps; ++it) {
        outLr[it - 1] = schedule.next(it, lossMon);
        moveMon.observe(relDisp[it - 1]);
        lossMon.observe(loss[it - 1]);
        outRate[it - 1] = lossMon.rate();
        const bool finished = it >= opt.maxIterations || (opt.stopCriterion == StopDisplacement ? moveMon.converged() : lossMon.converged());
        if (finished && finishedAt == 0) finishedAt = it;
    }
    return finishedAt;
}
