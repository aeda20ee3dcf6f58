// TEST INFRASTRUCTURE - CPU restatement ("port") of the reference's gradient-descent step.
//
// This file is the readable, independent re-statement of WembedEmbedder::calculateStep and
// the loop around it, in fp64 like the reference.  It is pinned against the reference's own
// sources (oracle/_ref/libwembed_ref.so, built by oracle/Makefile) by tests/test_oracle_*.py
// and against the golden vectors under tests/golden/ that were generated from that build.
// It exists so that (a) the GPU parity tests have a checker on the GPU box, where
// /root/reference does not exist, and (b) bench.py has a CPU baseline at sizes where the
// reference's SNN index (O(n^(1-1/d)) distance evaluations per query) cannot finish.
//
// The product never links, loads or calls this file.
//
// All citations are relative to the reference checkout (Vraier/wembed).  The arithmetic of
// every force / optimizer / monitor function follows the cited lines; the one free choice is
// the spatial index: the reference re-tests every candidate exactly (WembedEmbedder.cpp:
// 196-201), so any index returning a superset of the in-radius points gives the same
// forces up to summation order.  The port uses a Morton-sorted bounding-box hierarchy and
// visits candidates in ascending vertex id, which makes its result independent of the index.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

#include "oracle_api.h"

namespace {

// ---------------------------------------------------------------------------------------------
// util::deterministicSum (src/utilLib/include/ParallelReduce.hpp:18-37): blocks of 4096 summed
// sequentially, then the block sums summed sequentially.
template <typename F>
double blockSum(std::size_t n, F&& element, std::size_t blockSize = 4096) {
    if (n == 0) return 0.0;
    const std::size_t numBlocks = (n + blockSize - 1) / blockSize;
    std::vector<double> partial(numBlocks, 0.0);
#pragma omp parallel for schedule(static)
    for (std::size_t b = 0; b < numBlocks; b++) {
        const std::size_t end = std::min((b + 1) * blockSize, n);
        double s = 0.0;
        for (std::size_t i = b * blockSize; i < end; i++) s += element(i);
        partial[b] = s;
    }
    double total = 0.0;
    for (double p : partial) total += p;
    return total;
}

// ConvergenceMonitor (src/embeddingLib/src/embedder/ConvergenceMonitor.cpp:6-42)
struct LossMonitor {
    double relTol, alpha;
    int patience, window;
    std::vector<double> ring;
    int head = 0, count = 0, observed = 0, stagnant = 0;
    double smoothed = 0.0;
    double rate = std::numeric_limits<double>::infinity();
    LossMonitor(double tol, int pat, double a, int win)
        : relTol(tol), alpha(a), patience(pat), window(win < 1 ? 1 : win), ring(static_cast<std::size_t>(window) + 1, 0.0) {}
    void observe(double loss) {
        smoothed = observed == 0 ? loss : alpha * loss + (1.0 - alpha) * smoothed;
        observed++;
        ring[head] = smoothed;
        head = (head + 1) % static_cast<int>(ring.size());
        if (count < static_cast<int>(ring.size())) count++;
        if (count >= static_cast<int>(ring.size())) {
            const double start = ring[head];
            rate = (start - smoothed) / std::max(std::abs(start), 1e-12);
        } else {
            rate = std::numeric_limits<double>::infinity();
        }
        stagnant = rate < relTol ? stagnant + 1 : 0;
    }
    bool converged() const { return stagnant >= patience; }
};

// DisplacementMonitor (src/embeddingLib/src/embedder/DisplacementMonitor.cpp:5-14)
struct MoveMonitor {
    double relTol;
    int patience, settled = 0;
    void observe(double rel) { settled = rel < relTol ? settled + 1 : 0; }
    bool converged() const { return settled >= patience; }
};

// LRScheduler (src/embeddingLib/src/gradientOptimizer/LRScheduler.cpp:7-39)
struct Schedule {
    orc_options o;
    double current;
    int growth = 0, decay = 0;
    explicit Schedule(const orc_options& opts) : o(opts), current(opts.learningRate) {}
    double rate(int iteration, const LossMonitor& mon) {
        double lr;
        if (o.lrScheduleType == 0) {
            lr = o.learningRate * std::pow(o.lrCoolingFactor, static_cast<double>(iteration));
        } else {
            const double r = mon.rate;
            if (r > o.lrGrowthThreshold) {
                decay = 0;
                if (++growth >= o.lrAdaptPatience) { current *= o.lrGrowthFactor; growth = 0; }
            } else if (r < o.lrDecayThreshold) {
                growth = 0;
                if (++decay >= o.lrAdaptPatience) { current *= o.lrDecayFactor; decay = 0; }
            } else {
                growth = decay = 0;
            }
            lr = current;
        }
        if (iteration < o.warmupSteps) return lr * static_cast<double>(iteration) / static_cast<double>(o.warmupSteps);
        return lr;
    }
};

// ---------------------------------------------------------------------------------------------
// Morton-sorted box hierarchy over the current positions (the port's free index choice).
struct BoxTree {
    static constexpr int LEAF = 8, FAN = 8;
    int n = 0, d = 0;
    std::vector<int> order;                    // sorted position -> vertex id
    std::vector<double> pts;                   // n x d, in sorted order
    std::vector<double> bound;                 // per sorted point: weight bound used for pruning
    struct Level { int count; std::vector<double> lo, hi, wmax; };
    std::vector<Level> levels;                 // levels[0] = leaves (LEAF points each)

    void build(int n_, int d_, const double* x, const double* weightBound) {
        n = n_; d = d_;
        std::vector<double> mean(d, 0.0), sd(d, 0.0), lo(d), hi(d);
        for (int k = 0; k < d; k++) {
            double s = 0, s2 = 0, mn = 1e300, mx = -1e300;
            for (int v = 0; v < n; v++) { const double e = x[(size_t)v * d + k]; s += e; s2 += e * e; mn = std::min(mn, e); mx = std::max(mx, e); }
            mean[k] = s / n; sd[k] = std::sqrt(std::max(0.0, s2 / n - mean[k] * mean[k]));
            lo[k] = std::max(mn, mean[k] - 4 * sd[k]); hi[k] = std::min(mx, mean[k] + 4 * sd[k]);
            if (!(hi[k] > lo[k])) hi[k] = lo[k] + 1.0;
        }
        const int bits = std::max(1, std::min(16, 60 / d));
        const double cells = static_cast<double>(1u << bits);
        std::vector<uint64_t> key(n);
#pragma omp parallel for schedule(static)
        for (int v = 0; v < n; v++) {
            uint64_t code = 0;
            for (int k = 0; k < d; k++) {
                double t = (x[(size_t)v * d + k] - lo[k]) / (hi[k] - lo[k]) * cells;
                uint32_t q = t <= 0 ? 0u : (t >= cells - 1 ? static_cast<uint32_t>(cells - 1) : static_cast<uint32_t>(t));
                for (int b = 0; b < bits; b++) code |= static_cast<uint64_t>((q >> b) & 1u) << (b * d + k);
            }
            key[v] = code;
        }
        order.resize(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
        pts.resize((size_t)n * d); bound.resize(n);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            std::memcpy(&pts[(size_t)i * d], &x[(size_t)order[i] * d], sizeof(double) * d);
            bound[i] = weightBound[order[i]];
        }
        levels.clear();
        int count = (n + LEAF - 1) / LEAF;
        levels.push_back({count, std::vector<double>((size_t)count * d), std::vector<double>((size_t)count * d), std::vector<double>(count)});
        {
            Level& L = levels[0];
#pragma omp parallel for schedule(static)
            for (int b = 0; b < count; b++) {
                for (int k = 0; k < d; k++) { L.lo[(size_t)b * d + k] = 1e300; L.hi[(size_t)b * d + k] = -1e300; }
                double wm = 0;
                for (int i = b * LEAF; i < std::min(n, (b + 1) * LEAF); i++) {
                    for (int k = 0; k < d; k++) {
                        L.lo[(size_t)b * d + k] = std::min(L.lo[(size_t)b * d + k], pts[(size_t)i * d + k]);
                        L.hi[(size_t)b * d + k] = std::max(L.hi[(size_t)b * d + k], pts[(size_t)i * d + k]);
                    }
                    wm = std::max(wm, bound[i]);
                }
                L.wmax[b] = std::pow(wm, 2.0 / d);
            }
        }
        while (levels.back().count > FAN) {
            const Level& C = levels.back();
            const int pc = (C.count + FAN - 1) / FAN;
            Level P{pc, std::vector<double>((size_t)pc * d), std::vector<double>((size_t)pc * d), std::vector<double>(pc)};
            for (int b = 0; b < pc; b++) {
                for (int k = 0; k < d; k++) { P.lo[(size_t)b * d + k] = 1e300; P.hi[(size_t)b * d + k] = -1e300; }
                double wm = 0;
                for (int c = b * FAN; c < std::min(C.count, (b + 1) * FAN); c++) {
                    for (int k = 0; k < d; k++) {
                        P.lo[(size_t)b * d + k] = std::min(P.lo[(size_t)b * d + k], C.lo[(size_t)c * d + k]);
                        P.hi[(size_t)b * d + k] = std::max(P.hi[(size_t)b * d + k], C.hi[(size_t)c * d + k]);
                    }
                    wm = std::max(wm, C.wmax[c]);
                }
                P.wmax[b] = wm;
            }
            levels.push_back(std::move(P));
        }
    }

    // Calls emit(sortedPos) for every point u whose box chain satisfies
    //   dist(box, q)^2 <= scale2 * (wmax_box)^(2/d)      [scale2 = (L * w_q^(1/d))^2]
    // i.e. a superset of { u : ||x_u - q|| <= L (w_q * bound_u)^(1/d) }.
    template <typename Emit>
    void query(const double* q, double scale2, double invD2, Emit&& emit, uint64_t* visitedNodes = nullptr) const {
        if (n == 0) return;
        walk(static_cast<int>(levels.size()) - 1, 0, levels.back().count, q, scale2, invD2, emit, visitedNodes);
    }

   private:
    template <typename Emit>
    void walk(int lvl, int begin, int end, const double* q, double scale2, double invD2, Emit&& emit, uint64_t* visited) const {
        const Level& L = levels[lvl];
        for (int b = begin; b < end; b++) {
            if (visited) ++*visited;
            double d2 = 0;
            for (int k = 0; k < d; k++) {
                const double lo = L.lo[(size_t)b * d + k], hi = L.hi[(size_t)b * d + k];
                const double e = q[k] < lo ? lo - q[k] : (q[k] > hi ? q[k] - hi : 0.0);
                d2 += e * e;
            }
            // slack of a few ulps: pruning must never drop a point the exact test would keep
            if (d2 > scale2 * L.wmax[b] * (1.0 + 1e-12)) continue;   // wmax holds bound^(2/d)
            if (lvl == 0) {
                for (int i = b * LEAF; i < std::min(n, (b + 1) * LEAF); i++) emit(i);
            } else {
                walk(lvl - 1, b * FAN, std::min(levels[lvl - 1].count, (b + 1) * FAN), q, scale2, invD2, emit, visited);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
struct Port {
    orc_options o;
    uint32_t seed;
    int n = 0, d = 0;
    std::vector<int32_t> rowPtr, col;       // Graph CSR (Graph.hpp:24-85)
    std::vector<double> x, xprev, w, iw, force, m, v, lossPerNode;
    int64_t iter = 0;
    int adamT = 0;
    double lossA = 0, lossR = 0, lastLr = 0, relDisp = 0, relLoss = 0;
    uint64_t numRepPairs = 0, numCandidates = 0, numNodeVisits = 0;
    LossMonitor lossMon;
    MoveMonitor moveMon;
    Schedule sched;
    BoxTree tree;
    std::vector<double> classMaxOfVertex;   // maxWeightOfClass[class(v)] (WeightedIndex.cpp:25-32)

    explicit Port(const orc_options& opts, uint32_t s)
        : o(opts), seed(s), lossMon(opts.stopLossTol, opts.stopLossPatience, opts.lossSmoothingFactor, opts.lossRateWindow),
          moveMon{opts.stopDisplacementTol, opts.stopDisplacementPatience}, sched(opts) {}

    // Graph::constructFromEdges / constructFromMap (Graph.cpp:87-150): symmetric, deduplicated,
    // rows sorted ascending; n = largest id + 1.  All self loops are dropped (the reference drops
    // only the first one it meets and then overruns its edge array, Graph.cpp:124-128).
    void buildGraph(int nHint, int64_t mIn, const int32_t* src, const int32_t* dst) {
        std::vector<uint64_t> e;
        e.reserve(2 * mIn);
        int maxId = -1;
        for (int64_t i = 0; i < mIn; i++) {
            maxId = std::max(maxId, std::max(src[i], dst[i]));
            if (src[i] == dst[i]) continue;
            e.push_back((static_cast<uint64_t>(src[i]) << 32) | static_cast<uint32_t>(dst[i]));
            e.push_back((static_cast<uint64_t>(dst[i]) << 32) | static_cast<uint32_t>(src[i]));
        }
        std::sort(e.begin(), e.end());
        e.erase(std::unique(e.begin(), e.end()), e.end());
        n = std::max(nHint, maxId + 1);
        rowPtr.assign(n + 1, 0);
        col.resize(e.size());
        for (std::size_t i = 0; i < e.size(); i++) {
            rowPtr[(e[i] >> 32) + 1]++;
            col[i] = static_cast<int32_t>(e[i] & 0xffffffffu);
        }
        for (int v = 0; v < n; v++) rowPtr[v + 1] += rowPtr[v];
    }

    // Graph::areNeighbors (Graph.cpp:67-83); rows are sorted so a binary search is equivalent.
    bool areNeighbors(int a, int b) const {
        if (rowPtr[a + 1] - rowPtr[a] > rowPtr[b + 1] - rowPtr[b]) std::swap(a, b);
        return std::binary_search(col.begin() + rowPtr[a], col.begin() + rowPtr[a + 1], b);
    }

    void allocate() {
        d = o.embeddingDimension;
        const std::size_t nd = static_cast<std::size_t>(n) * d;
        x.assign(nd, 0.0); xprev.assign(nd, 0.0); force.assign(nd, 0.0); m.assign(nd, 0.0); v.assign(nd, 0.0);
        w.assign(n, 0.0); iw.assign(n, 0.0); lossPerNode.assign(n, 0.0);
        lastLr = o.learningRate;  // EmbedderInterface.hpp:36
    }

    // WembedEmbedder ctor (WembedEmbedder.hpp:109-124)
    void initState(std::mt19937& global) {
        // EmbedderInterface::constructRandomCoordinates (EmbedderInterface.hpp:61-65) + Rand::randomCoordinates
        // (Rand.cpp:101-109): cube side pow((float)n, 1/d), vertex-major draw order.
        const double side = std::pow(static_cast<float>(n), 1.0 / d);
        for (std::size_t i = 0; i < x.size(); i++) {
            std::uniform_real_distribution<double> dist(0.0, side);
            x[i] = dist(global);
        }
        std::vector<double> weights(n);
        if (o.weightType == 1) {
            // constructDegreeWeights + rescaleWeights (WembedEmbedder.cpp:359-390)
            for (int u = 0; u < n; u++) {
                const int deg = rowPtr[u + 1] - rowPtr[u];
                weights[u] = deg > 0 ? deg : 1;
                if (o.dimensionHint > 0) weights[u] = std::pow(weights[u], static_cast<double>(d) / o.dimensionHint);
            }
            double sum = 0.0;
            for (int u = 0; u < n; u++) sum += weights[u];
            for (int u = 0; u < n; u++) weights[u] = weights[u] * (static_cast<double>(n) / sum);
        } else {
            std::fill(weights.begin(), weights.end(), 1.0);
        }
        setWeights(weights.data());
    }

    // WembedEmbedder::setWeights (WembedEmbedder.cpp:121-131)
    void setWeights(const double* weights) {
        std::copy(weights, weights + n, w.begin());
        for (int u = 0; u < n; u++) iw[u] = 1.0 / std::pow(w[u], 1.0 / static_cast<double>(d));
        // WeightedIndex::getDoublingWeightBuckets + updateIndices class assignment
        // (WeightedIndex.cpp:51-63, 18-32): weights are constant during a run, so the classes are too.
        classMaxOfVertex.assign(n, 0.0);
        if (n == 0) return;
        const double minW = *std::min_element(w.begin(), w.end());
        const double maxW = *std::max_element(w.begin(), w.end());
        std::vector<double> buckets;
        for (double c = minW * o.doublingFactor; c < maxW; c *= o.doublingFactor) buckets.push_back(c);
        std::vector<double> classMax = buckets;
        classMax.push_back(maxW);
        for (int u = 0; u < n; u++) {
            const std::size_t c = std::upper_bound(buckets.begin(), buckets.end(), w[u]) - buckets.begin();
            classMaxOfVertex[u] = classMax[c];
        }
    }

    // Rand::localGenerator + setToRandomUnitVector (Rand.cpp:29-35, DVec.hpp:412-424)
    void addRandomUnit(int vtx, double* f) const {
        std::seed_seq seq{seed, static_cast<uint32_t>(vtx), static_cast<uint32_t>(iter)};
        std::mt19937 gen(seq);
        double buf[64];
        double norm = 0.0;
        for (int k = 0; k < d; k++) {
            std::normal_distribution<double> dist(0.0, 1.0);
            buf[k] = dist(gen);
            norm += buf[k] * buf[k];
        }
        norm = std::sqrt(norm);
        for (int k = 0; k < d; k++) f[k] += buf[k] / norm;
    }

    // calculateLPNorm (VectorOperations.hpp:5-11)
    double distance(const double* a, const double* b) const {
        double s = 0.0;
        for (int k = 0; k < d; k++) s += std::pow(std::abs(a[k] - b[k]), 2);
        return std::sqrt(s);
    }

    // attractionForce (WembedEmbedder.cpp:140-172)
    double attract(int vtx, int u) {
        if (vtx == u) return 0.0;
        const double* pv = &x[(size_t)vtx * d];
        const double* pu = &x[(size_t)u * d];
        double* f = &force[(size_t)vtx * d];
        const double dist = distance(pu, pv);
        if (dist <= 0) { addRandomUnit(vtx, f); return 0.0; }
        const double ws = iw[vtx] * iw[u];
        if (dist * ws <= o.edgeLength) return 0.0;  // result *= 0 -> adds +0
        const double scale = o.attractionScale * ws;
        for (int k = 0; k < d; k++) {
            // differentiateLPNormDifference (VectorOperations.hpp:13-25): |a-b|/dist * sign(a-b)
            const double diff = pu[k] - pv[k];
            const double g = std::abs(diff) / dist * (diff < 0 ? -1.0 : 1.0);
            f[k] += g * scale;
        }
        return dist - o.edgeLength / ws;
    }

    // repellingForce (WembedEmbedder.cpp:174-210), numNegativeSamples = -1
    double repel(int vtx, int u) {
        if (vtx == u) return 0.0;
        const double* pv = &x[(size_t)vtx * d];
        const double* pu = &x[(size_t)u * d];
        double* f = &force[(size_t)vtx * d];
        const double dist = distance(pv, pu);
        if (dist <= 0) { addRandomUnit(vtx, f); return 0.0; }
        const double ws = iw[vtx] * iw[u];
        if (dist * ws > o.edgeLength) return 0.0;
        const double scale = o.repulsionScale * ws;
        for (int k = 0; k < d; k++) {
            const double diff = pv[k] - pu[k];
            const double g = std::abs(diff) / dist * (diff < 0 ? -1.0 : 1.0);
            f[k] += g * scale;
        }
        return o.edgeLength / ws - dist;
    }

    // Candidate ids of vtx in ascending order.  useClassBound=true reproduces the reference's
    // candidate set (WeightedIndex.cpp:65-81: radius L*(w_v*maxW_class(u))^(1/d), v itself included);
    // false keeps only points that can pass the exact test (bound = w_u).
    void candidates(int vtx, bool useClassBound, std::vector<int>& out, uint64_t* visits = nullptr) const {
        out.clear();
        const double* q = &x[(size_t)vtx * d];
        const double rq = o.edgeLength * std::pow(w[vtx], 1.0 / d);
        const double scale2 = rq * rq;
        tree.query(q, scale2, 2.0 / d, [&](int pos) {
            const int u = tree.order[pos];
            const double bw = useClassBound ? classMaxOfVertex[u] : w[u];
            const double r = o.edgeLength * std::pow(w[vtx] * bw, 1.0 / d);
            double s = 0.0;
            for (int k = 0; k < d; k++) { const double e = tree.pts[(size_t)pos * d + k] - q[k]; s += e * e; }
            if (s <= r * r * (1.0 + 1e-12)) out.push_back(u);
        }, visits);
        std::sort(out.begin(), out.end());
    }

    void rebuildIndex(bool useClassBound) { tree.build(n, d, x.data(), useClassBound ? classMaxOfVertex.data() : w.data()); }

    // WembedEmbedder::calculateStep (WembedEmbedder.cpp:13-63)
    void step() {
        iter++;                                   // EmbedderState::nextStep (EmbedderState.hpp:46-51)
        std::fill(force.begin(), force.end(), 0.0);
        lossA = lossR = 0.0;
        if (n <= 1) return;
        xprev = x;                                // :25
        rebuildIndex(false);                      // updateIndex (:212-240)

        // calculateAllAttractingForces (:260-272)
#pragma omp parallel for schedule(dynamic, 256)
        for (int vtx = 0; vtx < n; vtx++) {
            double loss = 0.0;
            for (int e = rowPtr[vtx]; e < rowPtr[vtx + 1]; e++) loss += attract(vtx, col[e]);
            lossPerNode[vtx] = loss;
        }
        lossA = blockSum(n, [&](std::size_t i) { return lossPerNode[i]; });

        // calculateAllRepellingForces (:274-294)
        uint64_t pairs = 0, cands = 0, visits = 0;
#pragma omp parallel reduction(+ : pairs, cands, visits)
        {
            std::vector<int> cand;
#pragma omp for schedule(dynamic, 256)
            for (int vtx = 0; vtx < n; vtx++) {
                candidates(vtx, false, cand, &visits);
                double loss = 0.0;
                for (int u : cand) {
                    if (u == vtx || areNeighbors(vtx, u)) continue;   // :284 (colour test == same vertex)
                    loss += repel(vtx, u);
                    pairs++;
                }
                cands += cand.size();
                lossPerNode[vtx] = loss;
            }
        }
        numRepPairs = pairs; numCandidates = cands; numNodeVisits = visits;
        lossR = blockSum(n, [&](std::size_t i) { return lossPerNode[i]; });

        // calculateAllCentreForces (:296-301)
        if (o.centreScale != 0.0) {
#pragma omp parallel for schedule(static)
            for (std::size_t i = 0; i < force.size(); i++) force[i] += -1.0 * o.centreScale * x[i];
        }

        lastLr = sched.rate(static_cast<int>(iter), lossMon);   // :53-54
        if (o.optimizerType == 1) {
            // AdamOptimizer::update (AdamOptimizer.cpp:15-30), beta1=.9 beta2=.999 eps=1e-8 (WembedEmbedder.hpp:46)
            adamT++;
            const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
            const double c1 = 1.0 - std::pow(b1, adamT), c2 = 1.0 - std::pow(b2, adamT);
#pragma omp parallel for schedule(static)
            for (std::size_t i = 0; i < force.size(); i++) {
                m[i] = b1 * m[i] + (1.0 - b1) * force[i];
                v[i] = b2 * v[i] + (1.0 - b2) * force[i] * force[i];
                const double mHat = m[i] / c1, vHat = v[i] / c2;
                x[i] += lastLr * mHat / (std::sqrt(vHat) + eps);
            }
        } else {
            // SimpleOptimizer::update (SimpleOptimizer.cpp:13-30)
#pragma omp parallel for schedule(static)
            for (std::size_t i = 0; i < force.size(); i++) {
                double g = force[i];
                g = std::max(g, -o.simpleOptMaxDisplacement);
                g = std::min(g, o.simpleOptMaxDisplacement);
                x[i] += g * lastLr;
            }
        }

        // applyGravityCentre (:303-319)
        std::vector<double> centre(d);
        for (int k = 0; k < d; k++)
            centre[k] = blockSum(n, [&](std::size_t u) { return x[u * d + k]; }) / static_cast<double>(n);
#pragma omp parallel for schedule(static)
        for (int u = 0; u < n; u++)
            for (int k = 0; k < d; k++) x[(size_t)u * d + k] -= centre[k];

        // observeDisplacement (:321-352)
        std::vector<double> disp(n), rad(n);
#pragma omp parallel for schedule(static)
        for (int u = 0; u < n; u++) {
            disp[u] = distance(&x[(size_t)u * d], &xprev[(size_t)u * d]);
            double r2 = 0.0;
            for (int k = 0; k < d; k++) r2 += x[(size_t)u * d + k] * x[(size_t)u * d + k];
            rad[u] = r2;
        }
        const double invN = 1.0 / static_cast<double>(n);
        const double meanDisp = blockSum(n, [&](std::size_t i) { return disp[i]; }) * invN;
        const double radius = std::sqrt(blockSum(n, [&](std::size_t i) { return rad[i]; }) * invN);
        relDisp = radius > 0.0 ? meanDisp / radius : 0.0;
        moveMon.observe(relDisp);
        lossMon.observe(lossA + lossR);            // :61-62
        relLoss = lossMon.rate;
    }

    // WembedEmbedder::isFinished (:65-75)
    bool finished() const {
        if (iter >= o.maxIterations) return true;
        if (n <= 1) return true;
        return o.stopCriterion == 0 ? moveMon.converged() : lossMon.converged();
    }
};

}  // namespace

extern "C" {

void port_options_default(orc_options* o) {
    // defaults of EmbedderOptions (EmbedderOptions.hpp:31-88)
    std::memset(o, 0, sizeof(*o));
    o->embeddingDimension = 4; o->weightType = 1; o->optimizerType = 1; o->maxIterations = 10000;
    o->lrScheduleType = 0; o->warmupSteps = 20; o->lrAdaptPatience = 20; o->stopCriterion = 1;
    o->stopDisplacementPatience = 5; o->lossRateWindow = 30; o->stopLossPatience = 50; o->numThreads = 0;
    o->dimensionHint = -1.0; o->attractionScale = 1.0; o->repulsionScale = 1.0; o->centreScale = 0.0;
    o->edgeLength = 1.0; o->doublingFactor = 2.0; o->simpleOptMaxDisplacement = 1.0; o->learningRate = 10;
    o->lrCoolingFactor = 0.995; o->lrDecayFactor = 0.5; o->lrDecayThreshold = 1e-2; o->lrGrowthFactor = 1.0;
    o->lrGrowthThreshold = 1e-1; o->stopDisplacementTol = 3e-4; o->lossSmoothingFactor = 0.3; o->stopLossTol = 1e-3;
}

void* port_create(int32_t n_hint, int64_t m, const int32_t* src, const int32_t* dst, const orc_options* o, int32_t seed,
                  int32_t init_state) {
    if (o->numThreads > 0) omp_set_num_threads(o->numThreads);
    auto* p = new Port(*o, static_cast<uint32_t>(seed));
    p->buildGraph(n_hint, m, src, dst);
    p->allocate();
    if (init_state) {
        std::mt19937 global(static_cast<uint32_t>(seed));  // Rand::setSeed (Rand.cpp:22-25)
        p->initState(global);
    } else {
        std::vector<double> ones(p->n, 1.0);
        p->setWeights(ones.data());
        std::fill(p->w.begin(), p->w.end(), 0.0);  // the reference leaves weights at 0 until setWeights
    }
    return p;
}

void port_destroy(void* h) { delete static_cast<Port*>(h); }
int32_t port_num_vertices(void* h) { return static_cast<Port*>(h)->n; }
int64_t port_num_directed_edges(void* h) { return static_cast<int64_t>(static_cast<Port*>(h)->col.size()); }
void port_csr(void* h, int32_t* rp, int32_t* col) {
    auto* p = static_cast<Port*>(h);
    std::copy(p->rowPtr.begin(), p->rowPtr.end(), rp);
    std::copy(p->col.begin(), p->col.end(), col);
}
int32_t port_are_neighbors(void* h, int32_t v, int32_t u) { return static_cast<Port*>(h)->areNeighbors(v, u) ? 1 : 0; }
void port_set_coordinates(void* h, const double* c) { auto* p = static_cast<Port*>(h); std::copy(c, c + p->x.size(), p->x.begin()); }
void port_set_weights(void* h, const double* w) { static_cast<Port*>(h)->setWeights(w); }
void port_get_coordinates(void* h, double* c) { auto* p = static_cast<Port*>(h); std::copy(p->x.begin(), p->x.end(), c); }
void port_get_weights(void* h, double* w) { auto* p = static_cast<Port*>(h); std::copy(p->w.begin(), p->w.end(), w); }
void port_get_forces(void* h, double* f) { auto* p = static_cast<Port*>(h); std::copy(p->force.begin(), p->force.end(), f); }
void port_step(void* h) { static_cast<Port*>(h)->step(); }
int32_t port_is_finished(void* h) { return static_cast<Port*>(h)->finished() ? 1 : 0; }
int64_t port_run(void* h) {
    auto* p = static_cast<Port*>(h);
    p->iter = 0;  // calculateEmbedding (:77-86) resets only the iteration counter
    while (!p->finished()) p->step();
    return p->iter;
}
void port_get_stats(void* h, double* s) {
    auto* p = static_cast<Port*>(h);
    s[ORC_LOSS_ATTRACT] = p->lossA; s[ORC_LOSS_REPEL] = p->lossR; s[ORC_LR] = p->lastLr; s[ORC_REL_DISP] = p->relDisp;
    s[ORC_REL_LOSS_IMPROVEMENT] = p->relLoss; s[ORC_ITERATION] = static_cast<double>(p->iter);
    s[ORC_NUM_REP_PAIRS] = static_cast<double>(p->numRepPairs);
    s[ORC_RESERVED] = static_cast<double>(p->numNodeVisits);
}
// Test helper (port only): flags[v] = 1 if v owns a pair - neighbour or not - whose weighted distance dist*ws lies
// within tau*L of the hinge threshold L, i.e. a pair on which fp32 and fp64 may legitimately disagree.
void port_flag_near_threshold(void* h, double tau, uint8_t* flags) {
    auto* p = static_cast<Port*>(h);
    const int n = p->n, d = p->d;
    const double L = p->o.edgeLength;
    p->rebuildIndex(false);
#pragma omp parallel for schedule(dynamic, 256)
    for (int v = 0; v < n; v++) {
        bool near = false;
        const double* q = &p->x[(size_t)v * d];
        for (int e = p->rowPtr[v]; e < p->rowPtr[v + 1] && !near; e++) {
            const int u = p->col[e];
            near = std::abs(p->distance(q, &p->x[(size_t)u * d]) * p->iw[v] * p->iw[u] - L) <= tau * L;
        }
        if (!near) {
            const double rq = L * (1.0 + 2.0 * tau) * std::pow(p->w[v], 1.0 / d);
            p->tree.query(q, rq * rq, 2.0 / d, [&](int pos) {
                const int u = p->tree.order[pos];
                if (u == v || near) return;
                const double dist = p->distance(q, &p->tree.pts[(size_t)pos * d]);
                if (std::abs(dist * p->iw[v] * p->iw[u] - L) <= tau * L) near = true;
            });
        }
        flags[v] = near ? 1 : 0;
    }
}

int64_t port_candidates(void* h, int32_t v, int32_t* out, int64_t cap) {
    auto* p = static_cast<Port*>(h);
    p->rebuildIndex(true);
    std::vector<int> c;
    p->candidates(v, true, c);
    for (std::size_t i = 0; i < c.size() && static_cast<int64_t>(i) < cap; i++) out[i] = c[i];
    return static_cast<int64_t>(c.size());
}

}  // extern "C"
