// C ABI (include/wembed_b200.h) and host orchestration of the device step.
//
// One wb_embedder owns one device-resident problem and one CUDA stream.  A step is a fixed sequence of kernel launches on that
// stream (launch_step); which of them do work - index rebuild + repulsion search, or only the evaluation of the stored pair list -
// is decided by the device itself (StepCtrl, params.cuh), so steps can be queued without the host waiting for anything.  The only
// host<->device traffic per step is one small H2D copy of the step's scalars and one D2H copy of the reduced sums.
// There is no CPU fallback anywhere in this file.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../include/wembed_b200.h"
#include "kernels.cuh"

#define WB_STRINGIFY_(x) #x
#define WB_STRINGIFY(x) WB_STRINGIFY_(x)

namespace {

thread_local std::string g_lastError;

// NCCL is bound lazily (dlopen) and only by the multi-GPU entry points: the single-GPU path has no NCCL dependency, and
// a host process that already carries an NCCL (e.g. the one bundled with PyTorch, same SONAME) keeps using that copy.
struct NcclApi {
    ncclResult_t (*getUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*commInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*allGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return a;
        a.getUniqueId = reinterpret_cast<decltype(a.getUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        a.commInitRank = reinterpret_cast<decltype(a.commInitRank)>(dlsym(lib, "ncclCommInitRank"));
        a.allGather = reinterpret_cast<decltype(a.allGather)>(dlsym(lib, "ncclAllGather"));
        a.commDestroy = reinterpret_cast<decltype(a.commDestroy)>(dlsym(lib, "ncclCommDestroy"));
        a.ok = a.getUniqueId && a.commInitRank && a.allGather && a.commDestroy;
        return a;
    }();
    return api;
}

int fail(int code, const std::string& msg) {
    g_lastError = msg;
    return code;
}

template <typename T>
T* dalloc(size_t count) {
    T* p = nullptr;
    WB_CUDA(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    return p;
}

// host copy of a step's scalars + where its sums land; lives in pinned memory until the step has been collected
struct StepSlotHost {
    wb::StepDyn dyn;
    double sums[wb::kMaxSums + 32];
};
struct PendingStep {
    cudaEvent_t done;
    StepSlotHost* host;   // pinned
    int64_t iteration;
    bool trivial;         // n <= 1: nothing was launched
};

}  // namespace

struct wb_embedder {
    int n = 0, dim = 0, V = 0, rowFloats = 0;
    int64_t numDirected = 0;
    wb_options opt{};
    cudaStream_t stream = nullptr;

    // graph (Graph.hpp:24-85): CSR, rows sorted ascending
    int *rowPtr = nullptr, *col = nullptr;

    // layout state
    float4 *x = nullptr, *xNew = nullptr, *mom1 = nullptr, *mom2 = nullptr, *force = nullptr;
    size_t rowsAlloc = 0;                 // rows of x / xNew / mom / force (n rounded up to whole tiles)
    float* iw = nullptr;
    double fixForce = 1.0, fixLoss = 1.0; // fixed-point scales of the repulsive terms (powers of two, chosen by wb_set_weights)
    double maxIw = 1.0, minIw = 1.0;
    std::vector<double> weights;          // state.currentWeights
    std::vector<int> hostDegree;          // CSR row lengths (host copy)
    std::vector<double> classMax;         // maxWeightOfClass[class(v)] (WeightedIndex.cpp:25-32), for the test hook
    int64_t iteration = 0;                // state.currentIteration
    int adamT = 0;                        // AdamOptimizer::t

    // device-side control block, per-step scalars, pair list
    wb::StepCtrl* ctrl = nullptr;
    wb::StepDyn* dyn = nullptr;
    int2* pairBuf = nullptr;              // world segments of pairCap pairs: segment p = what rank p found for this rank's vertices
    unsigned int pairCap = 0;
    unsigned int* pairCounts = nullptr;   // counts matrix [kMaxRanks][kMaxRanks] inside `mail`
    char* mail = nullptr;                 // [flags | counts | block sum rows | observation tiles | moment tiles] (step.cuh: k_exchange)
    size_t mailBytes = 0;
    int *repDeg = nullptr, *repCol = nullptr, *scanSums = nullptr, *longRows = nullptr, *longCount = nullptr;
    long long *repRowPtr = nullptr, *scanOffsets = nullptr;
    int scanBlocks = 0;
    float skinMax = 0.f, reuseTarget = 4.f;
    int nextRebuild = 1;                  // what the host knows about the next step: 1 rebuilds (or unknown), 0 reuses the list
    bool quantValid = false;              // the quantisation frame on the device belongs to the current positions

    wb::RepLayout repLayout{1, 0, 0};     // which sorted positions this rank's repulsion walk queries
    int* chunkCounter = nullptr;          // work counter of the persistent repulsion kernel
    int numHeavy = 0;                     // vertices of weight >= kHeavyWeight x mean, walked by k_repulse_heavy
    int *heavyVertex = nullptr, *heavySlot = nullptr, *heavyPos = nullptr;
    int numHubs = 0;                      // long CSR rows and heavy vertices, summed by k_hub_rows
    int *hubVertex = nullptr, *hubSlot = nullptr;
    double* hubD = nullptr;
    long long* hubF = nullptr;
    uint32_t* mtScratch = nullptr;        // tie-break generator state, one row of 624 words per warp of the fused kernel

    // spatial index
    int mortonBits = 0;                   // bits per dimension of the 32-bit Morton key
    uint32_t *keysIn = nullptr, *keysOut = nullptr;
    int *valsIn = nullptr, *valsOut = nullptr;
    void* cubTemp = nullptr;
    size_t cubBytes = 0;
    wb::QuantParams* quant = nullptr;
    float4* blkH = nullptr;               // half-precision copy of the array-of-blocks tree (box rounds of k_repulse_pairs)
    float halfSigmaLimit = 0.f;           // layouts with a larger per-dimension sd walk the fp32 boxes
    int halfMode = -1;                    // WB_HALF_BOXES: 0 never, 1 always, unset: by the layout
    float4* lvlLo[wb::kMaxLevels] = {};
    float4* lvlHi[wb::kMaxLevels] = {};
    float* lvlBound[wb::kMaxLevels] = {};
    int* ids = nullptr;
    float4* blk = nullptr;                // array-of-blocks copy of the boxes (wb::TreeView::blk)
    wb::TreeView tree{};
    wb::TreePlanes planes{};

    // reductions (all over GLOBAL block rows / tiles, see step.cuh)
    int passVerts = 0, vertsPerBlock = 0, numBlockRows = 0, cols = 0;   // fused kernel: vertices per pass, per block (fixed by n alone), rows, sums per row
    int fusedBlocks = 0, repBlocks = 0, numObsTiles = 0;
    double *blockPartials = nullptr, *forceSums = nullptr, *obsPartials = nullptr, *walkPartials = nullptr, *stats = nullptr;
    float* frameScratch = nullptr;        // per-tile moments of the layout (k_moments)
    int numMomentTiles = 1, momentCount = 1;
    int statsTotal = 0;

    // vertex-sharded multi-GPU step (wb_comm_init): this rank owns vertices [ownBegin, ownEnd); the other ranks' buffers are mapped
    // through CUDA IPC and written to directly by the kernels (peer stores over NVLink)
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, ownBegin = 0, ownEnd = 0, rowsPerRank = 0;
    char* peerMail[wb::kMaxRanks] = {};
    float4* peerX[wb::kMaxRanks] = {};
    int2* peerPairs[wb::kMaxRanks] = {};
    bool peersOpen = false, peerPairsOpen = false;
    bool localGroup = false;              // wb_comm_init_local: the "ranks" are handles of this process on one device (tests)
    int epoch = 0;                        // barrier counter (k_exchange)

    std::deque<PendingStep> pending;
    std::vector<PendingStep> freeSlots;

    // host <-> device staging of coordinate rows (wb_set_coordinates / wb_get_coordinates): per worker thread one stream and two
    // pinned chunks, allocated on first use and kept for the life of the handle
    struct StageLane { cudaStream_t stream = nullptr; float* pinned[2] = {nullptr, nullptr}; cudaEvent_t done[2] = {nullptr, nullptr}; };
    std::vector<StageLane> stageLanes;

    // the step as a CUDA graph (one GPU, no phase timing): [k_step_begin] -> IF(rebuild){index, search, list} -> [forces .. tail]
    cudaGraph_t stepGraph = nullptr;
    cudaGraphExec_t stepExec = nullptr;
    bool graphWanted = true, graphFailed = false;
    int graphKernels = 0;                 // kernels one replay launches when it rebuilds (for wb_launch_count)
    std::string graphNote;

    cudaEvent_t marks[8] = {};
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev[6] = {};
    double phaseMs[6] = {0, 0, 0, 0, 0, 0};
    bool havePhase = false;
};

namespace {

using wb::kFan;

#define WB_DISPATCH_V(V_, ...)                                   \
    switch (V_) {                                                \
        case 1: { constexpr int V = 1; __VA_ARGS__; } break;     \
        case 2: { constexpr int V = 2; __VA_ARGS__; } break;     \
        case 3: { constexpr int V = 3; __VA_ARGS__; } break;     \
        case 4: { constexpr int V = 4; __VA_ARGS__; } break;     \
        case 5: { constexpr int V = 5; __VA_ARGS__; } break;     \
        case 6: { constexpr int V = 6; __VA_ARGS__; } break;     \
        case 7: { constexpr int V = 7; __VA_ARGS__; } break;     \
        case 8: { constexpr int V = 8; __VA_ARGS__; } break;     \
        default: throw wb::CudaError{cudaErrorInvalidValue, "unsupported dimension", __FILE__, __LINE__}; \
    }

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

void free_all(wb_embedder* h) {
    auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    F(h->rowPtr); F(h->col); F(h->x); F(h->xNew); F(h->mom1); F(h->mom2); F(h->force); F(h->iw);
    if (h->world > 1) {
        for (int p = 0; p < h->world; ++p) {
            if (p == h->rank) continue;
            if (h->peersOpen) { cudaIpcCloseMemHandle(h->peerMail[p]); cudaIpcCloseMemHandle(h->peerX[p]); }
            if (h->peerPairsOpen) cudaIpcCloseMemHandle(h->peerPairs[p]);
        }
        h->peersOpen = h->peerPairsOpen = false;
    }
    F(h->ctrl); F(h->dyn); F(h->pairBuf); F(h->mail); F(h->repDeg); F(h->repRowPtr); F(h->repCol); F(h->scanSums); F(h->scanOffsets); F(h->longRows); F(h->longCount);
    F(h->chunkCounter); F(h->heavyVertex); F(h->heavySlot); F(h->heavyPos); F(h->hubVertex); F(h->hubSlot); F(h->hubD); F(h->hubF); F(h->mtScratch);
    F(h->keysIn); F(h->keysOut); F(h->valsIn); F(h->valsOut); F(h->cubTemp); F(h->quant); F(h->ids); F(h->blk); F(h->blkH);
    for (int l = 0; l < wb::kMaxLevels; ++l) { F(h->lvlLo[l]); if (l > 0) F(h->lvlHi[l]); F(h->lvlBound[l]); }
    F(h->forceSums); F(h->walkPartials); F(h->stats); F(h->frameScratch);
    if (h->comm) { nccl().commDestroy(h->comm); h->comm = nullptr; }
    for (auto& p : h->pending) { cudaEventDestroy(p.done); cudaFreeHost(p.host); }
    for (auto& p : h->freeSlots) { cudaEventDestroy(p.done); cudaFreeHost(p.host); }
    h->pending.clear(); h->freeSlots.clear();
    for (auto& l : h->stageLanes) {
        for (int b = 0; b < 2; ++b) { if (l.pinned[b]) cudaFreeHost(l.pinned[b]); if (l.done[b]) cudaEventDestroy(l.done[b]); }
        if (l.stream) cudaStreamDestroy(l.stream);
    }
    h->stageLanes.clear();
    if (h->stepExec) { cudaGraphExecDestroy(h->stepExec); h->stepExec = nullptr; }
    if (h->stepGraph) { cudaGraphDestroy(h->stepGraph); h->stepGraph = nullptr; }
    for (auto& e : h->ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : h->marks) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (h->stream) cudaStreamDestroy(h->stream);
    h->stream = nullptr;
}

// Fixed-point scales of the repulsive terms: the largest power of two such that n terms of the largest possible magnitude
// (|force component| <= |repulsionScale| * max ws, loss term <= L / min ws, ws = iw_v * iw_u) stay below 2^62.
void choose_fixed_scales(wb_embedder* h, double maxIw, double minIw) {
    auto scale = [&](double maxTerm) {
        const double bound = (double)std::max(h->n, 2) * maxTerm;
        if (!(bound > 0.0) || !std::isfinite(bound)) return 1.0;
        return std::ldexp(1.0, std::max(-900, std::min(900, 61 - (int)std::ceil(std::log2(bound)))));
    };
    h->maxIw = maxIw; h->minIw = minIw;
    h->fixForce = scale(std::fabs(h->opt.repulsion_scale) * maxIw * maxIw);
    h->fixLoss = scale(h->opt.edge_length / (minIw * minIw));
    // Box format of the repulsion walk (decided on the device whenever the quantisation frame is computed): half-precision boxes while
    // the layout's largest per-dimension sd is at most kHalfSpread smallest interaction radii (L / max ws); the rounding of a centred
    // coordinate is ~sd * 2^-11, i.e. ~3 % of that radius at the limit.  Simulated on the c3 layout with shortened mantissas: +3 % box
    // tests and +7 % point tests at 55 radii (the half-precision rounds are ~26 % cheaper), +0.2 % / +0.4 % at the 3.4 radii of c3 itself.
    // WB_HALF_BOXES=0 / 1 forces one format (A/B runs, tests).
    constexpr double kHalfSpread = 64.0;
    if (h->halfMode < 0) {
        const char* env = std::getenv("WB_HALF_BOXES");
        h->halfMode = env ? (std::atoi(env) != 0 ? 1 : 0) : 2;
    }
    const double rMin = h->opt.edge_length / (maxIw * maxIw), rMax = h->opt.edge_length * (1.0 + (double)h->skinMax) / (minIw * minIw);
    // squared gaps are summed in half precision (max 65504): radii beyond ~200 cannot be tested there at all
    const bool representable = rMax < 200.0;
    h->halfSigmaLimit = h->halfMode == 0 ? -1.f : (h->halfMode == 1 ? std::numeric_limits<float>::infinity()
                                                                    : (representable ? (float)(kHalfSpread * rMin) : -1.f));
}

// (re)writes the device control block: the pair list is void, the next step rebuilds index and list without a skin
void invalidate_list(wb_embedder* h) {
    wb::StepCtrl c{};
    c.overflow = 0; c.pairNeeded = 0; c.listValid = 0; c.rebuild = 1;
    c.skin = 0.f; c.dispAccum = 0.f;
    c.listL2 = (float)(h->opt.edge_length * h->opt.edge_length) * (1.0f + wb::kPruneSlack);
    c.pruneL = std::sqrt(c.listL2);
    c.skinMax = h->skinMax; c.reuseTarget = h->reuseTarget; c.skinCap = h->skinMax;
    c.pairBudget = (unsigned int)std::min<int64_t>((int64_t)4 * std::max(h->n, 1), 0x7fffffff);
    c.numBuilds = 0; c.numReused = 0;
    WB_CUDA(cudaMemcpyAsync(h->ctrl, &c, sizeof(c), cudaMemcpyHostToDevice, h->stream));   // pageable source: the copy is staged before the call returns
    h->nextRebuild = 1;
    // whatever made the list void (new weights, a grown pair buffer, a new policy, ..) may also have changed arguments a captured
    // step holds by value: capture it again on the next step
    if (h->stepExec) { cudaGraphExecDestroy(h->stepExec); h->stepExec = nullptr; }
    if (h->stepGraph) { cudaGraphDestroy(h->stepGraph); h->stepGraph = nullptr; }
}

// hub rows (long CSR rows; heavy vertices, whose rows of the pair list are long) and heavy vertices (walked by one block each)
void rebuild_hub_lists(wb_embedder* h) {
    const int n = h->n, V = h->V;
    auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    F(h->hubVertex); F(h->hubSlot); F(h->hubD); F(h->hubF); F(h->heavyVertex); F(h->heavySlot); F(h->heavyPos); F(h->walkPartials);
    double mean = 0.0;
    for (int v = 0; v < n; ++v) mean += h->weights[v];
    mean /= (double)std::max(n, 1);
    std::vector<int> hubs, hubSlot(n, -1), heavy, heavySlot(n, -1);
    for (int v = 0; v < n; ++v) {
        const bool isHeavy = h->weights[v] >= (double)wb::kHeavyWeight * mean;
        if (isHeavy) { heavySlot[v] = (int)heavy.size(); heavy.push_back(v); }
        if (isHeavy || h->hostDegree[v] > wb::kHubThreshold) { hubSlot[v] = (int)hubs.size(); hubs.push_back(v); }
    }
    h->numHubs = (int)hubs.size();
    h->numHeavy = (int)heavy.size();
    if (h->numHubs) {
        h->hubVertex = dalloc<int>(hubs.size());
        h->hubSlot = dalloc<int>(n);
        h->hubD = dalloc<double>(hubs.size() * wb::hub_doubles(V));
        h->hubF = dalloc<long long>(hubs.size() * wb::hub_fixed(V));
        WB_CUDA(cudaMemcpyAsync(h->hubVertex, hubs.data(), sizeof(int) * hubs.size(), cudaMemcpyHostToDevice, h->stream));
        WB_CUDA(cudaMemcpyAsync(h->hubSlot, hubSlot.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    }
    if (h->numHeavy) {
        h->heavyVertex = dalloc<int>(heavy.size());
        h->heavySlot = dalloc<int>(n);
        h->heavyPos = dalloc<int>(heavy.size());
        WB_CUDA(cudaMemcpyAsync(h->heavyVertex, heavy.data(), sizeof(int) * heavy.size(), cudaMemcpyHostToDevice, h->stream));
        WB_CUDA(cudaMemcpyAsync(h->heavySlot, heavySlot.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
        WB_CUDA(cudaMemsetAsync(h->heavyPos, 0, sizeof(int) * heavy.size(), h->stream));
    }
    h->walkPartials = dalloc<double>(((size_t)h->repBlocks * 8 + heavy.size()) * 3);
    WB_CUDA(cudaStreamSynchronize(h->stream));
}

// pair buffer + CSR entries for `cap` unordered pairs
void allocate_pair_list(wb_embedder* h, unsigned int cap) {
    auto F = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    F(h->pairBuf); F(h->repCol);
    size_t freeBytes = 0, totalBytes = 0;
    WB_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
    const size_t need = (size_t)cap * h->world * 16;             // pairs + both directions of the CSR
    if (need > freeBytes - std::min(freeBytes, (size_t)2 << 30))
        throw std::runtime_error("repulsion pair list: " + std::to_string(need >> 20) + " MiB needed, " + std::to_string(freeBytes >> 20) + " MiB free on the device");
    h->pairCap = cap;                                            // per segment
    h->pairBuf = dalloc<int2>((size_t)cap * h->world);
    h->repCol = dalloc<int>((size_t)2 * cap * h->world + 8);
}

void allocate(wb_embedder* h, const int32_t* rowPtr, const int32_t* col) {
    const int n = h->n, V = h->V;
    WB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // policy of the pair list (A/B runs and tests; results never depend on it)
        const char* e = std::getenv("WB_SKIN_MAX");
        h->skinMax = e ? (float)std::atof(e) : 0.3f;
        e = std::getenv("WB_REUSE_STEPS");
        h->reuseTarget = e ? std::max(1.f, (float)std::atof(e)) : 4.f;
        e = std::getenv("WB_GRAPH");
        h->graphWanted = !(e && std::atoi(e) == 0);
    }
    // (+ 8: the fused kernel copies whole 16-byte groups of these arrays)
    h->rowPtr = dalloc<int>(n + 1 + 8);
    h->col = dalloc<int>(h->numDirected + 8);
    WB_CUDA(cudaMemsetAsync(h->rowPtr, 0, sizeof(int) * (n + 1 + 8), h->stream));
    WB_CUDA(cudaMemcpyAsync(h->rowPtr, rowPtr, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, h->stream));
    if (h->numDirected) WB_CUDA(cudaMemcpyAsync(h->col, col, sizeof(int) * h->numDirected, cudaMemcpyHostToDevice, h->stream));
    h->hostDegree.resize(n);
    for (int v = 0; v < n; ++v) h->hostDegree[v] = rowPtr[v + 1] - rowPtr[v];

    // Block rows of the fused kernel: every block owns `vertsPerBlock` consecutive vertices, a multiple of one pass and a function of n
    // alone, so the per-block sums are the same array on any grid and any number of GPUs.  ~16 blocks per SM of a B200.
    h->passVerts = wb::pass_vertices(V);
    {
        const int passes = div_up(std::max(n, 1), h->passVerts);
        h->vertsPerBlock = h->passVerts * std::max(1, div_up(passes, 148 * 16));
    }
    h->numBlockRows = div_up(std::max(n, 1), h->vertsPerBlock);
    h->cols = wb::block_sums(V) + 1;
    h->numObsTiles = div_up(std::max(n, 1), wb::kObsTile);
    // the first kMomentSample vertices of every tile feed the next quantisation frame
    h->numMomentTiles = h->numObsTiles;
    h->momentCount = 0;
    for (int t = 0; t < h->numObsTiles; ++t) h->momentCount += std::max(0, std::min(wb::kMomentSample, n - t * wb::kObsTile));
    h->momentCount = std::max(h->momentCount, 1);
    h->rowsAlloc = (size_t)h->numBlockRows * h->vertsPerBlock;
    const size_t rows = h->rowsAlloc * V;
    for (float4** p : {&h->x, &h->xNew, &h->mom1, &h->mom2, &h->force}) {
        *p = dalloc<float4>(rows);
        WB_CUDA(cudaMemsetAsync(*p, 0, std::max<size_t>(rows, 1) * sizeof(float4), h->stream));
    }
    h->iw = dalloc<float>(h->rowsAlloc + 8);
    wb::k_fill<float><<<div_up(h->rowsAlloc + 8, 256), 256, 0, h->stream>>>(h->iw, (int64_t)h->rowsAlloc + 8, 1.0f);
    h->weights.assign(n, 1.0);
    h->classMax.assign(n, 1.0);

    h->ctrl = dalloc<wb::StepCtrl>(1);
    h->dyn = dalloc<wb::StepDyn>(1);
    {   // mail: one allocation, so a sharded run exports it with one IPC handle
        size_t off = wb::kMailData;
        auto carve = [&](size_t bytes) { const size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
        const size_t rowsAt = carve(sizeof(double) * h->numBlockRows * h->cols);
        const size_t obsAt = carve(sizeof(double) * h->numObsTiles * 2);
        h->mailBytes = off;
        h->mail = dalloc<char>(off);
        WB_CUDA(cudaMemsetAsync(h->mail, 0, off, h->stream));
        h->pairCounts = reinterpret_cast<unsigned int*>(h->mail + wb::kMailCounts);
        h->blockPartials = reinterpret_cast<double*>(h->mail + rowsAt);
        h->obsPartials = reinterpret_cast<double*>(h->mail + obsAt);
    }
    {   // pair list: room for 8 unordered pairs per vertex (never more than all pairs); grown on demand (collect_step)
        const int64_t all = (int64_t)n * (n - 1) / 2;
        int64_t cap = std::max<int64_t>(1024, std::min<int64_t>({all + 8, (int64_t)8 * n, (int64_t)0x3fffffff}));
        if (const char* e = std::getenv("WB_PAIR_CAP")) cap = std::max<int64_t>(1, std::atoll(e));      // tests: force the growth path
        allocate_pair_list(h, (unsigned int)cap);
    }
    h->repDeg = dalloc<int>(h->rowsAlloc + 8);
    h->repRowPtr = dalloc<long long>(h->rowsAlloc + 1 + 8);
    WB_CUDA(cudaMemsetAsync(h->repDeg, 0, sizeof(int) * (h->rowsAlloc + 8), h->stream));
    WB_CUDA(cudaMemsetAsync(h->repRowPtr, 0, sizeof(long long) * (h->rowsAlloc + 1 + 8), h->stream));
    h->longRows = dalloc<int>(std::max(n, 1));
    h->longCount = dalloc<int>(2);        // [rows queued, cursor of the sorting warps]
    h->scanBlocks = div_up(std::max(n, 1), wb::kScanItems);
    h->scanSums = dalloc<int>(h->scanBlocks + 1);
    h->scanOffsets = dalloc<long long>(h->scanBlocks + 1);
    choose_fixed_scales(h, 1.0, 1.0);

    // Morton keys: as many bits per dimension as fit a 32-bit key
    // 32-bit keys: floor(32 / d) bits per dimension.  Measured alternatives that did not pay and were removed: 64-bit keys (d = 16:
    // 4 instead of 2 bits per dimension, same test counts, dearer sort) and a weight band in the top key bits (c4: separate subtrees
    // per band cost more spatial coherence than the tighter pruning bounds saved, 150 vs 110 ms per step).
    h->mortonBits = std::max(1, std::min(16, 32 / h->dim));
    if (const char* e = std::getenv("WB_MORTON_BITS")) h->mortonBits = std::max(1, std::min(h->mortonBits, std::atoi(e)));
    h->keysIn = dalloc<uint32_t>(n); h->keysOut = dalloc<uint32_t>(n);
    h->valsIn = dalloc<int>(n); h->valsOut = dalloc<int>(n);
    h->cubBytes = 0;
    WB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, h->cubBytes, h->keysIn, h->keysOut, h->valsIn, h->valsOut, std::max(n, 1), 0, 32, h->stream));
    h->cubTemp = dalloc<char>(h->cubBytes);
    h->frameScratch = dalloc<float>((size_t)h->numObsTiles * 4 * wb::kMaxDim);
    h->quant = dalloc<wb::QuantParams>(1);
    WB_CUDA(cudaMemsetAsync(h->quant, 0, sizeof(wb::QuantParams), h->stream));

    // hierarchy: level 0 = points (stride = n rounded up to kFan); level l >= 1 = ceil(count[l-1] / kFan) boxes
    wb::TreeView& t = h->tree;
    std::memset(&t, 0, sizeof(t));
    int count = n, level = 0;
    while (true) {
        t.count[level] = count;
        t.stride[level] = std::max(kFan, div_up(count, kFan) * kFan);
        const size_t planes = (size_t)t.stride[level] * V;
        h->lvlLo[level] = dalloc<float4>(planes);
        wb::k_fill<float4><<<div_up(planes, 256), 256, 0, h->stream>>>(h->lvlLo[level], (int64_t)planes,
                                                                          make_float4(wb::kPadCoord, wb::kPadCoord, wb::kPadCoord, wb::kPadCoord));
        if (level > 0) {
            h->lvlHi[level] = dalloc<float4>(planes);
            wb::k_fill<float4><<<div_up(planes, 256), 256, 0, h->stream>>>(h->lvlHi[level], (int64_t)planes,
                                                                              make_float4(wb::kPadCoord, wb::kPadCoord, wb::kPadCoord, wb::kPadCoord));
        } else {
            h->lvlHi[level] = h->lvlLo[level];
        }
        h->lvlBound[level] = dalloc<float>(t.stride[level]);
        wb::k_fill<float><<<div_up(t.stride[level], 256), 256, 0, h->stream>>>(h->lvlBound[level], t.stride[level], 1.0f);
        t.lo[level] = h->lvlLo[level]; t.hi[level] = h->lvlHi[level]; t.bound[level] = h->lvlBound[level];
        h->planes.lo[level] = h->lvlLo[level]; h->planes.hi[level] = h->lvlHi[level]; h->planes.bound[level] = h->lvlBound[level];
        if (level >= 1 && count <= kFan) break;
        count = std::max(1, div_up(count, kFan));
        ++level;
        if (level >= wb::kMaxLevels - 1) throw wb::CudaError{cudaErrorInvalidValue, "graph too large for the index", __FILE__, __LINE__};
    }
    t.numLevels = level;
    {   // array-of-blocks copy of levels >= 1 (block 0 = null block); padding nodes keep the fill value, which no query passes
        int blocks = 1;
        for (int l = 1; l <= level; ++l) { t.blockOff[l] = blocks; blocks += t.stride[l] / kFan; }
        const size_t f4 = (size_t)blocks * wb::block_float4s(V);
        h->blk = dalloc<float4>(f4);
        WB_DISPATCH_V(V, wb::k_init_blocks<V><<<div_up(f4, 256), 256, 0, h->stream>>>(h->blk, (int64_t)f4));
        t.blk = h->blk;
        // half-precision copy: all-zero records never pass (end position 0)
        const size_t h4 = (size_t)blocks * wb::half_block_float4s(V);
        h->blkH = dalloc<float4>(h4);
        WB_CUDA(cudaMemsetAsync(h->blkH, 0, h4 * sizeof(float4), h->stream));
        t.blkH = h->blkH;
        t.quant = h->quant;
        // the walk's per-warp queries + stacks exceed the 48 KB static limit for the wider rows
        WB_DISPATCH_V(V, WB_CUDA(cudaFuncSetAttribute(wb::k_repulse_pairs<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb::repulse_smem_bytes(V, false)));
                         WB_CUDA(cudaFuncSetAttribute(wb::k_repulse_pairs<V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wb::repulse_smem_bytes(V, true))));
    }
    h->ids = dalloc<int>(t.stride[0]);
    WB_CUDA(cudaMemsetAsync(h->ids, 0xff, sizeof(int) * t.stride[0], h->stream));
    t.ids = h->ids;

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->opt.device);
    // persistent repulsion grid: enough resident blocks to fill every SM, never more than there are chunks
    h->repBlocks = std::max(1, std::min(div_up(div_up(n, 8), 8), sms * 4));
    h->chunkCounter = dalloc<int>(1);
    h->repLayout = wb::RepLayout{1, 0, div_up(std::max(n, 1), 32) * 32};
    h->ownBegin = 0; h->ownEnd = n; h->rowsPerRank = (int)h->rowsAlloc;
    h->fusedBlocks = h->numBlockRows;
    h->forceSums = dalloc<double>(h->cols);
    h->statsTotal = h->cols + wb::kTailStats;
    h->stats = dalloc<double>(h->statsTotal);
    WB_CUDA(cudaMemsetAsync(h->stats, 0, sizeof(double) * h->statsTotal, h->stream));
    for (auto& e : h->ev) WB_CUDA(cudaEventCreate(&e));
    for (auto& e : h->marks) WB_CUDA(cudaEventCreate(&e));
    rebuild_hub_lists(h);
    invalidate_list(h);
    WB_CUDA(cudaStreamSynchronize(h->stream));
}

wb::ForceParams force_params(const wb_embedder* h) {
    wb::ForceParams fp{};
    fp.edgeLength = (float)h->opt.edge_length;
    fp.attractionScale = (float)h->opt.attraction_scale;
    fp.repulsionScale = (float)h->opt.repulsion_scale;
    fp.centreScale = (float)h->opt.centre_scale;
    fp.optimizer = h->opt.optimizer;
    fp.beta1 = 0.9f; fp.beta2 = 0.999f; fp.eps = 1e-8f;      // WembedEmbedder.hpp:46
    fp.maxDisplacement = (float)h->opt.simple_max_displacement;
    fp.seed = h->opt.seed;
    fp.dim = h->dim;
    fp.keepForces = h->opt.keep_forces;
    fp.fixForce = h->fixForce; fp.invFixForce = 1.0 / h->fixForce;
    fp.fixLoss = h->fixLoss; fp.invFixLoss = 1.0 / h->fixLoss;
    fp.dispScale = (float)(h->maxIw / h->opt.edge_length) * 1.000001f;
    return fp;
}

// quantisation frame of the CURRENT positions (first step after wb_set_coordinates, test hook); inside a run it comes out of the
// previous step's recentre pass
void enqueue_frame(wb_embedder* h) {
    cudaStream_t s = h->stream;
    WB_DISPATCH_V(h->V, wb::k_moments<V><<<h->numObsTiles, 256, 0, s>>>(h->x, h->n, wb::kObsTile, h->frameScratch));
    wb::k_quant_params<<<1, 1024, 0, s>>>(h->frameScratch, h->numObsTiles, h->n, h->dim, h->mortonBits, h->halfSigmaLimit, h->quant);
    h->launches += 2;
    h->quantValid = true;
}

// Rebuild the index from the current positions (WembedEmbedder::updateIndex).  pointBound[v] = the pruning weight factor of v: iw[v] for
// forces, iw of v's class maximum for the test hook.  Inside a step every kernel returns at once unless the device decided to rebuild.
void enqueue_index(wb_embedder* h, const float* pointBound, int always) {
    const int n = h->n, V = h->V;
    cudaStream_t s = h->stream;
    const wb::TreeView& t = h->tree;
    WB_DISPATCH_V(V, wb::k_morton_keys<V><<<div_up(n, 256), 256, 0, s>>>(h->x, n, h->dim, h->mortonBits, h->quant, h->keysIn, h->valsIn, h->ctrl, always));
    WB_CUDA(cub::DeviceRadixSort::SortPairs(h->cubTemp, h->cubBytes, h->keysIn, h->keysOut, h->valsIn, h->valsOut, n, 0, h->mortonBits * h->dim, s));
    WB_DISPATCH_V(V, wb::k_build_low<V><<<div_up(t.stride[0], wb::kBuildThreads), wb::kBuildThreads, 0, s>>>(
                         h->x, pointBound, h->valsOut, n, t, h->planes, h->ids, h->heavySlot, h->heavyPos, h->blk, h->blkH, h->ctrl, always));
    h->launches += 3;
    if (t.numLevels >= 4) {
        WB_DISPATCH_V(V, wb::k_build_top<V><<<1, 1024, 0, s>>>(t, h->planes, h->blk, h->blkH, h->ctrl, always));
        h->launches += 1;
    }
    WB_CUDA(cudaGetLastError());
}

PendingStep take_slot(wb_embedder* h) {
    if (!h->freeSlots.empty()) {
        PendingStep p = h->freeSlots.back();
        h->freeSlots.pop_back();
        return p;
    }
    PendingStep p{};
    WB_CUDA(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
    WB_CUDA(cudaMallocHost(&p.host, sizeof(StepSlotHost)));
    return p;
}

// WembedEmbedder::calculateStep (WembedEmbedder.cpp:13-63) as a stream of launches.  `parts` selects which of them are issued, so that
// the same code serves direct launches (everything) and the capture of the step graph (build and rest separately).
// The step in launchable pieces.  The step graph takes Begin | [Build = Search + List] | Rest = Force + Move + Tail (one GPU: no
// exchanges); a local group (wb_step_group) runs piece by piece over all its handles, the exchanges in between without waiting.
enum : int {
    kPartBegin = 1, kPartSearch = 2, kPartXPairs = 4, kPartList = 8, kPartForce = 16, kPartXRows = 32, kPartMove = 64, kPartXCoords = 128, kPartTail = 256,
    kPartBuild = kPartSearch | kPartXPairs | kPartList, kPartRest = kPartForce | kPartXRows | kPartMove | kPartXCoords | kPartTail, kPartAll = 511
};
void launch_parts(wb_embedder* h, int parts, bool assumeBuild) {
    const int n = h->n, V = h->V;
    cudaStream_t s = h->stream;
    const wb::ForceParams fp = force_params(h);
    const bool timing = h->timing && parts == kPartAll;
    // What the host knows: after a blocking step it has read the device's decision for the next one and leaves out the launches of a
    // build that will not happen; with steps in flight it does not know, queues everything, and the kernels of a build return at once
    // on a reuse step (correctness never depends on this knowledge: the device flags alone decide what runs).
    const bool build = assumeBuild || h->nextRebuild != 0 || !h->pending.empty();
    const int waitPeers = h->localGroup ? 0 : 1;    // a local group shares one stream: stream order is the barrier

    if (timing) WB_CUDA(cudaEventRecord(h->ev[0], s));
    if (parts & kPartBegin) {
        wb::k_step_begin<<<1, 32, 0, s>>>(h->ctrl, h->pairCounts + h->rank * wb::kMaxRanks, h->world, h->chunkCounter, h->longCount, cudaGraphConditionalHandle{}, 0);
        h->launches += 1;
    }
    if (build && (parts & kPartSearch)) enqueue_index(h, h->iw, 0);
    if (timing) WB_CUDA(cudaEventRecord(h->ev[1], s));
    const bool sharded = h->world > 1;
    wb::Peers peers{};
    peers.world = h->world; peers.rank = h->rank;
    wb::PairSink sink{};
    wb::PairSource src{};
    wb::Replicas<double> rowsOut{}, obsOut{};
    wb::Replicas<float4> xOut{};
    rowsOut.world = obsOut.world = xOut.world = h->world;
    for (int p = 0; p < h->world; ++p) {
        char* mail = p == h->rank ? h->mail : h->peerMail[p];
        peers.mail[p] = mail;
        rowsOut.at[p] = reinterpret_cast<double*>(mail + (reinterpret_cast<char*>(h->blockPartials) - h->mail));
        obsOut.at[p] = reinterpret_cast<double*>(mail + (reinterpret_cast<char*>(h->obsPartials) - h->mail));
        xOut.at[p] = p == h->rank ? h->x : h->peerX[p];
        // what this rank finds for rank p's vertices goes into segment `rank` of p's buffer; it reads segment p of its own
        sink.seg[p] = (p == h->rank ? h->pairBuf : h->peerPairs[p]) + (size_t)h->rank * h->pairCap;
        src.seg[p] = h->pairBuf + (size_t)p * h->pairCap;
    }
    sink.count = h->pairCounts + h->rank * wb::kMaxRanks; sink.cap = h->pairCap; sink.world = h->world; sink.rowsPerRank = std::max(h->rowsPerRank, 1);
    src.counts = h->pairCounts; src.cap = h->pairCap; src.world = h->world; src.rank = h->rank; src.ownBegin = h->ownBegin; src.ownEnd = h->ownEnd;
    const int repWarps = h->repBlocks * wb::repulse_warps(V);
    if (build && (parts & kPartSearch)) {
        // at least ~8 work units per resident warp, else the tail of the dynamic schedule dominates
        const int64_t residentWarps = (int64_t)h->repBlocks * wb::repulse_warps(V);
        const int queriesPerUnit = h->repLayout.segRows / 32 >= 8 * residentWarps ? 32 : (h->repLayout.segRows / 16 >= 8 * residentWarps ? 16 : 8);
        // both box formats are launched; the one QuantParams::halfBoxes does not name returns at once (the choice is made on the device
        // from the layout, the host never waits for it)
        WB_DISPATCH_V(V, (wb::k_repulse_pairs<V, false><<<h->repBlocks, 32 * wb::repulse_warps(V), wb::repulse_smem_bytes(V, false), s>>>(
                             h->tree, h->rowPtr, h->col, n, sink, h->repLayout, queriesPerUnit, h->heavySlot, h->chunkCounter, h->walkPartials, h->ctrl)));
        WB_DISPATCH_V(V, (wb::k_repulse_pairs<V, true><<<h->repBlocks, 32 * wb::repulse_warps(V), wb::repulse_smem_bytes(V, true), s>>>(
                             h->tree, h->rowPtr, h->col, n, sink, h->repLayout, queriesPerUnit, h->heavySlot, h->chunkCounter, h->walkPartials, h->ctrl)));
        h->launches += 2;
        if (h->numHeavy) {
            WB_DISPATCH_V(V, wb::k_repulse_heavy<V><<<h->numHeavy, 256, 0, s>>>(h->tree, h->rowPtr, h->col, n, sink, h->repLayout, h->heavyVertex, h->heavySlot,
                                                                                 h->heavyPos, h->walkPartials + (size_t)repWarps * 3, h->ctrl));
            h->launches += 1;
        }
    }
    if (sharded && (parts & kPartXPairs)) {   // every rank's pairs have landed in their owners' buffers, and every rank knows all counts
        wb::k_exchange<<<1, 32, 0, s>>>(peers, ++h->epoch, 1, waitPeers, h->ctrl);
        h->launches += 1;
    }
    if (build && (parts & kPartList)) {
        // pair list -> CSR of partners
        const int own = std::max(1, h->ownEnd - h->ownBegin);
        const int pairBlocks = std::max(1, std::min(div_up((int64_t)h->pairCap * h->world, 256), 148 * 8));
        const int scanBlocks = div_up(own, wb::kScanItems);
        wb::k_rep_count<<<pairBlocks, 256, 0, s>>>(src, h->repDeg, h->ctrl);
        wb::k_scan_sums<<<scanBlocks, 256, 0, s>>>(h->repDeg + h->ownBegin, own, h->scanSums, h->ctrl);
        wb::k_scan_offsets<<<1, 1024, 0, s>>>(h->scanSums, h->scanOffsets, scanBlocks, h->ctrl);
        wb::k_scan_apply<<<scanBlocks, 256, 0, s>>>(h->repDeg + h->ownBegin, own, h->scanOffsets, scanBlocks, h->repRowPtr + h->ownBegin, h->ctrl);
        wb::k_rep_fill<<<pairBlocks, 256, 0, s>>>(src, h->repDeg, h->repRowPtr, h->repCol, h->ctrl);
        wb::k_rep_sort_rows<<<div_up(own, 256), 256, 0, s>>>(h->repRowPtr, h->repCol, h->ownBegin, h->ownEnd, h->hubSlot, h->longRows, h->longCount, h->ctrl);
        wb::k_rep_sort_long<<<148, 256, 0, s>>>(h->repRowPtr, h->repCol, reinterpret_cast<int*>(h->pairBuf), h->longRows, h->longCount, h->longCount + 1, h->ctrl);
        h->launches += 7;
    }
    if (timing) WB_CUDA(cudaEventRecord(h->ev[2], s));
    if (h->numHubs && (parts & kPartForce)) {
        WB_DISPATCH_V(V, wb::k_hub_rows<V><<<h->numHubs, 256, 0, s>>>(h->x, h->iw, h->rowPtr, h->col, h->repRowPtr, h->repCol, h->hubVertex, h->ownBegin, h->ownEnd,
                                                                       fp, h->hubD, h->hubF, h->ctrl));
        h->launches += 1;
    }
    const int ownBlocks = div_up(std::max(0, h->ownEnd - h->ownBegin), h->vertsPerBlock);
    if (ownBlocks > 0 && (parts & kPartForce)) {
        WB_DISPATCH_V(V, wb::k_step_fused<V><<<ownBlocks, 256, 0, s>>>(h->x, h->iw, h->rowPtr, h->col, h->repRowPtr, h->repCol, h->ownBegin, h->ownEnd, h->vertsPerBlock, fp,
                                                                        h->dyn, h->hubSlot, h->hubD, h->hubF, h->xNew, h->mom1, h->mom2, h->force,
                                                                        rowsOut, h->ctrl));
        h->launches += 1;
    }
    if (sharded && (parts & kPartXRows)) {   // every rank's sum rows have arrived everywhere
        wb::k_exchange<<<1, 32, 0, s>>>(peers, ++h->epoch, 0, waitPeers, h->ctrl);
        h->launches += 1;
    }
    if (parts & kPartMove) {
        wb::k_reduce_rows<<<h->cols, 256, 0, s>>>(h->blockPartials, h->numBlockRows, h->cols, h->forceSums, h->ctrl);
        h->launches += 1;
    }
    if (timing) WB_CUDA(cudaEventRecord(h->ev[3], s));
    // (a rank that owns no vertices - a small graph on many GPUs - has ownBegin = ownEnd = n, which need not be a tile boundary: it must
    // not touch the last tile, which belongs to the rank before it)
    const int obsBegin = h->ownBegin / wb::kObsTile, obsEnd = h->ownEnd > h->ownBegin ? div_up(h->ownEnd, wb::kObsTile) : obsBegin;
    if (obsEnd > obsBegin && (parts & kPartMove)) {
        h->launches += 1;
        if (sharded) {
            WB_DISPATCH_V(V, (wb::k_recentre_observe<V, true><<<obsEnd - obsBegin, 256, 0, s>>>(h->x, h->xNew, n, obsBegin, h->dim, h->forceSums, h->obsPartials, obsOut,
                                                                                               h->rank, h->ctrl)));
            // the recentred rows -> every other replica of x
            wb::k_publish_rows<<<148 * 4, 256, 0, s>>>(h->x, xOut, h->rank, (int64_t)h->ownBegin * V, (int64_t)(h->ownEnd - h->ownBegin) * V, h->ctrl);
            h->launches += 1;
        } else {
            WB_DISPATCH_V(V, (wb::k_recentre_observe<V, false><<<obsEnd - obsBegin, 256, 0, s>>>(h->x, h->xNew, n, obsBegin, h->dim, h->forceSums, h->obsPartials, obsOut,
                                                                                                0, h->ctrl)));
        }
    }
    if (sharded && (parts & kPartXCoords)) {   // every replica of x is complete, every rank holds all observation tiles
        wb::k_exchange<<<1, 32, 0, s>>>(peers, ++h->epoch, 0, waitPeers, h->ctrl);
        h->launches += 1;
    }
    if (parts & kPartTail) {
        // moments of a sample of the final layout -> the next build's quantisation frame (k_step_tail)
        WB_DISPATCH_V(V, wb::k_moments<V><<<h->numObsTiles, 256, 0, s>>>(h->x, n, wb::kMomentSample, h->frameScratch));
        wb::TailPolicy pol{(float)h->opt.edge_length, h->halfSigmaLimit, h->dim, h->mortonBits};
        wb::k_step_tail<<<1, 1024, 0, s>>>(h->forceSums, h->cols, h->obsPartials, h->numObsTiles, h->frameScratch, h->numMomentTiles, h->momentCount, n, h->walkPartials,
                                            repWarps + h->numHeavy, pol, h->quant, h->ctrl, h->stats);
        h->launches += 2;
    }
    if (timing) WB_CUDA(cudaEventRecord(h->ev[4], s));
    WB_CUDA(cudaGetLastError());
}

// Captures the step into a graph: root = k_step_begin (arms the conditional), IF node = the kernels of a build, child graph = the rest.
// Any failure leaves the handle on direct launches for good (graphNote says why).
void capture_step_graph(wb_embedder* h) {
    cudaStream_t s = h->stream;
    cudaGraph_t g = nullptr, rest = nullptr;
    const int64_t launchesBefore = h->launches;
    auto bail = [&](const char* what, cudaError_t e) {
        h->graphFailed = true;
        h->graphNote = std::string(what) + ": " + cudaGetErrorString(e);
        cudaStreamCaptureStatus st;
        if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(s, &junk); if (junk && junk != g) cudaGraphDestroy(junk); }
        cudaGetLastError();
        if (rest) cudaGraphDestroy(rest);
        if (g) cudaGraphDestroy(g);
        h->launches = launchesBefore;
    };
    cudaError_t e;
    if ((e = cudaGraphCreate(&g, 0)) != cudaSuccess) return bail("cudaGraphCreate", e);
    cudaGraphConditionalHandle handle{};
    if ((e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault)) != cudaSuccess) return bail("cudaGraphConditionalHandleCreate", e);
    cudaGraphNode_t root = nullptr, cond = nullptr, tail = nullptr;
    {
        wb::StepCtrl* ctrl = h->ctrl;
        unsigned int* counts = h->pairCounts + h->rank * wb::kMaxRanks;
        int world = h->world, inGraph = 1;
        int *chunk = h->chunkCounter, *longCount = h->longCount;
        void* args[] = {&ctrl, &counts, &world, &chunk, &longCount, &handle, &inGraph};
        cudaKernelNodeParams kp{};
        kp.func = reinterpret_cast<void*>(wb::k_step_begin);
        kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        if ((e = cudaGraphAddKernelNode(&root, g, nullptr, 0, &kp)) != cudaSuccess) return bail("cudaGraphAddKernelNode", e);
    }
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeIf;
    cp.conditional.size = 1;
    if ((e = cudaGraphAddNode(&cond, g, &root, 1, &cp)) != cudaSuccess) return bail("cudaGraphAddNode (conditional)", e);
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    if ((e = cudaStreamBeginCaptureToGraph(s, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess) return bail("cudaStreamBeginCaptureToGraph", e);
    try { launch_parts(h, kPartBuild, true); } catch (const wb::CudaError& err) { return bail(err.what, err.code); }
    cudaGraph_t got = nullptr;
    if ((e = cudaStreamEndCapture(s, &got)) != cudaSuccess) return bail("cudaStreamEndCapture (build)", e);
    if ((e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed)) != cudaSuccess) return bail("cudaStreamBeginCapture", e);
    try { launch_parts(h, kPartRest, true); } catch (const wb::CudaError& err) { return bail(err.what, err.code); }
    if ((e = cudaStreamEndCapture(s, &rest)) != cudaSuccess) return bail("cudaStreamEndCapture (rest)", e);
    if ((e = cudaGraphAddChildGraphNode(&tail, g, &cond, 1, rest)) != cudaSuccess) return bail("cudaGraphAddChildGraphNode", e);
    cudaGraphExec_t exec = nullptr;
    if ((e = cudaGraphInstantiate(&exec, g, 0)) != cudaSuccess) return bail("cudaGraphInstantiate", e);
    cudaGraphDestroy(rest);
    h->graphKernels = (int)(h->launches - launchesBefore) + 1;
    h->launches = launchesBefore;
    h->stepGraph = g;
    h->stepExec = exec;
}

void launch_step(wb_embedder* h, const PendingStep& slot) {
    cudaStream_t s = h->stream;
    WB_CUDA(cudaMemcpyAsync(h->dyn, &slot.host->dyn, sizeof(wb::StepDyn), cudaMemcpyHostToDevice, s));
    if (!h->quantValid) enqueue_frame(h);
    const bool wantGraph = h->graphWanted && !h->graphFailed && h->world == 1 && !h->timing;
    if (wantGraph && !h->stepExec) capture_step_graph(h);
    if (wantGraph && h->stepExec) {
        WB_CUDA(cudaGraphLaunch(h->stepExec, s));
        h->launches += h->graphKernels;
    } else {
        launch_parts(h, kPartAll, false);
    }
    WB_CUDA(cudaMemcpyAsync(slot.host->sums, h->stats, sizeof(double) * h->statsTotal, cudaMemcpyDeviceToHost, s));
    WB_CUDA(cudaEventRecord(slot.done, s));
}

// host-visible barrier of the ranks of a sharded run (a one-byte all-gather)
void comm_barrier(wb_embedder* h) {
    char* token = dalloc<char>(h->world);
    const bool ok = nccl().allGather(token + h->rank, token, 1, ncclChar, h->comm, h->stream) == ncclSuccess;
    const cudaError_t err = cudaStreamSynchronize(h->stream);
    cudaFree(token);
    if (!ok) throw std::runtime_error("ncclAllGather (barrier) failed");
    WB_CUDA(err);
}

// Sharded run: export this rank's buffers and map everybody else's (CUDA IPC; the handles travel through one small NCCL all-gather,
// which is all NCCL is used for).  all = false: only the pair buffers (they are reallocated when they overflow).
void map_peers(wb_embedder* h, bool all) {
    struct Handles { cudaIpcMemHandle_t mail, x, pairs; };
    Handles mine{};
    if (all) {
        WB_CUDA(cudaIpcGetMemHandle(&mine.mail, h->mail));
        WB_CUDA(cudaIpcGetMemHandle(&mine.x, h->x));
    }
    WB_CUDA(cudaIpcGetMemHandle(&mine.pairs, h->pairBuf));
    char* dSend = dalloc<char>(sizeof(Handles));
    char* dRecv = dalloc<char>(sizeof(Handles) * h->world);
    std::vector<Handles> got(h->world);
    cudaError_t err = cudaMemcpyAsync(dSend, &mine, sizeof(Handles), cudaMemcpyHostToDevice, h->stream);
    bool ncclOk = true;
    if (err == cudaSuccess) ncclOk = nccl().allGather(dSend, dRecv, sizeof(Handles), ncclChar, h->comm, h->stream) == ncclSuccess;
    if (err == cudaSuccess && ncclOk) err = cudaMemcpyAsync(got.data(), dRecv, sizeof(Handles) * h->world, cudaMemcpyDeviceToHost, h->stream);
    if (err == cudaSuccess && ncclOk) err = cudaStreamSynchronize(h->stream);
    cudaFree(dSend); cudaFree(dRecv);
    if (!ncclOk) throw std::runtime_error("ncclAllGather (IPC handles) failed");
    WB_CUDA(err);
    for (int p = 0; p < h->world; ++p) {
        if (p == h->rank) continue;
        void* ptr = nullptr;
        if (all) {
            WB_CUDA(cudaIpcOpenMemHandle(&ptr, got[p].mail, cudaIpcMemLazyEnablePeerAccess)); h->peerMail[p] = static_cast<char*>(ptr);
            WB_CUDA(cudaIpcOpenMemHandle(&ptr, got[p].x, cudaIpcMemLazyEnablePeerAccess)); h->peerX[p] = static_cast<float4*>(ptr);
        }
        if (h->peerPairsOpen) cudaIpcCloseMemHandle(h->peerPairs[p]);
        WB_CUDA(cudaIpcOpenMemHandle(&ptr, got[p].pairs, cudaIpcMemLazyEnablePeerAccess)); h->peerPairs[p] = static_cast<int2*>(ptr);
    }
    if (all) h->peersOpen = true;
    h->peerPairsOpen = true;
    // Opening mappings takes the driver anything from milliseconds to seconds per handle when 8 processes do it at once: nobody may
    // start stepping (and spinning in k_exchange, which gives up after a minute) before everybody has finished
    comm_barrier(h);
}

void collect_step(wb_embedder* h, wb_step_stats* out);

// EmbedderState::nextStep + the scalars of the step (host side)
PendingStep begin_step(wb_embedder* h, double learningRate) {
    h->iteration++;
    PendingStep slot = take_slot(h);
    slot.iteration = h->iteration;
    slot.trivial = h->n <= 1;
    if (slot.trivial) return slot;                            // "Abort in the case of the first hierarchy layer" (:19-21)
    if (h->opt.optimizer == WB_OPT_ADAM) h->adamT++;          // AdamOptimizer.cpp:19
    slot.host->dyn.lr = (float)learningRate;
    slot.host->dyn.invBias1 = (float)(1.0 / (1.0 - std::pow(0.9, h->adamT)));
    slot.host->dyn.invBias2 = (float)(1.0 / (1.0 - std::pow(0.999, h->adamT)));
    slot.host->dyn.iteration = (uint32_t)h->iteration;
    return slot;
}

void enqueue_step(wb_embedder* h, double learningRate) {
    if (h->localGroup) throw std::runtime_error("the handles of a local group step together: wb_step_group");
    PendingStep slot = begin_step(h, learningRate);
    if (slot.trivial) WB_CUDA(cudaEventRecord(slot.done, h->stream));
    else launch_step(h, slot);
    h->pending.push_back(slot);
    h->nextRebuild = 1;                                       // unknown until this step has been collected
}

// One step of a local group (wb_comm_init_local), queued from this one thread on ONE stream, piece by piece over all handles: when a
// handle's kernels of one piece run, every handle's kernels of the piece before have finished, which is what the barrier kernels of
// a real sharded run establish (here they only publish counts and flags).
void step_local_group(wb_embedder** hs, int world, double learningRate, wb_step_stats* out) {
    std::vector<cudaStream_t> own(world);
    cudaStream_t s = hs[0]->stream;
    for (int r = 0; r < world; ++r) { own[r] = hs[r]->stream; hs[r]->stream = s; }
    try {
        std::vector<PendingStep> slots;
        for (int r = 0; r < world; ++r) {
            wb_embedder* h = hs[r];
            slots.push_back(begin_step(h, learningRate));
            WB_CUDA(cudaMemcpyAsync(h->dyn, &slots[r].host->dyn, sizeof(wb::StepDyn), cudaMemcpyHostToDevice, s));
            if (!h->quantValid) enqueue_frame(h);
        }
        const int pieces[] = {kPartBegin | kPartSearch, kPartXPairs, kPartList | kPartForce, kPartXRows, kPartMove, kPartXCoords, kPartTail};
        for (int piece : pieces)
            for (int r = 0; r < world; ++r) launch_parts(hs[r], piece, false);
        for (int r = 0; r < world; ++r) {
            wb_embedder* h = hs[r];
            WB_CUDA(cudaMemcpyAsync(slots[r].host->sums, h->stats, sizeof(double) * h->statsTotal, cudaMemcpyDeviceToHost, s));
            WB_CUDA(cudaEventRecord(slots[r].done, s));
        }
        for (int r = 0; r < world; ++r) { hs[r]->pending.push_back(slots[r]); hs[r]->nextRebuild = 1; }
        for (int r = 0; r < world; ++r) collect_step(hs[r], out ? out + r : nullptr);
    } catch (...) {
        for (int r = 0; r < world; ++r) hs[r]->stream = own[r];
        throw;
    }
    for (int r = 0; r < world; ++r) hs[r]->stream = own[r];
}

// The pair buffer was too small for a build (StepCtrl::overflow): every kernel of that step and of all later ones returned at once, so
// the device state is that of the step before.  Grow the buffer and run the pending steps again, with the scalars they were queued with.
void recover_from_overflow(wb_embedder* h, double needed) {
    WB_CUDA(cudaStreamSynchronize(h->stream));
    // twice what the overflowing build wanted (a dense phase keeps growing for a step or two), or just above it if memory is short
    double want = std::max(2.0 * needed, 2.0 * (double)h->pairCap);
    {
        size_t freeBytes = 0, totalBytes = 0;
        WB_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
        const double room = ((double)freeBytes + 16.0 * (double)h->pairCap * h->world) * 0.9 / (16.0 * h->world);   // pairs that fit once the old buffers are gone
        if (want > room) want = std::max(1.1 * needed, std::min(want, room));
    }
    if (want > 4.0e9) throw std::runtime_error("repulsion pair list exceeds 4e9 pairs per producer");
    if (h->localGroup) throw std::runtime_error("the pair buffer of a local group cannot grow: create the handles with a larger WB_PAIR_CAP");
    if (h->world > 1 && h->peerPairsOpen) {      // nobody may still hold a mapping of a buffer that is about to be freed
        for (int p = 0; p < h->world; ++p)
            if (p != h->rank) cudaIpcCloseMemHandle(h->peerPairs[p]);
        h->peerPairsOpen = false;
        comm_barrier(h);                             // everybody has closed before anybody frees
    }
    allocate_pair_list(h, (unsigned int)want);
    if (h->world > 1) map_peers(h, false);       // collective: every rank sees the same overflow at the same step
    if (std::getenv("WB_DEBUG")) std::fprintf(stderr, "[wb rank %d] pair buffer grown to %u pairs per segment, replaying %zu step(s)\n", h->rank, h->pairCap, h->pending.size());
    invalidate_list(h);
    std::deque<PendingStep> again;
    again.swap(h->pending);
    for (const PendingStep& p : again) {
        launch_step(h, p);
        h->pending.push_back(p);
    }
    h->nextRebuild = 1;
}

void collect_step(wb_embedder* h, wb_step_stats* out) {
    PendingStep slot = h->pending.front();
    const int cols = h->cols;
    for (;;) {
        WB_CUDA(cudaEventSynchronize(slot.done));
        if (slot.trivial || slot.host->sums[cols + 8] == 0.0) break;
        if (slot.host->sums[cols + 8] != 1.0) {
            const long long code = (long long)slot.host->sums[cols + 9];
            throw std::runtime_error("sharded step: rank " + std::to_string(h->rank) + " waited a minute for rank " + std::to_string(code / 1000000) + " at barrier " +
                                     std::to_string(code % 1000000) + " (of " + std::to_string(h->epoch) + " queued) and gave up");
        }
        if (std::getenv("WB_DEBUG")) std::fprintf(stderr, "[wb rank %d] step %lld: pair buffer overflow, %.0f pairs needed, segment capacity %u\n", h->rank, (long long)slot.iteration, slot.host->sums[cols + 9], h->pairCap);
        recover_from_overflow(h, slot.host->sums[cols + 9]);
    }
    h->pending.pop_front();
    wb_step_stats st;
    std::memset(&st, 0, sizeof(st));
    st.iteration = slot.iteration;
    if (!slot.trivial) {
        const double* s = slot.host->sums;
        st.loss_attract = s[0];
        st.loss_repel = s[1];
        st.num_repulsion_pairs = s[2];
        for (int k = 0; k < h->dim; ++k) st.centroid[k] = s[4 + k] / (double)h->n;
        st.max_displacement_ratio = s[cols - 1];
        st.num_listed_pairs = s[cols + 0];
        st.num_candidates = s[cols + 1];
        st.num_box_tests = s[cols + 2];
        st.sum_displacement = s[cols + 3];
        st.sum_radius_sq = s[cols + 4];
        st.list_rebuilt = s[cols + 5];
        st.list_skin = s[cols + 6];
        const double invN = 1.0 / (double)h->n;                              // observeDisplacement (:341-350)
        const double radius = std::sqrt(st.sum_radius_sq * invN);
        st.rel_displacement = radius > 0.0 ? (st.sum_displacement * invN) / radius : 0.0;
        if (h->pending.empty()) h->nextRebuild = s[cols + 7] != 0.0 ? 1 : 0;
        if (h->timing && h->pending.empty()) {
            // events: 0 start | 1 index built | 2 pair list built | 3 forces + optimizer done | 4 recentred
            float ms;
            WB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1])); h->phaseMs[0] = ms;
            WB_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3])); h->phaseMs[1] = ms;
            WB_CUDA(cudaEventElapsedTime(&ms, h->ev[1], h->ev[2])); h->phaseMs[2] = ms;
            h->phaseMs[3] = 0.0;
            WB_CUDA(cudaEventElapsedTime(&ms, h->ev[3], h->ev[4])); h->phaseMs[4] = ms;
            WB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[4])); h->phaseMs[5] = ms;
            h->havePhase = true;
        }
    }
    h->freeSlots.push_back(slot);
    if (out) *out = st;
}

// Coordinates cross the C ABI as row-major n x d doubles in caller-owned (pageable) memory; the device keeps fp32 rows padded to
// 4V floats.  The conversion happens on the host while the rows pass through pinned chunks: a few worker threads each own a
// contiguous share of the rows, one stream and two pinned chunks, so the host pass (read 8 B, write 4 B per value), the PCIe
// copies (half the bytes of the double rows) and the other workers overlap.  Nothing is allocated per call.
constexpr size_t kStageChunkFloats = (size_t)1 << 18;       // 1 MiB per pinned chunk

int stage_lanes(wb_embedder* h, int64_t rows) {
    const int64_t rowsPerChunk = std::max<int64_t>(1, (int64_t)(kStageChunkFloats / (size_t)h->rowFloats));
    const int want = (int)std::max<int64_t>(1, std::min<int64_t>({8, (int64_t)std::max(1u, std::thread::hardware_concurrency()), (rows + rowsPerChunk - 1) / rowsPerChunk}));
    while ((int)h->stageLanes.size() < want) {
        wb_embedder::StageLane l;
        WB_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            WB_CUDA(cudaMallocHost(&l.pinned[b], kStageChunkFloats * sizeof(float)));
            WB_CUDA(cudaEventCreateWithFlags(&l.done[b], cudaEventDisableTiming));
        }
        h->stageLanes.push_back(l);
    }
    return want;
}

// body(lane index, first row, end row) on `lanes` threads over [0, rows); the first error wins
template <typename Body>
void run_lanes(wb_embedder* h, int lanes, int64_t rows, Body&& body) {
    std::vector<cudaError_t> err(lanes, cudaSuccess);
    auto work = [&](int t) {
        cudaSetDevice(h->opt.device);
        const int64_t r0 = rows * t / lanes, r1 = rows * (t + 1) / lanes;
        err[t] = body(t, r0, r1);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < lanes; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (cudaError_t e : err) WB_CUDA(e);
}

void upload_rows(wb_embedder* h, const double* src, float4* dst) {
    const int64_t rows = h->n;
    if (rows == 0) return;
    const int dim = h->dim, rf = h->rowFloats;
    const int64_t rowsPerChunk = std::max<int64_t>(1, (int64_t)(kStageChunkFloats / (size_t)rf));
    WB_CUDA(cudaStreamSynchronize(h->stream));                // nothing on the main stream may still be reading or writing dst
    const int lanes = stage_lanes(h, rows);
    float* out = reinterpret_cast<float*>(dst);
    run_lanes(h, lanes, rows, [&](int t, int64_t r0, int64_t r1) -> cudaError_t {
        auto& l = h->stageLanes[t];
        int b = 0;
        for (int64_t r = r0; r < r1; r += rowsPerChunk, b ^= 1) {
            const int64_t cnt = std::min(rowsPerChunk, r1 - r);
            cudaError_t e = cudaEventSynchronize(l.done[b]);   // the copy that last used this chunk has finished
            if (e != cudaSuccess) return e;
            float* p = l.pinned[b];
            const double* s = src + r * dim;
            if (rf == dim) {
                for (int64_t i = 0; i < cnt * dim; ++i) p[i] = (float)s[i];
            } else {
                for (int64_t v = 0; v < cnt; ++v) {
                    for (int k = 0; k < dim; ++k) p[v * rf + k] = (float)s[v * dim + k];
                    for (int k = dim; k < rf; ++k) p[v * rf + k] = 0.f;
                }
            }
            e = cudaMemcpyAsync(out + r * rf, p, sizeof(float) * cnt * rf, cudaMemcpyHostToDevice, l.stream);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(l.done[b], l.stream);
            if (e != cudaSuccess) return e;
        }
        return cudaStreamSynchronize(l.stream);
    });
}

void download_rows(wb_embedder* h, const float4* src, double* dst) {
    const int64_t rows = h->n;
    if (rows == 0) return;
    const int dim = h->dim, rf = h->rowFloats;
    const int64_t rowsPerChunk = std::max<int64_t>(1, (int64_t)(kStageChunkFloats / (size_t)rf));
    WB_CUDA(cudaStreamSynchronize(h->stream));                // the rows are final
    const int lanes = stage_lanes(h, rows);
    const float* in = reinterpret_cast<const float*>(src);
    run_lanes(h, lanes, rows, [&](int t, int64_t r0, int64_t r1) -> cudaError_t {
        auto& l = h->stageLanes[t];
        auto convert = [&](int b, int64_t r, int64_t cnt) {
            const float* p = l.pinned[b];
            double* d = dst + r * dim;
            if (rf == dim) {
                for (int64_t i = 0; i < cnt * dim; ++i) d[i] = (double)p[i];
            } else {
                for (int64_t v = 0; v < cnt; ++v)
                    for (int k = 0; k < dim; ++k) d[v * dim + k] = (double)p[v * rf + k];
            }
        };
        // two chunks in flight: while chunk i is converted on the host, chunk i + 1 crosses PCIe
        int64_t prevR = -1, prevCnt = 0;
        int b = 0;
        for (int64_t r = r0; r < r1; r += rowsPerChunk, b ^= 1) {
            const int64_t cnt = std::min(rowsPerChunk, r1 - r);
            cudaError_t e = cudaMemcpyAsync(l.pinned[b], in + r * rf, sizeof(float) * cnt * rf, cudaMemcpyDeviceToHost, l.stream);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(l.done[b], l.stream);
            if (e != cudaSuccess) return e;
            if (prevR >= 0) {
                e = cudaEventSynchronize(l.done[b ^ 1]);
                if (e != cudaSuccess) return e;
                convert(b ^ 1, prevR, prevCnt);
            }
            prevR = r; prevCnt = cnt;
        }
        if (prevR >= 0) {
            cudaError_t e = cudaEventSynchronize(l.done[b ^ 1]);
            if (e != cudaSuccess) return e;
            convert(b ^ 1, prevR, prevCnt);
        }
        return cudaSuccess;
    });
}

template <typename F>
int guarded(wb_embedder* h, F&& body) {
    if (!h) return fail(WB_ERR_INVALID, "null handle");
    try {
        WB_CUDA(cudaSetDevice(h->opt.device));
        body();
        return WB_OK;
    } catch (const wb::CudaError& e) {
        return fail(e.code == cudaErrorInvalidValue ? WB_ERR_INVALID : WB_ERR_CUDA,
                    std::string(e.what) + ": " + cudaGetErrorString(e.code) + " (" + e.file + ":" + std::to_string(e.line) + ")");
    } catch (const std::exception& e) {
        return fail(WB_ERR_INVALID, e.what());
    }
}

}  // namespace

extern "C" {

int wb_abi_version(void) { return WB_ABI_VERSION; }

const char* wb_build_info(void) {
    return "wembed_b200 sm_100a fp32 | index: morton-sorted 8-ary box hierarchy | repulsion: pair list + fused pull kernel (cp.async.bulk staged) | cuda " WB_STRINGIFY(CUDART_VERSION);
}

const char* wb_last_error(void) { return g_lastError.c_str(); }

int wb_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

void wb_options_default(wb_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->embedding_dimension = 4;          // EmbedderOptions.hpp:32
    o->optimizer = WB_OPT_ADAM;          // :49
    o->device = 0;
    o->keep_forces = 0;
    o->attraction_scale = 1.0;           // :40
    o->repulsion_scale = 1.0;            // :41
    o->centre_scale = 0.0;               // :43
    o->edge_length = 1.0;                // :44
    o->doubling_factor = 2.0;            // :39
    o->simple_max_displacement = 1.0;    // :51
    o->seed = 0;
}

int wb_create(wb_embedder** out, int32_t n, const int32_t* row_ptr, const int32_t* col, const wb_options* opts) {
    if (!out || !opts || n < 0 || (n > 0 && !row_ptr)) return fail(WB_ERR_INVALID, "wb_create: bad arguments");
    if (opts->embedding_dimension < 1 || opts->embedding_dimension > wb::kMaxDim)
        return fail(WB_ERR_UNSUPPORTED, "wb_create: embedding_dimension must be in 1..32");
    if (wb_device_count() <= opts->device) return fail(WB_ERR_NO_DEVICE, "wb_create: no CUDA device (there is no CPU fallback)");
    static const int32_t zeroRow[1] = {0};
    if (n == 0) row_ptr = zeroRow;
    // the invariants of Graph (Graph.cpp:87-150): monotone offsets, rows strictly ascending, ids in range, no self loops, symmetric
    if (row_ptr[0] != 0) return fail(WB_ERR_INVALID, "wb_create: row_ptr[0] != 0");
    for (int v = 0; v < n; ++v)
        if (row_ptr[v + 1] < row_ptr[v]) return fail(WB_ERR_INVALID, "wb_create: row_ptr not monotone");
    if (row_ptr[n] > 0 && !col) return fail(WB_ERR_INVALID, "wb_create: col is null but row_ptr[n] > 0");
    for (int v = 0; v < n; ++v) {
        for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
            if (col[e] < 0 || col[e] >= n || col[e] == v) return fail(WB_ERR_INVALID, "wb_create: neighbour id out of range or self loop");
            if (e > row_ptr[v] && col[e] <= col[e - 1]) return fail(WB_ERR_INVALID, "wb_create: rows must be strictly ascending");
        }
    }
    // symmetric (Graph stores every undirected edge from both sides): the repulsion's neighbour filter looks at one endpoint's row only
    for (int v = 0; v < n; ++v) {
        for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
            const int u = col[e];
            if (!std::binary_search(col + row_ptr[u], col + row_ptr[u + 1], v)) return fail(WB_ERR_INVALID, "wb_create: the CSR is not symmetric");
        }
    }
    // the walk packs a block / leaf index into 27 bits of a stack entry (walk.cuh: kRefMask)
    if ((int64_t)n > ((int64_t)1 << 29)) return fail(WB_ERR_UNSUPPORTED, "wb_create: graph too large for the index (n > 2^29)");
    auto* h = new wb_embedder();
    h->n = n;
    h->dim = opts->embedding_dimension;
    h->V = (h->dim + 3) / 4;
    h->rowFloats = 4 * h->V;
    h->numDirected = row_ptr[n];
    h->opt = *opts;
    const int rc = guarded(h, [&] { allocate(h, row_ptr, col); });
    if (rc != WB_OK) { free_all(h); delete h; return rc; }
    *out = h;
    return WB_OK;
}

int wb_destroy(wb_embedder* h) {
    if (!h) return WB_OK;
    cudaSetDevice(h->opt.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_all(h);
    delete h;
    return WB_OK;
}

int wb_set_coordinates(wb_embedder* h, const double* coords) {
    if (h && h->n > 0 && !coords) return fail(WB_ERR_INVALID, "wb_set_coordinates: null buffer");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_set_coordinates: collect the asynchronous steps first");
    return guarded(h, [&] {
        upload_rows(h, coords, h->x);
        h->quantValid = false;               // the frame on the device belongs to the old layout
        invalidate_list(h);
        WB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int wb_set_weights(wb_embedder* h, const double* weights) {
    if (h && h->n > 0 && !weights) return fail(WB_ERR_INVALID, "wb_set_weights: null buffer");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_set_weights: collect the asynchronous steps first");
    return guarded(h, [&] {
        const int n = h->n;
        for (int v = 0; v < n; ++v)
            if (!(weights[v] > 0.0) || !std::isfinite(weights[v])) throw std::runtime_error("wb_set_weights: weights must be finite and > 0");
        h->weights.assign(weights, weights + n);
        if (n == 0) return;
        // invExpWeights[v] = 1 / pow(w, 1/d), computed in double like the reference (WembedEmbedder.cpp:127-130)
        std::vector<float> iw(n);
        for (int v = 0; v < n; ++v) iw[v] = (float)(1.0 / std::pow(weights[v], 1.0 / (double)h->dim));
        WB_CUDA(cudaMemcpyAsync(h->iw, iw.data(), sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
        WB_CUDA(cudaStreamSynchronize(h->stream));
        const double minW = *std::min_element(h->weights.begin(), h->weights.end());
        const double maxW = *std::max_element(h->weights.begin(), h->weights.end());
        choose_fixed_scales(h, *std::max_element(iw.begin(), iw.end()), *std::min_element(iw.begin(), iw.end()));
        rebuild_hub_lists(h);
        // weight classes (WeightedIndex::getDoublingWeightBuckets + updateIndices, WeightedIndex.cpp:51-63, 18-32)
        std::vector<double> buckets;
        if (h->opt.doubling_factor > 1.0)
            for (double c = minW * h->opt.doubling_factor; c < maxW; c *= h->opt.doubling_factor) buckets.push_back(c);
        std::vector<double> classMaxOf = buckets;
        classMaxOf.push_back(maxW);
        for (int v = 0; v < n; ++v)
            h->classMax[v] = classMaxOf[std::upper_bound(buckets.begin(), buckets.end(), h->weights[v]) - buckets.begin()];
        h->quantValid = false;               // halfSigmaLimit may have changed
        invalidate_list(h);
        WB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int wb_get_coordinates(wb_embedder* h, double* coords) {
    if (h && h->n > 0 && !coords) return fail(WB_ERR_INVALID, "wb_get_coordinates: null buffer");
    return guarded(h, [&] { download_rows(h, h->x, coords); });
}

int wb_get_weights(wb_embedder* h, double* weights) {
    if (!h) return fail(WB_ERR_INVALID, "null handle");
    std::copy(h->weights.begin(), h->weights.end(), weights);
    return WB_OK;
}

int wb_get_forces(wb_embedder* h, double* forces) {
    if (h && !h->opt.keep_forces) return fail(WB_ERR_INVALID, "wb_get_forces: create the handle with keep_forces = 1");
    return guarded(h, [&] { download_rows(h, h->force, forces); });
}

int wb_reset_optimizer(wb_embedder* h) {
    return guarded(h, [&] {
        const size_t bytes = h->rowsAlloc * h->V * sizeof(float4);
        if (bytes) { WB_CUDA(cudaMemsetAsync(h->mom1, 0, bytes, h->stream)); WB_CUDA(cudaMemsetAsync(h->mom2, 0, bytes, h->stream)); }
        WB_CUDA(cudaStreamSynchronize(h->stream));
        h->adamT = 0;
        h->iteration = 0;
    });
}

int wb_set_iteration(wb_embedder* h, int64_t iteration) {
    if (!h) return fail(WB_ERR_INVALID, "null handle");
    h->iteration = iteration;
    return WB_OK;
}

int wb_step_async(wb_embedder* h, double learning_rate) {
    if (h && (int)h->pending.size() >= WB_MAX_INFLIGHT) return fail(WB_ERR_INVALID, "wb_step_async: too many steps in flight");
    return guarded(h, [&] { enqueue_step(h, learning_rate); });
}

int wb_step_collect(wb_embedder* h, wb_step_stats* stats) {
    if (h && h->pending.empty()) return fail(WB_ERR_INVALID, "wb_step_collect: no step in flight");
    return guarded(h, [&] { collect_step(h, stats); });
}

int wb_step(wb_embedder* h, double learning_rate, wb_step_stats* stats) {
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_step: collect the asynchronous steps first");
    return guarded(h, [&] { enqueue_step(h, learning_rate); collect_step(h, stats); });
}

int wb_step_group(wb_embedder** hs, int32_t world, double learning_rate, wb_step_stats* stats) {
    if (!hs || world < 2 || world > wb::kMaxRanks) return fail(WB_ERR_INVALID, "wb_step_group: 2 <= world <= 8");
    for (int r = 0; r < world; ++r)
        if (!hs[r] || !hs[r]->localGroup || hs[r]->world != world || hs[r]->rank != r || !hs[r]->pending.empty())
            return fail(WB_ERR_INVALID, "wb_step_group: the handles of one wb_comm_init_local call, in its order, nothing in flight");
    return guarded(hs[0], [&] { step_local_group(hs, world, learning_rate, stats); });
}

int wb_synchronize(wb_embedder* h) {
    return guarded(h, [&] { WB_CUDA(cudaStreamSynchronize(h->stream)); });
}

int wb_enable_timing(wb_embedder* h, int enable) {
    if (!h) return fail(WB_ERR_INVALID, "null handle");
    h->timing = enable != 0;
    return WB_OK;
}

int wb_get_phase_times(wb_embedder* h, double* ms6) {
    if (!h || !ms6) return fail(WB_ERR_INVALID, "null argument");
    if (!h->havePhase) return fail(WB_ERR_INVALID, "wb_get_phase_times: enable timing and run a synchronous step first");
    std::copy(h->phaseMs, h->phaseMs + 6, ms6);
    return WB_OK;
}

int wb_set_list_policy(wb_embedder* h, double skin_max, double reuse_steps) {
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_set_list_policy: collect the asynchronous steps first");
    if (h && (!(skin_max >= 0.0) || skin_max > 4.0 || !(reuse_steps >= 1.0))) return fail(WB_ERR_INVALID, "wb_set_list_policy: skin_max in [0, 4], reuse_steps >= 1");
    return guarded(h, [&] {
        h->skinMax = (float)skin_max;
        h->reuseTarget = (float)reuse_steps;
        choose_fixed_scales(h, h->maxIw, h->minIw);      // halfSigmaLimit depends on the largest list radius
        h->quantValid = false;
        invalidate_list(h);
        WB_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int wb_comm_unique_id(char* id128) {
    if (!id128) return fail(WB_ERR_INVALID, "wb_comm_unique_id: null buffer");
    static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId does not fit the ABI buffer");
    if (!nccl().ok) return fail(WB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    if (nccl().getUniqueId(&id) != ncclSuccess) return fail(WB_ERR_CUDA, "ncclGetUniqueId failed");
    std::memset(id128, 0, 128);
    std::memcpy(id128, &id, sizeof(id));
    return WB_OK;
}

int wb_comm_init(wb_embedder* h, const char* id128, int32_t rank, int32_t world) {
    if (h && (!id128 || world < 1 || world > wb::kMaxRanks || rank < 0 || rank >= world)) return fail(WB_ERR_INVALID, "wb_comm_init: bad arguments (1 <= world <= 8)");
    if (h && h->comm) return fail(WB_ERR_INVALID, "wb_comm_init: already initialised");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_comm_init: steps in flight");
    return guarded(h, [&] {
        if (world == 1) return;
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        if (!nccl().ok) throw std::runtime_error("libnccl.so.2 could not be loaded");
        if (nccl().commInitRank(&h->comm, world, id, rank) != ncclSuccess) throw std::runtime_error("ncclCommInitRank failed");
        h->world = world;
        h->rank = rank;
        // whole block rows and whole observation tiles per rank, so that the global rows / tiles have exactly one writer
        const int64_t align = std::lcm((int64_t)h->vertsPerBlock, (int64_t)wb::kObsTile);
        h->rowsPerRank = (int)(((int64_t)div_up(std::max(h->n, 1), world) + align - 1) / align * align);
        h->ownBegin = (int)std::min<int64_t>(h->n, (int64_t)rank * h->rowsPerRank);
        h->ownEnd = (int)std::min<int64_t>(h->n, (int64_t)h->ownBegin + h->rowsPerRank);
        // repulsion queries: blocks of kRepBlockChunks chunks of the sorted order dealt round-robin to the ranks
        const int blocksPerRank = div_up(div_up(div_up(std::max(h->n, 1), 32), wb::kRepBlockChunks), world);
        h->repLayout = wb::RepLayout{world, rank, blocksPerRank * wb::kRepBlockChunks * 32};
        // one segment per producing rank; a rank receives ~1 / world of all pairs, spread over `world` segments
        allocate_pair_list(h, std::max(1024u, (unsigned int)(((uint64_t)h->pairCap * 2 + world - 1) / world)));
        WB_CUDA(cudaMemsetAsync(h->mail, 0, wb::kMailData, h->stream));
        WB_CUDA(cudaStreamSynchronize(h->stream));
        map_peers(h, true);
        invalidate_list(h);
        WB_CUDA(cudaStreamSynchronize(h->stream));
        if (std::getenv("WB_DEBUG")) std::fprintf(stderr, "[wb rank %d/%d] owns [%d, %d), %d rows per rank, %u pairs per segment\n", rank, world, h->ownBegin, h->ownEnd, h->rowsPerRank, h->pairCap);
    });
}

// Test hook: the sharded step with all `world` ranks as handles of THIS process on ONE device (same problem on every handle): the peers'
// buffers are plain device pointers, no IPC, no NCCL.  The handles then step together through wb_step_group (one stream, the pieces
// of the step in lockstep).  Lets the sharded logic be tested where only one GPU is available.
int wb_comm_init_local(wb_embedder** hs, int32_t world) {
    if (!hs || world < 2 || world > wb::kMaxRanks) return fail(WB_ERR_INVALID, "wb_comm_init_local: 2 <= world <= 8");
    for (int r = 0; r < world; ++r) {
        if (!hs[r] || hs[r]->comm || hs[r]->world != 1 || !hs[r]->pending.empty()) return fail(WB_ERR_INVALID, "wb_comm_init_local: handles must be fresh");
        if (hs[r]->n != hs[0]->n || hs[r]->V != hs[0]->V || hs[r]->opt.device != hs[0]->opt.device || hs[r]->n < 2) return fail(WB_ERR_INVALID, "wb_comm_init_local: handles differ");
    }
    return guarded(hs[0], [&] {
        for (int r = 0; r < world; ++r) {
            wb_embedder* h = hs[r];
            h->world = world;
            h->rank = r;
            h->localGroup = true;
            const int64_t align = std::lcm((int64_t)h->vertsPerBlock, (int64_t)wb::kObsTile);
            h->rowsPerRank = (int)(((int64_t)div_up(std::max(h->n, 1), world) + align - 1) / align * align);
            h->ownBegin = (int)std::min<int64_t>(h->n, (int64_t)r * h->rowsPerRank);
            h->ownEnd = (int)std::min<int64_t>(h->n, (int64_t)h->ownBegin + h->rowsPerRank);
            const int blocksPerRank = div_up(div_up(div_up(std::max(h->n, 1), 32), wb::kRepBlockChunks), world);
            h->repLayout = wb::RepLayout{world, r, blocksPerRank * wb::kRepBlockChunks * 32};
            allocate_pair_list(h, std::max(1024u, (unsigned int)(((uint64_t)h->pairCap * 2 + world - 1) / world)));
            WB_CUDA(cudaMemsetAsync(h->mail, 0, wb::kMailData, h->stream));
            WB_CUDA(cudaStreamSynchronize(h->stream));
        }
        for (int r = 0; r < world; ++r) {
            for (int p = 0; p < world; ++p) {
                if (p == r) continue;
                hs[r]->peerMail[p] = hs[p]->mail;
                hs[r]->peerX[p] = hs[p]->x;
                hs[r]->peerPairs[p] = hs[p]->pairBuf;
            }
            invalidate_list(hs[r]);
            WB_CUDA(cudaStreamSynchronize(hs[r]->stream));
        }
    });
}

int wb_get_partition(wb_embedder* h, int32_t* begin, int32_t* end) {
    if (!h || !begin || !end) return fail(WB_ERR_INVALID, "wb_get_partition: null argument");
    *begin = h->ownBegin;
    *end = h->ownEnd;
    return WB_OK;
}

int wb_reconstruction(wb_embedder* h, int32_t count, const int32_t* nodes, double* out2) {
    if (h && (count < 0 || (count > 0 && !nodes) || !out2)) return fail(WB_ERR_INVALID, "wb_reconstruction: bad arguments");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_reconstruction: steps in flight");
    return guarded(h, [&] {
        out2[0] = out2[1] = 0.0;
        const int n = h->n, d = h->dim;
        if (count == 0 || n == 0) return;
        std::vector<int> deg(count);
        int maxDeg = 1;
        for (int i = 0; i < count; ++i) {
            if (nodes[i] < 0 || nodes[i] >= n) throw std::runtime_error("wb_reconstruction: node id out of range");
            deg[i] = h->hostDegree[nodes[i]];
            maxDeg = std::max(maxDeg, deg[i]);
        }
        int capacity = 1;
        while (capacity < maxDeg) capacity <<= 1;
        const int batch = (int)std::max<int64_t>(1, std::min<int64_t>(count, (int64_t)(64 << 20) / capacity));   // <= 1 Gi scratch bytes
        std::vector<double> wroot(n);
        for (int v = 0; v < n; ++v) wroot[v] = std::pow(h->weights[v], 1.0 / (double)d);
        double* dW = dalloc<double>(n);
        int* dNodes = dalloc<int>(count);
        double* dOut = dalloc<double>((size_t)count * 3);
        wb::SimKey* dKeys = dalloc<wb::SimKey>((size_t)batch * capacity);
        int* dCnt = dalloc<int>((size_t)batch * (capacity + 1));
        auto cleanup = [&] { cudaFree(dW); cudaFree(dNodes); cudaFree(dOut); cudaFree(dKeys); cudaFree(dCnt); };
        try {
            cudaStream_t s = h->stream;
            WB_CUDA(cudaMemcpyAsync(dW, wroot.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dNodes, nodes, sizeof(int) * count, cudaMemcpyHostToDevice, s));
            for (int first = 0; first < count; first += batch) {
                WB_DISPATCH_V(h->V, wb::k_reconstruction<V><<<std::min(batch, count - first), 256, 0, s>>>(h->x, dW, h->rowPtr, h->col, n, dNodes, first,
                                                                                                          count, capacity, dKeys, dCnt, dOut));
                h->launches += 1;
            }
            std::vector<double> res((size_t)count * 3);
            WB_CUDA(cudaMemcpyAsync(res.data(), dOut, sizeof(double) * res.size(), cudaMemcpyDeviceToHost, s));
            WB_CUDA(cudaStreamSynchronize(s));
            double a = 0.0, b = 0.0, k = 0.0;      // Toolkit::averageFromVector over the sampled nodes, in sample order
            for (int i = 0; i < count; ++i)
                if (res[3 * i + 2] != 0.0) { a += res[3 * i]; b += res[3 * i + 1]; k += 1.0; }
            if (k > 0.0) { out2[0] = a / k; out2[1] = b / k; }
        } catch (...) { cleanup(); throw; }
        cleanup();
    });
}

int wb_edge_detection(wb_embedder* h, int64_t count, const int32_t* v, const int32_t* w, const uint8_t* is_edge, double* out3) {
    if (h && (count < 0 || count > 0x7fffffff || (count > 0 && (!v || !w || !is_edge)) || !out3))
        return fail(WB_ERR_INVALID, "wb_edge_detection: bad arguments");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_edge_detection: steps in flight");
    return guarded(h, [&] {
        out3[0] = out3[1] = out3[2] = -1.0;                    // the reference's values when nothing was sampled (EdgeDetection.cpp:23-26)
        const int n = h->n, d = h->dim;
        if (count == 0 || n == 0) return;
        std::vector<int> flags((size_t)count);
        double numEdges = 0.0;
        for (int64_t i = 0; i < count; ++i) {
            if (v[i] < 0 || v[i] >= n || w[i] < 0 || w[i] >= n) throw std::runtime_error("wb_edge_detection: vertex id out of range");
            flags[(size_t)i] = is_edge[i] ? 1 : 0;
            numEdges += flags[(size_t)i];
        }
        const double N = (double)n, M = (double)(h->numDirected / 2), noM = N * (N - 1.0) / 2.0 - M;   // EdgeDetection.cpp:7-9
        std::vector<double> wroot(n);
        for (int u = 0; u < n; ++u) wroot[u] = std::pow(h->weights[u], 1.0 / (double)d);
        const int items = (int)count, blocks = std::max(1, std::min(div_up(count, 256), 148 * 8));
        double *dW = dalloc<double>(n), *dSimIn = dalloc<double>(count), *dSimOut = dalloc<double>(count);
        int *dV = dalloc<int>(count), *dU = dalloc<int>(count), *dFlagIn = dalloc<int>(count), *dFlagOut = dalloc<int>(count), *dPrefix = dalloc<int>(count);
        wb::F1Best* dBest = dalloc<wb::F1Best>(blocks + 1);
        void* dTemp = nullptr;
        auto cleanup = [&] {
            for (void* p : {(void*)dW, (void*)dSimIn, (void*)dSimOut, (void*)dV, (void*)dU, (void*)dFlagIn, (void*)dFlagOut, (void*)dPrefix, (void*)dBest, dTemp})
                if (p) cudaFree(p);
        };
        try {
            cudaStream_t s = h->stream;
            size_t sortBytes = 0, scanBytes = 0;
            WB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, dSimIn, dSimOut, dFlagIn, dFlagOut, items, 0, 64, s));
            WB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scanBytes, dFlagOut, dPrefix, items, s));
            WB_CUDA(cudaMalloc(&dTemp, std::max<size_t>(std::max(sortBytes, scanBytes), 1)));
            WB_CUDA(cudaMemcpyAsync(dW, wroot.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dV, v, sizeof(int) * count, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dU, w, sizeof(int) * count, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dFlagIn, flags.data(), sizeof(int) * count, cudaMemcpyHostToDevice, s));
            WB_DISPATCH_V(h->V, wb::k_pair_similarity<V><<<div_up(count, 256), 256, 0, s>>>(h->x, dW, dV, dU, count, dSimIn));
            // std::sort by similarity (EdgeSampler.cpp:62); the radix sort is stable, so ties keep the sampler's order
            WB_CUDA(cub::DeviceRadixSort::SortPairs(dTemp, sortBytes, dSimIn, dSimOut, dFlagIn, dFlagOut, items, 0, 64, s));
            WB_CUDA(cub::DeviceScan::InclusiveSum(dTemp, scanBytes, dFlagOut, dPrefix, items, s));
            wb::k_f1_curve<<<blocks, 256, 0, s>>>(dPrefix, count, numEdges, (double)count - numEdges, M, noM, dBest);
            wb::k_f1_best<<<1, 256, 0, s>>>(dBest, blocks, dBest + blocks);
            h->launches += 3;
            wb::F1Best best{};
            WB_CUDA(cudaMemcpyAsync(&best, dBest + blocks, sizeof(best), cudaMemcpyDeviceToHost, s));
            WB_CUDA(cudaStreamSynchronize(s));
            WB_CUDA(cudaGetLastError());
            out3[0] = best.precision; out3[1] = best.recall; out3[2] = best.f1;
        } catch (...) { cleanup(); throw; }
        cleanup();
    });
}

int wb_mark(wb_embedder* h, int slot) {
    if (h && (slot < 0 || slot >= 8)) return fail(WB_ERR_INVALID, "wb_mark: slot out of range");
    return guarded(h, [&] { WB_CUDA(cudaEventRecord(h->marks[slot], h->stream)); });
}

int wb_elapsed_ms(wb_embedder* h, int from, int to, double* ms) {
    if (h && (from < 0 || from >= 8 || to < 0 || to >= 8 || !ms)) return fail(WB_ERR_INVALID, "wb_elapsed_ms: bad arguments");
    return guarded(h, [&] {
        WB_CUDA(cudaEventSynchronize(h->marks[to]));
        float f = 0.f;
        WB_CUDA(cudaEventElapsedTime(&f, h->marks[from], h->marks[to]));
        *ms = f;
    });
}

int64_t wb_launch_count(wb_embedder* h) { return h ? h->launches : 0; }

int wb_exec_mode(wb_embedder* h, char* note, int32_t note_cap) {
    if (!h) return fail(WB_ERR_INVALID, "null handle");
    if (note && note_cap > 0) { std::strncpy(note, h->graphNote.c_str(), (size_t)note_cap - 1); note[note_cap - 1] = 0; }
    return h->stepExec ? 1 : 0;
}

int wb_query_candidates(wb_embedder* h, int32_t nq, const int32_t* queries, int64_t* out_offsets, int32_t* out_ids, int64_t cap) {
    if (h && (nq < 0 || (nq > 0 && (!queries || !out_offsets)))) return fail(WB_ERR_INVALID, "wb_query_candidates: bad arguments");
    if (h && !h->pending.empty()) return fail(WB_ERR_INVALID, "wb_query_candidates: steps in flight");
    int rc = WB_OK;
    const int g = guarded(h, [&] {
        const int n = h->n;
        for (int i = 0; i < nq; ++i)
            if (queries[i] < 0 || queries[i] >= n) throw std::runtime_error("wb_query_candidates: query id out of range");
        out_offsets[0] = 0;
        if (nq == 0 || n == 0) { for (int i = 0; i < nq; ++i) out_offsets[i + 1] = 0; return; }
        cudaStream_t s = h->stream;
        // index over class-maximum weights: the radius of class c is L * (w_v * maxW_c)^(1/d) (WeightedIndex.cpp:78-80)
        std::vector<float> iwClass(n);
        for (int v = 0; v < n; ++v) iwClass[v] = (float)(1.0 / std::pow(h->classMax[v], 1.0 / (double)h->dim)) * (1.0f - 1e-6f);
        float* dIwClass = dalloc<float>(n);
        double* dW = dalloc<double>(n);
        double* dClassMax = dalloc<double>(n);
        int* dQueries = dalloc<int>(nq);
        int64_t* dCounts = dalloc<int64_t>(nq);
        int64_t* dOffsets = dalloc<int64_t>(nq + 1);
        int* dCursor = dalloc<int>(nq);
        int* dOut = nullptr;
        auto cleanup = [&] { cudaFree(dIwClass); cudaFree(dW); cudaFree(dClassMax); cudaFree(dQueries); cudaFree(dCounts); cudaFree(dOffsets); cudaFree(dCursor); if (dOut) cudaFree(dOut); };
        try {
            WB_CUDA(cudaMemcpyAsync(dIwClass, iwClass.data(), sizeof(float) * n, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dW, h->weights.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dClassMax, h->classMax.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemcpyAsync(dQueries, queries, sizeof(int) * nq, cudaMemcpyHostToDevice, s));
            WB_CUDA(cudaMemsetAsync(dCursor, 0, sizeof(int) * nq, s));
            enqueue_frame(h);
            enqueue_index(h, dIwClass, 1);
            const float pruneL2 = (float)(h->opt.edge_length * h->opt.edge_length) * (1.0f + wb::kPruneSlack);
            const int blocks = div_up((int64_t)nq * kFan, 256);
            WB_DISPATCH_V(h->V, wb::k_candidates<V><<<blocks, 256, 0, s>>>(h->tree, h->x, dW, dClassMax, h->dim, h->opt.edge_length, pruneL2, h->iw,
                                                                            dQueries, nq, dCounts, nullptr, dCursor, nullptr, 0));
            std::vector<int64_t> counts(nq);
            WB_CUDA(cudaMemcpyAsync(counts.data(), dCounts, sizeof(int64_t) * nq, cudaMemcpyDeviceToHost, s));
            WB_CUDA(cudaStreamSynchronize(s));
            for (int i = 0; i < nq; ++i) out_offsets[i + 1] = out_offsets[i] + counts[i];
            const int64_t total = out_offsets[nq];
            if (total > cap || (total > 0 && !out_ids)) { rc = WB_ERR_INVALID; g_lastError = "wb_query_candidates: output capacity too small"; cleanup(); return; }
            if (total > 0) {
                dOut = dalloc<int>(total);
                WB_CUDA(cudaMemcpyAsync(dOffsets, out_offsets, sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, s));
                WB_DISPATCH_V(h->V, wb::k_candidates<V><<<blocks, 256, 0, s>>>(h->tree, h->x, dW, dClassMax, h->dim, h->opt.edge_length, pruneL2, h->iw,
                                                                                dQueries, nq, dCounts, dOffsets, dCursor, dOut, 1));
                WB_CUDA(cudaMemcpyAsync(out_ids, dOut, sizeof(int) * total, cudaMemcpyDeviceToHost, s));
                WB_CUDA(cudaStreamSynchronize(s));
                for (int i = 0; i < nq; ++i) std::sort(out_ids + out_offsets[i], out_ids + out_offsets[i + 1]);
            }
        } catch (...) { cleanup(); throw; }
        cleanup();
        invalidate_list(h);                  // the tree now carries the class bounds: the next step builds its own
        WB_CUDA(cudaStreamSynchronize(h->stream));
    });
    return g != WB_OK ? g : rc;
}

}  // extern "C"
