// Parameter blocks shared by the kernels of the device step.
#pragma once
#include "common.cuh"

namespace wb {

struct QuantParams {          // Morton quantisation frame of the next index build, written at the end of every step
    float lo[kMaxDim];
    float invCell[kMaxDim];
    float centre[kMaxDim];    // per-dimension mean: origin of the half-precision copy of the boxes (0 for padding dimensions)
    int halfBoxes;            // 1: the walk tests the half-precision boxes (layout narrow enough, see k_step_tail)
};

struct TreeView {             // implicit 8-ary box hierarchy, structure-of-planes (see common.cuh)
    int numLevels;            // top level index; level 0 = points
    int count[kMaxLevels];    // real nodes per level
    int stride[kMaxLevels];   // plane stride (count rounded up to kFan)
    const float4* lo[kMaxLevels];   // lo[l][c * stride[l] + node]
    const float4* hi[kMaxLevels];   // hi[0] == lo[0] (points)
    const float* bound[kMaxLevels]; // min over the subtree of the pruning weight factor (iw of points)
    const int* ids;           // sorted position -> vertex id (-1 for padding)
    // The same boxes once more, as array-of-blocks for the repulsion walk: a block = the 8 children of one node,
    // [lo planes: V x 8 float4 | hi planes: V x 8 float4 | meta: 8 x BoxMeta], so all loads of a test share one address register.
    // Blocks of all levels >= 1 live in one buffer; level l starts at blockOff[l]; block 0 is a null block nothing passes.
    const float4* blk;
    int blockOff[kMaxLevels];
    // and once more in half precision (same block numbering), relative to QuantParams::centre, rounded outwards:
    // [lo: HV x 8 chunks of 8 halves | hi: HV x 8 | meta: 8 x BoxMeta], HV = ceil(V / 2); see k_repulse_pairs
    const float4* blkH;
    const QuantParams* quant;
};

// per-child record of a block (16 bytes, read with one 128-bit load)
struct BoxMeta {
    float bound;              // min over the subtree of the pruning weight factor
    uint32_t childRef;        // level >= 2: block holding this node's children; level 1: kLeafFlag | leaf index
    uint32_t endPos;          // one past the last sorted position of the subtree
    float invBound;           // 1 / bound, rounded up (threshold factor of the half-precision box rounds)
};
constexpr uint32_t kLeafFlag = 0x80000000u;
__host__ __device__ constexpr int block_float4s(int V) { return (2 * V + 1) * kFan; }
__host__ __device__ constexpr int half_chunks(int V) { return (V + 1) / 2; }            // 16-byte chunks of 8 halves per row
__host__ __device__ constexpr int half_block_float4s(int V) { return (2 * half_chunks(V) + 1) * kFan; }
// relative slack of a half-precision sum of squares over 8 * half_chunks(V) dimensions, as a factor on the threshold BEFORE it is
// squared: (4 HV + 2) roundings of 2^-11 each, doubled, first-order square root rounded up
__host__ __device__ constexpr float half_margin_root(int V) { return 1.f + (float)(4 * half_chunks(V) + 6) * 4.9e-4f + 1.0e-6f; }

// Options of a run: constant between wb_set_weights calls, passed to the kernels by value.
struct ForceParams {
    float edgeLength;         // L
    float attractionScale, repulsionScale, centreScale;
    int optimizer;            // wb_optimizer (AdamOptimizer.cpp:19-28 / SimpleOptimizer.cpp:13-30)
    float beta1, beta2, eps, maxDisplacement;
    uint32_t seed;            // tie-break generator key (Rand.cpp:29-35)
    int dim;                  // real embedding dimension (<= 4V)
    int keepForces;
    // Repulsive terms are accumulated as 64-bit fixed-point integers (value * 2^k, k chosen by wb_set_weights so that n terms of
    // the largest possible magnitude cannot overflow): integer addition is associative, so the sum of a vertex' terms is bit-identical
    // whatever order the (unordered) list of its partners is stored in.
    double fixForce, invFixForce, fixLoss, invFixLoss;
    float dispScale;          // max_u iw_u / L: ||dx_v|| * iw_v * dispScale = displacement of v in units of its smallest interaction radius
};

// What changes from step to step; lives in device memory (one slot, rewritten by a small copy ahead of every step) so that a
// captured graph of the step can be replayed unchanged.
struct StepDyn {
    float lr, invBias1, invBias2;      // LRScheduler value; 1 / (1 - beta^t) of AdamOptimizer.cpp:23-24
    uint32_t iteration;                // state.currentIteration of this step (tie-break generator key)
};

// Device-resident control block of a handle: state of the repulsion pair list and the decisions the device takes by itself.
//
// The repulsion walk does not apply forces; it appends every unordered non-neighbour pair within the LIST radius
// L (1 + skin) / ws to pairBuf, the pairs are turned into a CSR of partners per vertex, and the fused step kernel evaluates the
// exact predicate of repellingForce (WembedEmbedder.cpp:196-201) on them, pull style like the reference's own loop.  With skin > 0 the
// list stays complete while no vertex has moved more than skin / 2 of its smallest interaction radius since the build
// (||x_v - x_u|| changes by at most the two displacements), so index rebuild and walk are skipped on such steps: the device keeps
// the running bound dispAccum = sum over steps of max_v displacement_v / radius_v and decides at the end of every step what the
// next one does.  Results never depend on skin: a listed pair beyond the exact threshold contributes exactly zero.
struct StepCtrl {
    int overflow;             // sticky: the pair buffer was too small; every kernel of this and later steps returns at once until
                              // the host has grown the buffer and replays the steps (wb_api.cu: collect_step)
    unsigned int pairNeeded;  // pairs the overflowing build wanted to store
    int listValid;            // the CSR of partners is complete for the current positions
    int rebuild;              // decision for the CURRENT step: 1 = index rebuild + walk + list build run, 0 = they return at once
    float skin;               // relative inflation of the list radius of the current / next build
    float dispAccum;          // see above; reset by a build
    float listL2;             // (L (1 + skin))^2 (1 + slack): what the walk prunes and prefilters with
    float pruneL;             // sqrt(listL2)
    float skinMax;            // policy: largest skin ever used (0 disables reuse), see k_step_tail
    float reuseTarget;        // policy: steps a list should live for
    float skinCap;            // policy: current ceiling of the skin (<= skinMax), lowered when a build listed more than pairBudget pairs
    unsigned int pairBudget;  // policy: listed pairs a build with skin > 0 should stay below
    int numBuilds, numReused; // statistics since wb_create
};

}  // namespace wb
