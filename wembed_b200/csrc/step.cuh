// The force / optimizer half of WembedEmbedder::calculateStep (WembedEmbedder.cpp:13-63):
//   pair list -> CSR of repulsion partners      k_rep_count, k_scan_*, k_rep_fill
//   attraction + repulsion + centre force + optimizer, one fused pull-style kernel (north_star's "fused step kernel")
//                                               k_step_fused (+ k_hub_rows for hub rows)          :260-301, AdamOptimizer.cpp:15-30
//   deterministic reductions                    k_reduce_tiles                                    ParallelReduce.hpp:18-37
//   recentre + displacement + next index frame  k_recentre_observe, k_step_tail                   :303-352
#pragma once
#include "mt19937.cuh"
#include "walk.cuh"

namespace wb {

// value -> fixed point (round to nearest even, symmetric in the sign, so a pair's two contributions cancel exactly)
__device__ __forceinline__ long long to_fixed(float term, double scale) { return __double2ll_rn((double)term * scale); }

// ---------------------------------------------------------------------------------------------
// First kernel of every step: resets the work counters of a build.
// In a captured step (CUDA graph) it also arms the conditional node that holds the kernels of a build: they only exist in the
// timeline of steps that rebuild.
__global__ void k_step_begin(StepCtrl* ctrl, unsigned int* pairCounts /* this rank's row of the counts matrix */, int world, int* chunkCounter,
                             int* longCount /* [2]: queue of long pair-list rows, cursor */, cudaGraphConditionalHandle buildNode, int inGraph) {
    const bool build = ctrl->overflow == 0 && ctrl->rebuild != 0;
    if (inGraph && threadIdx.x == 0) cudaGraphSetConditional(buildNode, build ? 1u : 0u);
    if (!build) return;
    if ((int)threadIdx.x < world) pairCounts[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { *chunkCounter = 0; longCount[0] = 0; longCount[1] = 0; }
}

// ---------------------------------------------------------------------------------------------
// Sharded run (wb_comm_init): every rank maps the other ranks' buffers (CUDA IPC over NVLink) and the kernels that produce data other
// ranks need store it straight into the consumers' memory - found pairs into the owners' inboxes (walk.cuh), a block's sums into every
// rank's copy of the sum rows, recentred positions into every replica of x.  What is left of the collectives is a barrier:
// k_exchange publishes "my kernels up to here are done" to every peer and waits until every peer has said the same.  The producing
// kernels end with a system-scope fence in every thread that stored into a peer: the flag travels separately and must not be seen
// before the data (the release half of the barrier; k_exchange fences again before it reads).
// Mail = one buffer per rank: [flags | counts matrix | block sum rows | observation tiles | moment tiles].
constexpr int kMailFlags = 0;                      // int[kMaxRanks]: last barrier epoch each peer has reached
constexpr int kMailCounts = 64;                    // unsigned[kMaxRanks][kMaxRanks]: pairs produced by rank p for rank d
constexpr int kMailData = 64 + 4 * kMaxRanks * kMaxRanks;
struct Peers {
    char* mail[kMaxRanks];
    int world, rank;
};
// where a kernel's output rows go: every rank's copy (one GPU: the only copy)
template <typename T>
struct Replicas {
    T* at[kMaxRanks];
    int world;
};

// wait = 0: only publish (a local group of handles on one stream: stream order already is the barrier).
__global__ void k_exchange(const Peers pm, int epoch, int withCounts, int wait, StepCtrl* ctrl) {
    const int p = threadIdx.x;
    if (p < pm.world) {
        if (withCounts) {                          // my row of the counts matrix -> every rank (mine included: it is the row I counted into)
            const unsigned int* mine = reinterpret_cast<const unsigned int*>(pm.mail[pm.rank] + kMailCounts) + pm.rank * kMaxRanks;
            unsigned int* dst = reinterpret_cast<unsigned int*>(pm.mail[p] + kMailCounts) + pm.rank * kMaxRanks;
            if (p != pm.rank)
                for (int d = 0; d < pm.world; ++d) dst[d] = mine[d];
        }
        __threadfence_system();
        *(reinterpret_cast<volatile int*>(pm.mail[p] + kMailFlags) + pm.rank) = epoch;
        volatile int* flag = reinterpret_cast<volatile int*>(pm.mail[pm.rank] + kMailFlags) + p;
        const long long t0 = clock64();
        while (wait && *flag < epoch) {
            if (clock64() - t0 > (110ll << 30)) { ctrl->pairNeeded = (unsigned int)(p * 1000000 + (epoch % 1000000)); ctrl->overflow = 2; break; }   // ~60 s: a peer died; the host reports which and when
        }
    }
    __syncthreads();
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// Pair list -> CSR of partners (both directions) for the vertices [ownBegin, ownEnd) of this rank.
struct PairSource {
    const int2* seg[kMaxRanks];        // one segment per producing rank (one GPU: the walk's own buffer)
    const unsigned int* counts;        // counts[p * kMaxRanks + d] = pairs rank p produced for rank d (one GPU: counts[0])
    unsigned int cap;
    int world, rank, ownBegin, ownEnd;
};

// degrees; also raises StepCtrl::overflow when a segment ran out of space (the rest of this step and all later ones then return at
// once; the host grows the buffer and replays them)
__global__ void __launch_bounds__(256) k_rep_count(const PairSource src, int* __restrict__ deg, StepCtrl* ctrl) {
    if (build_skipped(ctrl, 0)) return;
    // every rank looks at the whole matrix of counts, so all ranks of a sharded run take the same decision
    unsigned int worst = 0u;
    for (int p = 0; p < src.world; ++p)
        for (int d = 0; d < src.world; ++d) worst = max(worst, src.counts[p * kMaxRanks + d]);
    if (worst > src.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->pairNeeded = worst; ctrl->listValid = 0; ctrl->overflow = 1; }
        return;
    }
    for (int s = 0; s < src.world; ++s) {
        const unsigned int cnt = src.counts[s * kMaxRanks + src.rank];
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const int2 p = src.seg[s][i];
            if (p.x >= src.ownBegin && p.x < src.ownEnd) atomicAdd(deg + p.x, 1);
            if (p.y >= src.ownBegin && p.y < src.ownEnd) atomicAdd(deg + p.y, 1);
        }
    }
}

// exclusive scan of m integers (64-bit offsets: a dense phase can list billions of entries) in three small kernels: sums of 1024-item
// blocks, scan of those sums by one block, local scans
constexpr int kScanItems = 1024;
__global__ void __launch_bounds__(256) k_scan_sums(const int* __restrict__ in, int m, int* __restrict__ blockSums, const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ int sm[8];
    const int base = blockIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = base + k * 256 + threadIdx.x; s += i < m ? in[i] : 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 8; ++w) tot += sm[w]; blockSums[blockIdx.x] = tot; }
}
__global__ void __launch_bounds__(1024) k_scan_offsets(const int* __restrict__ blockSums, long long* __restrict__ blockOffsets, int numBlocks,
                                                       const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ long long sm[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < numBlocks; base += 1024) {
        const int i = base + threadIdx.x;
        const long long val = i < numBlocks ? (long long)blockSums[i] : 0ll;
        long long inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = sm[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            sm[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp > 0 ? sm[warp - 1] : 0ll) + inc - val;
        if (i < numBlocks) blockOffsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += sm[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) blockOffsets[numBlocks] = carry;    // grand total
}
__global__ void __launch_bounds__(256) k_scan_apply(const int* __restrict__ in, int m, const long long* __restrict__ blockOffsets, int numBlocks,
                                                    long long* __restrict__ out /* [m + 1] */, const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ int sm[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int first = blockIdx.x * kScanItems + threadIdx.x * 4;      // four consecutive items per thread
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = first + k < m ? in[first + k] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    long long before = blockOffsets[blockIdx.x] + inc - mine;
    for (int w = 0; w < warp; ++w) before += sm[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (first + k < m) out[first + k] = before; before += v[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[m] = blockOffsets[numBlocks];
}

// entries; leaves deg all zero again for the next build and declares the list valid
__global__ void __launch_bounds__(256) k_rep_fill(const PairSource src, int* __restrict__ deg, const long long* __restrict__ repRowPtr /* indexed by vertex */,
                                                  int* __restrict__ repCol, StepCtrl* ctrl) {
    if (build_skipped(ctrl, 0)) return;
    for (int s = 0; s < src.world; ++s) {
        const unsigned int cnt = src.counts[s * kMaxRanks + src.rank];
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const int2 p = src.seg[s][i];
            if (p.x >= src.ownBegin && p.x < src.ownEnd) repCol[repRowPtr[p.x] + atomicSub(deg + p.x, 1) - 1] = p.y;
            if (p.y >= src.ownBegin && p.y < src.ownEnd) repCol[repRowPtr[p.y] + atomicSub(deg + p.y, 1) - 1] = p.x;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->listValid = 1; ctrl->dispAccum = 0.f; ctrl->numBuilds += 1; }
}

// rows of the pair list -> ascending partner order (the list is appended to by whoever finds a pair first, so the order inside a
// row is arbitrary; sorting it makes the floating-point sum of a vertex' repulsive terms a fixed-order sum).  Rows of up to
// kShortRow entries - nearly all - are sorted by one thread each (insertion sort in registers); longer ones (dense phases early in a
// run: hundreds of partners per vertex for a few steps) are queued and sorted by one warp each in shared memory (bitonic network),
// or in global memory beyond kWarpSortMax entries.  Hub rows (summed in fixed point by k_hub_rows) are left alone.
constexpr int kShortRow = 8;
constexpr int kWarpSortMax = 1024;
__global__ void __launch_bounds__(256) k_rep_sort_rows(const long long* __restrict__ repRowPtr, int* __restrict__ repCol, int ownBegin, int ownEnd,
                                                       const int* __restrict__ hubSlot, int* __restrict__ longRows, int* __restrict__ numLong,
                                                       const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    const int v = ownBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= ownEnd || (hubSlot && hubSlot[v] >= 0)) return;
    const long long b = repRowPtr[v];
    const int len = (int)(repRowPtr[v + 1] - b);
    if (len < 2) return;
    if (len > kShortRow) { longRows[atomicAdd(numLong, 1)] = v; return; }
    int key[kShortRow];
#pragma unroll
    for (int i = 0; i < kShortRow; ++i) key[i] = i < len ? repCol[b + i] : 0x7fffffff;
    // odd-even transposition network on kShortRow registers (fixed indices: stays in registers)
#pragma unroll
    for (int round = 0; round < kShortRow; ++round) {
#pragma unroll
        for (int i = round & 1; i + 1 < kShortRow; i += 2) {
            const int lo = min(key[i], key[i + 1]), hi = max(key[i], key[i + 1]);
            key[i] = lo; key[i + 1] = hi;
        }
    }
#pragma unroll
    for (int i = 0; i < kShortRow; ++i)
        if (i < len) repCol[b + i] = key[i];
}
// one warp per queued row; rows are taken from the queue in any order (each row is sorted independently).  Up to kWarpSortMax entries:
// bitonic network on a padded copy in shared memory.  Beyond: rank sort through `scratch` (partner ids of a row are distinct, so
// the rank of an entry = the number of smaller entries), O(len^2 / 32) per warp - such rows only exist for a few steps of a dense phase.
__global__ void __launch_bounds__(256) k_rep_sort_long(const long long* __restrict__ repRowPtr, int* __restrict__ repCol, int* __restrict__ scratch,
                                                       const int* __restrict__ longRows, const int* __restrict__ numLong, int* __restrict__ cursor,
                                                       const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ int buf[8][kWarpSortMax];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *numLong;
    for (;;) {
        int at = 0;
        if (lane == 0) at = atomicAdd(cursor, 1);
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at >= total) break;
        const int v = longRows[at];
        const long long b = repRowPtr[v];
        const int len = (int)(repRowPtr[v + 1] - b);
        if (len <= kWarpSortMax) {
            int size = 32;
            while (size < len) size <<= 1;
            int* keys = buf[warp];
            for (int i = lane; i < size; i += 32) keys[i] = i < len ? repCol[b + i] : 0x7fffffff;
            __syncwarp();
            for (int k = 2; k <= size; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = lane; i < size; i += 32) {
                        const int partner = i ^ j;
                        if (partner > i) {
                            const bool up = (i & k) == 0;
                            const int a = keys[i], c = keys[partner];
                            if ((a > c) == up) { keys[i] = c; keys[partner] = a; }
                        }
                    }
                    __syncwarp();
                }
            }
            for (int i = lane; i < len; i += 32) repCol[b + i] = keys[i];
        } else {
            for (int i = lane; i < len; i += 32) {
                const int key = repCol[b + i];
                int rank = 0;
                for (int k = 0; k < len; ++k) rank += (int)(repCol[b + k] < key);
                scratch[b + rank] = key;
            }
            __syncwarp();
            __threadfence_block();
            for (int i = lane; i < len; i += 32) repCol[b + i] = scratch[b + i];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Fused step kernel.
//
// Layout of the work: G = V (rounded up to a power of two) lanes share one vertex and lane c owns float4 chunk c of every row it
// touches - its own row, the partners' rows, the force, the Adam moments.  For one partner the G lanes read its row with ONE
// coalesced access (16 B per lane), add their partial squared distances with log2(G) shuffles and each accumulates its own four
// force components; nothing has to be reduced at the end and every lane is busy in the optimizer epilogue.
// A vertex walks its CSR row (attractionForce, :140-172) and then its row of the repulsion pair list (repellingForce, :174-210); both
// loops gather the partner's row and weight (iw is 4n bytes and stays in L2) and share the arithmetic - distance, pair weight
// ws = iw_v iw_u, the hinge at dist ws = L.  Terms are fp32, the sums fp64, every sum is taken in a fixed order (rows of the pair list
// are sorted), so a step is reproducible bit for bit - and independent of the list's skin, because a listed pair beyond the hinge
// adds an exact zero.
// Per-pair arithmetic uses the single-instruction special functions (rsqrt.approx, rcp.approx: relative error <= 2^-22, the size of
// the fp32 rounding of the terms themselves); one-dimensional embeddings take an IEEE path with exact +-1 unit vectors.
//
// A block owns a fixed, GLOBAL run of `vertsPerBlock` consecutive vertices (a multiple of the 256 / G vertices of one pass) and emits
// one row of sums: {lossA, lossR, active pairs, list entries, sum xNew[k]} in double and the largest displacement ratio.
__host__ __device__ constexpr int attract_lanes(int V) { return V <= 1 ? 1 : (V <= 2 ? 2 : (V <= 4 ? 4 : 8)); }
__host__ __device__ constexpr int pass_vertices(int V) { return 256 / attract_lanes(V); }
__host__ __device__ constexpr int block_sums(int V) { return 4 + 4 * V; }         // doubles per block row; + one max column behind them
constexpr int kHubThreshold = 96;         // CSR rows longer than this are summed by one block each (k_hub_rows)

// per-hub record written by k_hub_rows: [attraction force (4V) | lossA | coincident partners | active pairs] as doubles and
// [repulsion force (4V) | lossR] as fixed-point integers
__host__ __device__ constexpr int hub_doubles(int V) { return 4 * V + 3; }
__host__ __device__ constexpr int hub_fixed(int V) { return 4 * V + 1; }

__device__ __forceinline__ float chunk_dist2(float4 a, float4 b) {
    float e, d2;
    e = a.x - b.x; d2 = e * e;
    e = a.y - b.y; d2 = fmaf(e, e, d2);
    e = a.z - b.z; d2 = fmaf(e, e, d2);
    e = a.w - b.w; d2 = fmaf(e, e, d2);
    return d2;
}

#ifndef WB_FUSED_MINBLOCKS
#define WB_FUSED_MINBLOCKS 4
#endif
template <int V>
__global__ void __launch_bounds__(256, WB_FUSED_MINBLOCKS)
k_step_fused(const float4* __restrict__ x, const float* __restrict__ iw, const int* __restrict__ rowPtr, const int* __restrict__ col,
             const long long* __restrict__ repRowPtr, const int* __restrict__ repCol, int rangeBegin, int rangeEnd, int vertsPerBlock,
             const ForceParams fp, const StepDyn* __restrict__ dynp, const int* __restrict__ hubSlot, const double* __restrict__ hubD,
             const long long* __restrict__ hubF, float4* __restrict__ xNew, float4* __restrict__ mom1, float4* __restrict__ mom2,
             float4* __restrict__ forceOut, const Replicas<double> blockPartials /* [block][K + 1] on every rank */,
             const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    constexpr int G = attract_lanes(V), VPW = 32 / G, VPB = 8 * VPW, K = block_sums(V), B = 4;
    __shared__ uint32_t mtState[8][624];
    __shared__ double unitBuf[8][VPW][4 * V];
    __shared__ double redBuf[8][K + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, gi = lane / G;
    const bool chunkLane = c < V;                                  // lanes G > V (V = 3, 5, 6, 7) only take part in the shuffles
    const int vBegin = rangeBegin + blockIdx.x * vertsPerBlock;     // rangeBegin and vertsPerBlock are multiples of VPB
    const int vEnd = min(rangeEnd, vBegin + vertsPerBlock);
    const float L = fp.edgeLength;
    const StepDyn dyn = *dynp;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* xc = x + c;
    double sumLossA = 0.0, sumLossR = 0.0, sumX[4] = {0.0, 0.0, 0.0, 0.0};
    int sumPairs = 0, sumEntries = 0;
    float maxMove = 0.f;

    for (int vBase = vBegin; vBase < vEnd; vBase += VPB) {
        const int v = vBase + warp * VPW + gi;
        const bool valid = v < vEnd;
        const int64_t at = (int64_t)v * V + c;
        float4 xv = zero4;
        float iwv = 1.f;
        double acc[4] = {0.0, 0.0, 0.0, 0.0}, lossA = 0.0, lossR = 0.0;
        int nCoincident = 0, nPairs = 0, e = 0, lenA = 0, lenR = 0, hub = -1;
        const int* repRow = repCol;                                // this vertex' row of the pair list
        if (valid) {
            if (chunkLane) xv = __ldg(x + at);
            iwv = __ldg(iw + v);
            hub = hubSlot ? __ldg(hubSlot + v) : -1;
            e = __ldg(rowPtr + v);
            const long long re = __ldg(repRowPtr + v);
            repRow = repCol + re;
            lenR = (int)(__ldg(repRowPtr + v + 1) - re);
            sumEntries += c == 0 ? lenR : 0;
            if (hub < 0) lenA = __ldg(rowPtr + v + 1) - e; else lenR = 0;
        }
        // ---- attraction over the CSR row (attractionForce, :140-172; neighbours ascending, B rows in flight).  All G lanes of a vertex
        // walk the same entries; the groups of a warp have different row lengths and the shuffles need every lane, so the warp iterates
        // to the longest row of its groups (hub rows are pre-summed).
        int len = lenA;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        for (int i = 0; i < len; i += B) {
            bool has[B];
            int u[B];
            float ws[B], dd[B];
            float4 r[B];
#pragma unroll
            for (int j = 0; j < B; ++j) {
                has[j] = i + j < lenA;
                u[j] = has[j] ? __ldg(col + e + i + j) : 0;
            }
#pragma unroll
            for (int j = 0; j < B; ++j) r[j] = (has[j] && chunkLane) ? __ldg(xc + (int64_t)u[j] * V) : xv;
#pragma unroll
            for (int j = 0; j < B; ++j) ws[j] = has[j] ? iwv * __ldg(iw + u[j]) : 0.f;
#pragma unroll
            for (int j = 0; j < B; ++j) dd[j] = chunkLane ? chunk_dist2(r[j], xv) : 0.f;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < B; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], o);
            }
            // the terms of a batch are added in fp32 (their sum carries the same relative error as each term), the batch sum goes into
            // the fp64 accumulator: one conversion + one DADD per component per batch.  (The CSR row is the same on every step, so
            // this grouping is fixed.)
            float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f, bl = 0.f;
            if (V > 1 || fp.dim > 1) {
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    // squared distances below FLT_MIN (dist < 1.1e-19) are neither coincident (that is d2 == 0 exactly, as with sqrtf)
                    // nor can they exceed the edge length: they contribute nothing and stay away from the flush-to-zero rsqrt
                    const float inv = rsqrt_approx(dd[j]);
                    const float dist = dd[j] * inv;
                    nCoincident += (int)(has[j] && dd[j] == 0.f);                        // :150-155, resolved below
                    const bool act = has[j] && dd[j] >= kFltMin && dist * ws[j] > L;     // :163-168
                    const float sc = act ? fp.attractionScale * ws[j] * inv : 0.f;
                    bx = fmaf(sc, r[j].x - xv.x, bx); by = fmaf(sc, r[j].y - xv.y, by);
                    bz = fmaf(sc, r[j].z - xv.z, bz); bw = fmaf(sc, r[j].w - xv.w, bw);
                    bl += act ? fmaf(-L, rcp_approx(ws[j]), dist) : 0.f;
                }
            } else {                                            // one dimension: exact +-1 unit vectors (VectorOperations.hpp:19-24), IEEE arithmetic
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    if (!has[j]) continue;
                    const float dist = sqrtf(dd[j]);
                    if (dist <= 0.f) { ++nCoincident; continue; }
                    if (dist * ws[j] > L) {
                        bx += copysignf(fp.attractionScale * ws[j], r[j].x - xv.x);
                        bl += dist - L / ws[j];
                    }
                }
            }
            acc[0] += (double)bx; acc[1] += (double)by; acc[2] += (double)bz; acc[3] += (double)bw;
            lossA += (double)bl;
        }
        // ---- repulsion over the vertex' row of the pair list (repellingForce, :174-210; partners ascending, the row is sorted).  Every
        // term goes into the fp64 accumulators by itself: an entry that is listed but beyond the hinge (lists are built with a skin)
        // adds an exact zero, so the sums do not depend on what else is listed.
        int rlen = lenR;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rlen = max(rlen, __shfl_xor_sync(0xffffffffu, rlen, o));
        float lossRf = 0.f;
        for (int i = 0; i < rlen; i += 2) {
            bool has[2];
            int u[2];
            float ws[2], dd[2];
            float4 r[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                has[j] = i + j < lenR;
                u[j] = has[j] ? __ldg(repRow + i + j) : 0;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) r[j] = (has[j] && chunkLane) ? __ldg(xc + (int64_t)u[j] * V) : xv;
#pragma unroll
            for (int j = 0; j < 2; ++j) ws[j] = has[j] ? iwv * __ldg(iw + u[j]) : 1.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) dd[j] = chunkLane ? chunk_dist2(xv, r[j]) : 0.f;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 2; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], o);
            }
            if (V > 1 || fp.dim > 1) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    // IEEE square root and divisions here (a handful of entries per vertex; a vertex near a heavy hub sums thousands of
                    // nearly parallel terms, where the 2^-22 of the approximate forms would show)
                    const float dist = sqrtf(dd[j]);
                    const bool inside = has[j] && dist * ws[j] <= L;                    // :196
                    nCoincident += (int)(has[j] && dist <= 0.f);                        // :183-188, resolved below
                    nPairs += (int)inside;
                    const bool act = inside && dist > 0.f;
                    const float sc = act ? fp.repulsionScale * ws[j] / dist : 0.f;
                    acc[0] += (double)(sc * (xv.x - r[j].x)); acc[1] += (double)(sc * (xv.y - r[j].y));
                    acc[2] += (double)(sc * (xv.z - r[j].z)); acc[3] += (double)(sc * (xv.w - r[j].w));
                    lossRf += act ? L / ws[j] - dist : 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (!has[j]) continue;
                    const float dist = sqrtf(dd[j]);
                    if (dist <= 0.f) { ++nCoincident; ++nPairs; continue; }
                    if (!(dist * ws[j] <= L)) continue;
                    ++nPairs;
                    acc[0] += (double)copysignf(fp.repulsionScale * ws[j], xv.x - r[j].x);
                    lossRf += L / ws[j] - dist;
                }
            }
        }
        lossR = (double)lossRf;
        if (valid && hub >= 0) {                               // hub rows were summed by k_hub_rows
            const double* hd = hubD + (int64_t)hub * hub_doubles(V);
            const long long* hf = hubF + (int64_t)hub * hub_fixed(V);
            if (chunkLane) {
                acc[0] = hd[4 * c] + (double)hf[4 * c] * fp.invFixForce; acc[1] = hd[4 * c + 1] + (double)hf[4 * c + 1] * fp.invFixForce;
                acc[2] = hd[4 * c + 2] + (double)hf[4 * c + 2] * fp.invFixForce; acc[3] = hd[4 * c + 3] + (double)hf[4 * c + 3] * fp.invFixForce;
            }
            lossA = hd[4 * V];
            nCoincident = (int)hd[4 * V + 1];
            nPairs = (int)hd[4 * V + 2];
            lossR = (double)hf[4 * V] * fp.invFixLoss;
        }
        // coincident partners: every one of them adds the same unit vector (generator re-created per pair, :150-155, :183-188).
        // The generator state (624 words) lives in per-warp shared memory; the rare vertices that need it take turns.
        uint32_t todo = __ballot_sync(0xffffffffu, nCoincident > 0 && c == 0);
        if (todo) {
            uint32_t rest = todo;
            while (rest) {
                const int l = __ffs(rest) - 1;
                rest &= rest - 1u;
                if (lane == l) random_unit_vector(mtState[warp], fp.seed, (uint32_t)v, dyn.iteration, fp.dim, unitBuf[warp][gi]);
                __syncwarp();
            }
            if (nCoincident > 0 && chunkLane) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * c + i < fp.dim) acc[i] += nCoincident * unitBuf[warp][gi][4 * c + i];
            }
            __syncwarp();
        }
        float disp2 = 0.f;
        if (valid && chunkLane) {
            float4 f = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
            if (fp.centreScale != 0.f) {                   // :296-301
                f.x = fmaf(-fp.centreScale, xv.x, f.x); f.y = fmaf(-fp.centreScale, xv.y, f.y);
                f.z = fmaf(-fp.centreScale, xv.z, f.z); f.w = fmaf(-fp.centreScale, xv.w, f.w);
            }
            if (fp.keepForces) forceOut[at] = f;
            float4 xn;
            if (fp.optimizer == 1) {
                const float4 m = mom1[at], s = mom2[at];
                const float fe[4] = {f.x, f.y, f.z, f.w};
                float me[4] = {m.x, m.y, m.z, m.w}, se[4] = {s.x, s.y, s.z, s.w};
                float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    me[i] = fp.beta1 * me[i] + (1.f - fp.beta1) * fe[i];
                    se[i] = fp.beta2 * se[i] + (1.f - fp.beta2) * fe[i] * fe[i];
                    const float mHat = me[i] * dyn.invBias1, vHat = se[i] * dyn.invBias2;
                    xe[i] = fmaf(dyn.lr * mHat, rcp_approx(sqrt_approx(vHat) + fp.eps), xe[i]);
                }
                mom1[at] = make_float4(me[0], me[1], me[2], me[3]);
                mom2[at] = make_float4(se[0], se[1], se[2], se[3]);
                xn = make_float4(xe[0], xe[1], xe[2], xe[3]);
            } else {
                const float cap = fp.maxDisplacement;
                xn.x = xv.x + fminf(fmaxf(f.x, -cap), cap) * dyn.lr;
                xn.y = xv.y + fminf(fmaxf(f.y, -cap), cap) * dyn.lr;
                xn.z = xv.z + fminf(fmaxf(f.z, -cap), cap) * dyn.lr;
                xn.w = xv.w + fminf(fmaxf(f.w, -cap), cap) * dyn.lr;
            }
            xNew[at] = xn;
            sumX[0] += (double)xn.x; sumX[1] += (double)xn.y; sumX[2] += (double)xn.z; sumX[3] += (double)xn.w;
            disp2 = chunk_dist2(xn, xv);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) disp2 += __shfl_xor_sync(0xffffffffu, disp2, o);
        if (valid && c == 0) {
            sumLossA += lossA;
            sumLossR += lossR;
            sumPairs += nPairs;
            // how far the vertex moved, in units of its smallest possible interaction radius (StepCtrl)
            maxMove = fmaxf(maxMove, sqrtf(disp2) * iwv * fp.dispScale);
        }
    }
    // fixed-order block reduction: lanes that own the same chunk combine (xor offsets G, 2G, ..), then the 8 warps in order
    double dPairs = (double)sumPairs, dEntries = (double)sumEntries;
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
        sumLossA += __shfl_xor_sync(0xffffffffu, sumLossA, o);
        sumLossR += __shfl_xor_sync(0xffffffffu, sumLossR, o);
        dPairs += __shfl_xor_sync(0xffffffffu, dPairs, o);
        dEntries += __shfl_xor_sync(0xffffffffu, dEntries, o);
        maxMove = fmaxf(maxMove, __shfl_xor_sync(0xffffffffu, maxMove, o));
#pragma unroll
        for (int i = 0; i < 4; ++i) sumX[i] += __shfl_xor_sync(0xffffffffu, sumX[i], o);
    }
    if (lane == 0) { redBuf[warp][0] = sumLossA; redBuf[warp][1] = sumLossR; redBuf[warp][2] = dPairs; redBuf[warp][3] = dEntries; redBuf[warp][K] = (double)maxMove; }
    if (lane < G && chunkLane) {
#pragma unroll
        for (int i = 0; i < 4; ++i) redBuf[warp][4 + 4 * c + i] = sumX[i];
    }
    __syncthreads();
    const int blockRow = vBegin / vertsPerBlock;                    // GLOBAL row: block ranges tile the vertex range from vertex 0
    if (threadIdx.x <= K) {
        double sacc = redBuf[0][threadIdx.x];
        if (threadIdx.x < K) { for (int w = 1; w < 8; ++w) sacc += redBuf[w][threadIdx.x]; }
        else { for (int w = 1; w < 8; ++w) sacc = fmax(sacc, redBuf[w][threadIdx.x]); }
        for (int q = 0; q < blockPartials.world; ++q) blockPartials.at[q][(int64_t)blockRow * (K + 1) + threadIdx.x] = sacc;
        if (blockPartials.world > 1) __threadfence_system();     // stores into peers' memory must have arrived before the kernel counts as done
    }
}

// Hub rows (degree > kHubThreshold, or a heavy vertex with thousands of repulsion partners): one block per hub strides over both
// rows, sums attraction in double and repulsion in fixed point, reduces in a fixed order; k_step_fused picks the record up instead
// of walking the rows itself.
template <int V>
__global__ void __launch_bounds__(256) k_hub_rows(const float4* __restrict__ x, const float* __restrict__ iw, const int* __restrict__ rowPtr,
                                                  const int* __restrict__ col, const long long* __restrict__ repRowPtr, const int* __restrict__ repCol,
                                                  const int* __restrict__ hubVertex, int ownBegin, int ownEnd, const ForceParams fp,
                                                  double* __restrict__ hubD, long long* __restrict__ hubF, const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    constexpr int KD = hub_doubles(V), KF = hub_fixed(V);
    __shared__ double redBuf[8 * KD];
    __shared__ long long redF[8][KF];
    const int v = hubVertex[blockIdx.x];
    if (v < ownBegin || v >= ownEnd) return;                     // another rank's vertex
    float4 xv[V];
    load_row<V>(x, v, xv);
    const float iwv = __ldg(iw + v), L = fp.edgeLength;
    double vals[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) vals[k] = 0.0;
    long long fix[KF];
#pragma unroll
    for (int k = 0; k < KF; ++k) fix[k] = 0ll;
    const int end = __ldg(rowPtr + v + 1);
    for (int e = __ldg(rowPtr + v) + threadIdx.x; e < end; e += 256) {
        const int u = __ldg(col + e);
        const float ws = iwv * __ldg(iw + u);
        float4 xu[V];
        load_row<V>(x, u, xu);
        const float dist = sqrtf(point_dist2<V>(xu, xv));
        if (dist <= 0.f) { vals[4 * V + 1] += 1.0; continue; }                  // :150-155
        if (dist * ws > L) {                                                     // :163-168
            vals[4 * V] += (double)(dist - L / ws);
            if (fp.dim == 1) { vals[0] += (double)copysignf(fp.attractionScale * ws, xu[0].x - xv[0].x); continue; }
            const float s = fp.attractionScale * ws / dist;
#pragma unroll
            for (int c = 0; c < V; ++c) {
                vals[4 * c] += (double)(s * (xu[c].x - xv[c].x)); vals[4 * c + 1] += (double)(s * (xu[c].y - xv[c].y));
                vals[4 * c + 2] += (double)(s * (xu[c].z - xv[c].z)); vals[4 * c + 3] += (double)(s * (xu[c].w - xv[c].w));
            }
        }
    }
    const long long rend = repRowPtr[v + 1];
    for (long long e = repRowPtr[v] + threadIdx.x; e < rend; e += 256) {
        const int u = repCol[e];
        const float ws = iwv * __ldg(iw + u);
        float4 xu[V];
        load_row<V>(x, u, xu);
        const float dist = sqrtf(point_dist2<V>(xv, xu));
        if (dist <= 0.f) { vals[4 * V + 1] += 1.0; vals[4 * V + 2] += 1.0; continue; }   // :183-188
        if (!(dist * ws <= L)) continue;                                                  // :196
        vals[4 * V + 2] += 1.0;
        if (fp.dim == 1) {
            fix[0] += to_fixed(copysignf(fp.repulsionScale * ws, xv[0].x - xu[0].x), fp.fixForce);
        } else {
            const float sc = fp.repulsionScale * ws / dist;
#pragma unroll
            for (int c = 0; c < V; ++c) {
                fix[4 * c] += to_fixed(sc * (xv[c].x - xu[c].x), fp.fixForce); fix[4 * c + 1] += to_fixed(sc * (xv[c].y - xu[c].y), fp.fixForce);
                fix[4 * c + 2] += to_fixed(sc * (xv[c].z - xu[c].z), fp.fixForce); fix[4 * c + 3] += to_fixed(sc * (xv[c].w - xu[c].w), fp.fixForce);
            }
        }
        fix[4 * V] += to_fixed(L / ws - dist, fp.fixLoss);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < KF; ++k) {
        long long s = fix[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) redF[warp][k] = s;
    }
    block_sum<KD, 256>(vals, redBuf, hubD + (int64_t)blockIdx.x * KD);       // (contains the barriers that publish redF)
    if (threadIdx.x < KF) {
        long long s = 0ll;
        for (int w = 0; w < 8; ++w) s += redF[w][threadIdx.x];
        hubF[(int64_t)blockIdx.x * KF + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Deterministic reduction of the per-block sums (util::deterministicSum's role, ParallelReduce.hpp:18-37): block rows are GLOBAL
// (row i = vertices [i S, (i + 1) S), S fixed by n alone), thread t of the reducing block adds the rows t, t + 256, .. in order and
// the 256 thread sums are combined by a fixed tree.  Neither scheduling nor the number of GPUs (whole rows per rank; the rows are
// exchanged) can change a bit of the result.  Block k reduces column k; the last column is a maximum.
__global__ void __launch_bounds__(256) k_reduce_rows(const double* __restrict__ rowSums, int numRows, int cols, double* __restrict__ out,
                                                     const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    __shared__ double sm[256];
    const int k = blockIdx.x;
    const bool isMax = k == cols - 1;
    double s = 0.0;
    for (int g = threadIdx.x; g < numRows; g += 256) { const double v = rowSums[(int64_t)g * cols + k]; s = isMax ? fmax(s, v) : s + v; }
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = isMax ? fmax(sm[threadIdx.x], sm[threadIdx.x + o]) : sm[threadIdx.x] + sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
}

// ---------------------------------------------------------------------------------------------
// applyGravityCentre + observeDisplacement (WembedEmbedder.cpp:303-352): x = xnew - centroid and the sums of ||x - xprev|| and
// ||x||^2.  One block per tile of kObsTile vertices (global tiles: the partial sums do not depend on the grid or on the number of
// GPUs).  forceSums = output of k_reduce_rows ({lossA, lossR, pairs, list entries, sum xnew[k], max displacement}).
// MULTI: a sharded run stores the tile sums into every rank's copy (peers over NVLink); the new rows themselves are published to the
// other replicas of x by k_publish_rows right after this kernel (one coalesced 512-byte store per warp and peer: scattering 16-byte
// stores to 7 peers from inside this kernel ran at a quarter of that).
constexpr int kMomentSample = 256;        // vertices per tile that feed the next quantisation frame (k_moments after this pass)
template <int V, bool MULTI>
__global__ void __launch_bounds__(256) k_recentre_observe(float4* x, const float4* __restrict__ xNew, int n, int tileBegin, int dim,
                                                          const double* __restrict__ forceSums, double* __restrict__ obsPartials /* [tile][2] */,
                                                          const Replicas<double> obsPeers, int rank, const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    __shared__ double redBuf[8 * 2];
    float cen[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) cen[k] = (k < dim) ? (float)(forceSums[4 + k] / (double)n) : 0.f;
    const int tile = tileBegin + blockIdx.x;
    const int vEnd = min(n, (tile + 1) * kObsTile);
    double sums[2] = {0.0, 0.0};
    for (int v = tile * kObsTile + threadIdx.x; v < vEnd; v += 256) {
        float disp2 = 0.f, rad2 = 0.f;
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const int64_t at = (int64_t)v * V + c;
            const float4 a = xNew[at], o = x[at];
            const float4 r = make_float4(a.x - cen[4 * c], a.y - cen[4 * c + 1], a.z - cen[4 * c + 2], a.w - cen[4 * c + 3]);
            x[at] = r;
            disp2 = fmaf(r.x - o.x, r.x - o.x, disp2); disp2 = fmaf(r.y - o.y, r.y - o.y, disp2);
            disp2 = fmaf(r.z - o.z, r.z - o.z, disp2); disp2 = fmaf(r.w - o.w, r.w - o.w, disp2);
            rad2 = fmaf(r.x, r.x, rad2); rad2 = fmaf(r.y, r.y, rad2); rad2 = fmaf(r.z, r.z, rad2); rad2 = fmaf(r.w, r.w, rad2);
        }
        sums[0] += (double)sqrtf(disp2);
        sums[1] += (double)rad2;
    }
    block_sum<2, 256>(sums, redBuf, obsPartials + (int64_t)tile * 2);
    if constexpr (MULTI) {
        if (threadIdx.x < 2) {
            const double v = obsPartials[(int64_t)tile * 2 + threadIdx.x];
            for (int q = 0; q < obsPeers.world; ++q)
                if (q != rank) obsPeers.at[q][(int64_t)tile * 2 + threadIdx.x] = v;
        }
        __threadfence_system();                                    // see k_publish_rows
    }
}

// Sharded run: the rows [first, first + count) of x this rank has just recentred -> every other replica of x.  Every thread ends with a
// system-scope fence: the stores must have ARRIVED in the peers' memory before this kernel counts as done, because the barrier's flags
// (k_exchange) travel separately and must not overtake them in the NVLink fabric.
__global__ void __launch_bounds__(256) k_publish_rows(const float4* __restrict__ x, const Replicas<float4> peers, int rank, int64_t first, int64_t count,
                                                      const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = x[first + i];
        for (int q = 0; q < peers.world; ++q)
            if (q != rank) peers.at[q][first + i] = v;
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// Last kernel of a step (one block): the observation sums, the quantisation frame of the next index build, the device's decision
// about the next step (rebuild the pair list or reuse it, and with which skin) and the record the host reads back.
//
// stats layout (doubles): [0] lossA [1] lossR [2] active pairs [3] list entries [4 .. 4+4V) sum xnew [K] max displacement ratio of this
//   step | then kTailStats values: listed pairs, point tests, box tests, sum displacement, sum radius^2, rebuilt (this step), skin of
//   the current list, next step rebuilds, overflow, pairs needed
constexpr int kTailStats = 10;
struct TailPolicy { float edgeLength; float halfSigmaLimit; int dim; int mortonBits; };

__global__ void __launch_bounds__(1024) k_step_tail(const double* __restrict__ forceSums, int cols, const double* __restrict__ obsPartials, int numObsTiles,
                                                    const float* __restrict__ momentPartials, int numMomentTiles, int momentCount, int n,
                                                    const double* __restrict__ walkPartials, int walkRows, const TailPolicy pol, QuantParams* __restrict__ qp,
                                                    StepCtrl* ctrl, double* __restrict__ stats) {
    __shared__ QuantScratch sc;
    __shared__ double sm[1024];
    __shared__ double res[5];
    if (ctrl->overflow != 0) {
        if (threadIdx.x == 0) { stats[cols + 8] = (double)ctrl->overflow; stats[cols + 9] = (double)ctrl->pairNeeded; }
        return;
    }
    const bool rebuilt = ctrl->rebuild != 0;
    // one pass: thread t adds the observation tiles t, t + 1024, .. (columns 0, 1) and the walk's warps t, t + 1024, .. (integer
    // statistics, columns 2..4), then the 1024 thread sums are combined by a fixed tree of shuffles and one pass over the 32 warps
    double part[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = threadIdx.x; r < numObsTiles; r += 1024) { part[0] += obsPartials[(int64_t)r * 2]; part[1] += obsPartials[(int64_t)r * 2 + 1]; }
    if (rebuilt)
        for (int r = threadIdx.x; r < walkRows; r += 1024) { part[2] += walkPartials[(int64_t)r * 3]; part[3] += walkPartials[(int64_t)r * 3 + 1]; part[4] += walkPartials[(int64_t)r * 3 + 2]; }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
        if ((threadIdx.x & 31) == 0) sm[(threadIdx.x >> 5) * 5 + k] = part[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sm[w * 5 + threadIdx.x];
        res[threadIdx.x] = t;
    }
    __syncthreads();
    // frame of the final layout, from the sample of tiles the recentre pass took moments of (momentCount vertices)
    quant_from_partials(momentPartials, numMomentTiles, momentCount, pol.dim, pol.mortonBits, pol.halfSigmaLimit, qp, sc);
    if (threadIdx.x == 0) {
        const float r = (float)forceSums[cols - 1];                // largest displacement of this step, in smallest interaction radii
        float accum = rebuilt ? r : ctrl->dispAccum + r;           // a list built this step saw the positions BEFORE the step's move
        const float skin = ctrl->skin;
        const bool reuse = ctrl->listValid != 0 && skin > 0.f && accum <= 0.5f * skin * 0.999f && isfinite(accum);
        const double listed = 0.5 * forceSums[3];                  // every listed pair is an entry in both vertices' rows
        for (int k = 0; k < cols; ++k) stats[k] = forceSums[k];
        stats[cols + 0] = listed;
        stats[cols + 1] = res[3]; stats[cols + 2] = res[4];
        stats[cols + 3] = res[0]; stats[cols + 4] = res[1];
        stats[cols + 5] = rebuilt ? 1.0 : 0.0;
        stats[cols + 6] = (double)skin;
        stats[cols + 7] = reuse ? 0.0 : 1.0;
        stats[cols + 8] = 0.0; stats[cols + 9] = 0.0;
        if (reuse) {
            ctrl->numReused += 1;
        } else {
            // skin of the next build: large enough for reuseTarget steps at the current pace, at most skinMax; none at all if even
            // one step would outrun the largest allowed skin (then inflating the radius only costs)
            // (a build whose inflated radius listed more pairs than the budget lowers the ceiling; it recovers slowly)
            float cap = ctrl->skinCap;
            if (rebuilt && skin > 0.f && listed > (double)ctrl->pairBudget) cap = fmaxf(0.02f, 0.75f * skin);
            else cap = fminf(ctrl->skinMax, cap * 1.1f);
            ctrl->skinCap = cap;
            // a skin pays only if the list then lives for reuseTarget steps at the current pace: the search with an inflated radius costs
            // ~(1 + skin)^2 searches (measured at c3: 3.9 x at skin 1) and every listed pair is evaluated every step
            float s = 0.f;
            if (cap > 0.f && isfinite(r) && 2.2f * ctrl->reuseTarget * r <= cap) s = fminf(cap, fmaxf(2.2f * ctrl->reuseTarget * r, 0.05f));
            const float Ls = pol.edgeLength * (1.f + s);
            ctrl->skin = s;
            ctrl->listL2 = Ls * Ls * (1.f + kPruneSlack);
            ctrl->pruneL = sqrtf(ctrl->listL2);
        }
        ctrl->dispAccum = accum;
        ctrl->rebuild = reuse ? 0 : 1;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T* p, int64_t count, T value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = value;
}

}  // namespace wb
