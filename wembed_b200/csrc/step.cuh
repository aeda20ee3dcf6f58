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
__global__ void k_step_begin(StepCtrl* ctrl, unsigned int* pairCounts, int world, int* chunkCounter) {
    if (ctrl->overflow != 0 || ctrl->rebuild == 0) return;
    if ((int)threadIdx.x < world) pairCounts[threadIdx.x] = 0u;
    if (threadIdx.x == 0) *chunkCounter = 0;
}

// ---------------------------------------------------------------------------------------------
// Pair list -> CSR of partners (both directions) for the vertices [ownBegin, ownEnd) of this rank.
struct PairSource {
    const int2* seg[kMaxRanks];        // one segment per producing rank (one GPU: the walk's own buffer)
    const unsigned int* count;         // [world] pairs in each segment
    unsigned int cap;
    int world, ownBegin, ownEnd;
};

// degrees; also raises StepCtrl::overflow when a segment ran out of space (the rest of this step and all later ones then return at
// once; the host grows the buffer and replays them)
__global__ void __launch_bounds__(256) k_rep_count(const PairSource src, int* __restrict__ deg, StepCtrl* ctrl) {
    if (build_skipped(ctrl, 0)) return;
    unsigned int worst = 0u;
    for (int s = 0; s < src.world; ++s) worst = max(worst, src.count[s]);
    if (worst > src.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->pairNeeded = worst; ctrl->listValid = 0; ctrl->overflow = 1; }
        return;
    }
    for (int s = 0; s < src.world; ++s) {
        const unsigned int cnt = src.count[s];
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const int2 p = src.seg[s][i];
            if (p.x >= src.ownBegin && p.x < src.ownEnd) atomicAdd(deg + p.x, 1);
            if (p.y >= src.ownBegin && p.y < src.ownEnd) atomicAdd(deg + p.y, 1);
        }
    }
}

// exclusive scan of m integers in three small kernels: sums of 1024-item blocks, scan of those sums by one block, local scans
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
__global__ void __launch_bounds__(1024) k_scan_offsets(int* __restrict__ blockSums, int numBlocks, const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ int sm[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < numBlocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int val = i < numBlocks ? blockSums[i] : 0;
        int inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sm[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = sm[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            sm[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? sm[warp - 1] : 0) + inc - val;
        if (i < numBlocks) blockSums[i] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += sm[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) blockSums[numBlocks] = carry;       // grand total
}
__global__ void __launch_bounds__(256) k_scan_apply(const int* __restrict__ in, int m, const int* __restrict__ blockSums, int numBlocks,
                                                    int* __restrict__ out /* [m + 1] */, const StepCtrl* __restrict__ ctrl) {
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
    int before = blockSums[blockIdx.x] + inc - mine;
    for (int w = 0; w < warp; ++w) before += sm[w];
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (first + k < m) out[first + k] = before; before += v[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[m] = blockSums[numBlocks];
}

// entries; leaves deg all zero again for the next build and declares the list valid
__global__ void __launch_bounds__(256) k_rep_fill(const PairSource src, int* __restrict__ deg, const int* __restrict__ repRowPtr /* indexed by vertex */,
                                                  int* __restrict__ repCol, StepCtrl* ctrl) {
    if (build_skipped(ctrl, 0)) return;
    for (int s = 0; s < src.world; ++s) {
        const unsigned int cnt = src.count[s];
        for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const int2 p = src.seg[s][i];
            if (p.x >= src.ownBegin && p.x < src.ownEnd) repCol[repRowPtr[p.x] + atomicSub(deg + p.x, 1) - 1] = p.y;
            if (p.y >= src.ownBegin && p.y < src.ownEnd) repCol[repRowPtr[p.y] + atomicSub(deg + p.y, 1) - 1] = p.x;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->listValid = 1; ctrl->dispAccum = 0.f; ctrl->numBuilds += 1; }
}

// ---------------------------------------------------------------------------------------------
// Fused step kernel.
//
// Layout of the work: G = V (rounded up to a power of two) lanes share one vertex and lane c owns float4 chunk c of every row it
// touches - its own row, the partners' rows, the force, the Adam moments.  For one partner the G lanes read its row with ONE
// coalesced access (16 B per lane), add their partial squared distances with log2(G) shuffles and each accumulates its own four
// force components; nothing has to be reduced at the end and every lane is busy in the optimizer epilogue.  A vertex first walks
// its CSR row (attractionForce, :140-172), then its row of the repulsion pair list (repellingForce, :174-210).
//
// Memory: a block works through a contiguous run of tiles of 256 / G vertices.  Everything that is read exactly once - the tile's
// rows of x, m, v, its windows of both row-pointer arrays and of iw, and the entries of both of its CSR rows - is brought to shared
// memory by bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) one tile ahead of the warps, so the only loads that
// occupy registers and scoreboard slots are the gathers of partner rows and partner weights (L2-resident: x is 4nd bytes).
// Sums: attraction in double (terms are fp32); repulsion in 64-bit fixed point, because the pair list is unordered and integer
// addition does not care.  Every tile emits {lossA, lossR, pairs, sum xNew[k], max displacement ratio} for the tile reducer; tiles
// are global (tile i = vertices [i VPB, (i+1) VPB)), so the reduced values do not depend on the grid or on the number of GPUs.
__host__ __device__ constexpr int attract_lanes(int V) { return V <= 1 ? 1 : (V <= 2 ? 2 : (V <= 4 ? 4 : 8)); }
__host__ __device__ constexpr int tile_vertices(int V) { return 256 / attract_lanes(V); }
__host__ __device__ constexpr int tile_sums(int V) { return 3 + 4 * V; }          // + one max column behind them
constexpr int kStageEdges = 2048;         // CSR entries staged per tile (c3: ~1 280 per 128 vertices); the rest is read from global memory
constexpr int kStageRep = 1024;           // pair-list entries staged per tile
constexpr int kHubThreshold = 96;         // CSR rows longer than this are summed by one block each (k_hub_rows)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WB_DONE_%=;\n"
        "bra WB_WAIT_%=;\n"
        "WB_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completion is counted on `bar`
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <int V>
struct StepStage {                        // one tile of one block
    static constexpr int G = attract_lanes(V), VPB = 256 / G;
    float4 x[VPB * V], m[VPB * V], s[VPB * V];
    int col[kStageEdges + 8];
    int rcol[kStageRep + 8];
    int rowPtr[VPB + 8];
    int rrowPtr[VPB + 8];
    float iw[VPB + 8];
};
template <int V>
constexpr size_t step_fused_smem() {
    return 2 * sizeof(StepStage<V>) + 2 * sizeof(uint64_t) + 8 * (4 * V) * sizeof(double) + 2 * 8 * (tile_sums(V) + 1) * sizeof(double);
}

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
#define WB_FUSED_MINBLOCKS 3
#endif
template <int V>
__global__ void __launch_bounds__(256, WB_FUSED_MINBLOCKS)
k_step_fused(const float4* __restrict__ x, const float* __restrict__ iw, const int* __restrict__ rowPtr, const int* __restrict__ col,
             const int* __restrict__ repRowPtr, const int* __restrict__ repCol, int rangeBegin, int rangeEnd, int tilesPerBlock,
             const ForceParams fp, const StepDyn* __restrict__ dynp, const int* __restrict__ hubSlot, const double* __restrict__ hubD,
             const long long* __restrict__ hubF, float4* __restrict__ xNew, float4* __restrict__ mom1, float4* __restrict__ mom2,
             float4* __restrict__ forceOut, double* __restrict__ tilePartials, uint32_t* __restrict__ mtScratch,
             const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    using Stage = StepStage<V>;
    constexpr int G = Stage::G, VPW = 32 / G, VPB = Stage::VPB, K = tile_sums(V), B = 4;
    extern __shared__ __align__(128) unsigned char smemStep[];
    Stage* stage = reinterpret_cast<Stage*>(smemStep);                                 // [2]
    uint64_t* full = reinterpret_cast<uint64_t*>(smemStep + 2 * sizeof(Stage));        // [2]
    double (*unitBuf)[4 * V] = reinterpret_cast<double (*)[4 * V]>(full + 2);          // [8]
    double (*redBuf)[8][K + 1] = reinterpret_cast<double (*)[8][K + 1]>(unitBuf + 8);  // [2][8][K + 1], by tile parity
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, gi = lane / G;
    const bool chunkLane = c < V;                                  // lanes G > V (V = 3, 5, 6, 7) only take part in the shuffles
    const int vBegin = rangeBegin + blockIdx.x * tilesPerBlock * VPB;      // rangeBegin is a multiple of VPB
    const int vEnd = min(rangeEnd, vBegin + tilesPerBlock * VPB);
    const int passes = vBegin < vEnd ? (vEnd - vBegin + VPB - 1) / VPB : 0;
    const float L = fp.edgeLength;
    const StepDyn dyn = *dynp;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* xc = x + c;

    // producer state (thread 0 only): bounds of both CSR rows of the tile to be copied next
    int nextLo = 0, nextHi = 0, nextRLo = 0, nextRHi = 0;
    auto passBounds = [&](int p, int& lo, int& hi, int& rlo, int& rhi) {
        const int v0 = vBegin + p * VPB, v1 = min(v0 + VPB, vEnd);
        lo = __ldg(rowPtr + v0); hi = __ldg(rowPtr + v1);
        rlo = repRowPtr[v0]; rhi = repRowPtr[v1];
    };
    auto issue = [&](int p, int lo, int hi, int rlo, int rhi) {      // bulk copies of tile p into stage p & 1
        Stage& st = stage[p & 1];
        uint64_t* bar = full + (p & 1);
        const int v0 = vBegin + p * VPB, rows = min(VPB, vEnd - v0);
        const uint32_t rowBytes = (uint32_t)rows * V * 16u;
        const uint32_t rpBytes = (uint32_t)((rows + 1 + 3) & ~3) * 4u;          // v0 is a multiple of 4: the windows start aligned
        const uint32_t iwBytes = (uint32_t)((rows + 3) & ~3) * 4u;
        const int e0 = lo & ~3, r0 = rlo & ~3;                                  // entries from an aligned entry
        const uint32_t edgeBytes = (uint32_t)((max(min(hi - e0, kStageEdges + 4), 0) + 3) & ~3) * 4u;
        const uint32_t repBytes = (uint32_t)((max(min(rhi - r0, kStageRep + 4), 0) + 3) & ~3) * 4u;
        mbar_expect_tx(bar, 3u * rowBytes + 2u * rpBytes + iwBytes + edgeBytes + repBytes);
        bulk_copy(st.x, x + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.m, mom1 + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.s, mom2 + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.rowPtr, rowPtr + v0, rpBytes, bar);
        bulk_copy(st.rrowPtr, repRowPtr + v0, rpBytes, bar);
        bulk_copy(st.iw, iw + v0, iwBytes, bar);
        if (edgeBytes) bulk_copy(st.col, col + e0, edgeBytes, bar);
        if (repBytes) bulk_copy(st.rcol, repCol + r0, repBytes, bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && passes > 0) {
        int lo, hi, rlo, rhi;
        passBounds(0, lo, hi, rlo, rhi);
        issue(0, lo, hi, rlo, rhi);
        if (passes > 1) passBounds(1, nextLo, nextHi, nextRLo, nextRHi);
    }

    for (int p = 0; p < passes; ++p) {
        // every warp has left stage (p + 1) & 1 (barrier at the end of tile p - 1): refill it, and fetch the bounds after that
        if (threadIdx.x == 0 && p + 1 < passes) {
            issue(p + 1, nextLo, nextHi, nextRLo, nextRHi);
            if (p + 2 < passes) passBounds(p + 2, nextLo, nextHi, nextRLo, nextRHi);
        }
        mbar_wait(full + (p & 1), (uint32_t)(p >> 1) & 1u);
        const Stage& st = stage[p & 1];
        const int v0 = vBegin + p * VPB;
        const int slot = warp * VPW + gi, v = v0 + slot;
        const bool valid = v < vEnd;
        const int e0 = st.rowPtr[0] & ~3, r0 = st.rrowPtr[0] & ~3;             // global index of staged entry 0 of either row
        float4 xv = zero4;
        float iwv = 1.f;
        double acc[4] = {0.0, 0.0, 0.0, 0.0}, loss = 0.0;
        int nCoincident = 0, nPairs = 0, e = 0, end = 0, re = 0, rend = 0, hub = -1;
        if (valid) {
            if (chunkLane) xv = st.x[slot * V + c];
            iwv = st.iw[slot];
            hub = hubSlot ? __ldg(hubSlot + v) : -1;
            if (hub < 0) { e = st.rowPtr[slot]; end = st.rowPtr[slot + 1]; re = st.rrowPtr[slot]; rend = st.rrowPtr[slot + 1]; }
        }
        // ---- attraction over the CSR row (neighbours ascending, B rows in flight).  All G lanes of a vertex walk the same entries;
        // the groups of a warp have different row lengths and the shuffles need every lane, so the warp iterates to the longest row.
        int len = end - e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        for (int i = 0; i < len; i += B) {
            bool has[B];
            int u[B];
            float wsE[B], dd[B];
            float4 r[B];
#pragma unroll
            for (int j = 0; j < B; ++j) {
                const int idx = e + i + j, at = idx - e0;
                has[j] = idx < end;
                u[j] = has[j] ? (at < kStageEdges + 4 ? st.col[at] : __ldg(col + idx)) : 0;
            }
#pragma unroll
            for (int j = 0; j < B; ++j) r[j] = (has[j] && chunkLane) ? __ldg(xc + (int64_t)u[j] * V) : xv;
#pragma unroll
            for (int j = 0; j < B; ++j) wsE[j] = has[j] ? iwv * __ldg(iw + u[j]) : 0.f;
#pragma unroll
            for (int j = 0; j < B; ++j) dd[j] = chunkLane ? chunk_dist2(r[j], xv) : 0.f;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < B; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], o);
            }
            // the terms of a batch are added in fp32 (their sum carries the same relative error as each term), the batch
            // sum goes into the double accumulator: one conversion + one DADD per component per batch
            float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f, bl = 0.f;
            if (V > 1 || fp.dim > 1) {
                // Branch-free pair arithmetic with single-instruction rsqrt / rcp (relative error <= 2^-22, the size of the fp32
                // rounding of the terms themselves).  Squared distances below FLT_MIN are neither coincident (that is d2 == 0
                // exactly, as with sqrtf) nor can they exceed the edge length: they contribute nothing and stay away from the
                // flush-to-zero rsqrt.
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const float inv = rsqrt_approx(dd[j]);
                    const float dist = dd[j] * inv;
                    nCoincident += (int)(has[j] && dd[j] == 0.f);                        // :150-155, resolved below
                    const bool act = has[j] && dd[j] >= kFltMin && dist * wsE[j] > L;     // :163-168
                    const float sc = act ? fp.attractionScale * wsE[j] * inv : 0.f;
                    bx = fmaf(sc, r[j].x - xv.x, bx); by = fmaf(sc, r[j].y - xv.y, by);
                    bz = fmaf(sc, r[j].z - xv.z, bz); bw = fmaf(sc, r[j].w - xv.w, bw);
                    bl += act ? fmaf(-L, rcp_approx(wsE[j]), dist) : 0.f;
                }
            } else {                                            // one dimension: exact +-1 unit vectors (VectorOperations.hpp:19-24), IEEE arithmetic
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    if (!has[j]) continue;
                    const float dist = sqrtf(dd[j]);
                    if (dist <= 0.f) { ++nCoincident; continue; }
                    if (dist * wsE[j] > L) {
                        bx += copysignf(fp.attractionScale * wsE[j], r[j].x - xv.x);
                        bl += dist - L / wsE[j];
                    }
                }
            }
            acc[0] += (double)bx; acc[1] += (double)by; acc[2] += (double)bz; acc[3] += (double)bw;
            loss += (double)bl;
        }
        // ---- repulsion over the vertex' row of the pair list (unordered; exact predicate of repellingForce, :183-201)
        long long rep[4] = {0ll, 0ll, 0ll, 0ll}, lossR = 0ll;
        int rlen = rend - re;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rlen = max(rlen, __shfl_xor_sync(0xffffffffu, rlen, o));
        for (int i = 0; i < rlen; i += 2) {
            bool has[2];
            int u[2];
            float iwu[2], dd[2];
            float4 r[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int idx = re + i + j, at = idx - r0;
                has[j] = idx < rend;
                u[j] = has[j] ? (at < kStageRep + 4 ? st.rcol[at] : repCol[idx]) : 0;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) r[j] = (has[j] && chunkLane) ? __ldg(xc + (int64_t)u[j] * V) : xv;
#pragma unroll
            for (int j = 0; j < 2; ++j) iwu[j] = has[j] ? __ldg(iw + u[j]) : 1.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) dd[j] = chunkLane ? chunk_dist2(xv, r[j]) : 0.f;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 2; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], o);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (!has[j]) continue;
                const float dist = sqrtf(dd[j]), ws = iwv * iwu[j];
                if (dist <= 0.f) { ++nCoincident; ++nPairs; continue; }          // :183-188
                if (!(dist * ws <= L)) continue;                                 // :196 (listed with a skin: most entries end here)
                ++nPairs;
                if (fp.dim == 1) {
                    rep[0] += to_fixed(copysignf(fp.repulsionScale * ws, xv.x - r[j].x), fp.fixForce);
                } else {
                    const float sc = fp.repulsionScale * ws / dist;
                    rep[0] += to_fixed(sc * (xv.x - r[j].x), fp.fixForce); rep[1] += to_fixed(sc * (xv.y - r[j].y), fp.fixForce);
                    rep[2] += to_fixed(sc * (xv.z - r[j].z), fp.fixForce); rep[3] += to_fixed(sc * (xv.w - r[j].w), fp.fixForce);
                }
                lossR += to_fixed(L / ws - dist, fp.fixLoss);
            }
        }
        if (valid && hub >= 0) {                               // hub rows were summed by k_hub_rows
            const double* hd = hubD + (int64_t)hub * hub_doubles(V);
            const long long* hf = hubF + (int64_t)hub * hub_fixed(V);
            if (chunkLane) {
                acc[0] = hd[4 * c]; acc[1] = hd[4 * c + 1]; acc[2] = hd[4 * c + 2]; acc[3] = hd[4 * c + 3];
                rep[0] = hf[4 * c]; rep[1] = hf[4 * c + 1]; rep[2] = hf[4 * c + 2]; rep[3] = hf[4 * c + 3];
            }
            loss = hd[4 * V];
            nCoincident = (int)hd[4 * V + 1];
            nPairs = (int)hd[4 * V + 2];
            lossR = hf[4 * V];
        }
        // coincident partners: every one of them adds the same unit vector (generator re-created per pair, :150-155, :183-188)
        uint32_t todo = __ballot_sync(0xffffffffu, nCoincident > 0 && c == 0);
        while (todo) {                                          // one vertex at a time; the generator state lives in global scratch
            const int l = __ffs(todo) - 1;
            todo &= todo - 1u;
            if (lane == l)
                random_unit_vector(mtScratch + ((size_t)blockIdx.x * 8 + warp) * 624, fp.seed, (uint32_t)v, dyn.iteration, fp.dim, unitBuf[warp]);
            __syncwarp();
            if (lane / G == l / G && chunkLane) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * c + i < fp.dim) acc[i] += nCoincident * unitBuf[warp][4 * c + i];
            }
            __syncwarp();
        }
        // ---- optimizer epilogue and the tile's sums
        double sumX[4] = {0.0, 0.0, 0.0, 0.0}, sumLossA = 0.0, sumLossR = 0.0, sumPairs = 0.0;
        float maxMove = 0.f, disp2 = 0.f;
        if (valid && chunkLane) {
            const int64_t at = (int64_t)v * V + c;
            float4 f = make_float4((float)(acc[0] + (double)rep[0] * fp.invFixForce), (float)(acc[1] + (double)rep[1] * fp.invFixForce),
                                   (float)(acc[2] + (double)rep[2] * fp.invFixForce), (float)(acc[3] + (double)rep[3] * fp.invFixForce));
            if (fp.centreScale != 0.f) {                   // :296-301
                f.x = fmaf(-fp.centreScale, xv.x, f.x); f.y = fmaf(-fp.centreScale, xv.y, f.y);
                f.z = fmaf(-fp.centreScale, xv.z, f.z); f.w = fmaf(-fp.centreScale, xv.w, f.w);
            }
            if (fp.keepForces) forceOut[at] = f;
            float4 xn;
            if (fp.optimizer == 1) {
                const float4 m = st.m[slot * V + c], s = st.s[slot * V + c];
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
            sumX[0] = (double)xn.x; sumX[1] = (double)xn.y; sumX[2] = (double)xn.z; sumX[3] = (double)xn.w;
            disp2 = chunk_dist2(xn, xv);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) disp2 += __shfl_xor_sync(0xffffffffu, disp2, o);
        if (valid && c == 0) {
            sumLossA = loss;
            sumLossR = (double)lossR * fp.invFixLoss;
            sumPairs = (double)nPairs;
            // how far the vertex moved, in units of its smallest possible interaction radius (StepCtrl)
            maxMove = sqrtf(disp2) * iwv * fp.dispScale;
        }
        // fixed-order tile reduction: lanes that own the same chunk add up (xor offsets G, 2G, ..), then the 8 warps in order
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
            sumLossA += __shfl_xor_sync(0xffffffffu, sumLossA, o);
            sumLossR += __shfl_xor_sync(0xffffffffu, sumLossR, o);
            sumPairs += __shfl_xor_sync(0xffffffffu, sumPairs, o);
            maxMove = fmaxf(maxMove, __shfl_xor_sync(0xffffffffu, maxMove, o));
#pragma unroll
            for (int i = 0; i < 4; ++i) sumX[i] += __shfl_xor_sync(0xffffffffu, sumX[i], o);
        }
        if (lane == 0) { redBuf[p & 1][warp][0] = sumLossA; redBuf[p & 1][warp][1] = sumLossR; redBuf[p & 1][warp][2] = sumPairs; redBuf[p & 1][warp][K] = (double)maxMove; }
        if (lane < G && chunkLane) {
#pragma unroll
            for (int i = 0; i < 4; ++i) redBuf[p & 1][warp][3 + 4 * c + i] = sumX[i];
        }
        __syncthreads();                                        // stage p & 1 may be refilled (tile p + 2) from here on
        if (threadIdx.x <= K) {
            double sacc = redBuf[p & 1][0][threadIdx.x];
            if (threadIdx.x < K) { for (int w = 1; w < 8; ++w) sacc += redBuf[p & 1][w][threadIdx.x]; }
            else { for (int w = 1; w < 8; ++w) sacc = fmax(sacc, redBuf[p & 1][w][threadIdx.x]); }
            tilePartials[(int64_t)(v0 / VPB) * (K + 1) + threadIdx.x] = sacc;
        }
    }
}

// Hub rows (degree > kHubThreshold, or a heavy vertex with thousands of repulsion partners): one block per hub strides over both
// rows, sums attraction in double and repulsion in fixed point, reduces in a fixed order; k_step_fused picks the record up instead
// of walking the rows itself.
template <int V>
__global__ void __launch_bounds__(256) k_hub_rows(const float4* __restrict__ x, const float* __restrict__ iw, const int* __restrict__ rowPtr,
                                                  const int* __restrict__ col, const int* __restrict__ repRowPtr, const int* __restrict__ repCol,
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
    const int rend = repRowPtr[v + 1];
    for (int e = repRowPtr[v] + threadIdx.x; e < rend; e += 256) {
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
// Deterministic reduction of the tile sums (util::deterministicSum's role, ParallelReduce.hpp:18-37): tiles are grouped 64 by 64 in
// GLOBAL tile order; a group is summed sequentially, thread t of the reducing block adds the groups t, t + 256, .. in order, and
// the 256 thread sums are combined by a fixed tree.  Neither the grid of the step kernel nor the number of GPUs (whole groups per
// rank, the group sums are exchanged) can change a bit of the result.  Column K (the last) is a maximum.
constexpr int kTileGroup = 64;
__global__ void __launch_bounds__(256) k_reduce_tile_groups(const double* __restrict__ tilePartials, int tileBegin, int tileEnd, int cols,
                                                            double* __restrict__ groupSums, const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    // one thread per (group, column)
    const int numGroups = (tileEnd - tileBegin + kTileGroup - 1) / kTileGroup;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)numGroups * cols) return;
    const int g = (int)(i / cols), k = (int)(i % cols);
    const int t0 = tileBegin + g * kTileGroup, t1 = min(tileEnd, t0 + kTileGroup);
    double s = tilePartials[(int64_t)t0 * cols + k];
    if (k == cols - 1) { for (int t = t0 + 1; t < t1; ++t) s = fmax(s, tilePartials[(int64_t)t * cols + k]); }
    else { for (int t = t0 + 1; t < t1; ++t) s += tilePartials[(int64_t)t * cols + k]; }
    groupSums[(int64_t)(tileBegin / kTileGroup + g) * cols + k] = s;
}
// block k reduces column k of the group sums
__global__ void __launch_bounds__(256) k_reduce_groups(const double* __restrict__ groupSums, int numGroups, int cols, double* __restrict__ out,
                                                       const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    __shared__ double sm[256];
    const int k = blockIdx.x;
    const bool isMax = k == cols - 1;
    double s = 0.0;
    for (int g = threadIdx.x; g < numGroups; g += 256) { const double v = groupSums[(int64_t)g * cols + k]; s = isMax ? fmax(s, v) : s + v; }
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = isMax ? fmax(sm[threadIdx.x], sm[threadIdx.x + o]) : sm[threadIdx.x] + sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
}

// ---------------------------------------------------------------------------------------------
// applyGravityCentre + observeDisplacement (WembedEmbedder.cpp:303-352): x = xnew - centroid, the sums of ||x - xprev|| and ||x||^2,
// and - because this pass streams the final layout anyway - the per-dimension moments the next index build takes its quantisation
// frame from.  One block per tile of kObsTile vertices (global tiles: the partial sums do not depend on the grid or on the number
// of GPUs).  forceSums = output of k_reduce_groups ({lossA, lossR, pairs, sum xnew[k], max displacement}).
template <int V>
__global__ void __launch_bounds__(256) k_recentre_observe(float4* __restrict__ x, const float4* __restrict__ xNew, int n, int tileBegin, int dim,
                                                          const double* __restrict__ forceSums, double* __restrict__ obsPartials /* [tile][2] */,
                                                          float* __restrict__ momentPartials /* [tile][4][kMaxDim] */,
                                                          const StepCtrl* __restrict__ ctrl) {
    if (ctrl->overflow != 0) return;
    __shared__ double redBuf[8 * 2];
    __shared__ float smMom[8][4][4 * V];
    float cen[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) cen[k] = (k < dim) ? (float)(forceSums[3 + k] / (double)n) : 0.f;
    const int tile = tileBegin + blockIdx.x;
    const int vEnd = min(n, (tile + 1) * kObsTile);
    double sums[2] = {0.0, 0.0};
    float mn[4 * V], mx[4 * V], s1[4 * V], s2[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) { mn[k] = 3.0e38f; mx[k] = -3.0e38f; s1[k] = 0.f; s2[k] = 0.f; }
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
            const float e[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = 4 * c + i;
                mn[k] = fminf(mn[k], e[i]); mx[k] = fmaxf(mx[k], e[i]);
                s1[k] += e[i]; s2[k] = fmaf(e[i], e[i], s2[k]);
            }
        }
        sums[0] += (double)sqrtf(disp2);
        sums[1] += (double)rad2;
    }
    block_sum<2, 256>(sums, redBuf, obsPartials + (int64_t)tile * 2);
    moments_block_reduce<V>(mn, mx, s1, s2, smMom, momentPartials + (int64_t)tile * 4 * kMaxDim);
}

// ---------------------------------------------------------------------------------------------
// Last kernel of a step (one block): the observation sums, the quantisation frame of the next index build, the device's decision
// about the next step (rebuild the pair list or reuse it, and with which skin) and the record the host reads back.
//
// stats layout (doubles): [0] lossA [1] lossR [2] active pairs [3 .. 3+4V) sum xnew [K] max displacement ratio of this step |
//   then kTailStats values: listed pairs, point tests, box tests, sum displacement, sum radius^2, rebuilt (this step), skin of the
//   current list, next step rebuilds, overflow, pairs needed
constexpr int kTailStats = 10;
struct TailPolicy { float edgeLength; float halfSigmaLimit; int dim; int mortonBits; };

__global__ void __launch_bounds__(1024) k_step_tail(const double* __restrict__ forceSums, int cols, const double* __restrict__ obsPartials, int numObsTiles,
                                                    const float* __restrict__ momentPartials, int n, const double* __restrict__ walkPartials, int walkRows,
                                                    const unsigned int* __restrict__ pairCounts, int world, const TailPolicy pol, QuantParams* __restrict__ qp,
                                                    StepCtrl* ctrl, double* __restrict__ stats) {
    __shared__ QuantScratch sc;
    __shared__ double sm[1024];
    __shared__ double res[5];
    if (ctrl->overflow != 0) {
        if (threadIdx.x == 0) { stats[cols + 8] = 1.0; stats[cols + 9] = (double)ctrl->pairNeeded; }
        return;
    }
    const bool rebuilt = ctrl->rebuild != 0;
    // columns: 0, 1 = observation sums over the tiles; 2, 3, 4 = walk statistics over the walk's warps (integers)
    for (int col = 0; col < 5; ++col) {
        const double* src = col < 2 ? obsPartials + col : walkPartials + (col - 2);
        const int rows = col < 2 ? numObsTiles : (rebuilt ? walkRows : 0), stride = col < 2 ? 2 : 3;
        double s = 0.0;
        for (int r = threadIdx.x; r < rows; r += 1024) s += src[(int64_t)r * stride];
        sm[threadIdx.x] = s;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) res[col] = sm[0];
        __syncthreads();
    }
    quant_from_partials(momentPartials, numObsTiles, n, pol.dim, pol.mortonBits, pol.halfSigmaLimit, qp, sc);
    if (threadIdx.x == 0) {
        const float r = (float)forceSums[cols - 1];                // largest displacement of this step, in smallest interaction radii
        float accum = rebuilt ? r : ctrl->dispAccum + r;           // a list built this step saw the positions BEFORE the step's move
        const float skin = ctrl->skin;
        const bool reuse = ctrl->listValid != 0 && skin > 0.f && accum <= 0.5f * skin * 0.999f && isfinite(accum);
        unsigned int listed = 0u;
        for (int s = 0; s < world; ++s) listed += pairCounts[s];
        for (int k = 0; k < cols; ++k) stats[k] = forceSums[k];
        stats[cols + 0] = (double)listed;
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
            if (rebuilt && skin > 0.f && listed > ctrl->pairBudget) cap = fmaxf(0.02f, 0.75f * skin);
            else cap = fminf(ctrl->skinMax, cap * 1.1f);
            ctrl->skinCap = cap;
            float s = 0.f;
            if (cap > 0.f && isfinite(r) && 2.1f * r <= cap) s = fminf(cap, fmaxf(2.2f * ctrl->reuseTarget * r, 0.05f));
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
