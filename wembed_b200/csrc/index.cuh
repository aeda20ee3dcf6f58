// Spatial index of the repulsion search: Morton-sorted points under an implicit 8-ary hierarchy of tight boxes
// (replaces WembedEmbedder::updateIndex / WeightedIndex / SnnModel: WembedEmbedder.cpp:212-240, WeightedIndex.cpp:10-81, snn.cpp:97-160).
//
// Build (only on steps whose pair list has expired, see StepCtrl): the quantisation frame of the layout comes out of the previous
// step's recentre pass (k_step_tail) -> k_morton_keys -> radix sort -> k_build_low (sorted copy of the points + levels 1..3, one
// block per 512 points) -> k_build_top (the few nodes above, one block).
#pragma once
#include <cuda_fp16.h>

#include "params.cuh"

namespace wb {

// kernels of a step return at once when the device decided to reuse the pair list (or after a list overflow); `always` is set by
// callers outside the step (test hook)
__device__ __forceinline__ bool build_skipped(const StepCtrl* ctrl, int always) { return !always && (ctrl->rebuild == 0 || ctrl->overflow != 0); }

// initial state of the block buffer: coordinates far from everything, meta records all zero (endPos = 0: never passes)
template <int V>
__global__ void k_init_blocks(float4* __restrict__ blk, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const bool meta = (int)(i % block_float4s(V)) >= 2 * V * kFan;
    blk[i] = meta ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(kPadCoord, kPadCoord, kPadCoord, kPadCoord);
}

// writes node `idx` of level `lv` into its block; called by the 8 lanes (j = 0..7) that hold the node's reduced box
template <int V>
__device__ __forceinline__ void store_block_node(float4* __restrict__ blk, float4* __restrict__ blkH, const QuantParams* __restrict__ qp,
                                                 int blockOffLv, int blockOffBelow, int lv, int idx, int j,
                                                 const float4 (&lo)[V], const float4 (&hi)[V], float bound) {
    float4* b = blk + ((int64_t)blockOffLv + (idx >> kFanLog2)) * block_float4s(V) + (idx & (kFan - 1));
#pragma unroll
    for (int c = 0; c < V; ++c)
        if (j == c) { b[c * kFan] = lo[c]; b[(V + c) * kFan] = hi[c]; }
    const uint32_t childRef = lv == 1 ? (kLeafFlag | (uint32_t)idx) : (uint32_t)(blockOffBelow + idx);
    const uint32_t endPos = (uint32_t)min((int64_t)(idx + 1) << (kFanLog2 * lv), (int64_t)0x7fffffff);
    // .w = 1 / bound, rounded up (the half-precision box rounds turn it into a distance threshold; 0 for the empty boxes' inf bound)
    const float4 meta = make_float4(bound, __uint_as_float(childRef), __uint_as_float(endPos), __frcp_ru(bound));
    if (j == kFan - 1) b[2 * V * kFan] = meta;
    // half-precision copy: the box relative to the frame centre, lo rounded down and hi rounded up (subtraction and conversion both
    // directed), so it contains the fp32 box; lanes 0..HV-1 pack the lo chunks, lanes HV..2HV-1 the hi chunks
    constexpr int HV = half_chunks(V);
    float4* bh = blkH + ((int64_t)blockOffLv + (idx >> kFanLog2)) * half_block_float4s(V) + (idx & (kFan - 1));
    if (j < 2 * HV) {
        const bool up = j >= HV;
        const int k = up ? j - HV : j;
        __half2 h[4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float e[4] = {0.f, 0.f, 0.f, 0.f}, ctr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < V; ++c) {
                if (c == 2 * k + half) {
                    const float4 src = up ? hi[c] : lo[c];
                    e[0] = src.x; e[1] = src.y; e[2] = src.z; e[3] = src.w;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ctr[i] = qp->centre[4 * c + i];
                }
            }
            if (up) {
                h[2 * half] = __halves2half2(__float2half_ru(__fsub_ru(e[0], ctr[0])), __float2half_ru(__fsub_ru(e[1], ctr[1])));
                h[2 * half + 1] = __halves2half2(__float2half_ru(__fsub_ru(e[2], ctr[2])), __float2half_ru(__fsub_ru(e[3], ctr[3])));
            } else {
                h[2 * half] = __halves2half2(__float2half_rd(__fsub_rd(e[0], ctr[0])), __float2half_rd(__fsub_rd(e[1], ctr[1])));
                h[2 * half + 1] = __halves2half2(__float2half_rd(__fsub_rd(e[2], ctr[2])), __float2half_rd(__fsub_rd(e[3], ctr[3])));
            }
        }
        float4 packed;
        packed.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[0]));
        packed.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[1]));
        packed.z = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[2]));
        packed.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[3]));
        bh[j * kFan] = packed;
    }
    if (j == kFan - 1) bh[2 * HV * kFan] = meta;
}

// ---------------------------------------------------------------------------------------------
// Per-dimension min / max / sum / sum of squares of a layout (fixed-order reduction); partial layout [tile][4][kMaxDim] floats.
// One block per tile of kObsTile vertices.  Inside a step it runs after the recentre pass on a 1-in-4 sample (the first 256 vertices
// of every tile) and k_step_tail turns the partials into the frame of the next build.
constexpr int kObsTile = 1024;

template <int V>
__device__ __forceinline__ void moments_block_reduce(float (&mn)[4 * V], float (&mx)[4 * V], float (&s1)[4 * V], float (&s2)[4 * V],
                                                     float (*sm)[4][4 * V] /* [8] */, float* __restrict__ out /* [4][kMaxDim] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
        if (lane == 0) { sm[warp][0][k] = mn[k]; sm[warp][1][k] = mx[k]; sm[warp][2][k] = s1[k]; sm[warp][3][k] = s2[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 4 * V) {
        const int k = threadIdx.x;
        float a = sm[0][0][k], b = sm[0][1][k], c = sm[0][2][k], d = sm[0][3][k];
        for (int w = 1; w < 8; ++w) { a = fminf(a, sm[w][0][k]); b = fmaxf(b, sm[w][1][k]); c += sm[w][2][k]; d += sm[w][3][k]; }
        out[0 * kMaxDim + k] = a; out[1 * kMaxDim + k] = b; out[2 * kMaxDim + k] = c; out[3 * kMaxDim + k] = d;
    }
}

// perTile: how many vertices of every tile are looked at (the first ones); the frame only has to be representative
template <int V>
__global__ void __launch_bounds__(256) k_moments(const float4* __restrict__ x, int n, int perTile, float* __restrict__ partial) {
    float mn[4 * V], mx[4 * V], s1[4 * V], s2[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) { mn[k] = 3.0e38f; mx[k] = -3.0e38f; s1[k] = 0.f; s2[k] = 0.f; }
    const int vEnd = min(n, blockIdx.x * kObsTile + perTile);
    for (int v = blockIdx.x * kObsTile + threadIdx.x; v < vEnd; v += 256) {
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const float4 p = __ldg(x + (int64_t)v * V + c);
            const float e[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = 4 * c + i;
                mn[k] = fminf(mn[k], e[i]); mx[k] = fmaxf(mx[k], e[i]);
                s1[k] += e[i]; s2[k] = fmaf(e[i], e[i], s2[k]);
            }
        }
    }
    __shared__ float sm[8][4][4 * V];
    moments_block_reduce<V>(mn, mx, s1, s2, sm, partial + (int64_t)blockIdx.x * 4 * kMaxDim);
}

// Quantisation frame = [mean - 4 sd, mean + 4 sd] clipped to [min, max] per dimension, so a few far outliers do not eat the key
// resolution of the bulk.  Only locality depends on this frame, never results.  Also decides whether the next walk may test the
// half-precision copy of the boxes: the rounding of a centred coordinate to half precision is ~sd * 2^-11, which has to stay small
// against the smallest interaction radius or the outward-rounded boxes stop pruning; halfSigmaLimit = that radius times a constant
// (wb_set_weights), <= 0 disables, +inf forces.  Called by all 1024 threads of a block; smem = the scratch declared by the caller.
struct QuantScratch {
    float sMin[32][kMaxDim], sMax[32][kMaxDim];
    double sS1[32][kMaxDim], sS2[32][kMaxDim];
    float sSd[kMaxDim];
};
// partial layout: [row][4][kMaxDim] floats = {min, max, sum, sum of squares} over `count` vertices in all
__device__ __forceinline__ void quant_from_partials(const float* __restrict__ partial, int numRows, int count, int dim, int bits,
                                                    float halfSigmaLimit, QuantParams* __restrict__ qp, QuantScratch& sc) {
    // thread (k, j) = (dimension, slice): slice j folds rows j, j+32, .. in order; the 32 slices are combined in slice order
    const int k = threadIdx.x & 31, j = threadIdx.x >> 5;
    float mn = 3.0e38f, mx = -3.0e38f; double s1 = 0.0, s2 = 0.0;
    if (k < dim) {
        for (int b = j; b < numRows; b += 32) {
            const float* p = partial + (int64_t)b * 4 * kMaxDim;
            mn = fminf(mn, p[k]); mx = fmaxf(mx, p[kMaxDim + k]); s1 += p[2 * kMaxDim + k]; s2 += p[3 * kMaxDim + k];
        }
    }
    sc.sMin[j][k] = mn; sc.sMax[j][k] = mx; sc.sS1[j][k] = s1; sc.sS2[j][k] = s2;
    __syncthreads();
    if (j == 0) {
        float sd = 0.f;
        if (k < dim) {
            for (int t = 1; t < 32; ++t) { mn = fminf(mn, sc.sMin[t][k]); mx = fmaxf(mx, sc.sMax[t][k]); s1 += sc.sS1[t][k]; s2 += sc.sS2[t][k]; }
            const double mean = s1 / count;
            const double var = fmax(0.0, s2 / count - mean * mean);
            sd = (float)sqrt(var);
            float lo = fmaxf(mn, (float)mean - 4.f * sd), hi = fminf(mx, (float)mean + 4.f * sd);
            if (!(hi > lo)) hi = lo + 1.f;
            qp->lo[k] = lo;
            qp->invCell[k] = (float)(1u << bits) / (hi - lo);
            qp->centre[k] = (float)mean;
        } else {
            qp->centre[k] = 0.f;
        }
        sc.sSd[k] = sd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sdMax = 0.f;
        for (int t = 0; t < dim; ++t) sdMax = fmaxf(sdMax, sc.sSd[t]);
        qp->halfBoxes = (sdMax <= halfSigmaLimit) ? 1 : 0;      // false for NaN layouts as well
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_quant_params(const float* __restrict__ partial, int numTiles, int n, int dim, int bits,
                                                      float halfSigmaLimit, QuantParams* __restrict__ qp) {
    __shared__ QuantScratch sc;
    quant_from_partials(partial, numTiles, n, dim, bits, halfSigmaLimit, qp, sc);
}

// Morton key of every vertex (bit b of dimension k -> key bit b*dim + k).
template <int V>
__global__ void __launch_bounds__(256) k_morton_keys(const float4* __restrict__ x, int n, int dim, int bits, const QuantParams* __restrict__ qp,
                                                     uint32_t* __restrict__ keys, int* __restrict__ vals, const StepCtrl* __restrict__ ctrl, int always) {
    if (build_skipped(ctrl, always)) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t qmax = (1u << bits) - 1u;
    uint32_t key = 0;
#pragma unroll
    for (int c = 0; c < V; ++c) {
        const float4 p = __ldg(x + (int64_t)v * V + c);
        const float e[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 4 * c + i;
            if (k < dim) {
                const float t = (e[i] - qp->lo[k]) * qp->invCell[k];
                const uint32_t q = t <= 0.f ? 0u : (t >= (float)qmax ? qmax : (uint32_t)t);
                for (int b = 0; b < bits; ++b) key |= ((q >> b) & 1u) << (b * dim + k);
            }
        }
    }
    keys[v] = key;
    vals[v] = v;
}

// mutable view of the level planes for the builders
struct TreePlanes {
    float4* lo[kMaxLevels];
    float4* hi[kMaxLevels];
    float* bound[kMaxLevels];
};

// Sorted copy of the points (plane layout) and levels 1..3 of the hierarchy.  One block of 512 threads owns 512 consecutive sorted
// positions = 64 leaves = 8 level-2 nodes = 1 level-3 node; thread t owns position 512 b + t, 8-lane groups reduce a node's
// children with shuffles, the nodes of one level travel to the next through shared memory.
constexpr int kBuildThreads = 512;
template <int V>
struct BuildBox { float4 lo[V], hi[V]; float bound; };

template <int V>
__device__ __forceinline__ void reduce_children(float4 (&lo)[V], float4 (&hi)[V], float& b) {
#pragma unroll
    for (int o = kFan / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = min4(lo[c], shfl_xor4(lo[c], o)); hi[c] = max4(hi[c], shfl_xor4(hi[c], o)); }
        b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
}
template <int V>
__device__ __forceinline__ void empty_box(float4 (&lo)[V], float4 (&hi)[V]) {
#pragma unroll
    for (int c = 0; c < V; ++c) { lo[c] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f); hi[c] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f); }
}
template <int V>
__device__ __forceinline__ void store_plane_node(const TreePlanes& tp, const TreeView& t, int lv, int node, int j, const float4 (&lo)[V],
                                                 const float4 (&hi)[V], float b) {
#pragma unroll
    for (int c = 0; c < V; ++c)
        if (j == c) { tp.lo[lv][(int64_t)c * t.stride[lv] + node] = lo[c]; tp.hi[lv][(int64_t)c * t.stride[lv] + node] = hi[c]; }
    if (j == kFan - 1) tp.bound[lv][node] = b;
}

template <int V>
__global__ void __launch_bounds__(kBuildThreads) k_build_low(const float4* __restrict__ x, const float* __restrict__ pointBound,
                                                             const int* __restrict__ order, int n, const TreeView t, const TreePlanes tp,
                                                             int* __restrict__ ids, const int* __restrict__ heavySlot, int* __restrict__ heavyPos,
                                                             float4* __restrict__ blk, float4* __restrict__ blkH,
                                                             const StepCtrl* __restrict__ ctrl, int always) {
    if (build_skipped(ctrl, always)) return;
    __shared__ BuildBox<V> sBox[kBuildThreads / kFan];           // this block's 64 leaves, then (first 8 entries) its level-2 nodes
    const int tid = threadIdx.x;
    const int i = blockIdx.x * kBuildThreads + tid;              // sorted position
    const int leaf = i >> kFanLog2, j = i & (kFan - 1);
    const bool real = i < n;
    const int stride0 = t.stride[0];
    float4 lo[V], hi[V];
    float b = 3.0e38f;
    if (real) {
        const int src = order[i];
        b = __ldg(pointBound + src);
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const float4 p = __ldg(x + (int64_t)src * V + c);
            tp.lo[0][(int64_t)c * stride0 + i] = p;
            lo[c] = p; hi[c] = p;
        }
        tp.bound[0][i] = b;
        ids[i] = src;
        if (heavySlot) { const int hs = __ldg(heavySlot + src); if (hs >= 0) heavyPos[hs] = i; }
    } else {
        empty_box<V>(lo, hi);
    }
    reduce_children<V>(lo, hi, b);
    if (leaf < t.count[1]) {
        store_plane_node<V>(tp, t, 1, leaf, j, lo, hi, b);
        store_block_node<V>(blk, blkH, t.quant, t.blockOff[1], 0, 1, leaf, j, lo, hi, b);
    }
    if (t.numLevels < 2) return;
    if (j == 0) {
        BuildBox<V>& s = sBox[tid >> kFanLog2];
#pragma unroll
        for (int c = 0; c < V; ++c) { s.lo[c] = lo[c]; s.hi[c] = hi[c]; }
        s.bound = b;
    }
    __syncthreads();
    // level 2: threads 0..63, child = local leaf tid
    const int kids = kBuildThreads / kFan;                       // 64
    if (tid < kids) {
        const int child = blockIdx.x * kids + tid, parent = child >> kFanLog2, jj = tid & (kFan - 1);
        if (child < t.count[1]) {
#pragma unroll
            for (int c = 0; c < V; ++c) { lo[c] = sBox[tid].lo[c]; hi[c] = sBox[tid].hi[c]; }
            b = sBox[tid].bound;
        } else {
            empty_box<V>(lo, hi);
            b = 3.0e38f;
        }
        reduce_children<V>(lo, hi, b);
        if (parent < t.count[2]) {
            store_plane_node<V>(tp, t, 2, parent, jj, lo, hi, b);
            store_block_node<V>(blk, blkH, t.quant, t.blockOff[2], t.blockOff[1], 2, parent, jj, lo, hi, b);
        }
    }
    if (t.numLevels < 3) return;
    __syncthreads();                                             // everybody has read the leaves
    if (tid < kids && (tid & (kFan - 1)) == 0) {
        BuildBox<V>& s = sBox[tid >> kFanLog2];
#pragma unroll
        for (int c = 0; c < V; ++c) { s.lo[c] = lo[c]; s.hi[c] = hi[c]; }
        s.bound = b;
    }
    __syncthreads();
    // level 3: the first warp, child = this block's level-2 node `lane` (lanes 8..31 only take part in the shuffles)
    if (tid < 32) {
        const int child = blockIdx.x * kFan + tid, parent = blockIdx.x;
        if (tid < kFan && child < t.count[2]) {
#pragma unroll
            for (int c = 0; c < V; ++c) { lo[c] = sBox[tid].lo[c]; hi[c] = sBox[tid].hi[c]; }
            b = sBox[tid].bound;
        } else {
            empty_box<V>(lo, hi);
            b = 3.0e38f;
        }
        reduce_children<V>(lo, hi, b);
        if (tid < kFan && parent < t.count[3]) {
            store_plane_node<V>(tp, t, 3, parent, tid, lo, hi, b);
            store_block_node<V>(blk, blkH, t.quant, t.blockOff[3], t.blockOff[2], 3, parent, tid, lo, hi, b);
        }
    }
}

// Levels 4 .. top: at most n / 4096 nodes, built level by level by one block (the planes written by the level below are read back
// after a block barrier, so no pointer here is `const __restrict__`).
template <int V>
__global__ void __launch_bounds__(1024) k_build_top(const TreeView t, const TreePlanes tp, float4* blk, float4* blkH,
                                                    const StepCtrl* __restrict__ ctrl, int always) {
    if (build_skipped(ctrl, always)) return;
    for (int l = 4; l <= t.numLevels; ++l) {
        const int cCount = t.count[l - 1], cStride = t.stride[l - 1], pCount = t.count[l];
        for (int base = 0; base < pCount * kFan; base += 1024) {
            const int i = base + threadIdx.x;                    // child index
            const int parent = i >> kFanLog2, j = i & (kFan - 1);
            float4 lo[V], hi[V];
            float b = 3.0e38f;
            if (i < cCount) {
#pragma unroll
                for (int c = 0; c < V; ++c) { lo[c] = tp.lo[l - 1][(int64_t)c * cStride + i]; hi[c] = tp.hi[l - 1][(int64_t)c * cStride + i]; }
                b = tp.bound[l - 1][i];
            } else {
                empty_box<V>(lo, hi);
            }
            reduce_children<V>(lo, hi, b);
            if (parent < pCount) {
                store_plane_node<V>(tp, t, l, parent, j, lo, hi, b);
                store_block_node<V>(blk, blkH, t.quant, t.blockOff[l], t.blockOff[l - 1], l, parent, j, lo, hi, b);
            }
        }
        __syncthreads();
    }
}

}  // namespace wb
