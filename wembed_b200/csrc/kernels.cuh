// Device kernels of the WEmbed gradient-descent step for sm_100a.
//
// Reference semantics (paths relative to the Vraier/wembed checkout):
//   index rebuild      WembedEmbedder::updateIndex            src/embeddingLib/src/embedder/WembedEmbedder.cpp:212-240
//   repulsion          calculateAllRepellingForces/repellingForce                       :274-294, :174-210
//   attraction         calculateAllAttractingForces/attractionForce                     :260-272, :140-172
//   centre force       calculateAllCentreForces                                         :296-301
//   optimizer          AdamOptimizer::update / SimpleOptimizer::update   src/embeddingLib/src/gradientOptimizer/*.cpp
//   recentre+observe   applyGravityCentre / observeDisplacement                         :303-352
//
// All kernels are pull style: one owner computes force[v]; there are no floating-point atomics and
// every reduction has a fixed shape, so a step is bit-reproducible from run to run
// (tests/TestDeterminism.cpp protocol).  V = number of float4 chunks per position row.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "mt19937.cuh"

namespace wb {

// ---------------------------------------------------------------------------------------------
// Parameter blocks

struct QuantParams {          // Morton quantisation frame, rebuilt every step on the device
    float lo[kMaxDim];
    float invCell[kMaxDim];
    float centre[kMaxDim];    // per-dimension mean: origin of the half-precision copy of the boxes (0 for padding dimensions)
    int halfBoxes;            // 1: this step's repulsion walk tests the half-precision boxes (layout narrow enough, see k_quant_params)
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
    // WB_POINT_HALF: the sorted points once more as half_chunks(V) planes of 8 halves (point - centre, rounded to nearest) and per
    // point {iw, 1 / iw rounded up, length of the rounding displacement * margin, -}
    const float4* ptsH;
    const float4* pmeta;
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
#ifndef WB_HIT_BATCH
#define WB_HIT_BATCH 0             // 1: hits of the point rounds are resolved 32 at a time, one per lane (A/B candidate, not yet measured)
#endif
#ifndef WB_POINT_HALF
#define WB_POINT_HALF 0            // 1: the point rounds prefilter in half precision too (A/B candidate, not yet measured)
#endif

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

struct ForceParams {
    float edgeLength;         // L
    float pruneL2;            // L^2 * (1 + slack): conservative bound for box pruning
    float attractionScale, repulsionScale, centreScale;
    // optimizer (AdamOptimizer.cpp:19-28 / SimpleOptimizer.cpp:13-30)
    int optimizer;            // wb_optimizer
    float lr, beta1, beta2, eps, invBias1, invBias2, maxDisplacement;
    uint32_t seed, iteration; // tie-break generator key (Rand.cpp:29-35)
    int dim;                  // real embedding dimension (<= 4V)
    int keepForces;
    // Repulsion results are accumulated as 64-bit fixed-point integers (value * 2^k, k chosen by wb_set_weights so that n terms of
    // the largest possible magnitude cannot overflow): integer addition is associative, so the atomics that scatter a pair's
    // term to both of its vertices give bit-identical sums in any order.
    double fixForce, invFixForce, fixLoss, invFixLoss;
};

// ---------------------------------------------------------------------------------------------
// Index, stage 1: per-dimension min / max / sum / sum of squares (fixed-order reduction).
// partial layout: [block][4][kMaxDim] floats.
template <int V>
__global__ void __launch_bounds__(256) k_moments(const float4* __restrict__ x, int n, float* __restrict__ partial) {
    float mn[4 * V], mx[4 * V], s1[4 * V], s2[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) { mn[k] = 3.0e38f; mx[k] = -3.0e38f; s1[k] = 0.f; s2[k] = 0.f; }
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
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
        float* out = partial + (int64_t)blockIdx.x * 4 * kMaxDim;
        out[0 * kMaxDim + k] = a; out[1 * kMaxDim + k] = b; out[2 * kMaxDim + k] = c; out[3 * kMaxDim + k] = d;
    }
}

// Index, stage 2: quantisation frame = [mean - 4 sd, mean + 4 sd] clipped to [min, max] per dimension, so a few
// far outliers do not eat the key resolution of the bulk.  Only locality depends on this frame, never results.
// It also decides whether this step's repulsion walk may test the half-precision copy of the boxes: the rounding of a
// centred coordinate to half precision is ~sd * 2^-11, which has to stay small against the smallest interaction radius or the
// outward-rounded boxes stop pruning; halfSigmaLimit = that radius times a constant (wb_set_weights), <= 0 disables, +inf forces.
__global__ void __launch_bounds__(1024) k_quant_params(const float* __restrict__ partial, int numBlocks, int n, int dim, int bits,
                                                      float halfSigmaLimit, QuantParams* __restrict__ qp) {
    // thread (k, j) = (dimension, slice): slice j folds blocks j, j+32, .. in order; the 32 slices are combined in slice order
    __shared__ float sMin[32][kMaxDim], sMax[32][kMaxDim];
    __shared__ double sS1[32][kMaxDim], sS2[32][kMaxDim];
    const int k = threadIdx.x & 31, j = threadIdx.x >> 5;
    float mn = 3.0e38f, mx = -3.0e38f; double s1 = 0.0, s2 = 0.0;
    if (k < dim) {
        for (int b = j; b < numBlocks; b += 32) {
            const float* p = partial + (int64_t)b * 4 * kMaxDim;
            mn = fminf(mn, p[k]); mx = fmaxf(mx, p[kMaxDim + k]); s1 += p[2 * kMaxDim + k]; s2 += p[3 * kMaxDim + k];
        }
    }
    sMin[j][k] = mn; sMax[j][k] = mx; sS1[j][k] = s1; sS2[j][k] = s2;
    __shared__ float sSd[kMaxDim];
    __syncthreads();
    if (j == 0) {
        float sd = 0.f;
        if (k < dim) {
            for (int t = 1; t < 32; ++t) { mn = fminf(mn, sMin[t][k]); mx = fmaxf(mx, sMax[t][k]); s1 += sS1[t][k]; s2 += sS2[t][k]; }
            const double mean = s1 / n;
            const double var = fmax(0.0, s2 / n - mean * mean);
            sd = (float)sqrt(var);
            float lo = fmaxf(mn, (float)mean - 4.f * sd), hi = fminf(mx, (float)mean + 4.f * sd);
            if (!(hi > lo)) hi = lo + 1.f;
            qp->lo[k] = lo;
            qp->invCell[k] = (float)(1u << bits) / (hi - lo);
            qp->centre[k] = (float)mean;
        } else {
            qp->centre[k] = 0.f;
        }
        sSd[k] = sd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float sdMax = 0.f;
        for (int t = 0; t < dim; ++t) sdMax = fmaxf(sdMax, sSd[t]);
        qp->halfBoxes = (sdMax <= halfSigmaLimit) ? 1 : 0;      // false for NaN layouts as well
    }
}

// Index, stage 3: Morton key of every vertex (bit b of dimension k -> key bit b*dim + k).
template <int V>
__global__ void __launch_bounds__(256) k_morton_keys(const float4* __restrict__ x, int n, int dim, int bits,
                                                     const QuantParams* __restrict__ qp, uint32_t* __restrict__ keys, int* __restrict__ vals) {
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

// Index, stage 4: gather the positions into sorted order (plane layout) and build the level-1 boxes.
// One 8-lane group per leaf; lane j owns sorted point leaf*8 + j.
template <int V>
__global__ void __launch_bounds__(256) k_build_leaves(const float4* __restrict__ x, const float* __restrict__ pointBound,
                                                      const int* __restrict__ order, int n, float4* __restrict__ pts,
                                                      int stride0, float* __restrict__ bound0, int* __restrict__ ids, int* __restrict__ invOrder,
                                                      float4* __restrict__ lo1, float4* __restrict__ hi1,
                                                      float* __restrict__ bound1, int stride1, float4* __restrict__ blk, int blockOff1,
                                                      float4* __restrict__ blkH, const QuantParams* __restrict__ qp,
                                                      float4* __restrict__ ptsH, float4* __restrict__ pmeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // sorted position
    const int leaf = i >> kFanLog2, j = i & (kFan - 1);
    const bool real = i < n;
    float4 lo[V], hi[V];
    float b = 3.0e38f;
    if (real) {
        const int src = order[i];
        b = __ldg(pointBound + src);
#pragma unroll
        for (int c = 0; c < V; ++c) {
            const float4 p = __ldg(x + (int64_t)src * V + c);
            pts[(int64_t)c * stride0 + i] = p;
            lo[c] = p; hi[c] = p;
        }
        bound0[i] = b;
        ids[i] = src;
        invOrder[src] = i;
    } else {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f); hi[c] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f); }
    }
#if WB_POINT_HALF
    if (i < stride0) {
        // the point as the half-precision point rounds see it (padding points: +inf, they fail every test) and how far the rounding moved it
        constexpr int HV = half_chunks(V);
        float d2 = 0.f;
#pragma unroll
        for (int k = 0; k < HV; ++k) {
            __half2 h[4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int ch = 2 * k + half;
                float e[4] = {0.f, 0.f, 0.f, 0.f};
                if (ch < V) {
                    const float4 p = real ? lo[ch] : make_float4(kPadCoord, kPadCoord, kPadCoord, kPadCoord);
                    e[0] = p.x - qp->centre[4 * ch]; e[1] = p.y - qp->centre[4 * ch + 1];
                    e[2] = p.z - qp->centre[4 * ch + 2]; e[3] = p.w - qp->centre[4 * ch + 3];
                }
                h[2 * half] = __floats2half2_rn(e[0], e[1]);
                h[2 * half + 1] = __floats2half2_rn(e[2], e[3]);
                const float2 b0 = __half22float2(h[2 * half]), b1 = __half22float2(h[2 * half + 1]);
                d2 = fmaf(e[0] - b0.x, e[0] - b0.x, d2); d2 = fmaf(e[1] - b0.y, e[1] - b0.y, d2);
                d2 = fmaf(e[2] - b1.x, e[2] - b1.x, d2); d2 = fmaf(e[3] - b1.y, e[3] - b1.y, d2);
            }
            float4 packed;
            packed.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[0]));
            packed.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[1]));
            packed.z = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[2]));
            packed.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[3]));
            ptsH[(int64_t)k * stride0 + i] = packed;
        }
        float delta = real ? sqrtf(d2) * 1.001f * half_margin_root(V) : 0.f;
        if (!(delta >= 0.f)) delta = __int_as_float(0x7f800000);          // beyond the half range: every query keeps this point
        const float iwp = real ? b : 1.f;
        pmeta[i] = make_float4(iwp, __frcp_ru(iwp), delta, 0.f);
    }
#endif
#pragma unroll
    for (int o = kFan / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = min4(lo[c], shfl_xor4(lo[c], o)); hi[c] = max4(hi[c], shfl_xor4(hi[c], o)); }
        b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (leaf * kFan < n) {
#pragma unroll
        for (int c = 0; c < V; ++c)
            if (j == c) { lo1[(int64_t)c * stride1 + leaf] = lo[c]; hi1[(int64_t)c * stride1 + leaf] = hi[c]; }
        if (j == kFan - 1) bound1[leaf] = b;
        store_block_node<V>(blk, blkH, qp, blockOff1, 0, 1, leaf, j, lo, hi, b);
    }
}

// Index, stage 5: one level of the hierarchy from the level below (8 lanes per parent).
template <int V>
__global__ void __launch_bounds__(256) k_build_level(const float4* __restrict__ cLo, const float4* __restrict__ cHi,
                                                     const float* __restrict__ cBound, int cCount, int cStride,
                                                     float4* __restrict__ pLo, float4* __restrict__ pHi,
                                                     float* __restrict__ pBound, int pCount, int pStride, float4* __restrict__ blk,
                                                     int blockOffP, int blockOffC, int pLevel, float4* __restrict__ blkH,
                                                     const QuantParams* __restrict__ qp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // child index
    const int parent = i >> kFanLog2, j = i & (kFan - 1);
    float4 lo[V], hi[V];
    float b = 3.0e38f;
    if (i < cCount) {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = cLo[(int64_t)c * cStride + i]; hi[c] = cHi[(int64_t)c * cStride + i]; }
        b = cBound[i];
    } else {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f); hi[c] = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f); }
    }
#pragma unroll
    for (int o = kFan / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < V; ++c) { lo[c] = min4(lo[c], shfl_xor4(lo[c], o)); hi[c] = max4(hi[c], shfl_xor4(hi[c], o)); }
        b = fminf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (parent < pCount) {
#pragma unroll
        for (int c = 0; c < V; ++c)
            if (j == c) { pLo[(int64_t)c * pStride + parent] = lo[c]; pHi[(int64_t)c * pStride + parent] = hi[c]; }
        if (j == kFan - 1) pBound[parent] = b;
        store_block_node<V>(blk, blkH, qp, blockOffP, blockOffC, pLevel, parent, j, lo, hi, b);
    }
}

// ---------------------------------------------------------------------------------------------
// Hierarchy walk shared by the repulsion kernel and the candidate-set test hook.
//
// One 8-lane group per query (4 queries per warp).  The group keeps a depth-first cursor in registers
// (current level, index of the expanded parent, one 8-bit "pending children" mask per level) and in
// every iteration expands one node: lane j tests child j.  Children that pass at level 0 are points
// and are handed to `onPoint` by the lane that tested them; everything is visited in a fixed order.
struct WalkMasks {
    unsigned long long a = 0ull, b = 0ull;   // 8 bits per level, levels 0..7 in a, 8..15 in b
    __device__ __forceinline__ uint32_t get(int l) const { return (uint32_t)(((l < 8) ? (a >> (8 * l)) : (b >> (8 * (l - 8)))) & 0xffull); }
    __device__ __forceinline__ void set(int l, uint32_t m) {
        if (l < 8) a = (a & ~(0xffull << (8 * l))) | ((unsigned long long)m << (8 * l));
        else b = (b & ~(0xffull << (8 * (l - 8)))) | ((unsigned long long)m << (8 * (l - 8)));
    }
};

// passes(level, idx, d2, bound, lo[]) decides whether a child survives; onPoint consumes level-0 survivors.
template <int V, typename Pass, typename OnPoint>
__device__ __forceinline__ void walk_tree(const TreeView& t, const float4 (&q)[V], bool valid, Pass&& passes, OnPoint&& onPoint,
                                          int& pointTests) {
    const int lane = threadIdx.x & 31, j = lane & (kFan - 1), g = lane >> kFanLog2;
    const int top = t.numLevels;
    int lvl = top + 1, cur = 0;
    WalkMasks masks;
    masks.set(top + 1, 1u);      // virtual root
    bool done = !valid;
    while (__any_sync(0xffffffffu, !done)) {
        if (!done) {
            while (lvl <= top + 1 && masks.get(lvl) == 0u) { ++lvl; cur >>= kFanLog2; }
            if (lvl > top + 1) {
                done = true;
            } else {
                const uint32_t m = masks.get(lvl);
                const int bit = __ffs(m) - 1;
                masks.set(lvl, m & (m - 1u));
                cur = cur * kFan + bit;   // the node being expanded (index at level lvl)
                --lvl;                    // its children live one level down
            }
        }
        const int lv = done ? 0 : lvl;
        const int idx = done ? 0 : cur * kFan + j;
        float4 lo[V], hi[V];
        const int64_t st = t.stride[lv];
#pragma unroll
        for (int c = 0; c < V; ++c) lo[c] = __ldg(t.lo[lv] + c * st + idx);
        if (lv == 0) {
#pragma unroll
            for (int c = 0; c < V; ++c) hi[c] = lo[c];
        } else {
#pragma unroll
            for (int c = 0; c < V; ++c) hi[c] = __ldg(t.hi[lv] + c * st + idx);
        }
        const float bnd = __ldg(t.bound[lv] + idx);
        const float d2 = box_dist2<V>(q, lo, hi);
        const bool pass = !done && passes(lv, idx, d2, bnd);
        const uint32_t ball = __ballot_sync(0xffffffffu, pass);
        if (!done) {
            if (lvl == 0) {
                ++pointTests;
                if (pass) onPoint(idx, d2, bnd, lo);
            } else {
                masks.set(lvl, (ball >> (kFan * g)) & 0xffu);
            }
        }
    }
}

// u in N(v)?  Rows are sorted ascending (Graph.cpp:87-150), so a binary search equals Graph::areNeighbors (:67-83).
// (Measured and rejected: reading rows of <= 16 entries with four independent 128-bit loads and comparing in registers instead of
// the dependent search - c3 repel 3.06 -> 3.16 ms; the extra instructions cost more than the shorter latency chain saves.)
__device__ __forceinline__ bool is_neighbor(const int* __restrict__ col, int begin, int end, int u) {
    while (begin < end) {
        const int mid = (begin + end) >> 1;
        const int w = __ldg(col + mid);
        if (w == u) return true;
        if (w < u) begin = mid + 1; else end = mid;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Repulsion (WembedEmbedder.cpp:274-294 + 174-210): pair-stack walk, every unordered pair found once.
//
// A warp owns 32 consecutive queries and one LIFO stack of (level, node, query) pairs in shared memory.  Every
// round pops eight pairs, two per 8-lane group; lane c of the group tests child c of the pair's node against the
// pair's query (query coordinates come from shared memory, the child box from L1/L2).  Passing children are pushed
// as new pairs, ordered child-major so that the pairs popped together usually name the same node (one cache
// line serves the four groups).
//
// The repulsive term of a pair is antisymmetric bit for bit (ws and the distance are symmetric, x_v - x_u = -(x_u - x_v)
// exactly), so the query at sorted position p only searches positions > p - subtrees that end at or before p are cut
// by an integer comparison - and the lane that finds a partner applies the term to BOTH vertices.  That halves the walk.
// The scatter uses 64-bit integer atomics on fixed-point rows (see ForceParams): integer sums do not depend on the
// order of the additions, so the step stays bit-reproducible and there are still no floating-point atomics.
//
// The kernel is persistent: every warp fetches the next chunk of 32 queries from an integer counter until none is
// left, so warps whose queries need long walks do not hold finished warps of the same block hostage (measured: 28 %
// of all stall samples sat on the final block barrier before).  Chunks are handed out in ascending order, and late
// positions have short walks (few positions behind them), so the tail of the schedule is cheap by construction.
// Repulsion results: one row of 4V + 2 fixed-point integers per VERTEX: [force (4V) | loss | coincident partners].
// The queries are dealt out by SORTED position: the sorted order is cut into blocks of kRepBlockChunks chunks (a chunk = 32
// consecutive positions = one warp's queries) and the blocks are dealt round-robin to the ranks of a sharded run: whole blocks,
// because warps that run at the same time should work on neighbouring chunks (they share tree nodes in L1 / L2; dealing single
// chunks cost 1.6x in walk time), round-robin because the walk cost varies across space.  With world = 1 this is the identity.
constexpr int kRepBlockChunks = 32;
struct RepLayout {
    int world, rank, segRows;
    // l-th query this rank walks (l < segRows) -> sorted position
    __host__ __device__ __forceinline__ int position(int l) const {
        constexpr int blockRows = kRepBlockChunks * 32;
        return ((l / blockRows) * world + rank) * blockRows + l % blockRows;
    }
};

// value -> fixed point (round to nearest even, symmetric in the sign, so a pair's two contributions cancel exactly)
__device__ __forceinline__ long long to_fixed(float term, double scale) { return __double2ll_rn((double)term * scale); }
__device__ __forceinline__ void fixed_add(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

// warps per block of k_repulse_pairs: the per-warp shared memory (queries + stack) grows with V
__host__ __device__ constexpr int repulse_warps(int V) { return V <= 4 ? 8 : 4; }

// dynamic shared memory of k_repulse_pairs<V, HALF>
constexpr int kHitBuffer = 96;            // WB_HIT_BATCH: waiting hits per warp (<= 31 left over + 64 from one point round)
__host__ __device__ constexpr int repulse_smem_bytes(int V, bool half) {
    return repulse_warps(V) * (32 * (V + 1 + (half ? half_chunks(V) : 0)) * 16 + (8 + 56 * kMaxLevels + 72 + 80) * 4 + (WB_HIT_BATCH ? kHitBuffer * 8 : 0));
}

#ifndef WB_REPULSE_MINBLOCKS
#define WB_REPULSE_MINBLOCKS 4
#endif
// HALF selects the box format of the box rounds.  false: the fp32 array-of-blocks.  true: the half-precision copy (lo rounded
// down, hi rounded up, relative to the frame centre) tested with packed half2 arithmetic against the query rounded to half
// precision: 3 instead of 5 128-bit loads per child, one instead of V shared-memory loads for the query and ~5 instructions
// per PAIR of dimensions.  The test stays conservative: with e = the gap vector computed from the rounded operands and
// delta = |q - round(q)| (exact, kept per query), the true distance to the box is >= |e| (1 - eps) - |delta|, so a child is
// kept iff |e|^2 <= (L' / s + |delta|)^2 * margin, where margin covers the half-precision rounding of the sum (and always
// if that threshold is beyond the half range).  Only the
// number of boxes that pass changes (+1..2 % at c3), never the pair set: points are still tested exactly in fp32.
// Both instantiations are launched every step; QuantParams::halfBoxes (decided on the device from the layout's spread) says
// which one runs, the other returns at once.
template <int V, bool HALF>
__global__ void __launch_bounds__(256, (V <= 2 ? WB_REPULSE_MINBLOCKS : (V <= 4 ? 2 : 1)))
k_repulse_pairs(const TreeView t, const int* __restrict__ rowPtr, const int* __restrict__ col, int n, const ForceParams fp,
                long long* __restrict__ forceRep, const RepLayout lay, int queriesPerUnit, const int* __restrict__ heavySlot,
                int* __restrict__ chunkCounter, double* __restrict__ partials) {
    if ((t.quant->halfBoxes != 0) != HALF) return;
    constexpr int RS = 4 * V + 2;                // integers per result row
    constexpr int WARPS = repulse_warps(V), STACK = 56 * kMaxLevels + 72;   // LIFO bound: <= 56 leftovers per level + one push of 64
    constexpr int HV = half_chunks(V);
    constexpr int QROW = V + 1 + (HALF ? HV : 0), BLK = HALF ? half_block_float4s(V) : block_float4s(V);
    // relative slack of the half-precision sum of squares, applied to the threshold before it is squared
    constexpr float kHalfMarginRoot = half_margin_root(V);
    constexpr uint32_t kRefMask = 0x07ffffffu;   // low 27 bits of an entry: block (stack) or leaf (leaf queue); high 5 bits: query lane
    // dynamic shared memory (repulse_smem_bytes): per warp
    //   query rows [32][QROW]: V coordinate chunks + {iw, sorted position + 1, threshold factor, |delta|} (+ HV chunks of 8 halves: q - centre)
    //   stack [8 + STACK]: 8 null entries below the stack (a short pop reads them and nothing passes)
    //   leaf queue [80]: leaves waiting for their point round
    extern __shared__ float4 smemRep[];
    static_assert(repulse_smem_bytes(V, HALF) == WARPS * (32 * QROW * 16 + (8 + STACK + 80) * 4 + (WB_HIT_BATCH ? kHitBuffer * 8 : 0)),
                  "host and kernel disagree on the layout");
    float4 (*sQ)[32][QROW] = reinterpret_cast<float4 (*)[32][QROW]>(smemRep);
    uint32_t (*sStack)[8 + STACK] = reinterpret_cast<uint32_t (*)[8 + STACK]>(smemRep + WARPS * 32 * QROW);
    uint32_t (*sLeaf)[80] = reinterpret_cast<uint32_t (*)[80]>(reinterpret_cast<uint32_t*>(smemRep + WARPS * 32 * QROW) + WARPS * (8 + STACK));
#if WB_HIT_BATCH
    uint2* myHit = reinterpret_cast<uint2*>(&sLeaf[WARPS][0]) + (threadIdx.x >> 5) * kHitBuffer;   // {sorted position of the partner, query lane}
    int nHit = 0;
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane & (kFan - 1), g = lane >> kFanLog2;
    // lanes that precede this one in child-major order (c, g)
    uint32_t before = 0u;
#pragma unroll
    for (int l = 0; l < 32; ++l) {
        const int lc = l & (kFan - 1), lg = l >> kFanLog2;
        if (lc < c || (lc == c && lg < g)) before |= 1u << l;
    }
    float4* myQ = &sQ[warp][0][0];
    uint32_t* myStack = &sStack[warp][8];
    uint32_t* myLeaf = &sLeaf[warp][0];
    if (lane < 8) sStack[warp][lane] = 0u;       // entry 0 = (query 0, null block)
    const float4* myBlk = (HALF ? t.blkH : t.blk) + c;   // lane c tests child c of every block
    const float pruneL = sqrtf(fp.pruneL2);
    const float L = fp.edgeLength;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t rootBlock = (uint32_t)t.blockOff[t.numLevels];
    // Work unit of a warp = queriesPerUnit (8, 16 or 32) consecutive rows of this rank's share of the sorted order.  Small units
    // keep the dynamic schedule balanced when a rank (or a small graph) has few queries per resident warp.
    const int numChunks = lay.segRows / queriesPerUnit;
    int nPairs = 0, nTests = 0, boxSlots = 0;

    // Box round: one (block, query) pair per 8-lane group and slot; lane c tests child box c of the block.
    struct BoxSlot { uint32_t entry, childRef; bool pass; };
    auto testBox = [&](uint32_t entry) {
        BoxSlot r;
        r.entry = entry;
        const float4* b = myBlk + (size_t)(entry & kRefMask) * BLK;
        const float4* qrow = myQ + (entry >> 27) * QROW;
        if constexpr (HALF) {
            float4 lo[HV], hi[HV], qv[HV];
#pragma unroll
            for (int k = 0; k < HV; ++k) lo[k] = __ldg(b + k * kFan);
#pragma unroll
            for (int k = 0; k < HV; ++k) hi[k] = __ldg(b + (HV + k) * kFan);
            const float4 meta = __ldg(b + 2 * HV * kFan);
#pragma unroll
            for (int k = 0; k < HV; ++k) qv[k] = qrow[V + 1 + k];
            const float4 qm = qrow[V];
            const __half2 zero2 = __float2half2_rn(0.f);
            __half2 acc0 = zero2, acc1 = zero2;
#pragma unroll
            for (int k = 0; k < HV; ++k) {
                const float lw[4] = {lo[k].x, lo[k].y, lo[k].z, lo[k].w}, hw[4] = {hi[k].x, hi[k].y, hi[k].z, hi[k].w};
                const float qw[4] = {qv[k].x, qv[k].y, qv[k].z, qv[k].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 l2 = *reinterpret_cast<const __half2*>(&lw[i]), h2 = *reinterpret_cast<const __half2*>(&hw[i]);
                    const __half2 q2 = *reinterpret_cast<const __half2*>(&qw[i]);
                    // a NaN gap (inf - inf: coordinates beyond the half range) is dropped by hmax2, and such a query has |delta| = inf
                    const __half2 e = __hmax2(__hmax2(__hsub2(l2, q2), __hsub2(q2, h2)), zero2);
                    if (i & 1) acc1 = __hfma2(e, e, acc1); else acc0 = __hfma2(e, e, acc0);
                }
            }
            // the two accumulators and then their two halves are added in half precision (two more roundings, inside the margin)
            const __half2 acc = __hadd2(acc0, acc1);
            const float sum = __half2float(__hadd(__low2half(acc), __high2half(acc)));
            // qm.z = pruneL / iw_q and qm.w = |delta|, both scaled by sqrt(margin); meta.w = 1 / bound of the child
            const float thr = fmaf(qm.z, meta.w, qm.w);
            r.childRef = __float_as_uint(meta.y);
            // a half-precision sum saturates at 65504: thresholds beyond that cannot be decided here, the child is kept
            const float lim = thr * thr;
            r.pass = (sum <= lim || lim >= 6.0e4f) && __float_as_uint(meta.z) > __float_as_uint(qm.y);
            return r;
        } else {
            float4 lo[V], hi[V], qv[V];
#pragma unroll
            for (int k = 0; k < V; ++k) lo[k] = __ldg(b + k * kFan);
#pragma unroll
            for (int k = 0; k < V; ++k) hi[k] = __ldg(b + (V + k) * kFan);
            const float4 meta = __ldg(b + 2 * V * kFan);
#pragma unroll
            for (int k = 0; k < V; ++k) qv[k] = qrow[k];
            const float4 qm = qrow[V];
            const float s = qm.x * meta.x;
            const float d2 = box_dist2<V>(qv, lo, hi);
            r.childRef = __float_as_uint(meta.y);
            // the child covers sorted positions [.., endPos): keep it only if some of them lie behind the query (endPos > qpos + 1);
            // null and padding children have endPos = 0
            r.pass = (d2 * s * s <= fp.pruneL2) && __float_as_uint(meta.z) > __float_as_uint(qm.y);
            return r;
        }
    };
    // Point round: one (leaf, query) pair per 8-lane group and slot; lane c tests point c of the leaf with the exact predicate.
    struct PointSlot { int idx; uint32_t qq; float d2, ws; bool hit; };
    auto testPoint = [&](uint32_t entry, bool active) {
        PointSlot r;
        r.qq = entry >> 27;
        r.idx = active ? (int)(entry & kRefMask) * kFan + c : 0;
        const float4* qrow = myQ + r.qq * QROW;
        if constexpr (HALF && WB_POINT_HALF) {
            // half-precision prefilter: |p - q| >= |p_h - q_h| (1 - eps) - |delta_q| - |delta_p|; survivors are tested exactly in resolveHit
            float4 ph[HV], qh[HV];
#pragma unroll
            for (int k = 0; k < HV; ++k) ph[k] = __ldg(t.ptsH + (int64_t)k * t.stride[0] + r.idx);
            const float4 pm = __ldg(t.pmeta + r.idx);
#pragma unroll
            for (int k = 0; k < HV; ++k) qh[k] = qrow[V + 1 + k];
            const float4 qm = qrow[V];
            const __half2 zero2 = __float2half2_rn(0.f);
            __half2 acc0 = zero2, acc1 = zero2;
#pragma unroll
            for (int k = 0; k < HV; ++k) {
                const float pw[4] = {ph[k].x, ph[k].y, ph[k].z, ph[k].w}, qw[4] = {qh[k].x, qh[k].y, qh[k].z, qh[k].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 e = __hsub2(*reinterpret_cast<const __half2*>(&pw[i]), *reinterpret_cast<const __half2*>(&qw[i]));
                    if (i & 1) acc1 = __hfma2(e, e, acc1); else acc0 = __hfma2(e, e, acc0);
                }
            }
            const __half2 acc = __hadd2(acc0, acc1);
            const float sum = __half2float(__hadd(__low2half(acc), __high2half(acc)));
            const float thr = fmaf(qm.z, pm.y, qm.w + pm.z);
            const float lim = thr * thr;
            r.ws = qm.x * pm.x;
            r.d2 = -1.f;                                       // computed exactly by resolveHit
            // NaN sums (inf - inf: both beyond the half range) must not prune: !(sum > lim)
            r.hit = active && (uint32_t)r.idx >= __float_as_uint(qm.y) && (!(sum > lim) || lim >= 6.0e4f);   // idx > qpos
            return r;
        } else {
            float4 pu[V], qv[V];
#pragma unroll
            for (int k = 0; k < V; ++k) pu[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + r.idx);
            const float iwu = __ldg(t.bound[0] + r.idx);
#pragma unroll
            for (int k = 0; k < V; ++k) qv[k] = qrow[k];
            const float4 qm = qrow[V];
            r.ws = qm.x * iwu;
            r.d2 = point_dist2<V>(qv, pu);
            r.hit = active && (uint32_t)r.idx >= __float_as_uint(qm.y) && r.d2 * r.ws * r.ws <= fp.pruneL2;   // idx > qpos
            return r;
        }
    };
    // A hit is resolved by the lane that found it: exact predicate, neighbour filter, then the term goes to the rows of both
    // vertices (hits are rare - a handful per query - so this branch is cold).
    auto resolveHit = [&](const PointSlot& r) {
        if (!r.hit) return;
        float d2 = r.d2, ws = r.ws;
        if constexpr ((HALF && WB_POINT_HALF) || WB_HIT_BATCH) {   // the exact squared distance, same operation order as the fp32 point round
            const float4* qrow = myQ + r.qq * QROW;
            float4 pu[V], qv[V];
#pragma unroll
            for (int k = 0; k < V; ++k) pu[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + r.idx);
#pragma unroll
            for (int k = 0; k < V; ++k) qv[k] = qrow[k];
            d2 = point_dist2<V>(qv, pu);
            ws = qrow[V].x * __ldg(t.bound[0] + r.idx);
        }
        const float dist = sqrtf(d2);
        if (dist > 0.f && !(dist * ws <= L)) return;         // exact predicate; dist <= 0 is the coincident case
        const float4* qrow = myQ + r.qq * QROW;
        const int u = __ldg(t.ids + r.idx);
        const int v = __ldg(t.ids + (__float_as_uint(qrow[V].y) - 1u));   // the query's vertex (hits are rare: looked up here, not carried)
        // pairs with a heavy vertex belong to that vertex' block (k_repulse_heavy)
        if ((heavySlot && __ldg(heavySlot + u) >= 0) || is_neighbor(col, __ldg(rowPtr + v), __ldg(rowPtr + v + 1), u)) return;
        long long* fv = forceRep + (int64_t)v * RS;
        long long* fu = forceRep + (int64_t)u * RS;
        if (dist <= 0.f) {
            fixed_add(fv + 4 * V + 1, 1ll);
            fixed_add(fu + 4 * V + 1, 1ll);
            return;
        }
        if (fp.dim == 1) {                                   // unit vector exactly +-1
            const long long f = to_fixed(copysignf(fp.repulsionScale * ws, qrow[0].x - __ldg(t.lo[0] + r.idx).x), fp.fixForce);
            fixed_add(fv, f);
            fixed_add(fu, -f);
        } else {
            const float sc = fp.repulsionScale * ws / dist;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const float4 q = qrow[k], pu = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + r.idx);
                const float e[4] = {sc * (q.x - pu.x), sc * (q.y - pu.y), sc * (q.z - pu.z), sc * (q.w - pu.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (4 * k + i < fp.dim) {
                        const long long f = to_fixed(e[i], fp.fixForce);
                        fixed_add(fv + 4 * k + i, f);
                        fixed_add(fu + 4 * k + i, -f);
                    }
                }
            }
        }
        const long long l = to_fixed(L / ws - dist, fp.fixLoss);
        fixed_add(fv + 4 * V, l);
        fixed_add(fu + 4 * V, l);
        nPairs += 2;                                         // counted per direction, like the reference's loop over all v
    };

#if WB_HIT_BATCH
    // Hits wait in the warp's buffer and are resolved 32 at a time, one per lane, so that the neighbour filter's dependent loads
    // of up to 32 hits overlap instead of one lane's search stalling the warp (integer rows: the order of the adds does not matter).
    auto queueHit = [&](const PointSlot& r) {
        const uint32_t hm = __ballot_sync(0xffffffffu, r.hit);
        if (r.hit) myHit[nHit + __popc(hm & ltMask)] = make_uint2((uint32_t)r.idx, r.qq);
        nHit += __popc(hm);
    };
    auto flushHits = [&](int keep) {                         // until at most `keep` hits are left
        __syncwarp();
        while (nHit > keep) {
            const int take = min(32, nHit);
            PointSlot r;
            r.hit = lane < take;
            const uint2 h = myHit[nHit - take + (r.hit ? lane : 0)];
            r.idx = (int)h.x; r.qq = h.y; r.d2 = 0.f; r.ws = 0.f;
            resolveHit(r);
            nHit -= take;
            __syncwarp();
        }
    };
#endif
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = atomicAdd(chunkCounter, 1);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= numChunks) break;
        const int qBase = lay.position(chunk * queriesPerUnit);   // sorted position of lane 0's query (a unit never straddles a block)
        if (qBase >= n) continue;                             // padding of the last block
        const int qi = qBase + lane;
        bool valid = lane < queriesPerUnit && qi < n;
        int vertex = valid ? __ldg(t.ids + qi) : 0;
        // heavy vertices (thousands of partners each) are walked by k_repulse_heavy, one block per vertex
        if (valid && heavySlot && __ldg(heavySlot + vertex) >= 0) valid = false;
        {
            float4* row = myQ + lane * QROW;
#pragma unroll
            for (int k = 0; k < V; ++k) row[k] = valid ? __ldg(t.lo[0] + (int64_t)k * t.stride[0] + qi) : make_float4(0.f, 0.f, 0.f, 0.f);
            float delta = 0.f;
            if constexpr (HALF) {
                // the query as the box rounds see it: q - centre rounded to half precision, and how far that moved it
                float d2 = 0.f;
#pragma unroll
                for (int k = 0; k < HV; ++k) {
                    __half2 h[4];
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int ch = 2 * k + half;
                        float e[4] = {0.f, 0.f, 0.f, 0.f};
                        if (ch < V) {
                            const float4 q = row[ch];
                            e[0] = q.x - t.quant->centre[4 * ch]; e[1] = q.y - t.quant->centre[4 * ch + 1];
                            e[2] = q.z - t.quant->centre[4 * ch + 2]; e[3] = q.w - t.quant->centre[4 * ch + 3];
                        }
                        h[2 * half] = __floats2half2_rn(e[0], e[1]);
                        h[2 * half + 1] = __floats2half2_rn(e[2], e[3]);
                        const float2 b0 = __half22float2(h[2 * half]), b1 = __half22float2(h[2 * half + 1]);
                        d2 = fmaf(e[0] - b0.x, e[0] - b0.x, d2); d2 = fmaf(e[1] - b0.y, e[1] - b0.y, d2);
                        d2 = fmaf(e[2] - b1.x, e[2] - b1.x, d2); d2 = fmaf(e[3] - b1.y, e[3] - b1.y, d2);
                    }
                    float4 packed;
                    packed.x = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[0]));
                    packed.y = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[1]));
                    packed.z = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[2]));
                    packed.w = __uint_as_float(*reinterpret_cast<const uint32_t*>(&h[3]));
                    row[V + 1 + k] = packed;
                }
                // rounded up generously; a coordinate beyond the half range gives inf - x = inf (or NaN): everything passes for it
                delta = sqrtf(d2) * 1.001f * kHalfMarginRoot;
                if (!(delta >= 0.f)) delta = __int_as_float(0x7f800000);
            }
            // {iw (point rounds), sorted position + 1, box-round threshold factor pruneL * sqrt(margin) / iw, |delta| * sqrt(margin)}
            const float iwq = valid ? __ldg(t.bound[0] + qi) : 1.f;
            row[V] = make_float4(iwq, __uint_as_float((uint32_t)qi + 1u), pruneL * kHalfMarginRoot * 1.000001f * __frcp_ru(iwq), delta);
        }
        const uint32_t validMask = __ballot_sync(0xffffffffu, valid);
        if (valid) myStack[__popc(validMask & ltMask)] = ((uint32_t)lane << 27) | rootBlock;
        int sp = __popc(validMask), nLeaf = 0;
        __syncwarp();
        while (sp > 0 || nLeaf > 0) {
            // a round pops up to eight pairs: two per 8-lane group, tested back to back so their loads overlap.  Point rounds run
            // whenever two full slots of leaves are waiting (which keeps the leaf queue below 16 + 64 entries) or nothing else is left.
            if (nLeaf >= 16 || sp == 0) {
                const int take = min(8, nLeaf);
                const bool activeA = g < take, activeB = g + 4 < take;
                const uint32_t entryA = myLeaf[activeA ? nLeaf - 1 - g : 0];
                const uint32_t entryB = myLeaf[activeB ? nLeaf - 5 - g : 0];
                nLeaf -= take;
                nTests += (int)activeA + (int)activeB;
                const PointSlot a = testPoint(entryA, activeA);
                const PointSlot b = testPoint(entryB, activeB);
#if WB_HIT_BATCH
                queueHit(a);
                queueHit(b);
                if (nHit >= 32) flushHits(31);
#else
                resolveHit(a);
                resolveHit(b);
#endif
            } else {
                // a pop of fewer than eight pairs reads the null entries below the stack (they fail the position test)
                const uint32_t entryA = myStack[sp - 1 - g];
                const uint32_t entryB = myStack[sp - 5 - g];
                const int take = min(8, sp);
                sp -= take;
                boxSlots += take;
                const BoxSlot a = testBox(entryA);
                const BoxSlot b = testBox(entryB);
                __syncwarp();                          // every lane has read its entries before the stack is overwritten
                // passing boxes of level >= 2 go back to the stack (as their children's block), passing leaves to the leaf queue
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const BoxSlot& r = h == 0 ? a : b;
                    const bool isLeaf = (r.childRef & kLeafFlag) != 0u;
                    const uint32_t pm = __ballot_sync(0xffffffffu, r.pass), lm = __ballot_sync(0xffffffffu, isLeaf);
                    const int rank = __popc((isLeaf ? (pm & lm) : (pm & ~lm)) & before);
                    const uint32_t e = (r.entry & ~kRefMask) | (r.childRef & kRefMask);
                    uint32_t* dst = isLeaf ? myLeaf + nLeaf : myStack + sp;
                    if (r.pass) dst[rank] = e;
                    const int leaves = __popc(pm & lm);
                    nLeaf += leaves;
                    sp += __popc(pm) - leaves;
                }
            }
            __syncwarp();
        }
#if WB_HIT_BATCH
        flushHits(0);                                         // before the next chunk overwrites the query rows
#endif
    }
    // per-warp statistics (integers, so the order in which warps took chunks cannot change the reduced value);
    // every box slot is 8 lane tests and all 32 lanes counted it: 8 / 32 per lane
    double totalPairs = (double)nPairs, totalTests = (double)nTests, totalBoxTests = 0.25 * (double)boxSlots;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        totalPairs += __shfl_xor_sync(0xffffffffu, totalPairs, o);
        totalTests += __shfl_xor_sync(0xffffffffu, totalTests, o);
        totalBoxTests += __shfl_xor_sync(0xffffffffu, totalBoxTests, o);
    }
    if (lane == 0) {
        const int64_t w = (int64_t)blockIdx.x * WARPS + warp;
        partials[3 * w] = totalPairs;
        partials[3 * w + 1] = totalTests;
        partials[3 * w + 2] = totalBoxTests;
    }
}

// ---------------------------------------------------------------------------------------------
// Repulsion for heavy vertices (weight >= kHeavyWeight x the mean: hubs of heavy-tailed graphs).  The interaction radius grows
// like w^(1/d) and the number of partners like w, so a hub of weight 6000 (c4) has tens of thousands of in-radius partners and
// its ball covers most of the layout.  One block per heavy vertex scans the level-2 boxes with a fixed thread <-> box
// assignment, descends into passing leaves and points, and every thread applies its own hits: the vertex' own side to private
// accumulators (fixed-order block reduction, then one fixed-point add per component), the partner's side straight to the
// partner's row.  Every pair that involves a heavy vertex is handled here and only here (two heavy vertices: by the one at the
// lower sorted position).  Same predicates as the pair-stack walk, so the same pair set.
constexpr float kHeavyWeight = 32.0f;

template <int V>
__global__ void __launch_bounds__(256) k_repulse_heavy(const TreeView t, const int* __restrict__ rowPtr, const int* __restrict__ col, int n,
                                                       const ForceParams fp, long long* __restrict__ forceRep, const RepLayout lay,
                                                       const int* __restrict__ heavyVertex, const int* __restrict__ heavySlot,
                                                       const int* __restrict__ invOrder, double* __restrict__ partials /* [block][3] */) {
    constexpr int RS = 4 * V + 2, K = RS + 3;
    __shared__ double redBuf[8 * K];
    const int v = heavyVertex[blockIdx.x];
    const int p = invOrder[v];
    double vals[K];
#pragma unroll
    for (int k = 0; k < K; ++k) vals[k] = 0.0;
    // in a sharded run every rank launches all heavy vertices and keeps those whose sorted position falls in its blocks
    const bool mine = ((p >> 5) / kRepBlockChunks) % lay.world == lay.rank;
    if (mine) {
        float4 q[V];
#pragma unroll
        for (int k = 0; k < V; ++k) q[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + p);
        const float iwq = __ldg(t.bound[0] + p);
        const int rowBegin = __ldg(rowPtr + v), rowEnd = __ldg(rowPtr + v + 1);
        const float L = fp.edgeLength;
        const int top = t.numLevels >= 2 ? 2 : 1;            // level the flat scan starts from
        auto box = [&](int lv, int idx, float& bnd) {
            float4 lo[V], hi[V];
            const int64_t st = t.stride[lv];
#pragma unroll
            for (int k = 0; k < V; ++k) { lo[k] = __ldg(t.lo[lv] + k * st + idx); hi[k] = __ldg(t.hi[lv] + k * st + idx); }
            bnd = __ldg(t.bound[lv] + idx);
            return box_dist2<V>(q, lo, hi);
        };
        auto passes = [&](float d2, float bnd) { const float s = iwq * bnd; return d2 * s * s <= fp.pruneL2; };
        auto leaf = [&](int leafIdx) {
            for (int j = 0; j < kFan; ++j) {
                const int idx = leafIdx * kFan + j;
                if (idx >= n) break;
                vals[RS + 1] += 1.0;
                float4 pu[V];
#pragma unroll
                for (int k = 0; k < V; ++k) pu[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + idx);
                const float iwu = __ldg(t.bound[0] + idx);
                const float d2 = box_dist2<V>(q, pu, pu);
                if (!passes(d2, iwu) || idx == p) continue;
                const float dist = sqrtf(d2);
                const float ws = iwq * iwu;
                if (dist > 0.f && !(dist * ws <= L)) continue;
                const int u = __ldg(t.ids + idx);
                if (idx < p && __ldg(heavySlot + u) >= 0) continue;      // two heavy vertices: the lower position owns the pair
                if (is_neighbor(col, rowBegin, rowEnd, u)) continue;
                long long* fu = forceRep + (int64_t)u * RS;
                if (dist <= 0.f) { vals[4 * V + 1] += 1.0; fixed_add(fu + 4 * V + 1, 1ll); continue; }
                const float sc = fp.repulsionScale * ws / dist;
                if (fp.dim == 1) {
                    const float e = copysignf(fp.repulsionScale * ws, q[0].x - pu[0].x);
                    vals[0] += (double)e;
                    fixed_add(fu, -to_fixed(e, fp.fixForce));
                } else {
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float e[4] = {sc * (q[k].x - pu[k].x), sc * (q[k].y - pu[k].y), sc * (q[k].z - pu[k].z), sc * (q[k].w - pu[k].w)};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (4 * k + i < fp.dim) {
                                vals[4 * k + i] += (double)e[i];
                                fixed_add(fu + 4 * k + i, -to_fixed(e[i], fp.fixForce));
                            }
                        }
                    }
                }
                const float l = L / ws - dist;
                vals[4 * V] += (double)l;
                fixed_add(fu + 4 * V, to_fixed(l, fp.fixLoss));
                vals[RS] += 2.0;
            }
        };
        for (int node = threadIdx.x; node < t.count[top]; node += 256) {
            float bnd;
            vals[RS + 2] += 1.0;
            if (!passes(box(top, node, bnd), bnd)) continue;
            if (top == 1) { leaf(node); continue; }
            for (int c = 0; c < kFan; ++c) {
                const int lf = node * kFan + c;
                if (lf >= t.count[1]) break;
                vals[RS + 2] += 1.0;
                if (passes(box(1, lf, bnd), bnd)) leaf(lf);
            }
        }
    }
    __shared__ double total[K];
    block_sum<K, 256>(vals, redBuf, total);
    if (threadIdx.x < RS && mine) {
        // other heavy blocks may be adding their side of a pair to this row at the same time
        const int k = threadIdx.x;
        const long long f = k < 4 * V ? __double2ll_rn(total[k] * fp.fixForce) : (k == 4 * V ? __double2ll_rn(total[k] * fp.fixLoss) : (long long)total[k]);
        fixed_add(forceRep + (int64_t)v * RS + k, f);
    }
    if (threadIdx.x < 3) partials[(int64_t)blockIdx.x * 3 + threadIdx.x] = total[RS + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// Attraction, centre force and optimizer (WembedEmbedder.cpp:260-272, 140-172, 296-301; AdamOptimizer.cpp:15-30).
//
// Layout of the work: G = V (rounded up to a power of two) lanes share one vertex and lane c owns float4 chunk c of every
// row it touches - its own row, the neighbours' rows, the force, the Adam moments.  For one edge the G lanes read the
// neighbour's row with ONE coalesced access (16 B per lane), add their partial squared distances with log2(G) shuffles and
// each accumulates its own four force components in double; nothing has to be reduced at the end and every lane is busy in the
// optimizer epilogue.  The pair weight ws(v,u) = iw_v * iw_u is read from a per-CSR-entry array (weights are constant during a
// run, WembedEmbedder.cpp:121-131), so the only gather per edge is the neighbour row.

// ws of every CSR entry (recomputed by wb_set_weights)
__global__ void k_edge_weights(const int* __restrict__ rowPtr, const int* __restrict__ col, const float* __restrict__ iw, int n,
                               float* __restrict__ edgeWs) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float iwv = iw[v];
    for (int e = rowPtr[v]; e < rowPtr[v + 1]; ++e) edgeWs[e] = iwv * iw[col[e]];
}

__host__ __device__ constexpr int attract_lanes(int V) { return V <= 1 ? 1 : (V <= 2 ? 2 : (V <= 4 ? 4 : 8)); }

__device__ __forceinline__ float chunk_dist2(float4 a, float4 b) {
    float e, d2;
    e = a.x - b.x; d2 = e * e;
    e = a.y - b.y; d2 = fmaf(e, e, d2);
    e = a.z - b.z; d2 = fmaf(e, e, d2);
    e = a.w - b.w; d2 = fmaf(e, e, d2);
    return d2;
}

// one attractive pair, chunk view (attractionForce, WembedEmbedder.cpp:140-172): d2 is the full squared distance
__device__ __forceinline__ void attract_chunk(float4 xv, float4 xu, float d2, float ws, float L, float scale, int dim, double (&acc)[4],
                                              double& loss, int& nCoincident) {
    const float dist = sqrtf(d2);
    if (dist <= 0.f) { ++nCoincident; return; }               // :150-155, resolved by the caller
    if (dist * ws > L) {                                       // :163-168
        loss += (double)(dist - L / ws);
        if (dim == 1) { acc[0] += (double)copysignf(scale * ws, xu.x - xv.x); return; }   // unit vector exactly +-1
        const float s = scale * ws / dist;
        acc[0] += (double)(s * (xu.x - xv.x));
        acc[1] += (double)(s * (xu.y - xv.y));
        acc[2] += (double)(s * (xu.z - xv.z));
        acc[3] += (double)(s * (xu.w - xv.w));
    }
}

// Hub rows (degree > hubThreshold; heavy-tailed graphs have rows of 1e4-1e5 entries): one block per hub strides over the row,
// sums in double and reduces in a fixed order; k_attract_update picks the result up instead of walking the row itself.
template <int V>
__global__ void __launch_bounds__(256) k_attract_hubs(const float4* __restrict__ x, const float* __restrict__ edgeWs, const int* __restrict__ rowPtr,
                                                      const int* __restrict__ col, const int* __restrict__ hubVertex, const ForceParams fp,
                                                      double* __restrict__ hubForce /* [hub][4V + 2] */) {
    constexpr int K = 4 * V + 2;
    __shared__ double redBuf[8 * K];
    const int v = hubVertex[blockIdx.x];
    float4 xv[V];
    load_row<V>(x, v, xv);
    double vals[K];
#pragma unroll
    for (int k = 0; k < K; ++k) vals[k] = 0.0;
    int nCoincident = 0;
    const int end = __ldg(rowPtr + v + 1);
    for (int e = __ldg(rowPtr + v) + threadIdx.x; e < end; e += 256) {
        const int u = __ldg(col + e);
        const float ws = __ldg(edgeWs + e);
        float4 xu[V];
        load_row<V>(x, u, xu);
        const float d2 = point_dist2<V>(xu, xv);
        int coincidentHere = 0;
#pragma unroll
        for (int c = 0; c < V; ++c) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0}, loss = 0.0;
            int nc = 0;
            attract_chunk(xv[c], xu[c], d2, ws, fp.edgeLength, fp.attractionScale, fp.dim, acc, loss, nc);
            vals[4 * c] += acc[0]; vals[4 * c + 1] += acc[1]; vals[4 * c + 2] += acc[2]; vals[4 * c + 3] += acc[3];
            if (c == 0) { vals[4 * V] += loss; coincidentHere = nc; }
        }
        nCoincident += coincidentHere;
    }
    vals[4 * V + 1] = (double)nCoincident;
    block_sum<K, 256>(vals, redBuf, hubForce + (int64_t)blockIdx.x * K);
}

#ifndef WB_ATTRACT_FAST
#define WB_ATTRACT_FAST 1          // 0: IEEE sqrtf / divisions behind per-edge branches (the round-1 kernel, kept for A/B builds)
#endif
#ifndef WB_ATTRACT_BATCH
#define WB_ATTRACT_BATCH 4         // neighbour rows in flight per lane (measured: 2, 6 and 8 are slower, profiles/r1_summary.md)
#endif
#ifndef WB_ATTRACT_MINBLOCKS
#define WB_ATTRACT_MINBLOCKS 4
#endif
// The north_star's "fused step kernel".  Each block owns a fixed contiguous vertex range and emits
// {lossA, lossR, sum_v xnew[v][k]} for the deterministic reducer.
template <int V>
__global__ void __launch_bounds__(256, WB_ATTRACT_MINBLOCKS) k_attract_update(const float4* __restrict__ x, const float* __restrict__ edgeWs,
                                                        const int* __restrict__ rowPtr, const int* __restrict__ col, int rangeBegin,
                                                        int rangeEnd, int vertsPerBlock, const ForceParams fp,
                                                        const long long* __restrict__ forceRep, const int* __restrict__ hubSlot,
                                                        const double* __restrict__ hubForce, float4* __restrict__ xNew,
                                                        float4* __restrict__ mom1, float4* __restrict__ mom2,
                                                        float4* __restrict__ forceOut, double* __restrict__ partials) {
    constexpr int G = attract_lanes(V), VPW = 32 / G, VPB = 8 * VPW, K = 2 + 4 * V, RS = 4 * V + 2;
    __shared__ uint32_t mtState[8][624];
    __shared__ double unitBuf[8][VPW][4 * V];
    __shared__ double redBuf[8][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, gi = lane / G;
    const bool chunkLane = c < V;                                  // lanes G > V (V = 3, 5, 6, 7) only take part in the shuffles
    const int vBegin = rangeBegin + blockIdx.x * vertsPerBlock;     // [rangeBegin, rangeEnd): the vertices this rank owns
    const int vEnd = min(rangeEnd, vBegin + vertsPerBlock);
    double sumLossA = 0.0, sumLossR = 0.0, sumX[4] = {0.0, 0.0, 0.0, 0.0};
    const float L = fp.edgeLength;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int vBase = vBegin; vBase < vEnd; vBase += VPB) {
        const int v = vBase + warp * VPW + gi;
        const bool valid = v < vEnd;
        const int64_t at = (int64_t)v * V + c;
        float4 xv = zero4;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};      // summed in double, see k_repulse_pairs
        double loss = 0.0;
        int nCoincident = 0;
        int e = 0, end = 0, hub = -1;
        const long long* rep = forceRep;                           // this vertex' row of repulsion results (fixed point)
        if (valid) {
            rep = forceRep + (int64_t)v * RS;
            if (chunkLane) xv = __ldg(x + at);
            hub = hubSlot ? __ldg(hubSlot + v) : -1;
            if (hub < 0) { e = __ldg(rowPtr + v); end = __ldg(rowPtr + v + 1); }
        }
        // all G lanes of a vertex walk the same edges; groups of one warp have different row lengths, the shuffles below need
        // every lane, so the warp iterates to the longest row of its groups (rows beyond hubThreshold are pre-summed)
        int len = end - e;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
        constexpr int B = WB_ATTRACT_BATCH;
        for (int i = 0; i < len; i += B) {                         // neighbours in ascending order, B rows in flight
            bool has[B];
            int u[B];
            float wsE[B], dd[B];
            float4 r[B];
#pragma unroll
            for (int j = 0; j < B; ++j) {
                has[j] = e + i + j < end;
                u[j] = has[j] ? __ldg(col + e + i + j) : 0;
                wsE[j] = has[j] ? __ldg(edgeWs + e + i + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < B; ++j) r[j] = (has[j] && chunkLane) ? __ldg(x + (int64_t)u[j] * V + c) : xv;
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
#if WB_ATTRACT_FAST
            // Branch-free pair arithmetic with single-instruction rsqrt / rcp (relative error <= 2^-22, the size of the fp32
            // rounding of the terms themselves): ~30 instructions per edge instead of ~85 with IEEE sqrtf and two IEEE divisions
            // behind per-edge branches.  One-dimensional embeddings keep the exact +-1 unit vectors below.
            if (V > 1 || fp.dim > 1) {
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    // squared distances below FLT_MIN (dist < 1.1e-19) are neither coincident (that is dist == 0 exactly, as with
                    // sqrtf) nor can they exceed the edge length: they contribute nothing and stay away from the .ftz rsqrt
                    const float inv = rsqrt_approx(dd[j]);
                    const float dist = dd[j] * inv;
                    nCoincident += (int)(has[j] && dd[j] == 0.f);                        // :150-155, resolved below
                    const bool act = has[j] && dd[j] >= kFltMin && dist * wsE[j] > L;     // :163-168
                    const float sc = act ? fp.attractionScale * wsE[j] * inv : 0.f;
                    bx = fmaf(sc, r[j].x - xv.x, bx); by = fmaf(sc, r[j].y - xv.y, by);
                    bz = fmaf(sc, r[j].z - xv.z, bz); bw = fmaf(sc, r[j].w - xv.w, bw);
                    bl += act ? fmaf(-L, rcp_approx(wsE[j]), dist) : 0.f;
                }
            } else
#endif
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if (!has[j]) continue;
                const float dist = sqrtf(dd[j]);
                if (dist <= 0.f) { ++nCoincident; continue; }           // :150-155, resolved below
                if (dist * wsE[j] > L) {                                 // :163-168
                    if (fp.dim == 1) {
                        // one dimension: the unit vector is exactly +-1 (VectorOperations.hpp:19-24), so symmetric neighbours
                        // cancel exactly as they do in the reference
                        bx += copysignf(fp.attractionScale * wsE[j], r[j].x - xv.x);
                    } else {
                        const float sc = fp.attractionScale * wsE[j] / dist;
                        bx = fmaf(sc, r[j].x - xv.x, bx); by = fmaf(sc, r[j].y - xv.y, by);
                        bz = fmaf(sc, r[j].z - xv.z, bz); bw = fmaf(sc, r[j].w - xv.w, bw);
                    }
                    bl += dist - L / wsE[j];
                }
            }
            acc[0] += (double)bx; acc[1] += (double)by; acc[2] += (double)bz; acc[3] += (double)bw;
            loss += (double)bl;
        }
        if (valid && hub >= 0) {
            const double* hf = hubForce + (int64_t)hub * (4 * V + 2);
            if (chunkLane) { acc[0] = hf[4 * c]; acc[1] = hf[4 * c + 1]; acc[2] = hf[4 * c + 2]; acc[3] = hf[4 * c + 3]; }
            loss = hf[4 * V];
            nCoincident = (int)hf[4 * V + 1];
        }
        if (valid) nCoincident += (int)__ldg(rep + 4 * V + 1);

        // coincident partners: every one of them adds the same unit vector (generator re-created per pair, :150-155, :183-188).
        // The generator state (624 words) lives in per-warp shared memory; the rare vertices that need it take turns.
        uint32_t need = __ballot_sync(0xffffffffu, nCoincident > 0 && c == 0);
        if (need) {
            uint32_t todo = need;
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1u;
                if (lane == l) random_unit_vector(mtState[warp], fp.seed, (uint32_t)v, fp.iteration, fp.dim, unitBuf[warp][gi]);
                __syncwarp();
            }
            if (nCoincident > 0 && chunkLane) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * c + i < fp.dim) acc[i] += nCoincident * unitBuf[warp][gi][4 * c + i];
            }
            __syncwarp();
        }

        if (valid && c == 0) {
            sumLossA += loss;
            sumLossR += (double)__ldg(rep + 4 * V) * fp.invFixLoss;
        }
        if (valid && chunkLane) {
            // rows are (32 V + 16) bytes long, so every chunk of four integers is 16-byte aligned
            const longlong2 f01 = __ldg(reinterpret_cast<const longlong2*>(rep + 4 * c)), f23 = __ldg(reinterpret_cast<const longlong2*>(rep + 4 * c) + 1);
            float4 f = make_float4((float)(acc[0] + (double)f01.x * fp.invFixForce), (float)(acc[1] + (double)f01.y * fp.invFixForce),
                                   (float)(acc[2] + (double)f23.x * fp.invFixForce), (float)(acc[3] + (double)f23.y * fp.invFixForce));
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
                    const float mHat = me[i] * fp.invBias1, vHat = se[i] * fp.invBias2;
#if WB_ATTRACT_FAST
                    xe[i] = fmaf(fp.lr * mHat, rcp_approx(sqrt_approx(vHat) + fp.eps), xe[i]);
#else
                    xe[i] += fp.lr * mHat / (sqrtf(vHat) + fp.eps);
#endif
                }
                mom1[at] = make_float4(me[0], me[1], me[2], me[3]);
                mom2[at] = make_float4(se[0], se[1], se[2], se[3]);
                xn = make_float4(xe[0], xe[1], xe[2], xe[3]);
            } else {
                const float cap = fp.maxDisplacement;
                xn.x = xv.x + fminf(fmaxf(f.x, -cap), cap) * fp.lr;
                xn.y = xv.y + fminf(fmaxf(f.y, -cap), cap) * fp.lr;
                xn.z = xv.z + fminf(fmaxf(f.z, -cap), cap) * fp.lr;
                xn.w = xv.w + fminf(fmaxf(f.w, -cap), cap) * fp.lr;
            }
            xNew[at] = xn;
            sumX[0] += (double)xn.x; sumX[1] += (double)xn.y; sumX[2] += (double)xn.z; sumX[3] += (double)xn.w;
        }
    }
    // fixed-order block reduction: lanes that own the same chunk add up (xor offsets G, 2G, ..), then the 8 warps in order
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
        sumLossA += __shfl_xor_sync(0xffffffffu, sumLossA, o);
        sumLossR += __shfl_xor_sync(0xffffffffu, sumLossR, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) sumX[i] += __shfl_xor_sync(0xffffffffu, sumX[i], o);
    }
    if (lane == 0) { redBuf[warp][0] = sumLossA; redBuf[warp][1] = sumLossR; }
    if (lane < G && chunkLane) {
#pragma unroll
        for (int i = 0; i < 4; ++i) redBuf[warp][2 + 4 * c + i] = sumX[i];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double sacc = 0.0;
        for (int w = 0; w < 8; ++w) sacc += redBuf[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * K + threadIdx.x] = sacc;
    }
}

// ---------------------------------------------------------------------------------------------
// k_attract_staged (WB_ATTRACT_STAGED=1; A/B candidate, NOT the default and not yet measured on a GPU):
// the same fused step with every once-read stream staged through shared memory by bulk asynchronous copies
// (cp.async.bulk + mbarrier), two stages deep, so that the only loads that occupy registers and scoreboards are the
// neighbour-row gathers.  profiles/r1_summary.md section 7: k_attract_update is bound by memory latency / memory-level
// parallelism - its dependent chain per pass is rowPtr -> {col, ws} -> rows -> {m, v, result row} - and every
// register-based prefetch lost to the register budget.  Here one elected thread copies, for the pass after the current
// one, the block's 256 / G own rows of x, m, v and forceRep, its window of rowPtr and the CSR entries {col, ws} of
// those rows (at most kStageEdges of them; a pass with more reads the rest from global memory), and the warps find all
// of it in shared memory when they get there.  Arithmetic and summation order are exactly those of k_attract_update, so
// results are bit-identical to it.
// Requirements on the host side (allocate(), WB_ATTRACT_STAGED): rowPtr, col and edgeWs padded by 8 entries (the copies
// move whole 16-byte groups), vertex ranges of a block a multiple of 256 / G.
#ifndef WB_ATTRACT_STAGED
#define WB_ATTRACT_STAGED 0
#endif
constexpr int kStageEdges = 2048;         // CSR entries staged per pass (c3: ~1 280 per 128 vertices)

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
struct AttractStage {                     // one pass of one block
    static constexpr int G = attract_lanes(V), VPB = 256 / G, RS = 4 * V + 2;
    float4 x[VPB * V], m[VPB * V], s[VPB * V];
    long long rep[VPB * RS];
    int col[kStageEdges + 8];
    float ws[kStageEdges + 8];
    int rowPtr[VPB + 8];
};

template <int V>
__global__ void __launch_bounds__(256, 2)
k_attract_staged(const float4* __restrict__ x, const float* __restrict__ edgeWs, const int* __restrict__ rowPtr, const int* __restrict__ col,
                 int rangeBegin, int rangeEnd, int vertsPerBlock, const ForceParams fp, const long long* __restrict__ forceRep,
                 const int* __restrict__ hubSlot, const double* __restrict__ hubForce, float4* __restrict__ xNew, float4* __restrict__ mom1,
                 float4* __restrict__ mom2, float4* __restrict__ forceOut, double* __restrict__ partials, uint32_t* __restrict__ mtScratch) {
    using Stage = AttractStage<V>;
    constexpr int G = Stage::G, VPW = 32 / G, VPB = Stage::VPB, K = 2 + 4 * V, RS = Stage::RS, B = 4;
    extern __shared__ __align__(128) unsigned char smemAtt[];
    Stage* stage = reinterpret_cast<Stage*>(smemAtt);                                  // [2]
    uint64_t* full = reinterpret_cast<uint64_t*>(smemAtt + 2 * sizeof(Stage));         // [2]
    double (*unitBuf)[4 * V] = reinterpret_cast<double (*)[4 * V]>(full + 2);          // [8]
    double (*redBuf)[K] = reinterpret_cast<double (*)[K]>(unitBuf + 8);                // [8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = lane % G, gi = lane / G;
    const bool chunkLane = c < V;
    const int vBegin = rangeBegin + blockIdx.x * vertsPerBlock;
    const int vEnd = min(rangeEnd, vBegin + vertsPerBlock);
    const int passes = vBegin < vEnd ? (vEnd - vBegin + VPB - 1) / VPB : 0;
    double sumLossA = 0.0, sumLossR = 0.0, sumX[4] = {0.0, 0.0, 0.0, 0.0};
    const float L = fp.edgeLength;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* xc = x + c;

    // producer state (thread 0 only): CSR bounds of the pass to be copied next
    int nextLo = 0, nextHi = 0;
    auto passBounds = [&](int p, int& lo, int& hi) {
        const int v0 = vBegin + p * VPB;
        lo = __ldg(rowPtr + v0);
        hi = __ldg(rowPtr + min(v0 + VPB, vEnd));
    };
    auto issue = [&](int p, int lo, int hi) {     // bulk copies of pass p into stage p & 1
        Stage& st = stage[p & 1];
        uint64_t* bar = full + (p & 1);
        const int v0 = vBegin + p * VPB, rows = min(VPB, vEnd - v0);
        const uint32_t rowBytes = (uint32_t)rows * V * 16u, repBytes = (uint32_t)rows * RS * 8u;
        const int rp0 = v0 & ~3;                                            // rowPtr window from an aligned entry
        const uint32_t rpBytes = (uint32_t)((v0 - rp0 + rows + 1 + 3) & ~3) * 4u;
        const int e0 = lo & ~3;                                             // entries from an aligned entry
        const int staged = min(hi - e0, kStageEdges + 4);
        const uint32_t edgeBytes = (uint32_t)((max(staged, 0) + 3) & ~3) * 4u;
        mbar_expect_tx(bar, 3u * rowBytes + repBytes + rpBytes + 2u * edgeBytes);
        bulk_copy(st.x, x + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.m, mom1 + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.s, mom2 + (int64_t)v0 * V, rowBytes, bar);
        bulk_copy(st.rep, forceRep + (int64_t)v0 * RS, repBytes, bar);
        bulk_copy(st.rowPtr, rowPtr + rp0, rpBytes, bar);
        if (edgeBytes) {
            bulk_copy(st.col, col + e0, edgeBytes, bar);
            bulk_copy(st.ws, edgeWs + e0, edgeBytes, bar);
        }
    };
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && passes > 0) {
        int lo, hi;
        passBounds(0, lo, hi);
        issue(0, lo, hi);
        if (passes > 1) passBounds(1, nextLo, nextHi);
    }

    for (int p = 0; p < passes; ++p) {
        // every warp has left stage (p + 1) & 1 (barrier at the end of pass p - 1): refill it, and fetch the bounds after that
        if (threadIdx.x == 0 && p + 1 < passes) {
            issue(p + 1, nextLo, nextHi);
            if (p + 2 < passes) passBounds(p + 2, nextLo, nextHi);
        }
        mbar_wait(full + (p & 1), (uint32_t)(p >> 1) & 1u);
        const Stage& st = stage[p & 1];
        const int v0 = vBegin + p * VPB;
        const int slot = warp * VPW + gi, v = v0 + slot;
        const bool valid = v < vEnd;
        const int rpOff = v0 - (v0 & ~3);
        const int e0 = st.rowPtr[rpOff] & ~3;                               // global index of staged entry 0
        float4 xv = zero4;
        double acc[4] = {0.0, 0.0, 0.0, 0.0}, loss = 0.0;
        int nCoincident = 0, e = 0, end = 0, hub = -1;
        if (valid) {
            if (chunkLane) xv = st.x[slot * V + c];
            hub = hubSlot ? __ldg(hubSlot + v) : -1;
            if (hub < 0) { e = st.rowPtr[rpOff + slot]; end = st.rowPtr[rpOff + slot + 1]; }
        }
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
                const bool inStage = at < kStageEdges + 4;
                u[j] = has[j] ? (inStage ? st.col[at] : __ldg(col + idx)) : 0;
                wsE[j] = has[j] ? (inStage ? st.ws[at] : __ldg(edgeWs + idx)) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < B; ++j) r[j] = (has[j] && chunkLane) ? __ldg(xc + (int64_t)u[j] * V) : xv;
#pragma unroll
            for (int j = 0; j < B; ++j) dd[j] = chunkLane ? chunk_dist2(r[j], xv) : 0.f;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < B; ++j) dd[j] += __shfl_xor_sync(0xffffffffu, dd[j], o);
            }
            float bx = 0.f, by = 0.f, bz = 0.f, bw = 0.f, bl = 0.f;
            if (V > 1 || fp.dim > 1) {                          // same arithmetic as k_attract_update (WB_ATTRACT_FAST)
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const float inv = rsqrt_approx(dd[j]);
                    const float dist = dd[j] * inv;
                    nCoincident += (int)(has[j] && dd[j] == 0.f);
                    const bool act = has[j] && dd[j] >= kFltMin && dist * wsE[j] > L;
                    const float sc = act ? fp.attractionScale * wsE[j] * inv : 0.f;
                    bx = fmaf(sc, r[j].x - xv.x, bx); by = fmaf(sc, r[j].y - xv.y, by);
                    bz = fmaf(sc, r[j].z - xv.z, bz); bw = fmaf(sc, r[j].w - xv.w, bw);
                    bl += act ? fmaf(-L, rcp_approx(wsE[j]), dist) : 0.f;
                }
            } else {                                            // one dimension: exact +-1 unit vectors, IEEE arithmetic
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
        const long long* rep = st.rep + slot * RS;
        if (valid && hub >= 0) {
            const double* hf = hubForce + (int64_t)hub * (4 * V + 2);
            if (chunkLane) { acc[0] = hf[4 * c]; acc[1] = hf[4 * c + 1]; acc[2] = hf[4 * c + 2]; acc[3] = hf[4 * c + 3]; }
            loss = hf[4 * V];
            nCoincident = (int)hf[4 * V + 1];
        }
        if (valid) nCoincident += (int)rep[4 * V + 1];
        uint32_t todo = __ballot_sync(0xffffffffu, nCoincident > 0 && c == 0);
        while (todo) {                                          // coincident partners (:150-155, :183-188), one vertex at a time
            const int l = __ffs(todo) - 1;
            todo &= todo - 1u;
            if (lane == l)
                random_unit_vector(mtScratch + ((size_t)blockIdx.x * 8 + warp) * 624, fp.seed, (uint32_t)v, fp.iteration, fp.dim, unitBuf[warp]);
            __syncwarp();
            if (lane / G == l / G && chunkLane) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * c + i < fp.dim) acc[i] += nCoincident * unitBuf[warp][4 * c + i];
            }
            __syncwarp();
        }
        if (valid && c == 0) {
            sumLossA += loss;
            sumLossR += (double)rep[4 * V] * fp.invFixLoss;
        }
        if (valid && chunkLane) {
            const int64_t at = (int64_t)v * V + c;
            const longlong2 f01 = *reinterpret_cast<const longlong2*>(rep + 4 * c), f23 = *(reinterpret_cast<const longlong2*>(rep + 4 * c) + 1);
            float4 f = make_float4((float)(acc[0] + (double)f01.x * fp.invFixForce), (float)(acc[1] + (double)f01.y * fp.invFixForce),
                                   (float)(acc[2] + (double)f23.x * fp.invFixForce), (float)(acc[3] + (double)f23.y * fp.invFixForce));
            if (fp.centreScale != 0.f) {
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
                    const float mHat = me[i] * fp.invBias1, vHat = se[i] * fp.invBias2;
                    xe[i] = fmaf(fp.lr * mHat, rcp_approx(sqrt_approx(vHat) + fp.eps), xe[i]);
                }
                mom1[at] = make_float4(me[0], me[1], me[2], me[3]);
                mom2[at] = make_float4(se[0], se[1], se[2], se[3]);
                xn = make_float4(xe[0], xe[1], xe[2], xe[3]);
            } else {
                const float cap = fp.maxDisplacement;
                xn.x = xv.x + fminf(fmaxf(f.x, -cap), cap) * fp.lr;
                xn.y = xv.y + fminf(fmaxf(f.y, -cap), cap) * fp.lr;
                xn.z = xv.z + fminf(fmaxf(f.z, -cap), cap) * fp.lr;
                xn.w = xv.w + fminf(fmaxf(f.w, -cap), cap) * fp.lr;
            }
            xNew[at] = xn;
            sumX[0] += (double)xn.x; sumX[1] += (double)xn.y; sumX[2] += (double)xn.z; sumX[3] += (double)xn.w;
        }
        __syncthreads();                                        // stage p & 1 may be refilled (pass p + 2) from here on
    }
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
        sumLossA += __shfl_xor_sync(0xffffffffu, sumLossA, o);
        sumLossR += __shfl_xor_sync(0xffffffffu, sumLossR, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) sumX[i] += __shfl_xor_sync(0xffffffffu, sumX[i], o);
    }
    if (lane == 0) { redBuf[warp][0] = sumLossA; redBuf[warp][1] = sumLossR; }
    if (lane < G && chunkLane) {
#pragma unroll
        for (int i = 0; i < 4; ++i) redBuf[warp][2 + 4 * c + i] = sumX[i];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double sacc = 0.0;
        for (int w = 0; w < 8; ++w) sacc += redBuf[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * K + threadIdx.x] = sacc;
    }
}
template <int V>
constexpr size_t attract_staged_smem() { return 2 * sizeof(AttractStage<V>) + 2 * sizeof(uint64_t) + 8 * (4 * V) * sizeof(double) + 8 * (2 + 4 * V) * sizeof(double); }


// ---------------------------------------------------------------------------------------------
// Deterministic reduction of per-block partial sums: block k reduces column k.
// Thread t adds rows t, t+256, ... in order, then the 256 thread sums are combined by a fixed tree.
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partials, int rows, int cols,
                                                         double* __restrict__ out) {
    __shared__ double sm[256];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int r = threadIdx.x; r < rows; r += 256) s += partials[(int64_t)r * cols + k];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[k] = sm[0];
}

// ---------------------------------------------------------------------------------------------
// applyGravityCentre + observeDisplacement (WembedEmbedder.cpp:303-352): x = xnew - centroid, and the sums of
// ||x - xprev|| and ||x||^2.  forceSums = output of the reducer for k_attract_update ({lossA, lossR, sum xnew[k]}).
template <int V>
__global__ void __launch_bounds__(256) k_recentre_observe(float4* __restrict__ x, const float4* __restrict__ xNew, int n, int rangeBegin,
                                                          int rangeEnd, int vertsPerBlock, int dim, const double* __restrict__ forceSums,
                                                          double* __restrict__ partials) {
    __shared__ double redBuf[8 * 2];
    float cen[4 * V];
#pragma unroll
    for (int k = 0; k < 4 * V; ++k) cen[k] = (k < dim) ? (float)(forceSums[2 + k] / (double)n) : 0.f;
    const int vBegin = rangeBegin + blockIdx.x * vertsPerBlock, vEnd = min(rangeEnd, vBegin + vertsPerBlock);
    double sums[2] = {0.0, 0.0};
    for (int v = vBegin + threadIdx.x; v < vEnd; v += 256) {
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
    block_sum<2, 256>(sums, redBuf, partials + (int64_t)blockIdx.x * 2);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU: every rank contributes `cols` partial sums; all ranks add them in rank order, so the totals are identical
// on every rank and do not depend on arrival order.
__global__ void k_sum_ranks(const double* __restrict__ gathered, int world, int cols, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= cols) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += gathered[r * cols + k];
    out[k] = s;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T* p, int64_t count, T value) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = value;
}

// ---------------------------------------------------------------------------------------------
// Reconstruction quality: evaluationLib's NodeSampler / Reconstruction
// (src/evaluationLib/src/metrics/NodeSampler.cpp:5-111, Reconstruction.cpp:6-23) on the WeightedGeometric similarity
// dist / (w_a w_b)^(1/d) (src/embeddingLib/src/embeddingSpace/WeightedGeometric.cpp:17-21), without sorting all n nodes:
// for a sampled vertex v with sorted neighbour keys S_0 < S_1 < .. (key = (similarity, id), the reference's tie order), every
// other node x bumps the counter of p = upper_bound(S, key_x); rank(S_j) = sum_{p <= j} cnt[p] is the number of nodes ranked
// before neighbour j, so precision at that neighbour = (j + 1) / (rank + 1).  One block per sampled vertex; all arithmetic
// in double on the fp32 positions; counters are integers, so the result does not depend on scheduling.
struct SimKey {
    double sim;
    int id;
};
__device__ __forceinline__ bool key_less(const SimKey& a, const SimKey& b) { return a.sim < b.sim || (a.sim == b.sim && a.id < b.id); }

template <int V>
__device__ __forceinline__ double similarity(const float4* __restrict__ x, const double* __restrict__ wroot, int a, const float4 (&xa)[V],
                                             double wra, int b) {
    double d2 = 0.0;
#pragma unroll
    for (int c = 0; c < V; ++c) {
        const float4 p = __ldg(x + (int64_t)b * V + c);
        double e;
        e = (double)p.x - (double)xa[c].x; d2 += e * e;
        e = (double)p.y - (double)xa[c].y; d2 += e * e;
        e = (double)p.z - (double)xa[c].z; d2 += e * e;
        e = (double)p.w - (double)xa[c].w; d2 += e * e;
    }
    (void)a;
    return sqrt(d2) / (wra * wroot[b]);
}

template <int V>
__global__ void __launch_bounds__(256) k_reconstruction(const float4* __restrict__ x, const double* __restrict__ wroot, const int* __restrict__ rowPtr,
                                                        const int* __restrict__ col, int n, const int* __restrict__ nodes, int first, int count,
                                                        int capacity, SimKey* __restrict__ keyScratch, int* __restrict__ cntScratch,
                                                        double* __restrict__ out /* [count][3]: precision@deg, AP, valid */) {
    const int s = first + blockIdx.x;
    if (s >= count) return;
    const int v = nodes[s];
    const int begin = rowPtr[v], deg = rowPtr[v + 1] - begin;
    double* o = out + (int64_t)s * 3;
    if (deg == 0) { if (threadIdx.x == 0) { o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; } return; }
    SimKey* keys = keyScratch + (int64_t)blockIdx.x * capacity;
    int* cnt = cntScratch + (int64_t)blockIdx.x * (capacity + 1);
    float4 xv[V];
    load_row<V>(x, v, xv);
    const double wrv = wroot[v];
    int pow2 = 1;
    while (pow2 < deg) pow2 <<= 1;
    for (int i = threadIdx.x; i < pow2; i += 256) {
        SimKey k;
        if (i < deg) { k.id = col[begin + i]; k.sim = similarity<V>(x, wroot, v, xv, wrv, k.id); }
        else { k.id = 0x7fffffff; k.sim = 1.0e300; }
        keys[i] = k;
    }
    for (int i = threadIdx.x; i <= deg; i += 256) cnt[i] = 0;
    __syncthreads();
    // bitonic sort of the neighbour keys (deg is ~10 for most vertices, up to 1e5 for hubs; scratch lives in L1/L2)
    for (int k = 2; k <= pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < pow2; i += 256) {
                const int partner = i ^ j;
                if (partner > i) {
                    const SimKey a = keys[i], b = keys[partner];
                    const bool up = (i & k) == 0;
                    if (key_less(b, a) == up) { keys[i] = b; keys[partner] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int xnode = threadIdx.x; xnode < n; xnode += 256) {
        if (xnode == v) continue;
        SimKey kx;
        kx.id = xnode;
        kx.sim = similarity<V>(x, wroot, v, xv, wrv, xnode);
        int lo = 0, hi = deg;                      // first neighbour key greater than kx
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (key_less(kx, keys[mid])) hi = mid; else lo = mid + 1;
        }
        atomicAdd(cnt + lo, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long before = 0;
        double ap = 0.0;
        int atDeg = 0;
        for (int j = 0; j < deg; ++j) {
            before += cnt[j];                      // nodes ranked before neighbour j (0-based rank)
            ap += (double)(j + 1) / (double)(before + 1);
            if (before < deg) ++atDeg;
        }
        o[0] = (double)atDeg / (double)deg;        // precisions[deg - 1] (NodeSampler.cpp:46)
        o[1] = ap / (double)deg;                   // getAveragePrecision (NodeSampler.cpp:95-111)
        o[2] = 1.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Test hook: the reference's candidate set (WeightedIndex.cpp:65-81), evaluated in double on the same walk.
// writePass 0 counts per query; writePass 1 writes the ids behind offsets[q] (slot order is arbitrary - an
// integer cursor - because the host sorts every query's ids before returning them).
template <int V>
__global__ void __launch_bounds__(256) k_candidates(const TreeView t, const float4* __restrict__ x, const double* __restrict__ w,
                                                    const double* __restrict__ classMax, int dim, double edgeLength,
                                                    float pruneL2, const float* __restrict__ iw, const int* __restrict__ queries,
                                                    int nq, int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                                    int* __restrict__ cursor, int* __restrict__ outIds, int writePass) {
    const int lane = threadIdx.x & 31, j = lane & (kFan - 1);
    const int qn = (blockIdx.x * blockDim.x + threadIdx.x) >> kFanLog2;
    const bool valid = qn < nq;
    const int v = valid ? queries[qn] : 0;
    float4 q[V];
    load_row<V>(x, v, q);
    const float iwq = __ldg(iw + v);
    const double wq = w[v];
    int found = 0, nTests = 0;
    const int64_t base = (valid && writePass) ? offsets[qn] : 0;
    walk_tree<V>(
        t, q, valid,
        [&](int, int, float d2, float bnd) {
            const float s = iwq * bnd;
            return d2 * s * s <= pruneL2;
        },
        [&](int idx, float, float, const float4 (&pu)[V]) {
            const int u = __ldg(t.ids + idx);
            double d2 = 0.0;
#pragma unroll
            for (int c = 0; c < V; ++c) {
                double e;
                e = (double)pu[c].x - (double)q[c].x; d2 += e * e;
                e = (double)pu[c].y - (double)q[c].y; d2 += e * e;
                e = (double)pu[c].z - (double)q[c].z; d2 += e * e;
                e = (double)pu[c].w - (double)q[c].w; d2 += e * e;
            }
            const double r = edgeLength * pow(wq * classMax[u], 1.0 / (double)dim);
            if (d2 <= r * r) {
                if (writePass) outIds[base + atomicAdd(cursor + qn, 1)] = u;
                ++found;
            }
        },
        nTests);
    found = group_sum<kFan>(found);
    if (valid && !writePass && j == 0) counts[qn] = found;
}

// ---------------------------------------------------------------------------------------------
// Edge detection quality: evaluationLib's EdgeDetection over the pairs an EdgeSampler drew
// (src/evaluationLib/src/metrics/EdgeDetection.cpp:6-66, EdgeSampler.cpp:7-66): similarity of every sampled pair, ascending
// sort, and the best F1 over all prefixes of the sorted list - prefix i classifies entries 0..i as edges.

// WeightedGeometric similarity of the sampled pairs (WeightedGeometric.cpp:17-21), in double on the fp32 positions
template <int V>
__global__ void __launch_bounds__(256) k_pair_similarity(const float4* __restrict__ x, const double* __restrict__ wroot, const int* __restrict__ pv,
                                                         const int* __restrict__ pw, int64_t count, double* __restrict__ sim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int a = pv[i], b = pw[i];
    float4 xa[V];
    load_row<V>(x, a, xa);
    sim[i] = similarity<V>(x, wroot, a, xa, wroot[a], b);
}

struct F1Best {        // best prefix so far; ties keep the lowest index (the reference updates on F1 > best only, :52-57)
    double f1, precision, recall;
    long long index;
};
__device__ __forceinline__ bool f1_better(const F1Best& a, const F1Best& b) { return a.f1 > b.f1 || (a.f1 == b.f1 && a.index < b.index); }

// edgePrefix[i] = number of edges among the sorted entries 0..i.  Closed form of the reference's running percentages
// (wrongEdgesPercent = 1 - e / numSampledEdges, wrongNonEdgesPercent = ne / numSampledNonEdges, :30-35), then its F1 (:39-45).
__global__ void __launch_bounds__(256) k_f1_curve(const int* __restrict__ edgePrefix, int64_t count, double numEdges, double numNonEdges,
                                                  double M, double noM, F1Best* __restrict__ partial) {
    __shared__ F1Best sm[256];
    F1Best best{-1.0, -1.0, -1.0, 0x7fffffffffffffffll};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const double e = (double)edgePrefix[i], ne = (double)(i + 1) - e;
        const double wrongEdges = 1.0 - (numEdges > 0.0 ? e / numEdges : 0.0);
        const double wrongNonEdges = numNonEdges > 0.0 ? ne / numNonEdges : 0.0;
        const double truePositives = (1.0 - wrongEdges) * M;
        const double retrieved = truePositives + wrongNonEdges * noM;
        const double precision = truePositives / retrieved, recall = truePositives / M;
        const F1Best cur{2.0 / (1.0 / precision + 1.0 / recall), precision, recall, (long long)i};
        if (f1_better(cur, best)) best = cur;
    }
    sm[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && f1_better(sm[threadIdx.x + o], sm[threadIdx.x])) sm[threadIdx.x] = sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256) k_f1_best(const F1Best* __restrict__ partial, int blocks, F1Best* __restrict__ out) {
    __shared__ F1Best sm[256];
    F1Best best{-1.0, -1.0, -1.0, 0x7fffffffffffffffll};
    for (int i = threadIdx.x; i < blocks; i += 256)
        if (f1_better(partial[i], best)) best = partial[i];
    sm[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && f1_better(sm[threadIdx.x + o], sm[threadIdx.x])) sm[threadIdx.x] = sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

}  // namespace wb
