// Shared definitions of the device code: layout constants, error plumbing, small vector helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace wb {

// ---------------------------------------------------------------------------------------------
// Layout constants
//
// Positions are stored row-major with the row padded to a multiple of 4 floats (V float4 chunks,
// padding = 0) so every row is read with 128-bit loads; d <= 32 -> V in 1..8.
constexpr int kMaxChunks = 8;                 // V <= 8  <=> embeddingDimension <= 32
constexpr int kMaxDim = 4 * kMaxChunks;

// Spatial index: Morton-sorted points, grouped 8 by 8 into a complete implicit 8-ary hierarchy of
// axis-aligned boxes.  Level 0 = the sorted points themselves (degenerate boxes), level 1 = boxes of
// kFan consecutive points, level l = boxes of kFan level-(l-1) boxes; the top level has <= kFan nodes.
constexpr int kFan = 8;
constexpr int kFanLog2 = 3;
constexpr int kMaxLevels = 12;                // 8^11 > 2^31

constexpr float kPadCoord = 1.0e18f;          // coordinates of padding points: far from everything, square is finite
constexpr float kPruneSlack = 1.0e-5f;        // relative slack of every conservative (pruning) comparison

// number of scalar sums a force / observe kernel hands to the deterministic reducer
constexpr int kMaxSums = kMaxDim * 4 + 8;

// ---------------------------------------------------------------------------------------------
struct CudaError {
    cudaError_t code;
    const char* what;
    const char* file;
    int line;
};

#define WB_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t wb_err_ = (expr);                                        \
        if (wb_err_ != cudaSuccess) throw ::wb::CudaError{wb_err_, #expr, __FILE__, __LINE__}; \
    } while (0)

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sq(float a) { return a * a; }

// Single-instruction special functions (MUFU, relative error <= 2^-22): used where the result feeds fp32 terms whose own rounding
// is of the same size; IEEE sqrtf / division cost ~12 instructions and a branch each.
// The .ftz forms are one MUFU without the subnormal pre-scaling; callers must keep subnormal arguments away from them.
__device__ __forceinline__ float rsqrt_approx(float a) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float sqrt_approx(float a) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float rcp_approx(float a) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
constexpr float kFltMin = 1.17549435e-38f;

// squared distance between a point q (V chunks in registers) and the box [lo, hi] (lo == hi for a point)
template <int V>
__device__ __forceinline__ float box_dist2(const float4 (&q)[V], const float4 (&lo)[V], const float4 (&hi)[V]) {
    float d2 = 0.f;
#pragma unroll
    for (int c = 0; c < V; ++c) {
        float e;
        e = fmaxf(fmaxf(lo[c].x - q[c].x, q[c].x - hi[c].x), 0.f); d2 = fmaf(e, e, d2);
        e = fmaxf(fmaxf(lo[c].y - q[c].y, q[c].y - hi[c].y), 0.f); d2 = fmaf(e, e, d2);
        e = fmaxf(fmaxf(lo[c].z - q[c].z, q[c].z - hi[c].z), 0.f); d2 = fmaf(e, e, d2);
        e = fmaxf(fmaxf(lo[c].w - q[c].w, q[c].w - hi[c].w), 0.f); d2 = fmaf(e, e, d2);
    }
    return d2;
}

// squared Euclidean distance, dimensions summed in ascending order (calculateLPNorm, VectorOperations.hpp:5-11)
template <int V>
__device__ __forceinline__ float point_dist2(const float4 (&a)[V], const float4 (&b)[V]) {
    float d2 = 0.f;
#pragma unroll
    for (int c = 0; c < V; ++c) {
        float e;
        e = a[c].x - b[c].x; d2 = fmaf(e, e, d2);
        e = a[c].y - b[c].y; d2 = fmaf(e, e, d2);
        e = a[c].z - b[c].z; d2 = fmaf(e, e, d2);
        e = a[c].w - b[c].w; d2 = fmaf(e, e, d2);
    }
    return d2;
}

template <int V>
__device__ __forceinline__ void load_row(const float4* __restrict__ base, int64_t row, float4 (&out)[V]) {
#pragma unroll
    for (int c = 0; c < V; ++c) out[c] = __ldg(base + row * V + c);
}

// acc += s * (a - b)
template <int V>
__device__ __forceinline__ void axpy_diff(float4 (&acc)[V], float s, const float4 (&a)[V], const float4 (&b)[V]) {
#pragma unroll
    for (int c = 0; c < V; ++c) {
        acc[c].x = fmaf(s, a[c].x - b[c].x, acc[c].x);
        acc[c].y = fmaf(s, a[c].y - b[c].y, acc[c].y);
        acc[c].z = fmaf(s, a[c].z - b[c].z, acc[c].z);
        acc[c].w = fmaf(s, a[c].w - b[c].w, acc[c].w);
    }
}

// acc (double) += s * (a - b), every term evaluated in fp32 like the rest of the pair arithmetic
template <int V>
__device__ __forceinline__ void axpy_diff_d(double (&acc)[4 * V], float s, const float4 (&a)[V], const float4 (&b)[V]) {
#pragma unroll
    for (int c = 0; c < V; ++c) {
        acc[4 * c + 0] += (double)(s * (a[c].x - b[c].x));
        acc[4 * c + 1] += (double)(s * (a[c].y - b[c].y));
        acc[4 * c + 2] += (double)(s * (a[c].z - b[c].z));
        acc[4 * c + 3] += (double)(s * (a[c].w - b[c].w));
    }
}

__device__ __forceinline__ float4 shfl_xor4(float4 v, int laneMask) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, laneMask);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, laneMask);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, laneMask);
    v.w = __shfl_xor_sync(0xffffffffu, v.w, laneMask);
    return v;
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 min4(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

// Fixed-order sum over the G (power of two) lanes of a lane group: butterfly, so every lane ends with
// the same value and the order of additions does not depend on scheduling.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ int group_sum(int v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float4 group_sum(float4 v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = add4(v, shfl_xor4(v, o));
    return v;
}

// Deterministic block-wide sum of K doubles per thread: fixed butterfly inside each warp, then the
// warp results are added in warp order by the first K threads.  Result valid in thread t < K as out.
template <int K, int THREADS>
__device__ __forceinline__ void block_sum(double (&val)[K], double* smem /* [THREADS/32][K] */, double* blockOut /* global [K] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = val[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) smem[warp * K + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) s += smem[w * K + threadIdx.x];
        blockOut[threadIdx.x] = s;
    }
    __syncthreads();
}

}  // namespace wb
