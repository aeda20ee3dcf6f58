// Quality metrics of evaluationLib on the device: Reconstruction and EdgeDetection.
#pragma once
#include "common.cuh"

namespace wb {

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
