// Repulsion search (WembedEmbedder.cpp:274-294, 242-258; WeightedIndex.cpp:65-81): finds every unordered pair of non-adjacent vertices
// within the list radius and appends it to the pair buffer.  Forces are NOT applied here: the fused step kernel evaluates the exact
// predicate of repellingForce on the listed pairs (step.cuh), see StepCtrl in params.cuh.
#pragma once
#include <cuda_fp16.h>

#include "index.cuh"

namespace wb {

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
// Pair-stack walk, every unordered pair found once.
//
// A warp owns 32 consecutive queries and one LIFO stack of (block, query) pairs in shared memory.  Every box round pops eight pairs,
// two per 8-lane group; lane c of the group tests child c of the pair's block against the pair's query (query coordinates come from
// shared memory, the child box from L1/L2 through loads off ONE address register).  Passing children are pushed child-major (pairs
// popped together tend to name the same block): boxes of level >= 2 back onto the stack, leaves into a leaf queue.  A point round
// pops eight (leaf, query) pairs and lane c tests point c of the leaf; it runs whenever 16 leaves are waiting or the stack is empty.
//
// The query at sorted position p only searches positions > p - subtrees that end at or before p are cut by an integer comparison -
// so every unordered pair is found by exactly one lane, which appends {v, u} to the pair buffer (one warp-aggregated integer atomic
// for the slot).  The list radius is L (1 + skin) / ws with a relative slack of 1e-5 on top: the exact predicate is evaluated later,
// on the listed pairs, by the kernel that applies the forces.
//
// The kernel is persistent: every warp fetches the next chunk of queries from an integer counter until none is left.
// The queries are dealt out by SORTED position: the sorted order is cut into blocks of kRepBlockChunks chunks (a chunk = 32
// consecutive positions = one warp's queries) and the blocks are dealt round-robin to the ranks of a sharded run: whole blocks,
// because warps that run at the same time should work on neighbouring chunks (they share tree nodes in L1 / L2), round-robin
// because the walk cost varies across space.  With world = 1 this is the identity.
constexpr int kRepBlockChunks = 32;
struct RepLayout {
    int world, rank, segRows;
    // l-th query this rank walks (l < segRows) -> sorted position
    __host__ __device__ __forceinline__ int position(int l) const {
        constexpr int blockRows = kRepBlockChunks * 32;
        return ((l / blockRows) * world + rank) * blockRows + l % blockRows;
    }
};

// Where found pairs go.  One GPU: one buffer.  A sharded run (wb_comm_init) delivers every pair straight into the inbox of the
// rank(s) owning its two vertices, over NVLink peer mappings: segment `rank` of the owner's inbox, slot from a LOCAL counter (one per
// destination), so no remote atomics are needed; the counts travel after the kernel (k_publish_counts).
constexpr int kMaxRanks = 8;
struct PairSink {
    int2* seg[kMaxRanks];             // seg[dest]: where this rank writes pairs for `dest` (its own buffer for dest == rank)
    unsigned int* count;              // [world] local counters, reset by k_step_begin
    unsigned int cap;                 // capacity of every segment
    int world, rowsPerRank;           // owner(v) = v / rowsPerRank
};

// One slot of segment `dest` per calling lane.  SHARDED = false: every caller has the same destination and the lanes that are here
// together share one atomic.  SHARDED = true: the lanes are first grouped by destination (match.any), one atomic per group.
// (The first version looped over the destinations with the aggregation inside, "lanes with different destinations take turns".  It
// lost 3 pairs in 10 000 at world = 8 and only there - on 8 GPUs and in the single-process group of tests/test_gpu_sharded_local.py
// alike: slots were reserved and never written.  nvcc unrolls that loop by four, so only world = 8 runs the unrolled body twice;
// the same loop under `#pragma unroll 1`, per-lane atomics, and this version all give the single-GPU pair set.)
template <bool SHARDED>
__device__ __forceinline__ void append_pair(const PairSink& sink, int dest, int v, int u) {
    unsigned m = __activemask();
    if constexpr (SHARDED) m = __match_any_sync(m, dest);
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(sink.count + dest, (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < sink.cap) sink.seg[dest][slot] = make_int2(v, u);   // beyond the capacity: counted only, k_rep_count raises `overflow`
}
__device__ __forceinline__ void emit_pair(const PairSink& sink, int v, int u) {
    if (sink.world == 1) { append_pair<false>(sink, 0, v, u); return; }
    const int ov = v / sink.rowsPerRank, ou = u / sink.rowsPerRank;
    append_pair<true>(sink, ov, v, u);
    if (ou != ov) append_pair<true>(sink, ou, v, u);
}

// warps per block of k_repulse_pairs: the per-warp shared memory (queries + stack) grows with V
__host__ __device__ constexpr int repulse_warps(int V) { return V <= 4 ? 8 : 4; }

// dynamic shared memory of k_repulse_pairs<V, HALF>
__host__ __device__ constexpr int repulse_smem_bytes(int V, bool half) {
    return repulse_warps(V) * (32 * (V + 1 + (half ? half_chunks(V) : 0)) * 16 + (8 + 56 * kMaxLevels + 72 + 80) * 4);
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
// if that threshold is beyond the half range).  Only the number of boxes that pass changes (+1..2 % at c3), never the pair set:
// points are still tested in fp32.  Both instantiations are launched; QuantParams::halfBoxes (decided on the device from the
// layout's spread) says which one runs, the other returns at once.
template <int V, bool HALF>
__global__ void __launch_bounds__(256, (V <= 2 ? WB_REPULSE_MINBLOCKS : (V <= 4 ? 2 : 1)))
k_repulse_pairs(const TreeView t, const int* __restrict__ rowPtr, const int* __restrict__ col, int n, const PairSink sink,
                const RepLayout lay, int queriesPerUnit, const int* __restrict__ heavySlot, int* __restrict__ chunkCounter,
                double* __restrict__ partials, const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    if ((t.quant->halfBoxes != 0) != HALF) return;
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
    static_assert(repulse_smem_bytes(V, HALF) == WARPS * (32 * QROW * 16 + (8 + STACK + 80) * 4), "host and kernel disagree on the layout");
    float4 (*sQ)[32][QROW] = reinterpret_cast<float4 (*)[32][QROW]>(smemRep);
    uint32_t (*sStack)[8 + STACK] = reinterpret_cast<uint32_t (*)[8 + STACK]>(smemRep + WARPS * 32 * QROW);
    uint32_t (*sLeaf)[80] = reinterpret_cast<uint32_t (*)[80]>(reinterpret_cast<uint32_t*>(smemRep + WARPS * 32 * QROW) + WARPS * (8 + STACK));
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
    const float listL2 = ctrl->listL2, pruneL = ctrl->pruneL;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t rootBlock = (uint32_t)t.blockOff[t.numLevels];
    // Work unit of a warp = queriesPerUnit (8, 16 or 32) consecutive rows of this rank's share of the sorted order.  Small units
    // keep the dynamic schedule balanced when a rank (or a small graph) has few queries per resident warp.
    const int numChunks = lay.segRows / queriesPerUnit;
    int nEmitted = 0, nTests = 0, boxSlots = 0;

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
            r.pass = (d2 * s * s <= listL2) && __float_as_uint(meta.z) > __float_as_uint(qm.y);
            return r;
        }
    };
    // Point round: one (leaf, query) pair per 8-lane group and slot; lane c tests point c of the leaf against the list radius.
    struct PointSlot { int idx; uint32_t qq; bool hit; };
    auto testPoint = [&](uint32_t entry, bool active) {
        PointSlot r;
        r.qq = entry >> 27;
        r.idx = active ? (int)(entry & kRefMask) * kFan + c : 0;
        const float4* qrow = myQ + r.qq * QROW;
        float4 pu[V], qv[V];
#pragma unroll
        for (int k = 0; k < V; ++k) pu[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + r.idx);
        const float iwu = __ldg(t.bound[0] + r.idx);
#pragma unroll
        for (int k = 0; k < V; ++k) qv[k] = qrow[k];
        const float4 qm = qrow[V];
        const float ws = qm.x * iwu;
        const float d2 = point_dist2<V>(qv, pu);
        r.hit = active && (uint32_t)r.idx >= __float_as_uint(qm.y) && d2 * ws * ws <= listL2;   // idx > qpos
        return r;
    };
    // A hit is resolved by the lane that found it: neighbour filter (Graph::areNeighbors, WembedEmbedder.cpp:284), then the pair
    // joins the list (hits are rare - a handful per query - so this branch is cold).
    auto resolveHit = [&](const PointSlot& r) {
        if (!r.hit) return;
        const float4* qrow = myQ + r.qq * QROW;
        const int u = __ldg(t.ids + r.idx);
        const int v = __ldg(t.ids + (__float_as_uint(qrow[V].y) - 1u));   // the query's vertex (hits are rare: looked up here, not carried)
        // pairs with a heavy vertex belong to that vertex' block (k_repulse_heavy)
        if ((heavySlot && __ldg(heavySlot + u) >= 0) || is_neighbor(col, __ldg(rowPtr + v), __ldg(rowPtr + v + 1), u)) return;
        emit_pair(sink, v, u);
        ++nEmitted;
    };

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
                resolveHit(a);
                resolveHit(b);
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
    }
    // per-warp statistics (integers, so the order in which warps took chunks cannot change the reduced value);
    // every box slot is 8 lane tests and all 32 lanes counted it: 8 / 32 per lane
    double totalPairs = (double)nEmitted, totalTests = (double)nTests, totalBoxTests = 0.25 * (double)boxSlots;
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
    // sharded run: this thread's pairs went into other GPUs' memory; they must have ARRIVED there before this kernel counts as done,
    // because the "done" flag travels separately (k_exchange) and must not overtake them in the NVLink fabric
    if (sink.world > 1) __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// Search for heavy vertices (weight >= kHeavyWeight x the mean: hubs of heavy-tailed graphs).  The interaction radius grows
// like w^(1/d) and the number of partners like w, so a hub of weight 6000 (c4) has tens of thousands of in-radius partners and
// its ball covers most of the layout.  One block per heavy vertex scans the level-2 boxes with a fixed thread <-> box
// assignment, descends into passing leaves and points, and every thread appends its own hits.  Every pair that involves a heavy
// vertex is found here and only here (two heavy vertices: by the one at the lower sorted position).  Same predicates as the
// pair-stack walk, so the same pair set.
constexpr float kHeavyWeight = 32.0f;

template <int V>
__global__ void __launch_bounds__(256) k_repulse_heavy(const TreeView t, const int* __restrict__ rowPtr, const int* __restrict__ col, int n,
                                                       const PairSink sink, const RepLayout lay, const int* __restrict__ heavyVertex,
                                                       const int* __restrict__ heavySlot, const int* __restrict__ heavyPos,
                                                       double* __restrict__ partials /* [block][3] */, const StepCtrl* __restrict__ ctrl) {
    if (build_skipped(ctrl, 0)) return;
    __shared__ double redBuf[8 * 3];
    const int v = heavyVertex[blockIdx.x];
    const int p = heavyPos[blockIdx.x];
    const float listL2 = ctrl->listL2;
    double vals[3] = {0.0, 0.0, 0.0};      // emitted pairs, point tests, box tests
    // in a sharded run every rank launches all heavy vertices and keeps those whose sorted position falls in its blocks
    const bool mine = ((p >> 5) / kRepBlockChunks) % lay.world == lay.rank;
    if (mine) {
        float4 q[V];
#pragma unroll
        for (int k = 0; k < V; ++k) q[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + p);
        const float iwq = __ldg(t.bound[0] + p);
        const int rowBegin = __ldg(rowPtr + v), rowEnd = __ldg(rowPtr + v + 1);
        const int top = t.numLevels >= 2 ? 2 : 1;            // level the flat scan starts from
        auto box = [&](int lv, int idx, float& bnd) {
            float4 lo[V], hi[V];
            const int64_t st = t.stride[lv];
#pragma unroll
            for (int k = 0; k < V; ++k) { lo[k] = __ldg(t.lo[lv] + k * st + idx); hi[k] = __ldg(t.hi[lv] + k * st + idx); }
            bnd = __ldg(t.bound[lv] + idx);
            return box_dist2<V>(q, lo, hi);
        };
        auto passes = [&](float d2, float bnd) { const float s = iwq * bnd; return d2 * s * s <= listL2; };
        auto leaf = [&](int leafIdx) {
            for (int j = 0; j < kFan; ++j) {
                const int idx = leafIdx * kFan + j;
                if (idx >= n) break;
                vals[1] += 1.0;
                float4 pu[V];
#pragma unroll
                for (int k = 0; k < V; ++k) pu[k] = __ldg(t.lo[0] + (int64_t)k * t.stride[0] + idx);
                const float iwu = __ldg(t.bound[0] + idx);
                const float d2 = point_dist2<V>(q, pu);
                if (!passes(d2, iwu) || idx == p) continue;
                const int u = __ldg(t.ids + idx);
                if (idx < p && __ldg(heavySlot + u) >= 0) continue;      // two heavy vertices: the lower position owns the pair
                if (is_neighbor(col, rowBegin, rowEnd, u)) continue;
                emit_pair(sink, v, u);
                vals[0] += 1.0;
            }
        };
        for (int node = threadIdx.x; node < t.count[top]; node += 256) {
            float bnd;
            vals[2] += 1.0;
            if (!passes(box(top, node, bnd), bnd)) continue;
            if (top == 1) { leaf(node); continue; }
            for (int c = 0; c < kFan; ++c) {
                const int lf = node * kFan + c;
                if (lf >= t.count[1]) break;
                vals[2] += 1.0;
                if (passes(box(1, lf, bnd), bnd)) leaf(lf);
            }
        }
    }
    block_sum<3, 256>(vals, redBuf, partials + (int64_t)blockIdx.x * 3);
    if (sink.world > 1) __threadfence_system();               // see k_repulse_pairs
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


}  // namespace wb
