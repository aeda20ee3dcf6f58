/*
 * wembed_b200.h - C ABI of the B200-native WEmbed gradient-descent step.
 *
 * This is the drop-in boundary of the project: plain pointers and sizes, no C++ or
 * torch types.  One handle owns one device-resident embedding problem (CSR graph,
 * positions, weights, Adam moments, spatial index) on one GPU and is driven by one
 * host thread.  The reference has no C ABI of its own for this path (its only FFI
 * precedent is the four sprk_* functions bound in
 * src/embeddingLib/src/spacialQuery/SprkQueries.cpp:22,27,59,64); each entry point
 * below names the reference C++ interface it replaces.  Paths are relative to the
 * reference checkout (Vraier/wembed).
 *
 * Conventions
 *   - every function returns 0 on success and a negative wb_status on failure;
 *     wb_last_error() returns a human readable message for the calling thread.
 *   - all host buffers are caller owned and only read / written during the call.
 *   - coordinates cross the boundary as row-major n x d doubles, exactly as
 *     EmbedderInterface::copyCoordinatesTo does (EmbedderInterface.hpp:94-96).
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef WEMBED_B200_H
#define WEMBED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_ABI_VERSION 2

typedef struct wb_embedder wb_embedder; /* opaque */

enum wb_status {
    WB_OK = 0,
    WB_ERR_INVALID = -1,  /* bad argument */
    WB_ERR_CUDA = -2,     /* CUDA runtime error, see wb_last_error */
    WB_ERR_NO_DEVICE = -3,
    WB_ERR_UNSUPPORTED = -4
};

enum wb_optimizer { WB_OPT_SIMPLE = 0, WB_OPT_ADAM = 1 };    /* EmbedderOptions.hpp:6  */

/*
 * The force / optimizer knobs of EmbedderOptions (EmbedderOptions.hpp:31-88) that the
 * device step reads.  Scheduling and stopping knobs stay on the host (they are scalar
 * logic layered over wb_step by the C++ facade, see wembed_b200/host).
 */
typedef struct wb_options {
    int32_t embedding_dimension;   /* EmbedderOptions::embeddingDimension, 1..32            */
    int32_t optimizer;             /* wb_optimizer; Adam: beta1=.9 beta2=.999 eps=1e-8      */
                                   /*   (WembedEmbedder.hpp:46)                             */
    int32_t reserved2;             /* must be 0.  Device state is fp32 (north_star: "within 1e-5 relative error in fp32");  */
                                   /*   sums are taken in fp64 / 64-bit fixed point.  The reference is fp64 throughout.      */
    int32_t device;                /* CUDA device ordinal                                   */
    int32_t keep_forces;           /* 1: keep state.force of every step for wb_get_forces   */
    int32_t reserved0;
    double attraction_scale;       /* EmbedderOptions::attractionScale                      */
    double repulsion_scale;        /* EmbedderOptions::repulsionScale                       */
    double centre_scale;           /* EmbedderOptions::centreScale (0 = off)                */
    double edge_length;            /* EmbedderOptions::edgeLength                           */
    double doubling_factor;        /* EmbedderOptions::doublingFactor (weight classes)      */
    double simple_max_displacement;/* EmbedderOptions::simpleOptMaxDisplacement             */
    uint32_t seed;                 /* base seed of the coincident-pair tie-break generator  */
                                   /*   (Rand::localGenerator, Rand.cpp:29-35)              */
    uint32_t reserved1[7];
} wb_options;

/*
 * Per-step observables, the values WembedEmbedder::calculateStep leaves in
 * EmbedderState (EmbedderState.hpp:30-35) plus the sums they are made of.
 */
typedef struct wb_step_stats {
    double loss_attract;       /* state.lastAttractLoss (WembedEmbedder.cpp:270-271) */
    double loss_repel;         /* state.lastRepelLoss   (WembedEmbedder.cpp:292-293) */
    double sum_displacement;   /* sum_v ||x_v - xprev_v||   (WembedEmbedder.cpp:330-343) */
    double sum_radius_sq;      /* sum_v ||x_v||^2, after recentring (:334-345)       */
    double rel_displacement;   /* state.lastRelDisplacement (:347-350)               */
    double num_repulsion_pairs;/* pairs that passed the neighbour filter and the     */
                               /*   exact weighted-distance test                      */
    double num_candidates;     /* point (exact distance) tests of the repulsion walk */
    double num_box_tests;      /* box tests of the repulsion walk                    */
    double centroid[32];       /* per-dimension mean removed by applyGravityCentre   */
    int64_t iteration;         /* state.currentIteration after the step              */
    /* state of the repulsion pair list (see wb_set_list_policy) */
    double num_listed_pairs;   /* unordered pairs in the list this step evaluated     */
    double list_rebuilt;       /* 1: this step rebuilt index + list, 0: it reused them */
    double list_skin;          /* relative inflation of the list radius               */
    double max_displacement_ratio; /* max_v ||dx_v|| / (smallest interaction radius of v) of this step */
} wb_step_stats;

/* -- life cycle ----------------------------------------------------------------- */

/* ABI / build information.  No device needed. */
int wb_abi_version(void);
const char* wb_build_info(void);
const char* wb_last_error(void);
/* Number of visible CUDA devices (0 when there is none).  No device needed. */
int wb_device_count(void);
/* Fills `o` with the defaults of EmbedderOptions (EmbedderOptions.hpp:31-88).  No device needed. */
void wb_options_default(wb_options* o);

/*
 * Replaces the WembedEmbedder constructor (WembedEmbedder.hpp:90-125) minus the random
 * initial layout: uploads the CSR (row_ptr[n+1], col[row_ptr[n]], neighbours sorted
 * ascending, no self loops, symmetric - the invariants of Graph, Graph.cpp:87-150),
 * allocates positions / weights / force / Adam moments.  Positions start at 0 and
 * weights at 1 until wb_set_coordinates / wb_set_weights are called.
 * Environment (read once per handle, for A/B runs and tests only - results never depend on it): WB_HALF_BOXES=0 / 1 forces the
 * fp32 / the half-precision box format of the repulsion walk; unset, the device picks per step from the layout's spread
 * (DESIGN.md section 3.1).
 */
int wb_create(wb_embedder** out, int32_t n, const int32_t* row_ptr, const int32_t* col, const wb_options* opts);
int wb_destroy(wb_embedder* h);

/* -- state ---------------------------------------------------------------------- */

/* WembedEmbedder::setCoordinates (WembedEmbedder.cpp:104-119): n x d row-major doubles. */
int wb_set_coordinates(wb_embedder* h, const double* coords);
/* WembedEmbedder::setWeights (WembedEmbedder.cpp:121-131): n doubles, all > 0;
 * recomputes invExpWeights[v] = 1 / w^(1/d) in double and the weight classes
 * (WeightedIndex::getDoublingWeightBuckets, WeightedIndex.cpp:51-63). */
int wb_set_weights(wb_embedder* h, const double* weights);
/* EmbedderInterface::copyCoordinatesTo (EmbedderInterface.hpp:94-96). */
int wb_get_coordinates(wb_embedder* h, double* coords);
/* WembedEmbedder::getWeights (WembedEmbedder.cpp:96-98). */
int wb_get_weights(wb_embedder* h, double* weights);
/* state.force of the most recent step (EmbedderState.hpp:26), n x d doubles.  Test hook. */
int wb_get_forces(wb_embedder* h, double* forces);
/* AdamOptimizer::reset (AdamOptimizer.cpp:32-36) and state.currentIteration = 0. */
int wb_reset_optimizer(wb_embedder* h);
int wb_set_iteration(wb_embedder* h, int64_t iteration);

/* -- the hot path ----------------------------------------------------------------- */

/*
 * One WembedEmbedder::calculateStep (WembedEmbedder.cpp:13-63) with the learning rate
 * the host-side LRScheduler produced for this iteration: index rebuild, attractive and
 * repulsive forces, optional centre force, optimizer update, recentring and the
 * displacement / loss sums.  Blocks until `stats` is valid.
 */
int wb_step(wb_embedder* h, double learning_rate, wb_step_stats* stats);

/*
 * Asynchronous variant: enqueues the step on the handle's stream and returns.
 * The k-th call's observables are fetched (in order) with wb_step_collect.
 * At most WB_MAX_INFLIGHT steps may be outstanding.
 */
#define WB_MAX_INFLIGHT 64
int wb_step_async(wb_embedder* h, double learning_rate);
int wb_step_collect(wb_embedder* h, wb_step_stats* stats);
int wb_synchronize(wb_embedder* h);

/*
 * The repulsion search does not apply forces: it lists every non-adjacent pair within L (1 + skin) / ws, and the step evaluates the
 * exact predicate of repellingForce (WembedEmbedder.cpp:196-201) on the listed pairs.  While no vertex has moved further than
 * skin / 2 of its smallest interaction radius since the search, the list is provably complete and index rebuild + search are
 * skipped; the device keeps that bound itself and picks the skin of each search from the current pace of the layout, between 0 (early
 * in a run: every step searches, nothing is inflated) and skin_max, aiming at lists that live for `reuse_steps` steps.  Results never
 * depend on the policy (a listed pair beyond the exact threshold contributes exactly zero); skin_max = 0 restores "search every step",
 * which is what WembedEmbedder::updateIndex does.  Defaults: 1.0 and 4 (environment, for A/B runs: WB_SKIN_MAX, WB_REUSE_STEPS).
 */
int wb_set_list_policy(wb_embedder* h, double skin_max, double reuse_steps);

/* -- test hooks for the spatial index ------------------------------------------------ */

/*
 * Rebuilds the device index from the current positions and evaluates, for each of the
 * `nq` query vertices, the reference's candidate set
 *   U_c { u in class c : ||x_u - x_v|| <= edgeLength * (w_v * maxW_c)^(1/d) }
 * (WeightedIndex::getNodesWithinWeightedDistance, WeightedIndex.cpp:65-81; includes v
 * itself and graph neighbours, as the reference's does).  Distances are evaluated in
 * double.  Results are written as a CSR: out_offsets[nq+1], out_ids[cap] sorted
 * ascending per query.  Returns WB_ERR_INVALID if cap is too small (out_offsets[nq]
 * then holds the required size).
 */
int wb_query_candidates(wb_embedder* h, int32_t nq, const int32_t* queries, int64_t* out_offsets,
                        int32_t* out_ids, int64_t cap);

/* Device time (ms, CUDA events) of each phase of the most recent wb_step:
 * [0] index rebuild  [1] forces + optimizer (k_hub_rows, k_step_fused, k_reduce_rows)  [2] repulsion search + pair list  [3] unused (0)
 * [4] recentre + observe + tail  [5] total.  [0] and [2] are ~0 on a step that reused its pair list.
 * Mirrors the util::Timer keys of WembedEmbedder.cpp:28-58. Requires wb_enable_timing(h,1), which also makes the step use direct
 * launches instead of the step graph. */
int wb_enable_timing(wb_embedder* h, int enable);
int wb_get_phase_times(wb_embedder* h, double* ms6);

/* -- multi-GPU: one graph sharded by vertex range over the GPUs of one node ----------------------- */

/*
 * One process per GPU creates the SAME problem (same CSR, weights, coordinates) on its device and then joins the group: rank 0 calls
 * wb_comm_unique_id and ships the 128 bytes to the other ranks by any means (bench.py uses torch.distributed), every rank calls
 * wb_comm_init (world <= 8, one node).  The ranks map each other's buffers through CUDA IPC (NCCL only carries the handles) and from
 * then on a step on rank r
 *   - searches the repulsion pairs of its blocks of the sorted order and stores every pair straight into the pair buffer of the
 *     rank(s) owning its two vertices (peer stores over NVLink);
 *   - runs the fused force + optimizer kernel for the vertices [r * rows, (r + 1) * rows) it owns (wb_get_partition) and stores each
 *     block's row of sums into every rank's copy of the sum rows;
 *   - recentres its rows and stores them into EVERY replica of the positions, together with its observation tiles;
 * with three flag barriers in between (k_exchange).  Positions stay replicated; every vertex is written by exactly one owner (the
 * multi-GPU form of the reference's OpenMP `parallel for` over vertices, WembedEmbedder.cpp:262,279).  All sums are taken over
 * global block rows / tiles in a fixed order, so the sharded step is bit-identical to the single-GPU step and all ranks return
 * identical statistics (except num_candidates / num_box_tests, which count the rank's own share of the search).
 * The reference has no distributed code.
 */
int wb_comm_unique_id(char* id128);
int wb_comm_init(wb_embedder* h, const char* id128, int32_t rank, int32_t world);
/* Test hook: the same sharded step with all `world` ranks as handles of the calling process on ONE device (plain device pointers instead of
 * IPC mappings, no NCCL).  Such handles only step together, through wb_step_group: it queues the pieces of the step for all handles in
 * lockstep on one stream, so stream order stands in for the barrier kernels (which then only publish).  stats: `world` records or NULL.
 * The pair buffers of a local group cannot grow (set WB_PAIR_CAP before wb_create). */
int wb_comm_init_local(wb_embedder** handles, int32_t world);
int wb_step_group(wb_embedder** handles, int32_t world, double learning_rate, wb_step_stats* stats);
/* [begin, end) of the vertices this handle owns (everything before wb_comm_init). */
int wb_get_partition(wb_embedder* h, int32_t* begin, int32_t* end);

/* -- evaluation ("next") ---------------------------------------------------------- */

/*
 * evaluationLib's Reconstruction metric (src/evaluationLib/src/metrics/Reconstruction.cpp:6-23, NodeSampler.cpp:5-111) of
 * the CURRENT layout on the WeightedGeometric similarity dist / (w_a w_b)^(1/d) (WeightedGeometric.cpp:17-21), for the given
 * sample of vertices: out2[0] = "constructDeg" (mean precision at k = degree), out2[1] = "MAP" (mean average precision).
 * Vertices without neighbours are skipped.  The caller chooses the sample (the reference draws <= 1000 random nodes).
 */
int wb_reconstruction(wb_embedder* h, int32_t count, const int32_t* nodes, double* out2);

/*
 * evaluationLib's EdgeDetection metric (src/evaluationLib/src/metrics/EdgeDetection.cpp:6-66) of the CURRENT layout on the
 * WeightedGeometric similarity, over the vertex pairs (v[i], w[i]) an EdgeSampler drew (EdgeSampler.cpp:7-66: every edge once,
 * plus non-edges at a chosen rate; is_edge[i] says which).  The pairs are sorted by similarity on the device (stable: ties keep
 * the caller's order) and every prefix of the sorted list is scored as "these are the edges"; out3 = {precision, recall, F1}
 * of the prefix with the best F1 (the first one on ties), with the sampled counts scaled to the graph's m edges and
 * n(n-1)/2 - m non-edges exactly as the reference does.  The caller chooses the sample (wembed_b200/metrics.py mirrors the
 * reference's sampler).
 */
int wb_edge_detection(wb_embedder* h, int64_t count, const int32_t* v, const int32_t* w, const uint8_t* is_edge, double* out3);

/* -- measurement ------------------------------------------------------------------------- */

/* Records CUDA event `slot` (0..7) on the handle's stream; wb_elapsed_ms waits for event `to` and returns the
 * device time between two recorded events.  This is how bench.py times steps on the launching stream. */
int wb_mark(wb_embedder* h, int slot);
int wb_elapsed_ms(wb_embedder* h, int from, int to, double* ms);
/* Number of kernels of this library launched through the handle since wb_create
 * (the CUB radix-sort call of the index rebuild is counted as one). */
int64_t wb_launch_count(wb_embedder* h);
/* How steps are issued: 1 = one CUDA graph launch per step (k_step_begin -> IF(rebuild){index, search, pair list} -> forces .. tail; one GPU,
 * phase timing off), 0 = direct launches (phase timing on, sharded runs, WB_GRAPH=0, or the capture failed: `note` then says why). */
int wb_exec_mode(wb_embedder* h, char* note, int32_t note_cap);

#ifdef __cplusplus
}
#endif
#endif /* WEMBED_B200_H */
