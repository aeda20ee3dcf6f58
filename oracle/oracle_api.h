/*
 * oracle_api.h - TEST INFRASTRUCTURE.  C API shared by the two CPU checkers:
 *
 *   ref_*  : the reference's own C++ (compiled in place from /root/reference by
 *            oracle/Makefile into oracle/_ref/libwembed_ref.so, third-party deps shimmed)
 *   port_* : oracle/wembed_port.cpp, an independent CPU restatement of the same step
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load these libraries.  The product (wembed_b200/) never does.
 *
 * Both expose the same functions (X = ref | port) so tests can drive them identically.
 */
#ifndef WEMBED_ORACLE_API_H
#define WEMBED_ORACLE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors EmbedderOptions (src/embeddingLib/include/embedder/EmbedderOptions.hpp:31-88). */
typedef struct orc_options {
    int32_t embeddingDimension;   /* 4 */
    int32_t weightType;           /* 0 Unit, 1 Degree (default) */
    int32_t optimizerType;        /* 0 Simple, 1 Adam (default) */
    int32_t maxIterations;        /* 10000 */
    int32_t lrScheduleType;       /* 0 ExponentialCooling (default), 1 LossAdaptive */
    int32_t warmupSteps;          /* 20 */
    int32_t lrAdaptPatience;      /* 20 */
    int32_t stopCriterion;        /* 0 Displacement, 1 Loss (default) */
    int32_t stopDisplacementPatience; /* 5 */
    int32_t lossRateWindow;       /* 30 */
    int32_t stopLossPatience;     /* 50 */
    int32_t numThreads;           /* OpenMP threads, <=0: leave default */
    double dimensionHint;         /* -1 */
    double attractionScale;       /* 1 */
    double repulsionScale;        /* 1 */
    double centreScale;           /* 0 */
    double edgeLength;            /* 1 */
    double doublingFactor;        /* 2 */
    double simpleOptMaxDisplacement; /* 1 */
    double learningRate;          /* 10 */
    double lrCoolingFactor;       /* 0.995 */
    double lrDecayFactor;         /* 0.5 */
    double lrDecayThreshold;      /* 1e-2 */
    double lrGrowthFactor;        /* 1.0 */
    double lrGrowthThreshold;     /* 1e-1 */
    double stopDisplacementTol;   /* 3e-4 */
    double lossSmoothingFactor;   /* 0.3 */
    double stopLossTol;           /* 1e-3 */
} orc_options;

/* stats[8] layout of X_get_stats */
enum { ORC_LOSS_ATTRACT = 0, ORC_LOSS_REPEL = 1, ORC_LR = 2, ORC_REL_DISP = 3, ORC_REL_LOSS_IMPROVEMENT = 4,
       ORC_ITERATION = 5, ORC_NUM_REP_PAIRS = 6, ORC_RESERVED = 7 };

#define ORC_DECLARE(X)                                                                                          \
    void X##_options_default(orc_options* o);                                                                   \
    /* edges: m (src,dst) pairs, undirected, each once or twice; seed: Rand::setSeed; init_state: run the     \
       constructor's random layout + degree/unit weights (1) or leave zeros for set_* (0) */                    \
    void* X##_create(int32_t n_hint, int64_t m, const int32_t* src, const int32_t* dst, const orc_options* o,   \
                     int32_t seed, int32_t init_state);                                                         \
    void X##_destroy(void* h);                                                                                  \
    int32_t X##_num_vertices(void* h);                                                                          \
    int64_t X##_num_directed_edges(void* h);                                                                    \
    void X##_csr(void* h, int32_t* row_ptr, int32_t* col);                                                      \
    int32_t X##_are_neighbors(void* h, int32_t v, int32_t u);                                                   \
    void X##_set_coordinates(void* h, const double* coords);                                                    \
    void X##_set_weights(void* h, const double* w);                                                             \
    void X##_get_coordinates(void* h, double* coords);                                                          \
    void X##_get_weights(void* h, double* w);                                                                   \
    void X##_get_forces(void* h, double* f);                                                                    \
    void X##_step(void* h);                                                                                     \
    int32_t X##_is_finished(void* h);                                                                           \
    int64_t X##_run(void* h); /* calculateEmbedding; returns iterations done */                                 \
    void X##_get_stats(void* h, double* stats8);                                                                \
    /* candidate set of v for the CURRENT coordinates (rebuilds the index); returns count, writes <= cap */     \
    int64_t X##_candidates(void* h, int32_t v, int32_t* out, int64_t cap);

ORC_DECLARE(ref)
ORC_DECLARE(port)

#ifdef __cplusplus
}
#endif
#endif
