// Device kernels of the WEmbed gradient-descent step for sm_100a.
//
// Reference semantics (paths relative to the Vraier/wembed checkout):
//   index rebuild      WembedEmbedder::updateIndex            src/embeddingLib/src/embedder/WembedEmbedder.cpp:212-240   index.cuh
//   repulsion search   calculateAllRepellingForces / getRepellingCandidatesForNode              :274-294, :242-258       walk.cuh
//   forces + optimizer attractionForce, repellingForce, calculateAllCentreForces, AdamOptimizer::update / SimpleOptimizer::update
//                                                                                   :140-210, :296-301                   step.cuh
//   recentre+observe   applyGravityCentre / observeDisplacement                                 :303-352                 step.cuh
//   quality metrics    evaluationLib Reconstruction / EdgeDetection                                                      metrics.cuh
//
// All force kernels are pull style: one owner computes force[v]; there are no floating-point atomics and every reduction has a
// fixed shape, so a step is bit-reproducible from run to run (tests/TestDeterminism.cpp protocol).
// V = number of float4 chunks per position row.
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "index.cuh"
#include "walk.cuh"
#include "step.cuh"
#include "metrics.cuh"
