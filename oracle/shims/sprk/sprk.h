// Test-infrastructure shim (NOT product code, NOT reference code).
//
// `sprk` is an external, un-vendored Rust crate (github.com/wembed-pdf/sprk @ main,
// fetched by src/sprk/CMakeLists.txt:10-17); neither its source nor a Rust toolchain
// exist here.  These four entry points are the ones the reference binds
// (src/embeddingLib/src/spacialQuery/SprkQueries.cpp:22,27,59,64); the stand-ins abort,
// so the oracle must always be driven with IndexType::SNN.  Sprk parity is unpinned.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

struct SprkHandle;

inline SprkHandle* sprk_create(const float*, std::size_t, std::size_t) {
    std::fprintf(stderr, "oracle shim: sprk is unavailable; select IndexType::SNN\n");
    std::abort();
}
inline void sprk_destroy(SprkHandle*) {}
inline void sprk_query_radius(const SprkHandle*, const float*, double, std::uint64_t**, std::size_t*) { std::abort(); }
inline void sprk_free_results(std::uint64_t*, std::size_t) {}
