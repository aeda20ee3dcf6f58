// Exact re-statement of the generator behind the reference's coincident-pair tie-break:
//   std::mt19937 gen(std::seed_seq{seed, v, iteration});           (Rand::localGenerator, Rand.cpp:29-35)
//   d draws of std::normal_distribution<double>(0,1), a fresh distribution object per draw,
//   then normalisation to unit length                               (setToRandomUnitVector, DVec.hpp:412-424)
// as libstdc++ (GCC 13) implements them: seed_seq::generate [rand.util.seedseq], the MT19937
// recurrence and tempering, generate_canonical<double,53> (two 32-bit draws per double) and the
// Marsaglia polar method.  Host+device so tests can check it against <random> on the CPU.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define WB_HD __host__ __device__
#else
#define WB_HD
#endif

namespace wb {

struct Mt19937 {
    static constexpr int N = 624, M = 397;
    uint32_t* s;   // 624 words of caller-provided scratch (shared memory on the device)
    int p;

    WB_HD explicit Mt19937(uint32_t* scratch) : s(scratch), p(N) {}

    // std::seed_seq{a0,a1,a2}.generate over 624 words followed by mersenne_twister_engine::seed(seq)
    WB_HD void seed3(uint32_t a0, uint32_t a1, uint32_t a2) {
        const uint32_t in[3] = {a0, a1, a2};
        const uint32_t n = N, sz = 3, t = 11, pp = (n - t) / 2, q = pp + t, m = n;  // m = max(s+1, n)
        for (uint32_t i = 0; i < n; ++i) s[i] = 0x8b8b8b8bu;
        for (uint32_t k = 0; k < m; ++k) {
            const uint32_t km1 = k == 0 ? 15u : (k - 1) % n;  // (size_t(0)-1) % 624 == 15
            uint32_t arg = s[k % n] ^ s[(k + pp) % n] ^ s[km1];
            uint32_t r1 = 1664525u * (arg ^ (arg >> 27));
            uint32_t r2 = r1;
            if (k == 0) r2 += sz;
            else if (k <= sz) r2 += k % n + in[k - 1];
            else r2 += k % n;
            s[(k + pp) % n] += r1;
            s[(k + q) % n] += r2;
            s[k % n] = r2;
        }
        for (uint32_t k = m; k < m + n; ++k) {
            uint32_t arg = s[k % n] + s[(k + pp) % n] + s[(k - 1) % n];
            uint32_t r3 = 1566083941u * (arg ^ (arg >> 27));
            uint32_t r4 = r3 - k % n;
            s[(k + pp) % n] ^= r3;
            s[(k + q) % n] ^= r4;
            s[k % n] = r4;
        }
        bool zero = (s[0] & 0x80000000u) == 0u;
        for (int i = 1; zero && i < N; ++i) zero = s[i] == 0u;
        if (zero) s[0] = 0x80000000u;
        p = N;
    }

    WB_HD void twist() {
        const uint32_t upper = 0x80000000u, lower = 0x7fffffffu, a = 0x9908b0dfu;
        for (int k = 0; k < N - M; ++k) {
            uint32_t y = (s[k] & upper) | (s[k + 1] & lower);
            s[k] = s[k + M] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        }
        for (int k = N - M; k < N - 1; ++k) {
            uint32_t y = (s[k] & upper) | (s[k + 1] & lower);
            s[k] = s[k + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        }
        uint32_t y = (s[N - 1] & upper) | (s[0] & lower);
        s[N - 1] = s[M - 1] ^ (y >> 1) ^ ((y & 1u) ? a : 0u);
        p = 0;
    }

    WB_HD uint32_t next() {
        if (p >= N) twist();
        uint32_t z = s[p++];
        z ^= (z >> 11);
        z ^= (z << 7) & 0x9d2c5680u;
        z ^= (z << 15) & 0xefc60000u;
        z ^= (z >> 18);
        return z;
    }

    // std::generate_canonical<double, 53>(mt19937)
    WB_HD double canonical() {
        const double lo = static_cast<double>(next());
        const double hi = static_cast<double>(next());
        double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
        if (r >= 1.0) r = 0.99999999999999988897769753748434595763683319091796875;  // nextafter(1, 0)
        return r;
    }

    // one draw of a freshly constructed std::normal_distribution<double>(0, 1)
    WB_HD double normal() {
        double x, y, r2;
        do {
            x = 2.0 * canonical() - 1.0;
            y = 2.0 * canonical() - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = sqrt(-2 * log(r2) / r2);
        return y * mult;
    }
};

// out[0..d) = the unit vector the reference adds for a coincident pair of vertex v at `iteration`.
WB_HD inline void random_unit_vector(uint32_t* scratch624, uint32_t seed, uint32_t v, uint32_t iteration, int d, double* out) {
    Mt19937 gen(scratch624);
    gen.seed3(seed, v, iteration);
    double norm = 0.0;
    for (int k = 0; k < d; ++k) {
        out[k] = gen.normal();
        norm += out[k] * out[k];
    }
    norm = sqrt(norm);
    for (int k = 0; k < d; ++k) out[k] /= norm;
}

}  // namespace wb
