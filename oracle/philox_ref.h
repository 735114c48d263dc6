/* oracle/philox_ref.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1,2,3",
 * SC'11; Random123 v1.09 `philox4x32_R(10, ctr, key)`).  The reference (shinmorino/sqaod) has no
 * counter-based generator: its CUDA path fills a pool with cuRAND MT19937
 * (sqaodc/cuda/DeviceRandomMT19937.cpp:43-113) and its CPU path uses a host MT19937
 * (sqaodc/common/Random.cpp:70-168).  BASELINE.json's north_star replaces the device pool with
 * per-(step, round, trotter) Philox, so this header pins that generator on the CPU for the
 * exact-chain parity tests.  Known-answer vectors: tests/test_oracle_rng.py.
 */
#pragma once
#include <stdint.h>

static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared by the oracle's "philox" mode and the CUDA kernels
 * (sqaod_b200/csrc/philox.cuh): key = 64-bit seed; counter =
 *   c0 = index within the step (round r for the dense sweep, spin index otherwise)
 *   c1 = trotter y
 *   c2 = low 32 bits of the step counter
 *   c3 = (domain << 24) | (step counter bits 32..55)
 */
enum { SQB_DOM_DENSE_SWEEP = 0, SQB_DOM_RANDOMIZE = 1, SQB_DOM_BG_SIDE0 = 2, SQB_DOM_BG_SIDE1 = 3,
       SQB_DOM_RANDOMIZE1 = 4 };

static inline void sqb_philox(uint64_t seed, uint64_t step, uint32_t domain, uint32_t idx, uint32_t y,
                              uint32_t out[4]) {
    uint32_t ctr[4] = { idx, y, (uint32_t)step, (domain << 24) | (uint32_t)((step >> 32) & 0xffffffu) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    philox4x32_10(ctr, key, out);
}

/* uniform in [0,1): same integer->real conversions as the reference's host generator
 * (sqaodc/common/Random.cpp:157-168) applied to Philox words. */
static inline float sqb_u01_f32(const uint32_t w[4]) { return (float)w[1] * (float)(1. / 4294967296.); }
static inline double sqb_u01_f64(const uint32_t w[4]) {
    uint32_t a = w[1] >> 5, b = w[2] >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
