/* philox.cuh -- counter-based Philox4x32-10 for the sweeps (device side).
 * Replaces the reference's device MT19937 pool + conversion kernels (sqaodc/cuda/DeviceRandomMT19937.cpp:43-113,
 * DeviceRandomBuffer.cu:54-136, DeviceRandom.cuh:8-20): every (step, round, trotter) draws its flip position and its
 * uniform from a pure function of the seed, so no random numbers ever touch HBM.  Stream layout and the
 * integer->real conversions are the ones restated in oracle/philox_ref.h (kept in lock-step by
 * tests/test_dense_annealer_gpu.py::test_exact_chain_*). */
#pragma once
#include <stdint.h>

namespace sqb {

enum { DOM_DENSE_SWEEP = 0, DOM_RANDOMIZE = 1, DOM_BG_SIDE0 = 2, DOM_BG_SIDE1 = 3, DOM_RANDOMIZE1 = 4, DOM_PROBLEM = 5 };

struct Philox4 { uint32_t w[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

__host__ __device__ __forceinline__ Philox4 sqbPhilox(uint64_t seed, uint64_t step, uint32_t domain, uint32_t idx, uint32_t y) {
    return philox4x32_10(idx, y, (uint32_t)step, (domain << 24) | (uint32_t)((step >> 32) & 0xffffffu),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

template <class real> __device__ __forceinline__ real philoxUniform(const Philox4 &p);
template <> __device__ __forceinline__ float philoxUniform<float>(const Philox4 &p) {
    return __uint2float_rn(p.w[1]) * 2.3283064365386963e-10f; /* u32 * 2^-32, as Random.cpp:157-161 */
}
template <> __device__ __forceinline__ double philoxUniform<double>(const Philox4 &p) {
    uint32_t a = p.w[1] >> 5, b = p.w[2] >> 6; /* 53 bits from two words, as Random.cpp:165-168 */
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

} // namespace sqb
