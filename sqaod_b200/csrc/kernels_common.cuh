/* kernels_common.cuh -- device helpers shared by the sm_100a kernels (internal). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sqb {

__device__ __forceinline__ int laneId() { return threadIdx.x & 31; }

template <class T> __device__ __forceinline__ T warpSum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* ---- shared-memory addresses, mbarrier and TMA bulk-copy wrappers (PTX ISA 8.x, sm_90+/sm_100a) ---- */
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarInitFence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbarArriveExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smemAddr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    while (!mbarTryWait(bar, parity)) {}
}
/* 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
 * dst/src 16-byte aligned, bytes a multiple of 16. */
__device__ __forceinline__ void tmaLoad1D(void *smemDst, const void *gmemSrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(smemDst)),
                 "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void tmaLoad1DHint(void *smemDst, const void *gmemSrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smemAddr(smemDst)),
        "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2PolicyEvictFirst() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2PolicyEvictLast() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

/* ---- explicit 32-bit shared-memory accesses (the accept chain keeps its table addresses in registers) ---- */
__device__ __forceinline__ uint32_t ldsU32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stsU32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void ldsReal(uint32_t a, float &v) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); }
__device__ __forceinline__ void ldsReal(uint32_t a, double &v) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); }
__device__ __forceinline__ void stsReal(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void stsReal(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
/* CTA-scope release / acquire on 32-bit counters in shared memory (hand-offs between the warps of the sweep kernel) */
__device__ __forceinline__ uint32_t ldAcquireCta(uint32_t a) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
/* plain (fence-free) poll of a shared-memory counter.  ld.acquire costs a fence, and a fence waits for everything the warp has
 * in flight -- on the accept chain that includes the asynchronous cross-term gathers of the last commit (HBM latency).  The
 * counters and the data they guard live in the shared memory of ONE SM, whose load/store unit executes a warp's accesses in
 * program order, and the producers publish with st.release / red.release, so the data read after a successful poll is current. */
__device__ __forceinline__ uint32_t ldVolatileCta(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stReleaseCta(uint32_t a, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void redAddReleaseCta(uint32_t a, uint32_t v) {
    asm volatile("red.release.cta.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
/* per-thread asynchronous 4- / 8-byte copies global -> shared (SASS: LDGSTS); waitAll covers every copy this thread issued */
__device__ __forceinline__ void cpAsyncReal(uint32_t dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncReal(uint32_t dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }
/* one instruction asks L2 for a whole contiguous block (TMA bulk prefetch, async proxy): addr 16-byte aligned, bytes a multiple of 16 */
__device__ __forceinline__ void prefetchL2Bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
/* named barrier over `count` threads (a multiple of 32) of the CTA; warps may arrive from different code paths */
__device__ __forceinline__ void namedBarSync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

/* ---- gpu-scope release/acquire on 64-bit flags in global memory (inter-CTA hand-off) ---- */
__device__ __forceinline__ void stRelease(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldAcquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxed(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldRelaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/* system scope: flags written / polled across GPUs over NVLink (ring sharding) */
__device__ __forceinline__ void stReleaseSys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldAcquireSys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelaxedSys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ldRelaxedSys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/* ---- packed spin rows ----
 * A trotter's N spins are kept as bits (1 = +1) in 64-bit words laid out for the sweep's dot product: the row is
 * cut into groups of 128 spins; inside a group, lane l of a warp owns spins 4l..4l+3 (one 128-bit load of J);
 * 16 consecutive groups (2048 spins) form a super-block whose 32 lanes x 64 bits are stored contiguously.
 *   spin j -> word64 = (j / 2048) * 32 + (j % 128) / 4,  bit = ((j / 128) % 16) * 4 + j % 4                      */
__host__ __device__ __forceinline__ int packedWords64(int N) { return ((N + 2047) / 2048) * 32; }
__host__ __device__ __forceinline__ void spinBitPos(int j, int &w64, int &bit) {
    int g = j >> 7;
    w64 = ((g >> 4) << 5) + ((j >> 2) & 31);
    bit = ((g & 15) << 2) + (j & 3);
}
__device__ __forceinline__ int spinAt(const unsigned long long *row, int j) { /* +1 / -1 */
    int w, b;
    spinBitPos(j, w, b);
    return ((row[w] >> b) & 1ull) ? 1 : -1;
}

/* flip the sign of v when bit 31 of m is set */
__device__ __forceinline__ float signFlip(float v, uint32_t m) { return __uint_as_float(__float_as_uint(v) ^ (m & 0x80000000u)); }
__device__ __forceinline__ double signFlip(double v, uint32_t m) {
    return __hiloint2double(__double2hiint(v) ^ (int)(m & 0x80000000u), __double2loint(v));
}

} // namespace sqb
