/* l2_split.cu -- does the 126 MB L2 of a B200 hold more of J when each die reads only half of J's columns?
 *
 * The dense sweep reads random 32 KB rows of a 256 MiB matrix from every SM; ncu shows 68 % of those bytes coming from
 * HBM, i.e. L2 behaves like ~84 MB.  This probe streams the same number of bytes with the sweep's access machinery (TMA
 * bulk copies into per-warp shared-memory rings, 14 warps per SM) in three ways:
 *   mode 0: every CTA reads full random rows                                   (the sweep's pattern)
 *   mode 1: CTAs of die 0 read only columns [0, N/2), CTAs of die 1 only [N/2, N)   (dies found by a latency probe)
 *   mode 2: the same split, but by CTA parity (both dies read both halves)      (control)
 *   mode 3: full rows with an L2 cache-hint policy: a fraction of the lines evict_last, the rest normal or evict_first
 * Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:stream` and compare the three launches.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_split l2_split.cu */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

/* per-SM latency of dependent loads from `nAddr` lines spread over the buffer: out[smid][k] in cycles */
__global__ void latencyProbe(const unsigned long long *buf, size_t strideWords, int nAddr, unsigned *out, unsigned *smids) {
    if (threadIdx.x != 0) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smids[blockIdx.x] = smid;
    for (int k = 0; k < nAddr; ++k) {
        const unsigned long long *p = buf + (size_t)k * strideWords;
        unsigned long long v;
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); /* warm L2 */
        unsigned best = 0xffffffffu;
        unsigned long long off = v; /* the buffer is zero filled: a chain of dependent loads of the same line */
        for (int rep = 0; rep < 8; ++rep) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) : "l"(off) : "memory");
#pragma unroll
            for (int c = 0; c < 8; ++c) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(off) : "l"(p + off) : "memory");
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "l"(off) : "memory");
            best = min(best, (unsigned)((t1 - t0) / 8));
        }
        if (off != 0) best = 0;
        out[blockIdx.x * nAddr + k] = best;
    }
}

enum { WARPS = 14, STAGES = 3, CHUNK = 4096 };

__global__ void __launch_bounds__(WARPS * 32, 1)
streamKernel(const float *J, int N, int rowsPerWarp, int mode, const int *dieOfSm, unsigned seed, float *sink, float fracLast, int secondary) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * STAGES * CHUNK);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < WARPS * STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(&bars[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int half = (mode == 0 || mode == 3) ? -1 : (mode == 1 ? dieOfSm[smid & 255] : (int)(blockIdx.x & 1));
    const size_t rowBytes = (size_t)N * 4;
    const size_t spanBytes = (half < 0) ? rowBytes : rowBytes / 2;       /* bytes read per row visit */
    const size_t spanOff = (half <= 0) ? 0 : rowBytes / 2;
    const int chunksPerSpan = (int)(spanBytes / CHUNK);
    const int visits = (half < 0) ? rowsPerWarp : 2 * rowsPerWarp;       /* same bytes per warp in every mode */
    unsigned char *ring = smem + (size_t)warp * STAGES * CHUNK;
    uint64_t *myBars = bars + warp * STAGES;
    const long long total = (long long)visits * chunksPerSpan;
    float acc = 0.f;
    long long issued = 0;
    /* mode 3: full rows with an L2 policy -- a fraction of the lines (by address) evict_last, the rest normal / evict_first */
    uint64_t policy = 0;
    if (mode == 3) {
        if (secondary == 0) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(policy) : "f"(fracLast));
        else asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(policy) : "f"(fracLast));
    }
    auto issue = [&]() {
        if (issued >= total) return;
        const int v = (int)(issued / chunksPerSpan), c = (int)(issued % chunksPerSpan);
        const uint32_t row = hash32(seed ^ (uint32_t)(blockIdx.x * 131071u + warp * 8191u + v * 7u)) % (uint32_t)N;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(J) + (size_t)row * rowBytes + spanOff + (size_t)c * CHUNK;
        const int s = (int)(issued % STAGES);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(&myBars[s])), "r"((uint32_t)CHUNK) : "memory");
        if (mode == 3)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                             smemAddr(ring + (size_t)s * CHUNK)),
                         "l"(src), "r"((uint32_t)CHUNK), "r"(smemAddr(&myBars[s])), "l"(policy)
                         : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(ring + (size_t)s * CHUNK)),
                         "l"(src), "r"((uint32_t)CHUNK), "r"(smemAddr(&myBars[s]))
                         : "memory");
        ++issued;
    };
    if (lane == 0) for (int s = 0; s < STAGES; ++s) issue();
    for (long long k = 0; k < total; ++k) {
        const int s = (int)(k % STAGES);
        const uint32_t parity = (uint32_t)((k / STAGES) & 1);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smemAddr(&myBars[s])), "r"(parity) : "memory");
        }
        const float4 *p = reinterpret_cast<const float4 *>(ring + (size_t)s * CHUNK);
        for (int i = lane; i < CHUNK / 16; i += 32) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
        if (lane == 0) issue();
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8192;
    const int rowsPerWarp = argc > 2 ? atoi(argv[2]) : 2048;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nSM = prop.multiProcessorCount;
    float *J;
    CK(cudaMalloc(&J, (size_t)N * N * 4));
    CK(cudaMemset(J, 0, (size_t)N * N * 4));
    /* ---- which SMs share a die: correlate per-SM load latencies over lines whose home die alternates at random ---- */
    const int nAddr = 64;
    unsigned *dLat, *dSm;
    CK(cudaMalloc(&dLat, sizeof(unsigned) * nSM * nAddr));
    CK(cudaMalloc(&dSm, sizeof(unsigned) * nSM));
    latencyProbe<<<nSM, 32>>>((const unsigned long long *)J, (size_t)(1 << 20) / 8 + 512, nAddr, dLat, dSm);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned> lat(nSM * nAddr), smid(nSM);
    CK(cudaMemcpy(lat.data(), dLat, sizeof(unsigned) * nSM * nAddr, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(smid.data(), dSm, sizeof(unsigned) * nSM, cudaMemcpyDeviceToHost));
    /* per address: median latency over SMs; sign pattern of (lat > median) per SM; die = agreement with SM 0's pattern */
    std::vector<int> sign(nSM * nAddr);
    for (int k = 0; k < nAddr; ++k) {
        std::vector<unsigned> col(nSM);
        for (int s = 0; s < nSM; ++s) col[s] = lat[s * nAddr + k];
        std::vector<unsigned> sorted = col;
        std::sort(sorted.begin(), sorted.end());
        const unsigned med = sorted[nSM / 2];
        for (int s = 0; s < nSM; ++s) sign[s * nAddr + k] = col[s] > med ? 1 : 0;
    }
    std::vector<int> die(nSM);
    int n1 = 0;
    double meanAgree = 0;
    for (int s = 0; s < nSM; ++s) {
        int agree = 0;
        for (int k = 0; k < nAddr; ++k) agree += (sign[s * nAddr + k] == sign[k]);
        die[s] = (agree * 2 >= nAddr) ? 0 : 1;
        n1 += die[s];
        meanAgree += (double)std::max(agree, nAddr - agree) / nAddr;
    }
    printf("die probe: %d SMs on die 0, %d on die 1; mean pattern agreement %.2f (0.5 = no signal); latency range %u..%u cycles\n", nSM - n1, n1,
           meanAgree / nSM, *std::min_element(lat.begin(), lat.end()), *std::max_element(lat.begin(), lat.end()));
    std::vector<int> dieBySmid(256, 0);
    for (int s = 0; s < nSM; ++s) dieBySmid[smid[s] & 255] = die[s];
    int *dDie;
    CK(cudaMalloc(&dDie, sizeof(int) * 256));
    CK(cudaMemcpy(dDie, dieBySmid.data(), sizeof(int) * 256, cudaMemcpyHostToDevice));
    float *sink;
    CK(cudaMalloc(&sink, 4));
    const size_t smemBytes = (size_t)WARPS * STAGES * CHUNK + WARPS * STAGES * 8 + 128;
    CK(cudaFuncSetAttribute(streamKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    struct Run { int mode; float frac; int secondary; const char *what; };
    const Run runs[] = {{0, 0.f, 0, "full rows"}, {1, 0.f, 0, "column halves by die"}, {2, 0.f, 0, "column halves by CTA parity"},
                        {3, 0.25f, 0, "evict_last 25% / normal"}, {3, 0.40f, 0, "evict_last 40% / normal"}, {3, 0.60f, 0, "evict_last 60% / normal"},
                        {3, 0.25f, 1, "evict_last 25% / evict_first"}, {3, 0.40f, 1, "evict_last 40% / evict_first"}, {3, 1.0f, 0, "evict_last 100%"},
                        {0, 0.f, 0, "full rows (again)"}};
    for (const Run &r : runs) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        streamKernel<<<nSM, WARPS * 32, smemBytes>>>(J, N, rowsPerWarp, r.mode, dDie, 12345u, sink, r.frac, r.secondary);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)nSM * WARPS * rowsPerWarp * N * 4;
        printf("%-32s: %.3f ms, %.2f TB/s of row bytes\n", r.what, ms, bytes / ms / 1e9);
    }
    return 0;
}
