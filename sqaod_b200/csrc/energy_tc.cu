/* energy_tc.cu -- the one tensor-core kernel: split-precision spin GEMM on tcgen05 (sm_100a).
 *
 *     C[y][i] = sum_k Q[y][k] * A[i][k]          Q: m x K spins/bits (int8, exactly representable in bf16)
 *                                                 A: NA x K fp32 couplings (J, J^T or W)
 * replaces the reference's cublasSgemm call sites on this path: bipartite `Jq = qFixed . J(^T)`
 * (CUDABipartiteGraphAnnealer.cu:392-401 -> DeviceMath.cpp:167-178, 289-309) and the batched energy
 * `xA = q . J^T` (DeviceMath.cpp:191-209).  A is split ONCE per problem into three bf16 matrices
 * A = hi + mid + lo (each residual is computed exactly in fp32, so |A - hi - mid - lo| <= 2^-27 |A|, below fp32 epsilon);
 * Q is widened to bf16 per call.  Each of the three partial products Q.hi^T, Q.mid^T, Q.lo^T gets its OWN fp32 TMEM
 * accumulator (3 x 128 columns) and the epilogue adds them with round-to-nearest: tensor-core accumulation truncates, and
 * adding the tiny mid/lo products into one large accumulator would bias every sum towards zero (measured: 1.1e-5
 * relative on the C2 energies).  Kept apart, the hi sum needs ~21 bits and the mid/lo sums are small, so each is
 * accumulated essentially exactly; the result is within an ulp or two of the fp32 GEMM and exact on quantised inputs.
 *
 * Structure (one CTA per 128 x 128 output tile, 192 threads):
 *   warp 0  : TMA producer -- per K block one 128x64 bf16 box (128B swizzle) of Q and one of each of the three planes of A
 *             (64 KiB per stage, 3 stages)
 *   warp 1  : TMEM allocation + single-thread tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) issue,
 *             tcgen05.commit -> mbarrier to recycle shared-memory stages and to hand the accumulator over
 *   warps 2-5: epilogue -- tcgen05.ld 32x32b.x32 of the accumulator, fp32 stores of C
 * Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (InstrDescriptor, SmemDescriptor version 1).
 */
#include "device.hpp"
#include "kernels_common.cuh"
#include "b200_solvers.hpp"
#include "tc_gemm.hpp"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace sqb {

enum { TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_STAGES = 3, TC_THREADS = 192, TC_TMEM_COLS = 512 /* 3 x 128 used */ };
/* one pipeline stage = one K block of Q and the same K block of the three planes of A: Q is fetched once, not once per plane */
enum { TC_TILE_BYTES = TC_BN * TC_BK * 2, TC_STAGE_BYTES = (TC_BM + 3 * TC_BN) * TC_BK * 2, TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 + 256 };

__device__ __forceinline__ void tma2D(void *smemDst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smemAddr(smemDst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t umdesc(const void *smemTile) { /* K-major, SWIZZLE_128B, 8-row groups 1024 B apart */
    const uint64_t addr = (uint64_t)((smemAddr(smemTile) >> 4) & 0x3fffu);
    return addr | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void tcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbarArrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tcSpinGemmKernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapB, float *C, int ldc, int m, int NA,
                 int rowsPadB, int kBlocksPerSplit) {
    extern __shared__ unsigned char smemRaw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023); /* 128B swizzle atoms: 1024-byte aligned */
    uint64_t *fullBar = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *emptyBar = fullBar + TC_STAGES;
    uint64_t *tmemFullBar = emptyBar + TC_STAGES;
    uint32_t *tmemBaseSlot = reinterpret_cast<uint32_t *>(tmemFullBar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * TC_BN, m0 = blockIdx.x * TC_BM; /* batch tiles on grid.x (may exceed 65535 / 128 rows) */
    const int numKB = kBlocksPerSplit;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbarInit(&fullBar[s], 1); mbarInit(&emptyBar[s], 1); }
        mbarInit(tmemFullBar, 1);
        mbarInitFence();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) { /* whole warp: allocate the accumulator columns */
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smemAddr(tmemBaseSlot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcFenceBefore();
    __syncthreads();
    tcFenceAfter();
    const uint32_t tmemBase = *tmemBaseSlot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < numKB; ++kb) {
                const int s = kb % TC_STAGES;
                if (kb >= TC_STAGES) mbarWait(&emptyBar[s], ((kb / TC_STAGES) - 1) & 1);
                unsigned char *a = smem + s * TC_STAGE_BYTES, *b = a + TC_BM * TC_BK * 2;
                const int kc = kb * TC_BK;
                mbarArriveExpectTx(&fullBar[s], TC_STAGE_BYTES);
                tma2D(a, &mapQ, kc, m0, &fullBar[s]);
                for (int split = 0; split < 3; ++split) tma2D(b + split * TC_TILE_BYTES, &mapB, kc, split * rowsPadB + n0, &fullBar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            /* c=F32 (bit 4), a=b=BF16 (bits 7, 10), K-major both, N>>3 at bit 17, M>>4 at bit 24 */
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < numKB; ++kb) {
                const int s = kb % TC_STAGES;
                mbarWait(&fullBar[s], (kb / TC_STAGES) & 1);
                tcFenceAfter();
                const unsigned char *a = smem + s * TC_STAGE_BYTES, *b = a + TC_BM * TC_BK * 2;
                const uint64_t ad = umdesc(a);
#pragma unroll
                for (int split = 0; split < 3; ++split) { /* hi / mid / lo -> accumulator columns 0 / 128 / 256 */
                    const uint64_t bd = umdesc(b + split * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) /* +32 bytes per K=16 step inside the 128-byte swizzle row */
                        umma(tmemBase + (uint32_t)(split * TC_BN), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                }
                ummaCommit(&emptyBar[s]);
            }
            ummaCommit(tmemFullBar);
        }
    } else {
        mbarWait(tmemFullBar, 0);
        tcFenceAfter();
        const int quarter = warp & 3; /* a warp may only touch TMEM lanes 32*(warp%4) .. +31 */
        const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < TC_BN; c += 32) {
            uint32_t r[32];
            float acc[32];
#pragma unroll
            for (int part = 2; part >= 0; --part) { /* lo, mid, hi: small terms first */
                const uint32_t taddr = tmemBase + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(part * TC_BN + c);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = (part == 2) ? __uint_as_float(r[j]) : acc[j] + __uint_as_float(r[j]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(acc[j]);
            if (row < m) {
                float *dst = C + (size_t)row * ldc + n0 + c;
                if (n0 + c + 32 <= NA && (ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(dst + j) =
                            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + c + j < NA) dst[j] = __uint_as_float(r[j]);
                }
            }
        }
        tcFenceBefore();
    }
    __syncthreads();
    if (warp == 1) {
        tcFenceAfter();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBase), "r"((uint32_t)TC_TMEM_COLS) : "memory");
    }
}

/* ---- operand preparation ---- */
__global__ void tcSplitKernel(__nv_bfloat16 *out, int rowsPad, int Kp, const float *A, int ldA, int rows, int K) {
    const int k = blockIdx.y * blockDim.x + threadIdx.x, r = blockIdx.x; /* rows on grid.x: no 65535 limit */
    if (k >= K || r >= rows) return;
    const float a = A[(size_t)r * ldA + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(a);
    const float r1 = a - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
    const size_t plane = (size_t)rowsPad * Kp, at = (size_t)r * Kp + k;
    out[at] = hi;
    out[plane + at] = mid;
    out[2 * plane + at] = lo;
}
__global__ void tcWidenSpinsKernel(__nv_bfloat16 *out, int Kp, const signed char *Q, int ldq, int m, int K) {
    const int k = blockIdx.y * blockDim.x + threadIdx.x, r = blockIdx.x;
    if (k >= Kp || r >= m) return;
    out[(size_t)r * Kp + k] = __float2bfloat16_rn(k < K ? (float)Q[(size_t)r * ldq + k] : 0.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encodeTiled() {
    static EncodeTiledFn fn = NULL;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = NULL;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
static void makeMap(CUtensorMap *map, void *base, int rows, int Kp, int boxRows) {
    EncodeTiledFn fn = encodeTiled();
    sqb_throwErrorIf(fn == NULL, "cuTensorMapEncodeTiled is not available from the driver.");
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)boxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    sqb_throwErrorIf(rc != CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d).", (int)rc);
}

bool tcEnabled() {
    const char *env = getenv("SQAOD_B200_NO_TC");
    if (env != NULL && *env != '0') return false;
    return encodeTiled() != NULL;
}

void tcPrepareOperand(const B200Device &dev, TcOperand &op, const float *d_A, int ldA, int rows, int K) {
    op.rows = rows;
    op.K = K;
    op.rowsPad = sq::roundUp(rows, TC_BN);
    op.Kp = sq::roundUp(K, TC_BK);
    op.data.alloc(&dev, (size_t)3 * op.rowsPad * op.Kp); /* zero filled: padding rows / columns contribute nothing */
    dim3 grid(rows, (K + 127) / 128);
    tcSplitKernel<<<grid, 128, 0, dev.stream()>>>((__nv_bfloat16 *)op.data.p, op.rowsPad, op.Kp, d_A, ldA, rows, K);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
    makeMap((CUtensorMap *)op.map, op.data.p, 3 * op.rowsPad, op.Kp, TC_BN);
    op.ready = true;
}

void tcWidenSpins(const B200Device &dev, DevBuf<unsigned short> &qbf, const TcOperand &B, const signed char *d_Q, int ldq, int m) {
    const int mPad = sq::roundUp(m, TC_BM); /* whole boxes: padding rows stay zero */
    const size_t need = (size_t)mPad * B.Kp;
    if (qbf.n < need || qbf.dev != &dev) qbf.alloc(&dev, need);
    dim3 wgrid(m, (B.Kp + 127) / 128);
    tcWidenSpinsKernel<<<wgrid, 128, 0, dev.stream()>>>((__nv_bfloat16 *)qbf.p, B.Kp, d_Q, ldq, m, B.K);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}

void tcSpinGemmBf16(const B200Device &dev, float *d_C, int ldc, const TcOperand &B, const unsigned short *d_Qbf, int m) {
    sqb_throwErrorIf(!B.ready, "tensor-core operand not prepared.");
    const int mPad = sq::roundUp(m, TC_BM);
    CUtensorMap mapQ;
    makeMap(&mapQ, const_cast<unsigned short *>(d_Qbf), mPad, B.Kp, TC_BM);
    static unsigned long long attrSetMask = 0ull; /* the attribute is per device: one bit per device number */
    dev.makeCurrent();
    const unsigned long long devBit = 1ull << (dev.devNo() & 63);
    if (!(attrSetMask & devBit)) {
        CUDA_CHECK(cudaFuncSetAttribute(tcSpinGemmKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
        attrSetMask |= devBit;
    }
    dim3 grid((m + TC_BM - 1) / TC_BM, (B.rows + TC_BN - 1) / TC_BN);
    tcSpinGemmKernel<<<grid, TC_THREADS, TC_SMEM_BYTES, dev.stream()>>>(mapQ, *(const CUtensorMap *)B.map, d_C, ldc, m, B.rows, B.rowsPad,
                                                                        B.Kp / TC_BK);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}

void tcSpinGemm(const B200Device &dev, float *d_C, int ldc, const TcOperand &B, const signed char *d_Q, int ldq, int m, TcWorkspace &ws) {
    ws.dev = &dev;
    tcWidenSpins(dev, ws.qbf, B, d_Q, ldq, m);
    tcSpinGemmBf16(dev, d_C, ldc, B, ws.qbf.p, m);
}

/* E_b = alpha * ( sum_i v_bi (g_i + C_bi) + sum_j f_j u_bj ) + beta0, C = u . A^T from tcSpinGemm */
__global__ void tcRowDotEnergyKernel(float *E, const float *C, int ldc, const signed char *v, int ldv, const float *g, int R,
                                     const float *f, const signed char *u, int ldu, int Ccols, float alpha, float beta0) {
    const int b = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) s += (float)v[(size_t)b * ldv + i] * ((g ? g[i] : 0.f) + C[(size_t)b * ldc + i]);
    if (f)
        for (int j = threadIdx.x; j < Ccols; j += blockDim.x) s += f[j] * (float)u[(size_t)b * ldu + j];
    __shared__ float sh[8];
    s = warpSum(s);
    if (laneId() == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        E[b] = alpha * t + beta0;
    }
}

void tcBatchedEnergy(const B200Device &dev, float *d_E, const TcOperand &A, const signed char *d_u, int ldu, const signed char *d_v, int ldv,
                     const float *d_g, const float *d_f, int nBatch, float alpha, float beta0, TcWorkspace &ws) {
    const int ldc = sq::roundUp(A.rows, 32);
    const size_t need = (size_t)nBatch * ldc;
    if (ws.cbuf.n < need || ws.cbuf.dev != &dev) ws.cbuf.alloc(&dev, need);
    tcSpinGemm(dev, ws.cbuf.p, ldc, A, d_u, ldu, nBatch, ws);
    tcRowDotEnergyKernel<<<nBatch, 256, 0, dev.stream()>>>(d_E, ws.cbuf.p, ldc, d_v, ldv, d_g, A.rows, d_f, d_u, ldu, A.K, alpha, beta0);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}

} // namespace sqb
