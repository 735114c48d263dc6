/* formulas.cu -- QUBO <-> Ising conversion and batched energy evaluation on the device.
 *
 * Replaces DeviceFormulas.cpp:23-140 / DeviceMath.cpp:154-220 (cuBLAS gemm/gemv + CUB reductions in the reference)
 * and CUDAFormulas.cpp:9-272 (host-matrix wrappers).  Math: sqaodc/cpu/SharedFormulas.cpp:9-202.
 *   dense:      h = -1/2 colsum(W), J = -1/4 W (zero diagonal), c = sum(J) + sum(diag J)
 *               E_b = -c - h.q_b - q_b^T J q_b            (spins)        E_b = x_b^T W x_b   (bits)
 *   bipartite:  J = -1/4 W, h0 = -1/4 colsum(W) - 1/2 b0, h1 = -1/4 rowsum(W) - 1/2 b1, c = -1/4 sum(W) - 1/2 (sum b0 + sum b1)
 *               E_b = -c - h0.q0_b - h1.q1_b - q1_b^T J q0_b      E_b = b0.x0_b + b1.x1_b + x1_b^T W x0_b
 * This file holds the CUDA-core (FFMA/DFMA) energy path; spins/bits are exact small integers, accumulation is in `real`.
 */
#include "device.hpp"
#include "kernels_common.cuh"
#include "b200_solvers.hpp"

namespace sqb {

/* ---------------- QUBO -> Ising ---------------- */
template <class real>
__global__ void colSumKernel(real *out, const real *A, int ldA, int rows, int cols, real scale, const real *addv, real addScale) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    real s = 0;
    for (int i = 0; i < rows; ++i) s += A[(size_t)i * ldA + j];
    real v = scale * s;
    if (addv) v += addScale * addv[j];
    out[j] = v;
}
template <class real>
__global__ void rowSumKernel(real *out, const real *A, int ldA, int rows, int cols, real scale, const real *addv, real addScale) {
    int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows) return;
    real s = 0;
    for (int j = laneId(); j < cols; j += 32) s += A[(size_t)i * ldA + j];
    s = warpSum(s);
    if (laneId() == 0) {
        real v = scale * s;
        if (addv) v += addScale * addv[i];
        out[i] = v;
    }
}
template <class real>
__global__ void scaleMatrixKernel(real *J, int ldJ, const real *W, int ldW, int rows, int cols, real scale, int zeroDiag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ldJ || i >= rows) return;
    real v = 0;
    if (j < cols) {
        v = scale * W[(size_t)i * ldW + j];
        if (zeroDiag && i == j) v = 0;
    }
    J[(size_t)i * ldJ + j] = v; /* also clears the row padding (reference: clearPadding, DeviceCopy) */
}
/* partial[b] = sum over a strided slice of scale*A (+ diag once more when diagToo) ; double accumulation */
template <class real>
__global__ void matrixSumPartialKernel(double *partial, const real *A, int ldA, int rows, int cols, int diagToo) {
    double s = 0;
    for (int i = blockIdx.x; i < rows; i += gridDim.x)
        for (int j = threadIdx.x; j < cols; j += blockDim.x) {
            double v = (double)A[(size_t)i * ldA + j];
            s += v;
            if (diagToo && i == j) s += v;
        }
    __shared__ double sh[32];
    s = warpSum(s);
    if (laneId() == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}
template <class real>
__global__ void finishConstKernel(real *c, const double *partial, int n, double scale, const real *v0, int n0, const real *v1, int n1,
                                  double vscale) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0;
    for (int i = 0; i < n; ++i) s += partial[i];
    s *= scale;
    double t = 0;
    for (int i = 0; i < n0; ++i) t += (double)v0[i];
    for (int i = 0; i < n1; ++i) t += (double)v1[i];
    *c = (real)(s + vscale * t);
}

template <class real>
void devDenseHamiltonian(const B200Device &dev, real *d_h, real *d_J, int ldJ, real *d_c, const real *d_W, int ldW, int N, real sign) {
    cudaStream_t st = dev.stream();
    colSumKernel<real><<<(N + 127) / 128, 128, 0, st>>>(d_h, d_W, ldW, N, N, real(-0.5) * sign, (const real *)NULL, real(0));
    DevBuf<double> part;
    const int nPart = 256;
    part.alloc(&dev, nPart);
    matrixSumPartialKernel<real><<<nPart, 256, 0, st>>>(part.p, d_W, ldW, N, N, 1);
    finishConstKernel<real><<<1, 32, 0, st>>>(d_c, part.p, nPart, -0.25 * (double)sign, (const real *)NULL, 0, (const real *)NULL, 0, 0.);
    dim3 grid((ldJ + 127) / 128, N);
    scaleMatrixKernel<real><<<grid, 128, 0, st>>>(d_J, ldJ, d_W, ldW, N, N, real(-0.25) * sign, 1);
    CUDA_CHECK(cudaGetLastError());
    dev.launchCount += 4;
}

template <class real>
void devBipartiteHamiltonian(const B200Device &dev, real *d_h0, real *d_h1, real *d_J, int ldJ, real *d_c, const real *d_b0,
                             const real *d_b1, const real *d_W, int ldW, int N0, int N1, real sign) {
    cudaStream_t st = dev.stream();
    colSumKernel<real><<<(N0 + 127) / 128, 128, 0, st>>>(d_h0, d_W, ldW, N1, N0, real(-0.25) * sign, d_b0, real(-0.5) * sign);
    rowSumKernel<real><<<(N1 + 3) / 4, 128, 0, st>>>(d_h1, d_W, ldW, N1, N0, real(-0.25) * sign, d_b1, real(-0.5) * sign);
    DevBuf<double> part;
    const int nPart = 256;
    part.alloc(&dev, nPart);
    matrixSumPartialKernel<real><<<nPart, 256, 0, st>>>(part.p, d_W, ldW, N1, N0, 0);
    finishConstKernel<real><<<1, 32, 0, st>>>(d_c, part.p, nPart, -0.25 * (double)sign, d_b0, N0, d_b1, N1, -0.5 * (double)sign);
    dim3 grid((ldJ + 127) / 128, N1);
    scaleMatrixKernel<real><<<grid, 128, 0, st>>>(d_J, ldJ, d_W, ldW, N1, N0, real(-0.25) * sign, 0);
    CUDA_CHECK(cudaGetLastError());
    dev.launchCount += 5;
}

/* ---------------- batched bilinear energy (CUDA cores) ---------------- */
enum { EN_BT = 8, EN_ROWS = 32, EN_CT = 2048, EN_THREADS = 256 };

/* block (rb, bb): rows [rb*32, +32) of A, batch entries [bb*8, +8).  Column tiles of u are staged in shared memory as
 * `real`; each warp owns 4 rows and keeps 4 x 8 partial sums in registers.  partial[rb][b] = sum_{i in rows} v_bi (g_i + A_i.u_b) */
template <class real>
__global__ void __launch_bounds__(EN_THREADS) energyPartialKernel(real *partial, const real *A, int ldA, int R, int C,
                                                                    const signed char *u, int ldu, const signed char *v, int ldv,
                                                                    const real *g, int nBatch) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    real *us = reinterpret_cast<real *>(smemRaw); /* [EN_BT][EN_CT] */
    __shared__ real blockAcc[EN_THREADS / 32][EN_BT];
    const int lane = laneId(), warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * EN_ROWS, b0 = blockIdx.y * EN_BT;
    const int nb = min(EN_BT, nBatch - b0);
    real acc[4][EN_BT];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int b = 0; b < EN_BT; ++b) acc[k][b] = real(0);

    for (int c0 = 0; c0 < C; c0 += EN_CT) {
        const int cw = min(EN_CT, C - c0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < EN_BT * EN_CT; idx += EN_THREADS) {
            int b = idx / EN_CT, j = idx % EN_CT;
            us[idx] = (b < nb && j < cw) ? (real)u[(size_t)(b0 + b) * ldu + c0 + j] : real(0);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = r0 + warp * 4 + k;
            if (i >= R) continue;
            const real *arow = A + (size_t)i * ldA + c0;
            for (int j = lane; j < cw; j += 32) {
                const real a = arow[j];
#pragma unroll
                for (int b = 0; b < EN_BT; ++b) acc[k][b] += a * us[b * EN_CT + j];
            }
        }
    }
    real mine[EN_BT];
#pragma unroll
    for (int b = 0; b < EN_BT; ++b) mine[b] = real(0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = r0 + warp * 4 + k;
#pragma unroll
        for (int b = 0; b < EN_BT; ++b) {
            real s = warpSum(acc[k][b]);
            if (i < R && b < nb) {
                real vi = (real)v[(size_t)(b0 + b) * ldv + i];
                mine[b] += vi * ((g ? g[i] : real(0)) + s);
            }
        }
    }
    if (lane == 0)
        for (int b = 0; b < EN_BT; ++b) blockAcc[warp][b] = mine[b];
    __syncthreads();
    if (threadIdx.x < EN_BT && threadIdx.x < nb) {
        real s = 0;
        for (int w = 0; w < EN_THREADS / 32; ++w) s += blockAcc[w][threadIdx.x];
        partial[(size_t)blockIdx.x * nBatch + b0 + threadIdx.x] = s;
    }
}
template <class real>
__global__ void energyFinishKernel(real *E, const real *partial, int nRowBlocks, int nBatch, const real *f, const signed char *u,
                                   int ldu, int C, real alpha, real beta0) {
    int b = blockIdx.x;
    real s = 0;
    if (f)
        for (int j = threadIdx.x; j < C; j += blockDim.x) s += f[j] * (real)u[(size_t)b * ldu + j];
    for (int k = threadIdx.x; k < nRowBlocks; k += blockDim.x) s += partial[(size_t)k * nBatch + b];
    __shared__ real sh[8];
    s = warpSum(s);
    if (laneId() == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        real t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        E[b] = alpha * t + beta0;
    }
}

template <class real>
void devBatchedEnergy(const B200Device &dev, real *d_E, const real *d_A, int ldA, int R, int C, const signed char *d_u, int ldu,
                      const signed char *d_v, int ldv, const real *d_g, const real *d_f, int nBatch, real alpha, real beta0) {
    if (nBatch <= 0) return;
    cudaStream_t st = dev.stream();
    const int nRowBlocks = (R + EN_ROWS - 1) / EN_ROWS;
    DevBuf<real> partial;
    partial.alloc(&dev, (size_t)nRowBlocks * nBatch);
    const size_t smem = (size_t)EN_BT * EN_CT * sizeof(real);
    CUDA_CHECK(cudaFuncSetAttribute(energyPartialKernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nRowBlocks, (nBatch + EN_BT - 1) / EN_BT);
    energyPartialKernel<real><<<grid, EN_THREADS, smem, st>>>(partial.p, d_A, ldA, R, C, d_u, ldu, d_v, ldv, d_g, nBatch);
    energyFinishKernel<real><<<nBatch, 256, 0, st>>>(d_E, partial.p, nRowBlocks, nBatch, d_f, d_u, ldu, C, alpha, beta0);
    CUDA_CHECK(cudaGetLastError());
    dev.launchCount += 2;
}

/* ---------------- 2-D bipartite QUBO energy: every (x1_i1, x0_i0) pair ---------------- */
template <class real>
__global__ void wx0Kernel(real *tmp, const real *W, int ldW, int N0, int N1, const signed char *x0, int ldx0, int n0) {
    /* tmp[i0][r] = sum_j W[r][j] x0[i0][j] */
    int r = blockIdx.x * blockDim.x + threadIdx.x, i0 = blockIdx.y;
    if (r >= N1 || i0 >= n0) return;
    real s = 0;
    for (int j = 0; j < N0; ++j) s += W[(size_t)r * ldW + j] * (real)x0[(size_t)i0 * ldx0 + j];
    tmp[(size_t)i0 * N1 + r] = s;
}
template <class real>
__global__ void energy2DKernel(real *E, int ldE, const real *tmp, const real *b0, const real *b1, int N0, int N1,
                               const signed char *x0, int ldx0, int n0, const signed char *x1, int ldx1, int n1) {
    int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y;
    if (i0 >= n0 || i1 >= n1) return;
    real e = 0;
    for (int j = 0; j < N0; ++j) e += b0[j] * (real)x0[(size_t)i0 * ldx0 + j];
    for (int r = 0; r < N1; ++r) {
        real xr = (real)x1[(size_t)i1 * ldx1 + r];
        e += xr * (b1[r] + tmp[(size_t)i0 * N1 + r]);
    }
    E[(size_t)i1 * ldE + i0] = e;
}
template <class real>
void devBipartiteEnergy2D(const B200Device &dev, real *d_E, int ldE, const real *d_b0, const real *d_b1, const real *d_W, int ldW,
                          int N0, int N1, const signed char *d_x0, int ldx0, int n0, const signed char *d_x1, int ldx1, int n1) {
    cudaStream_t st = dev.stream();
    DevBuf<real> tmp;
    tmp.alloc(&dev, (size_t)n0 * N1);
    wx0Kernel<real><<<dim3((N1 + 63) / 64, n0), 64, 0, st>>>(tmp.p, d_W, ldW, N0, N1, d_x0, ldx0, n0);
    energy2DKernel<real><<<dim3((n0 + 63) / 64, n1), 64, 0, st>>>(d_E, ldE, tmp.p, d_b0, d_b1, N0, N1, d_x0, ldx0, n0, d_x1, ldx1, n1);
    CUDA_CHECK(cudaGetLastError());
    dev.launchCount += 2;
}

#define INSTANTIATE(real)                                                                                                       \
    template void devDenseHamiltonian<real>(const B200Device &, real *, real *, int, real *, const real *, int, int, real);       \
    template void devBipartiteHamiltonian<real>(const B200Device &, real *, real *, real *, int, real *, const real *, const real *, \
                                                const real *, int, int, int, real);                                               \
    template void devBatchedEnergy<real>(const B200Device &, real *, const real *, int, int, int, const signed char *, int,         \
                                         const signed char *, int, const real *, const real *, int, real, real);                   \
    template void devBipartiteEnergy2D<real>(const B200Device &, real *, int, const real *, const real *, const real *, int, int,   \
                                             int, const signed char *, int, int, const signed char *, int, int);
INSTANTIATE(float)
INSTANTIATE(double)

/* =====================================================================================
 * Formulas objects with host matrices in / out (reference: CUDAFormulas.cpp:9-272)
 * ===================================================================================== */
namespace {

template <class real> struct HostToDev { /* uploads a host real matrix / vector; bit matrices are narrowed to int8 */
    static void matrix(const B200Device &dev, DevBuf<real> &d, int &ld, const sq::MatrixType<real> &M) {
        ld = sq::roundUp(M.cols, 32);
        d.alloc(&dev, (size_t)M.rows * ld);
        dev.h2d2D(d.p, sizeof(real) * ld, M.data, sizeof(real) * M.stride, sizeof(real) * M.cols, M.rows);
    }
    static void vector(const B200Device &dev, DevBuf<real> &d, const sq::VectorType<real> &v) {
        d.alloc(&dev, v.size);
        dev.h2d(d.p, v.data, sizeof(real) * v.size);
    }
    static void bits(const B200Device &dev, DevBuf<signed char> &d, int &ld, std::vector<signed char> &stage,
                     const real *data, int rows, int cols, int stride) {
        ld = sq::roundUp(cols, 16);
        stage.assign((size_t)rows * ld, 0);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) stage[(size_t)r * ld + c] = (signed char)data[(size_t)r * stride + c];
        d.alloc(&dev, stage.size());
        dev.h2d(d.p, stage.data(), stage.size());
    }
};

template <class real> class B200DenseGraphFormulas : public sq::cuda::DenseGraphFormulas<real> {
    typedef sq::MatrixType<real> Matrix;
    typedef sq::VectorType<real> Vector;
    B200Device *dev_;
public:
    B200DenseGraphFormulas() : dev_(NULL) {}
    void assignDevice(sq::cuda::Device &device) { dev_ = &asB200(device); }
    void check() const { sqb_throwErrorIf(dev_ == NULL, "Device not set."); }

    void energy(real *E, const Matrix &A, const real *g, const real *xdata, int rows, int cols, int stride, real alpha, real beta0) {
        check();
        sqb_throwErrorIf(A.rows != A.cols || cols != A.rows, "shape mismatch.");
        DevBuf<real> dA, dg, dE;
        DevBuf<signed char> dx;
        std::vector<signed char> stage;
        int ldA, ldx;
        HostToDev<real>::matrix(*dev_, dA, ldA, A);
        HostToDev<real>::bits(*dev_, dx, ldx, stage, xdata, rows, cols, stride);
        if (g) { dg.alloc(dev_, cols); dev_->h2d(dg.p, g, sizeof(real) * cols); }
        dE.alloc(dev_, rows);
        devBatchedEnergy<real>(*dev_, dE.p, dA.p, ldA, A.rows, A.cols, dx.p, ldx, dx.p, ldx, g ? dg.p : NULL, NULL, rows, alpha, beta0);
        dev_->d2h(E, dE.p, sizeof(real) * rows);
        dev_->synchronize();
    }
    void calculate_E(real *E, const Matrix &W, const Vector &x) { energy(E, W, NULL, x.data, 1, x.size, x.size, real(1), real(0)); }
    void calculate_E(Vector *E, const Matrix &W, const Matrix &x) {
        E->resize(x.rows);
        energy(E->data, W, NULL, x.data, x.rows, x.cols, x.stride, real(1), real(0));
    }
    void calculateHamiltonian(Vector *h, Matrix *J, real *c, const Matrix &W) {
        check();
        sqb_throwErrorIf(W.rows != W.cols, "W is not a sqare matrix.");
        const int N = W.rows;
        h->resize(N);
        J->resize(N, N);
        DevBuf<real> dW, dJ, dh, dc;
        int ld;
        HostToDev<real>::matrix(*dev_, dW, ld, W);
        dJ.alloc(dev_, (size_t)N * ld); dh.alloc(dev_, N); dc.alloc(dev_, 1);
        devDenseHamiltonian<real>(*dev_, dh.p, dJ.p, ld, dc.p, dW.p, ld, N, real(1));
        dev_->d2h(h->data, dh.p, sizeof(real) * N);
        dev_->d2h2D(J->data, sizeof(real) * J->stride, dJ.p, sizeof(real) * ld, sizeof(real) * N, N);
        dev_->d2h(c, dc.p, sizeof(real));
        dev_->synchronize();
    }
    void calculate_E(real *E, const Vector &h, const Matrix &J, real c, const Vector &q) {
        sqb_throwErrorIf(h.size != J.rows, "shape mismatch.");
        energy(E, J, h.data, q.data, 1, q.size, q.size, real(-1), -c);
    }
    void calculate_E(Vector *E, const Vector &h, const Matrix &J, real c, const Matrix &q) {
        sqb_throwErrorIf(h.size != J.rows, "shape mismatch.");
        E->resize(q.rows);
        energy(E->data, J, h.data, q.data, q.rows, q.cols, q.stride, real(-1), -c);
    }
};

template <class real> class B200BipartiteGraphFormulas : public sq::cuda::BipartiteGraphFormulas<real> {
    typedef sq::MatrixType<real> Matrix;
    typedef sq::VectorType<real> Vector;
    B200Device *dev_;
public:
    B200BipartiteGraphFormulas() : dev_(NULL) {}
    void assignDevice(sq::cuda::Device &device) { dev_ = &asB200(device); }
    void check() const { sqb_throwErrorIf(dev_ == NULL, "Device not set."); }

    /* E_b = alpha (v_b.(g + A u_b) + f.u_b) + beta0 with A N1 x N0, u = side 0, v = side 1 */
    void energy(real *E, const Vector &f, const Vector &g, const Matrix &A, const real *u, int ustride, const real *v, int vstride,
                int nBatch, real alpha, real beta0) {
        check();
        const int N0 = A.cols, N1 = A.rows;
        sqb_throwErrorIf(f.size != N0 || g.size != N1, "shape mismatch.");
        DevBuf<real> dA, df, dg, dE;
        DevBuf<signed char> du, dv;
        std::vector<signed char> s0, s1;
        int ldA, ldu, ldv;
        HostToDev<real>::matrix(*dev_, dA, ldA, A);
        HostToDev<real>::vector(*dev_, df, f);
        HostToDev<real>::vector(*dev_, dg, g);
        HostToDev<real>::bits(*dev_, du, ldu, s0, u, nBatch, N0, ustride);
        HostToDev<real>::bits(*dev_, dv, ldv, s1, v, nBatch, N1, vstride);
        dE.alloc(dev_, nBatch);
        devBatchedEnergy<real>(*dev_, dE.p, dA.p, ldA, N1, N0, du.p, ldu, dv.p, ldv, dg.p, df.p, nBatch, alpha, beta0);
        dev_->d2h(E, dE.p, sizeof(real) * nBatch);
        dev_->synchronize();
    }
    void calculate_E(real *E, const Vector &b0, const Vector &b1, const Matrix &W, const Vector &x0, const Vector &x1) {
        sqb_throwErrorIf(x0.size != W.cols || x1.size != W.rows, "shape mismatch.");
        energy(E, b0, b1, W, x0.data, x0.size, x1.data, x1.size, 1, real(1), real(0));
    }
    void calculate_E(Vector *E, const Vector &b0, const Vector &b1, const Matrix &W, const Matrix &x0, const Matrix &x1) {
        sqb_throwErrorIf(x0.cols != W.cols || x1.cols != W.rows || x0.rows != x1.rows, "shape mismatch.");
        E->resize(x0.rows);
        energy(E->data, b0, b1, W, x0.data, x0.stride, x1.data, x1.stride, x0.rows, real(1), real(0));
    }
    void calculate_E_2d(Matrix *E, const Vector &b0, const Vector &b1, const Matrix &W, const Matrix &x0, const Matrix &x1) {
        check();
        const int N0 = W.cols, N1 = W.rows;
        sqb_throwErrorIf(x0.cols != N0 || x1.cols != N1 || b0.size != N0 || b1.size != N1, "shape mismatch.");
        E->resize(x1.rows, x0.rows);
        DevBuf<real> dW, db0, db1, dE;
        DevBuf<signed char> dx0, dx1;
        std::vector<signed char> s0, s1;
        int ldW, ld0, ld1;
        HostToDev<real>::matrix(*dev_, dW, ldW, W);
        HostToDev<real>::vector(*dev_, db0, b0);
        HostToDev<real>::vector(*dev_, db1, b1);
        HostToDev<real>::bits(*dev_, dx0, ld0, s0, x0.data, x0.rows, N0, x0.stride);
        HostToDev<real>::bits(*dev_, dx1, ld1, s1, x1.data, x1.rows, N1, x1.stride);
        const int ldE = x0.rows;
        dE.alloc(dev_, (size_t)x1.rows * ldE);
        devBipartiteEnergy2D<real>(*dev_, dE.p, ldE, db0.p, db1.p, dW.p, ldW, N0, N1, dx0.p, ld0, x0.rows, dx1.p, ld1, x1.rows);
        dev_->d2h2D(E->data, sizeof(real) * E->stride, dE.p, sizeof(real) * ldE, sizeof(real) * x0.rows, x1.rows);
        dev_->synchronize();
    }
    void calculateHamiltonian(Vector *h0, Vector *h1, Matrix *J, real *c, const Vector &b0, const Vector &b1, const Matrix &W) {
        check();
        const int N0 = W.cols, N1 = W.rows;
        sqb_throwErrorIf(b0.size != N0 || b1.size != N1, "shape mismatch.");
        h0->resize(N0); h1->resize(N1); J->resize(N1, N0);
        DevBuf<real> dW, db0, db1, dJ, dh0, dh1, dc;
        int ld;
        HostToDev<real>::matrix(*dev_, dW, ld, W);
        HostToDev<real>::vector(*dev_, db0, b0);
        HostToDev<real>::vector(*dev_, db1, b1);
        dJ.alloc(dev_, (size_t)N1 * ld); dh0.alloc(dev_, N0); dh1.alloc(dev_, N1); dc.alloc(dev_, 1);
        devBipartiteHamiltonian<real>(*dev_, dh0.p, dh1.p, dJ.p, ld, dc.p, db0.p, db1.p, dW.p, ld, N0, N1, real(1));
        dev_->d2h(h0->data, dh0.p, sizeof(real) * N0);
        dev_->d2h(h1->data, dh1.p, sizeof(real) * N1);
        dev_->d2h2D(J->data, sizeof(real) * J->stride, dJ.p, sizeof(real) * ld, sizeof(real) * N0, N1);
        dev_->d2h(c, dc.p, sizeof(real));
        dev_->synchronize();
    }
    void calculate_E(real *E, const Vector &h0, const Vector &h1, const Matrix &J, real c, const Vector &q0, const Vector &q1) {
        sqb_throwErrorIf(q0.size != J.cols || q1.size != J.rows, "shape mismatch.");
        energy(E, h0, h1, J, q0.data, q0.size, q1.data, q1.size, 1, real(-1), -c);
    }
    void calculate_E(Vector *E, const Vector &h0, const Vector &h1, const Matrix &J, real c, const Matrix &q0, const Matrix &q1) {
        sqb_throwErrorIf(q0.cols != J.cols || q1.cols != J.rows || q0.rows != q1.rows, "shape mismatch.");
        E->resize(q0.rows);
        energy(E->data, h0, h1, J, q0.data, q0.stride, q1.data, q1.stride, q0.rows, real(-1), -c);
    }
};

} // namespace
} // namespace sqb

namespace sqaod { namespace cuda {
template <> DenseGraphFormulas<float> *newDenseGraphFormulas<float>() { return new sqb::B200DenseGraphFormulas<float>(); }
template <> DenseGraphFormulas<double> *newDenseGraphFormulas<double>() { return new sqb::B200DenseGraphFormulas<double>(); }
template <> BipartiteGraphFormulas<float> *newBipartiteGraphFormulas<float>() { return new sqb::B200BipartiteGraphFormulas<float>(); }
template <> BipartiteGraphFormulas<double> *newBipartiteGraphFormulas<double>() { return new sqb::B200BipartiteGraphFormulas<double>(); }
}} // namespace sqaod::cuda
