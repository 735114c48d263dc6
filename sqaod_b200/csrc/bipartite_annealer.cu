/* bipartite_annealer.cu -- bipartite-graph SQA / SA annealer for B200 (sm_100a).
 *
 * Replaces CUDABipartiteGraphAnnealer.cu:392-625 (cuBLAS gemm + transform2d flip kernels + MT19937 pool).  Chain as in
 * the reference CPU solver (sqaodc/cpu/CPUBipartiteGraphAnnealer.cpp:331-441, 485-542): one step = half step on side 1
 * then on side 0; a half step computes dEmat = qFixed . Jeff^T once (Jeff = J for side 1, J^T for side 0) and then
 * attempts EVERY spin of the annealed side, even trotters first, (odd m) trotter m-1, then odd trotters:
 *     dE = (2/m) q (h_i + dEmat[y][i]) - q (q[y-1][i] + q[y+1][i]) coef          (no factor 2: each edge counted once)
 * The uniform of attempt (side, y, i) is Philox(seed, step, domain=side, i, y) -- no pool, and unlike the reference's
 * CUDA kernel even and odd trotters never share a random number.
 * Spins are int8 on the device (the reference keeps them as `real`), so the contraction reads 1 byte per spin.
 *
 * This file holds the CUDA-core contraction (register-tiled FFMA/DFMA GEMM with q widened on the fly); the fp32 solver
 * switches to the tcgen05 split-precision GEMM in energy_tc.cu when that path is enabled.
 */
#include "device.hpp"
#include "kernels_common.cuh"
#include "philox.cuh"
#include "b200_solvers.hpp"
#include <math.h>
#include <type_traits>
#include <time.h>
#include <algorithm>

namespace sqb {

/* C[y][i] = sum_k Q[y][k] * A[i][k]   (y < m, i < NA, k < NF); A row-major NA x ldA, Q int8 m x ldq (zero padded) */
template <class real, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemmSpinKernel(real *C, int ldc, const real *A, int ldA, const signed char *Q,
                                                                         int ldq, int m, int NA, int NF) {
    enum { BK = 16, NT = (BM / TM) * (BN / TN) };
    __shared__ real As[BK][BN + 4];
    __shared__ real Qs[BK][BM + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int i0 = blockIdx.x * BN, y0 = blockIdx.y * BM;
    real acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = real(0);

    for (int k0 = 0; k0 < NF; k0 += BK) {
        for (int idx = tid; idx < BN * BK; idx += NT) { /* A tile: BN rows x BK */
            int r = idx / BK, k = idx % BK;
            int gi = i0 + r, gk = k0 + k;
            As[k][r] = (gi < NA && gk < NF) ? A[(size_t)gi * ldA + gk] : real(0);
        }
        for (int idx = tid; idx < BM * BK; idx += NT) {
            int r = idx / BK, k = idx % BK;
            int gy = y0 + r, gk = k0 + k;
            Qs[k][r] = (gy < m && gk < NF) ? (real)Q[(size_t)gy * ldq + gk] : real(0);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            real qa[TM], ab[TN];
#pragma unroll
            for (int a = 0; a < TM; ++a) qa[a] = Qs[k][ty * TM + a];
#pragma unroll
            for (int b = 0; b < TN; ++b) ab[b] = As[k][tx * TN + b];
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int b = 0; b < TN; ++b) acc[a][b] += qa[a] * ab[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
        int gy = y0 + ty * TM + a;
        if (gy >= m) continue;
#pragma unroll
        for (int b = 0; b < TN; ++b) {
            int gi = i0 + tx * TN + b;
            if (gi < NA) C[(size_t)gy * ldc + gi] = acc[a][b];
        }
    }
}

template <class real>
void devSpinGemm(const B200Device &dev, real *C, int ldc, const real *A, int ldA, const signed char *Q, int ldq, int m, int NA, int NF) {
    enum { BM = 64, BN = 64, TM = 4, TN = 4 };
    dim3 grid((NA + BN - 1) / BN, (m + BM - 1) / BM);
    gemmSpinKernel<real, BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, dev.stream()>>>(C, ldc, A, ldA, Q, ldq, m, NA, NF);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}
template void devSpinGemm<float>(const B200Device &, float *, int, const float *, int, const signed char *, int, int, int, int);
template void devSpinGemm<double>(const B200Device &, double *, int, const double *, int, const signed char *, int, int, int, int);

template <class real> __device__ __forceinline__ real expR(real v);
template <> __device__ __forceinline__ float expR<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double expR<double>(double v) { return exp(v); }

/* phase 0: even trotters (< m2), 1: trotter m-1 (odd m), 2: odd trotters, 3: every trotter (SA) */
template <class real, bool SQA>
__global__ void bgFlipKernel(signed char *Q, int ldq, const real *dEmat, int ldc, const real *h, int NA, int m, int phase,
                             unsigned long long seed, unsigned long long step, unsigned domain, real twoDivM, real coef, real beta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int y;
    if (phase == 0) y = 2 * blockIdx.y;
    else if (phase == 1) y = m - 1;
    else if (phase == 2) y = 2 * blockIdx.y + 1;
    else y = blockIdx.y;
    if (i >= NA || y >= m) return;
    signed char *row = Q + (size_t)y * ldq;
    const real q = (real)row[i];
    real dE;
    if (SQA) {
        dE = twoDivM * q * (h[i] + dEmat[(size_t)y * ldc + i]);
        const int n0 = (y + m - 1) % m, n1 = (y + 1) % m;
        const real nb = (real)((int)Q[(size_t)n0 * ldq + i] + (int)Q[(size_t)n1 * ldq + i]);
        dE -= q * nb * coef;
    } else {
        dE = real(2) * q * (h[i] + dEmat[(size_t)y * ldc + i]);
    }
    const real thr = (dE < real(0)) ? real(1) : expR<real>(-dE * beta);
    const Philox4 p = sqbPhilox(seed, step, domain, (uint32_t)i, (uint32_t)y);
    if (thr > philoxUniform<real>(p)) row[i] = (signed char)(-row[i]);
}

/* One launch per half step: a block owns blockDim.x (4..32) spin indices of EVERY trotter, so the phases of the reference order (even
 * trotters, [trotter m-1 of an odd ring], odd trotters) are separated by __syncthreads() instead of kernel boundaries.
 * Same Philox stream and arithmetic as bgFlipKernel.  When `qbf` is given the new spins are also written as bf16
 * ([.][ldbf]), the operand layout of the tcgen05 contraction of the next half step. */
enum { BGF_THREADS = 1024 };
template <class real, bool SQA>
__global__ void __launch_bounds__(BGF_THREADS)
bgFlipFusedKernel(signed char *Q, int ldq, const real *dEmat, int ldc, const real *h, int NA, int m, unsigned long long seed,
                  unsigned long long step, unsigned domain, real twoDivM, real coef, real beta, unsigned short *qbf, int ldbf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, rowLanes = blockDim.y;
    const bool ok = i < NA;
    const real hi = ok ? h[i] : real(0);
    auto attempt = [&](int y) {
        if (!ok) return;
        signed char *row = Q + (size_t)y * ldq;
        signed char qi = row[i];
        const real q = (real)qi;
        real dE;
        if (SQA) {
            dE = twoDivM * q * (hi + dEmat[(size_t)y * ldc + i]);
            const int n0 = (y + m - 1) % m, n1 = (y + 1) % m;
            const real nb = (real)((int)Q[(size_t)n0 * ldq + i] + (int)Q[(size_t)n1 * ldq + i]);
            dE -= q * nb * coef;
        } else {
            dE = real(2) * q * (hi + dEmat[(size_t)y * ldc + i]);
        }
        const real thr = (dE < real(0)) ? real(1) : expR<real>(-dE * beta);
        const Philox4 p = sqbPhilox(seed, step, domain, (uint32_t)i, (uint32_t)y);
        if (thr > philoxUniform<real>(p)) {
            qi = (signed char)(-qi);
            row[i] = qi;
        }
        if (qbf) qbf[(size_t)y * ldbf + i] = (qi > 0) ? (unsigned short)0x3f80 : (unsigned short)0xbf80; /* +-1 in bf16 */
    };
    if (!SQA) {
        for (int y = threadIdx.y; y < m; y += rowLanes) attempt(y);
        return;
    }
    const int m2 = (m / 2) * 2;
    for (int y = 2 * threadIdx.y; y < m2; y += 2 * rowLanes) attempt(y);
    __syncthreads();
    if ((m & 1) && threadIdx.y == 0) attempt(m - 1);
    __syncthreads();
    for (int y = 2 * threadIdx.y + 1; y < m2; y += 2 * rowLanes) attempt(y);
}

template <class real> __global__ void transposeMatKernel(real *T, int ldT, const real *A, int ldA, int rows, int cols) {
    __shared__ real tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
    if (r < rows && c < cols) tile[threadIdx.y][threadIdx.x] = A[(size_t)r * ldA + c];
    __syncthreads();
    int tr = blockIdx.x * 32 + threadIdx.y, tc = blockIdx.y * 32 + threadIdx.x;
    if (tr < cols && tc < rows) T[(size_t)tr * ldT + tc] = tile[threadIdx.x][threadIdx.y];
}

/* =====================================================================================
 * host class
 * ===================================================================================== */
template <class real> class B200BipartiteGraphAnnealer : public sq::cuda::BipartiteGraphAnnealer<real> {
    typedef sq::MatrixType<real> HostMatrix;
    typedef sq::VectorType<real> HostVector;
    typedef sq::BipartiteGraphAnnealer<real> Base;
    typedef B200BipartiteGraphAnnealer<real> This;
    using Base::N0_; using Base::N1_; using Base::m_; using Base::om_; using Base::algo_;

public:
    B200BipartiteGraphAnnealer() : dev_(NULL), ldJ_(0), ldJT_(0), ldq0_(0), ldq1_(0), ldc_(0), c_(0), seed_(0), step_(0), randomizeCount_(0) {
        m_ = -1;
        selectAlgorithm(sq::algoDefault);
    }
    void assignDevice(sq::cuda::Device &device) {
        sqb_throwErrorIf(dev_ != NULL, "Device assigned more than once.");
        dev_ = &asB200(device);
    }
    sq::Algorithm selectAlgorithm(sq::Algorithm algo) { /* CUDABipartiteGraphAnnealer.cu:102-113 */
        switch (algo) {
        case sq::algoColoring:
        case sq::algoSAColoring:
            algo_ = algo;
            break;
        default:
            this->selectDefaultAlgorithm(algo, sq::algoColoring, sq::algoSAColoring);
            break;
        }
        return algo_;
    }
    void seed(unsigned long long seed) {
        sqb_throwErrorIf(dev_ == NULL, "Device not set.");
        seed_ = seed; step_ = 0; randomizeCount_ = 0;
        this->setState(Base::solRandSeedGiven);
    }
    void allocProblem() {
        ldJ_ = sq::roundUp(N0_, 32);
        ldJT_ = sq::roundUp(N1_, 32);
        dJ_.alloc(dev_, (size_t)N1_ * ldJ_);
        dJT_.alloc(dev_, (size_t)N0_ * ldJT_);
        dh0_.alloc(dev_, N0_);
        dh1_.alloc(dev_, N1_);
    }
    void makeTranspose() {
        dim3 grid((N0_ + 31) / 32, (N1_ + 31) / 32);
        transposeMatKernel<real><<<grid, dim3(32, 32), 0, dev_->stream()>>>(dJT_.p, ldJT_, dJ_.p, ldJ_, N1_, N0_);
        CUDA_CHECK(cudaGetLastError());
        ++dev_->launchCount;
        /* fp32: bf16 hi/mid/lo splits of J and J^T for the tcgen05 contraction (energy_tc.cu) */
        tcJ_.ready = tcJT_.ready = false;
        qbfValid_[0] = qbfValid_[1] = false;
        if constexpr (std::is_same<real, float>::value) {
            if (tcEnabled()) {
                tcPrepareOperand(*dev_, tcJ_, dJ_.p, ldJ_, N1_, N0_);
                tcPrepareOperand(*dev_, tcJT_, dJT_.p, ldJT_, N0_, N1_);
            }
        }
    }
    void setQUBO(const HostVector &b0, const HostVector &b1, const HostMatrix &W, sq::OptimizeMethod om = sq::optMinimize) {
        sqb_throwErrorIf(W.cols != b0.size || W.rows != b1.size, "%s, shape mismatch between b0, b1 and W.", __func__);
        sqb_throwErrorIf(dev_ == NULL, "Device not set.");
        this->clearState(Base::solProblemSet);
        N0_ = b0.size; N1_ = b1.size;
        m_ = (N0_ + N1_) / 4;
        om_ = om;
        allocProblem();
        DevBuf<real> dW, db0, db1, dc;
        dW.alloc(dev_, (size_t)N1_ * ldJ_); db0.alloc(dev_, N0_); db1.alloc(dev_, N1_); dc.alloc(dev_, 1);
        dev_->h2d2D(dW.p, sizeof(real) * ldJ_, W.data, sizeof(real) * W.stride, sizeof(real) * N0_, N1_);
        dev_->h2d(db0.p, b0.data, sizeof(real) * N0_);
        dev_->h2d(db1.p, b1.data, sizeof(real) * N1_);
        /* maximize: b0, b1, W negated before the conversion (CUDABipartiteGraphAnnealer.cu:135-142) */
        devBipartiteHamiltonian<real>(*dev_, dh0_.p, dh1_.p, dJ_.p, ldJ_, dc.p, db0.p, db1.p, dW.p, ldJ_, N0_, N1_,
                                      om == sq::optMaximize ? real(-1) : real(1));
        makeTranspose();
        dev_->d2h(&c_, dc.p, sizeof(real));
        dev_->synchronize();
        this->setState(Base::solProblemSet);
    }
    void setHamiltonian(const HostVector &h0, const HostVector &h1, const HostMatrix &J, real c = real(0.)) {
        sqb_throwErrorIf(J.cols != h0.size || J.rows != h1.size, "%s, shape mismatch between h0, h1 and J.", __func__);
        sqb_throwErrorIf(dev_ == NULL, "Device not set.");
        this->clearState(Base::solProblemSet);
        N0_ = h0.size; N1_ = h1.size;
        m_ = (N0_ + N1_) / 4;
        om_ = sq::optMinimize;
        c_ = c;
        allocProblem();
        dev_->h2d2D(dJ_.p, sizeof(real) * ldJ_, J.data, sizeof(real) * J.stride, sizeof(real) * N0_, N1_);
        dev_->h2d(dh0_.p, h0.data, sizeof(real) * N0_);
        dev_->h2d(dh1_.p, h1.data, sizeof(real) * N1_);
        makeTranspose();
        dev_->synchronize();
        this->setState(Base::solProblemSet);
    }
    void getHamiltonian(HostVector *h0, HostVector *h1, HostMatrix *J, real *c) const {
        this->throwErrorIfProblemNotSet();
        h0->resize(N0_); h1->resize(N1_); J->resize(N1_, N0_);
        dev_->d2h(h0->data, dh0_.p, sizeof(real) * N0_);
        dev_->d2h(h1->data, dh1_.p, sizeof(real) * N1_);
        dev_->d2h2D(J->data, sizeof(real) * J->stride, dJ_.p, sizeof(real) * ldJ_, sizeof(real) * N0_, N1_);
        dev_->synchronize();
        *c = c_;
    }
    sq::Preferences getPreferences() const {
        sq::Preferences prefs = Base::getPreferences();
        prefs.pushBack(sq::Preference(sq::pnDevice, "cuda"));
        return prefs;
    }
    void prepare() {
        this->throwErrorIfProblemNotSet();
        sqb_throwErrorIf(m_ <= 0, "# trotters must be a positive integer.");
        if (!this->isRandSeedGiven()) seed((unsigned long long)time(NULL));
        this->setState(Base::solRandSeedGiven);
        if (m_ == 1) this->selectDefaultSAAlgorithm(algo_, sq::algoSAColoring);
        ldq0_ = sq::roundUp(N0_, 16);
        ldq1_ = sq::roundUp(N1_, 16);
        ldc_ = sq::roundUp(std::max(N0_, N1_), 32);
        dq0_.alloc(dev_, (size_t)m_ * ldq0_);
        dq1_.alloc(dev_, (size_t)m_ * ldq1_);
        ddE_.alloc(dev_, (size_t)m_ * ldc_);
        dE_.alloc(dev_, m_);
        E_.resize(m_);
        eBack_.alloc(dev_, m_);
        hq0_.assign((size_t)m_ * ldq0_, 0);
        hq1_.assign((size_t)m_ * ldq1_, 0);
        xPairs_.clear();
        qPairs_.clear();
        qbfValid_[0] = qbfValid_[1] = false;
        this->setState(Base::solPrepared);
    }
    void randomizeSpin() {
        this->throwErrorIfNotPrepared();
        launchRandomizeSpin(*dev_, dq0_.p, ldq0_, N0_, m_, seed_, randomizeCount_, DOM_RANDOMIZE);
        launchRandomizeSpin(*dev_, dq1_.p, ldq1_, N1_, m_, seed_, randomizeCount_, DOM_RANDOMIZE1);
        ++randomizeCount_;
        qbfValid_[0] = qbfValid_[1] = false;
        this->setState(Base::solQSet);
    }
    void uploadSpins() {
        dev_->h2d(dq0_.p, hq0_.data(), hq0_.size());
        dev_->h2d(dq1_.p, hq1_.data(), hq1_.size());
        dev_->synchronize();
        qbfValid_[0] = qbfValid_[1] = false;
    }
    void set_q(const sq::BitSetPair &qPair) {
        sqb_throwErrorIf(qPair.bits0.size != N0_ || qPair.bits1.size != N1_, "Dimension of q0/q1 does not match N0/N1.");
        this->throwErrorIfNotPrepared();
        for (int y = 0; y < m_; ++y) {
            memcpy(&hq0_[(size_t)y * ldq0_], qPair.bits0.data, N0_);
            memcpy(&hq1_[(size_t)y * ldq1_], qPair.bits1.data, N1_);
        }
        uploadSpins();
        this->setState(Base::solQSet);
    }
    void set_qset(const sq::BitSetPairArray &qPairs) {
        sqb_throwErrorIf(qPairs.size() == 0, "empty q set.");
        for (int i = 0; i < qPairs.size(); ++i)
            sqb_throwErrorIf(qPairs[i].bits0.size != N0_ || qPairs[i].bits1.size != N1_, "Dimension of q0/q1 does not match N0/N1.");
        m_ = qPairs.size();
        prepare(); /* CUDABipartiteGraphAnnealer.cu:219-220 */
        for (int y = 0; y < m_; ++y) {
            memcpy(&hq0_[(size_t)y * ldq0_], qPairs[y].bits0.data, N0_);
            memcpy(&hq1_[(size_t)y * ldq1_], qPairs[y].bits1.data, N1_);
        }
        uploadSpins();
        this->setState(Base::solQSet);
    }
    const HostVector &get_E() const {
        if (!this->isEAvailable()) const_cast<This *>(this)->calculate_E();
        const_cast<This *>(this)->eBack_.wait(const_cast<This *>(this)->E_.data, (size_t)m_);
        return E_;
    }
    const sq::BitSetPairArray &get_x() const {
        if (!this->isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
        return xPairs_;
    }
    const sq::BitSetPairArray &get_q() const {
        if (!this->isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
        return qPairs_;
    }
    void calculate_E() { calculateEnergy(); }
    void makeSolution() {
        this->throwErrorIfQNotSet();
        syncBits();
        this->setState(Base::solSolutionAvailable);
        calculateEnergy();
    }
    real getSystemE(real G, real beta) const {
        This *self = const_cast<This *>(this);
        self->calculateEnergy();
        self->eBack_.wait(self->E_.data, (size_t)m_);
        real E = E_.sum() / m_;
        if (sq::isSQAAlgorithm(algo_)) {
            real spinDotSum = (real)(ringSpinDot(*dev_, dq0_.p, ldq0_, N0_, m_) + ringSpinDot(*dev_, dq1_.p, ldq1_, N1_, m_));
            real coef = real(0.5) / beta * std::log(std::tanh(G * beta / m_));
            E -= spinDotSum * coef;
        }
        if (om_ == sq::optMaximize) E *= real(-1.);
        return E;
    }
    void annealOneStep(real G, real beta) {
        this->throwErrorIfQNotSet();
        this->clearState(Base::solSolutionAvailable);
        const bool sqa = (algo_ == sq::algoColoring);
        real twoDivM = real(2.) / real(m_), coef = real(0), b = beta;
        if (sqa) coef = std::log(std::tanh(G * beta / m_)) / beta;
        else b = real(1.) / G; /* annealOneStep(kT, _) */
        halfStep(1, sqa, twoDivM, coef, b);
        halfStep(0, sqa, twoDivM, coef, b);
        ++step_;
    }

private:
    void calculateEnergy() {
        this->throwErrorIfQNotSet();
        const real sign = (om_ == sq::optMaximize) ? real(-1) : real(1);
        /* E = -c - h0.q0 - h1.q1 - q1^T J q0 */
        bool done = false;
        if constexpr (std::is_same<real, float>::value) {
            if (tcJ_.ready && tcEnabled()) {
                tcBatchedEnergy(*dev_, dE_.p, tcJ_, dq0_.p, ldq0_, dq1_.p, ldq1_, dh1_.p, dh0_.p, m_, -sign, -sign * c_, tcWs_);
                done = true;
            }
        }
        if (!done)
            devBatchedEnergy<real>(*dev_, dE_.p, dJ_.p, ldJ_, N1_, N0_, dq0_.p, ldq0_, dq1_.p, ldq1_, dh1_.p, dh0_.p, m_, -sign, -sign * c_);
        eBack_.enqueue(dE_.p, (size_t)m_); /* asynchronous: get_E() waits for this copy's own event */
        this->setState(Base::solEAvailable);
    }
    void halfStep(int side, bool sqa, real twoDivM, real coef, real beta) {
        signed char *qA = side ? dq1_.p : dq0_.p;
        const signed char *qF = side ? dq0_.p : dq1_.p;
        const int ldqA = side ? ldq1_ : ldq0_, ldqF = side ? ldq0_ : ldq1_;
        const int NA = side ? N1_ : N0_, NF = side ? N0_ : N1_;
        const real *Je = side ? dJ_.p : dJT_.p;
        const int ldJe = side ? ldJ_ : ldJT_;
        const real *h = side ? dh1_.p : dh0_.p;
        const unsigned dom = side ? DOM_BG_SIDE1 : DOM_BG_SIDE0;
        bool done = false;
        unsigned short *bfOut = NULL; /* bf16 copy of the side being updated: operand of the next half step's contraction */
        int ldbf = 0;
        if constexpr (std::is_same<real, float>::value) {
            const TcOperand &op = side ? tcJ_ : tcJT_;   /* contracts over the fixed side */
            const TcOperand &opNext = side ? tcJT_ : tcJ_; /* the next half step contracts over this side */
            if (op.ready && opNext.ready && tcEnabled()) {
                const int f = 1 - side;
                if (!qbfValid_[f]) { /* after set_q / randomize: rebuild; afterwards the flip kernel keeps it current */
                    tcWidenSpins(*dev_, qbf_[f], op, qF, ldqF, m_);
                    qbfValid_[f] = true;
                }
                tcSpinGemmBf16(*dev_, ddE_.p, ldc_, op, qbf_[f].p, m_);
                const size_t need = (size_t)sq::roundUp(m_, 128) * opNext.Kp;
                if (qbf_[side].n < need || qbf_[side].dev != dev_) { qbf_[side].alloc(dev_, need); qbfValid_[side] = false; }
                bfOut = qbf_[side].p;
                ldbf = opNext.Kp;
                done = true;
            }
        }
        if (!done) devSpinGemm<real>(*dev_, ddE_.p, ldc_, Je, ldJe, qF, ldqF, m_, NA, NF);
        cudaStream_t st = dev_->stream();
        int cols = 32; /* spin indices per block: fewer when the side is short, so that the grid still covers the device */
        while (cols > 4 && (NA + cols - 1) / cols < dev_->numSMs()) cols >>= 1;
        const dim3 grid((NA + cols - 1) / cols), block(cols, BGF_THREADS / cols);
        if (!sqa)
            bgFlipFusedKernel<real, false><<<grid, block, 0, st>>>(qA, ldqA, ddE_.p, ldc_, h, NA, m_, seed_, step_, dom, twoDivM, coef, beta,
                                                                   bfOut, ldbf);
        else
            bgFlipFusedKernel<real, true><<<grid, block, 0, st>>>(qA, ldqA, ddE_.p, ldc_, h, NA, m_, seed_, step_, dom, twoDivM, coef, beta,
                                                                  bfOut, ldbf);
        ++dev_->launchCount;
        CUDA_CHECK(cudaGetLastError());
        if (bfOut) qbfValid_[side] = true; /* every spin of this side has just been rewritten */
    }
    void syncBits() {
        xPairs_.clear();
        qPairs_.clear();
        dev_->d2h(hq0_.data(), dq0_.p, hq0_.size());
        dev_->d2h(hq1_.data(), dq1_.p, hq1_.size());
        dev_->synchronize();
        for (int y = 0; y < m_; ++y) {
            sq::BitSet q0(N0_), q1(N1_), x0(N0_), x1(N1_);
            for (int i = 0; i < N0_; ++i) { char v = hq0_[(size_t)y * ldq0_ + i]; q0(i) = v; x0(i) = (char)((v + 1) / 2); }
            for (int i = 0; i < N1_; ++i) { char v = hq1_[(size_t)y * ldq1_ + i]; q1(i) = v; x1(i) = (char)((v + 1) / 2); }
            qPairs_.pushBack(sq::BitSetPair(q0, q1));
            xPairs_.pushBack(sq::BitSetPair(x0, x1));
        }
    }

    B200Device *dev_;
    DevBuf<real> dJ_, dJT_, dh0_, dh1_, ddE_, dE_;
    AsyncReadback<real> eBack_; /* energies: pinned landing buffer + completion event */
    DevBuf<signed char> dq0_, dq1_;
    int ldJ_, ldJT_, ldq0_, ldq1_, ldc_;
    real c_;
    unsigned long long seed_, step_, randomizeCount_;
    HostVector E_;
    std::vector<signed char> hq0_, hq1_;
    sq::BitSetPairArray xPairs_, qPairs_;
    TcOperand tcJ_, tcJT_;
    TcWorkspace tcWs_;
    DevBuf<unsigned short> qbf_[2]; /* bf16 copies of the two spin matrices (tcgen05 operand layout), kept current by the flip kernel */
    bool qbfValid_[2] = {false, false};
};

} // namespace sqb

namespace sqaod { namespace cuda {
template <> BipartiteGraphAnnealer<float> *newBipartiteGraphAnnealer<float>() { return new sqb::B200BipartiteGraphAnnealer<float>(); }
template <> BipartiteGraphAnnealer<double> *newBipartiteGraphAnnealer<double>() { return new sqb::B200BipartiteGraphAnnealer<double>(); }
}} // namespace sqaod::cuda
