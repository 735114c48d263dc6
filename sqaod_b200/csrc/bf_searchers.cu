/* bf_searchers.cu -- dense-graph and bipartite-graph brute-force searchers for B200 (sm_100a).
 *
 * Replaces CUDADenseGraphBFSearcher.cpp:52-185 + DeviceDenseGraphBatchSearch.cu:42-187 and
 * CUDABipartiteGraphBFSearcher.cpp:96-211 + DeviceBipartiteGraphBatchSearch.cu:38-217.  The reference expands every
 * packed x of a tile into a tile x N 0/1 matrix, runs a GEMM (2N^2+2N flop per state), a CUB min-reduce and a CUB
 * select, with a host sync per 2^18 states.  Here x never leaves registers:
 *     E(x) = x^T W x  is split over the bit index as  E = E_H(high bits) + E_L(low L bits) + sum_{b<L} x_b c_b(high),
 * a thread owns one value of the high bits and walks its 2^L low-bit states in Gray-code order, so one state costs
 *     A += +-c_b (one bit flips)  ;  E = A + E_L[s] (table shared by every thread)  ;  Emin = min(Emin, E)
 * -- three FP64 operations.  All arithmetic is double for both solver precisions: on inputs whose partial sums are
 * exactly representable (integers, the reference tests' 2^-14 grid) the result equals the CPU searcher's bit for bit
 * (CPUDenseGraphBatchSearch.cpp:25-50); on other inputs it is the correctly rounded minimum rather than a
 * summation-order-dependent one.  Solution lists follow the CPU searcher: all argmins, ascending, capped
 * (CPUDenseGraphBFSearcher.cpp:103-131).  The bipartite searcher maps (x0, x1) onto the same engine through the
 * symmetric form  E = z^T [[diag b1, W/2],[W^T/2, diag b0]] z  with a bit order that makes every 2-D tile a contiguous
 * range.
 */
#include "device.hpp"
#include "kernels_common.cuh"
#include "b200_solvers.hpp"
#include <float.h>
#include <math.h>
#include <algorithm>
#include <vector>

namespace sqb {

enum { BF_MAX_L = 13, BF_MAX_M = 9, BF_COLLECT_CAP = 1 << 17 };

struct BFParams {
    const double *Wb; /* [N][N], indexed by kernel bit */
    const double *EL; /* [2^L] E_L(gray(s)) */
    int N, L, M;
    unsigned long long rBegin, rEnd, spanBase; /* spanBase = rBegin >> (L+M) */
    double *ctaMin;
    unsigned long long *globalMin;
    double target;
    unsigned long long *outX;
    unsigned int *outCount;
    unsigned int outCap;
};

__host__ __device__ __forceinline__ unsigned long long orderedKey(double v) {
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double fromOrderedKey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

#define BF_VISIT(SIDX)                                                                        \
    {                                                                                         \
        const double E_ = A + EL[(SIDX)];                                                     \
        if (COLLECT) {                                                                        \
            if (E_ == P.target) {                                                             \
                unsigned int at = atomicAdd(P.outCount, 1u);                                  \
                unsigned long long s_ = (unsigned long long)(SIDX);                           \
                if (at < P.outCap) P.outX[at] = xBase | (s_ ^ (s_ >> 1));                     \
            }                                                                                 \
        } else                                                                                \
            Emin = (E_ < Emin) ? E_ : Emin; /* energies are finite: no NaN handling needed */   \
    }
#define BF_FLIP(D) { A += D; D = -D; }
/* bits 0..2 of the Gray code are 0 at every multiple of 16, so inside a 16-block their flips alternate 0->1, 1->0 */
#define BF_SET(D) { A += D; }
#define BF_CLR(D) { A -= D; }

template <bool COLLECT> __global__ void __launch_bounds__(512, 2) bfKernel(BFParams P) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int N = P.N, L = P.L, M = P.M;
    double *Wb = reinterpret_cast<double *>(smemRaw);
    double *EL = Wb + N * N;
    double *cH = EL + (1 << L); /* [L+M] + [1]: influence of the CTA's high bits, and their own energy */
    __shared__ double redMin[16];

    const unsigned long long span = P.spanBase + blockIdx.x;
    const unsigned long long hh = span << (L + M);
    if (COLLECT) {
        if (P.ctaMin[blockIdx.x] != P.target) return;
    }
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) Wb[i] = P.Wb[i];
    for (int i = threadIdx.x; i < (1 << L); i += blockDim.x) EL[i] = P.EL[i];
    __syncthreads();
    if (threadIdx.x < L + M) {
        const int b = threadIdx.x;
        double s = 0;
        for (int a = L + M; a < N; ++a)
            if ((hh >> a) & 1ull) s += Wb[a * N + b];
        cH[b] = 2.0 * s;
    } else if (threadIdx.x == L + M) {
        double s = 0;
        for (int a = L + M; a < N; ++a) {
            if (!((hh >> a) & 1ull)) continue;
            s += Wb[a * N + a];
            for (int a2 = a + 1; a2 < N; ++a2)
                if ((hh >> a2) & 1ull) s += 2.0 * Wb[a * N + a2];
        }
        cH[L + M] = s;
    }
    __syncthreads();

    double Emin = DBL_MAX;
    const unsigned int mid = threadIdx.x;
    if (mid < (1u << M)) {
        const unsigned long long xBase = hh | ((unsigned long long)mid << L);
        const unsigned long long xLast = xBase + ((1ull << L) - 1ull);
        if (xLast >= P.rBegin && xBase < P.rEnd) {
            /* this thread's constant part and the coefficients of its L low bits */
            double A = cH[L + M];
            for (int a = 0; a < M; ++a) {
                if (!((mid >> a) & 1u)) continue;
                const int ba = L + a;
                A += Wb[ba * N + ba] + cH[ba];
                for (int a2 = a + 1; a2 < M; ++a2)
                    if ((mid >> a2) & 1u) A += 2.0 * Wb[ba * N + (L + a2)];
            }
            double d[BF_MAX_L];
#pragma unroll
            for (int b = 0; b < BF_MAX_L; ++b) {
                double c = 0;
                if (b < L) {
                    c = cH[b];
                    for (int a = 0; a < M; ++a)
                        if ((mid >> a) & 1u) c += 2.0 * Wb[(L + a) * N + b];
                }
                d[b] = c;
            }
            const bool whole = (xBase >= P.rBegin) && (xLast < P.rEnd);
            if (whole && L >= 4) {
                const int nBlocks = 1 << (L - 4);
                for (int q = 0; q < nBlocks; ++q) {
                    const int s0 = q << 4;
                    BF_VISIT(s0 + 0)  BF_SET(d[0])
                    BF_VISIT(s0 + 1)  BF_SET(d[1])
                    BF_VISIT(s0 + 2)  BF_CLR(d[0])
                    BF_VISIT(s0 + 3)  BF_SET(d[2])
                    BF_VISIT(s0 + 4)  BF_SET(d[0])
                    BF_VISIT(s0 + 5)  BF_CLR(d[1])
                    BF_VISIT(s0 + 6)  BF_CLR(d[0])
                    BF_VISIT(s0 + 7)  BF_FLIP(d[3])
                    BF_VISIT(s0 + 8)  BF_SET(d[0])
                    BF_VISIT(s0 + 9)  BF_SET(d[1])
                    BF_VISIT(s0 + 10) BF_CLR(d[0])
                    BF_VISIT(s0 + 11) BF_CLR(d[2])
                    BF_VISIT(s0 + 12) BF_SET(d[0])
                    BF_VISIT(s0 + 13) BF_CLR(d[1])
                    BF_VISIT(s0 + 14) BF_CLR(d[0])
                    BF_VISIT(s0 + 15)
                    if (q + 1 < nBlocks) {
                        switch (__ffs(q + 1) + 3) { /* bit flipped by the step 16q+15 -> 16(q+1) */
                        case 4: BF_FLIP(d[4]) break;
                        case 5: BF_FLIP(d[5]) break;
                        case 6: BF_FLIP(d[6]) break;
                        case 7: BF_FLIP(d[7]) break;
                        case 8: BF_FLIP(d[8]) break;
                        case 9: BF_FLIP(d[9]) break;
                        case 10: BF_FLIP(d[10]) break;
                        case 11: BF_FLIP(d[11]) break;
                        default: BF_FLIP(d[12]) break;
                        }
                    }
                }
            } else {
                /* short rows (L < 4) and the partial rows at the edges of the range: same walk, one state at a time */
                const int nStates = 1 << L;
                for (int s = 0; s < nStates; ++s) {
                    const unsigned long long x = xBase | (unsigned long long)(s ^ (s >> 1));
                    if (x >= P.rBegin && x < P.rEnd) BF_VISIT(s)
                    if (s + 1 < nStates) {
                        const int k = __ffs(s + 1) - 1;
#pragma unroll
                        for (int b = 0; b < BF_MAX_L; ++b)
                            if (b == k) BF_FLIP(d[b])
                    }
                }
            }
        }
    }
    if (!COLLECT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Emin = fmin(Emin, __shfl_xor_sync(0xffffffffu, Emin, o));
        if (laneId() == 0) redMin[threadIdx.x >> 5] = Emin;
        __syncthreads();
        if (threadIdx.x == 0) {
            double e = DBL_MAX;
            for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) e = fmin(e, redMin[w]);
            P.ctaMin[blockIdx.x] = e;
            atomicMin(P.globalMin, orderedKey(e));
        }
    }
}

/* ---- the engine shared by both searchers: works on "kernel indices" r in [0, 2^N) ---- */
class BFEngine {
public:
    BFEngine() : dev_(NULL), N_(0), L_(0), M_(0), threads_(32), smem_(0), maxSpans_(0) {}
    void assign(B200Device *dev) { dev_ = dev; }
    B200Device *device() const { return dev_; }
    /* Wk: symmetric N x N in kernel bit order (entry [a][b] couples kernel bits a and b) */
    void setProblem(const std::vector<double> &Wk, int N) {
        N_ = N;
        M_ = std::min((int)BF_MAX_M, N);
        L_ = std::min((int)BF_MAX_L, std::max(0, N - M_ - 8));
        threads_ = std::max(32, 1 << M_);
        std::vector<double> EL((size_t)1 << L_);
        for (unsigned s = 0; s < (1u << L_); ++s) {
            unsigned g = s ^ (s >> 1);
            double e = 0;
            for (int b = 0; b < L_; ++b) {
                if (!((g >> b) & 1u)) continue;
                e += Wk[(size_t)b * N + b];
                for (int b2 = b + 1; b2 < L_; ++b2)
                    if ((g >> b2) & 1u) e += 2.0 * Wk[(size_t)b * N + b2];
            }
            EL[s] = e;
        }
        dW_.alloc(dev_, (size_t)N * N);
        dEL_.alloc(dev_, EL.size());
        dev_->h2d(dW_.p, Wk.data(), sizeof(double) * N * N);
        dev_->h2d(dEL_.p, EL.data(), sizeof(double) * EL.size());
        dGlobalMin_.alloc(dev_, 1);
        dCount_.alloc(dev_, 1);
        dOut_.alloc(dev_, BF_COLLECT_CAP);
        smem_ = sizeof(double) * ((size_t)N * N + ((size_t)1 << L_) + L_ + M_ + 1);
        CUDA_CHECK(cudaFuncSetAttribute(bfKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));
        CUDA_CHECK(cudaFuncSetAttribute(bfKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_));
        dev_->synchronize();
    }
    /* minimum over r in [rBegin, rEnd) */
    double rangeMin(unsigned long long rBegin, unsigned long long rEnd) {
        BFParams P = params(rBegin, rEnd);
        const unsigned long long key = orderedKey(DBL_MAX);
        dev_->h2d(dGlobalMin_.p, &key, sizeof(key));
        bfKernel<false><<<spans(rBegin, rEnd), threads_, smem_, dev_->stream()>>>(P);
        CUDA_CHECK(cudaGetLastError());
        ++dev_->launchCount;
        unsigned long long got = 0;
        dev_->d2h(&got, dGlobalMin_.p, sizeof(got));
        dev_->synchronize();
        return fromOrderedKey(got);
    }
    /* all r in the range of the last rangeMin() call with E(r) == target, ascending; stops once `want` are found */
    void collect(unsigned long long rBegin, unsigned long long rEnd, double target, size_t want, std::vector<unsigned long long> *out) {
        std::vector<unsigned long long> found;
        collectRec(rBegin, rEnd, rBegin, rEnd, target, want, &found);
        std::sort(found.begin(), found.end());
        for (size_t i = 0; i < found.size() && out->size() < want + 0; ++i) out->push_back(found[i]);
    }

private:
    unsigned int spans(unsigned long long rBegin, unsigned long long rEnd) const {
        const int sh = L_ + M_;
        return (unsigned int)(((rEnd - 1) >> sh) - (rBegin >> sh) + 1);
    }
    BFParams params(unsigned long long rBegin, unsigned long long rEnd) {
        BFParams P;
        P.Wb = dW_.p; P.EL = dEL_.p; P.N = N_; P.L = L_; P.M = M_;
        P.rBegin = rBegin; P.rEnd = rEnd; P.spanBase = rBegin >> (L_ + M_);
        unsigned int n = spans(rBegin, rEnd);
        if (n > maxSpans_) { dCtaMin_.alloc(dev_, n); maxSpans_ = n; }
        P.ctaMin = dCtaMin_.p; P.globalMin = dGlobalMin_.p;
        P.target = 0; P.outX = dOut_.p; P.outCount = dCount_.p; P.outCap = BF_COLLECT_CAP;
        return P;
    }
    /* [fullBegin, fullEnd) is the range whose per-CTA minima are in dCtaMin_; [b, e) the sub-range to gather from */
    void collectRec(unsigned long long fullBegin, unsigned long long fullEnd, unsigned long long b, unsigned long long e, double target,
                    size_t want, std::vector<unsigned long long> *found) {
        if (found->size() >= want || b >= e) return;
        BFParams P;
        P.Wb = dW_.p; P.EL = dEL_.p; P.N = N_; P.L = L_; P.M = M_;
        P.rBegin = b; P.rEnd = e; P.spanBase = b >> (L_ + M_);
        P.ctaMin = dCtaMin_.p + ((b >> (L_ + M_)) - (fullBegin >> (L_ + M_)));
        P.globalMin = dGlobalMin_.p; P.target = target; P.outX = dOut_.p; P.outCount = dCount_.p; P.outCap = BF_COLLECT_CAP;
        unsigned int zero = 0;
        dev_->h2d(dCount_.p, &zero, sizeof(zero));
        bfKernel<true><<<spans(b, e), threads_, smem_, dev_->stream()>>>(P);
        CUDA_CHECK(cudaGetLastError());
        ++dev_->launchCount;
        unsigned int count = 0;
        dev_->d2h(&count, dCount_.p, sizeof(count));
        dev_->synchronize();
        const unsigned long long spanSize = 1ull << (L_ + M_);
        if (count > BF_COLLECT_CAP && (e - b) > 1) {
            /* more ties than the gather buffer holds: halve the range, lower half first -- on a span boundary while the range covers
             * several spans, inside the span otherwise (the kernel takes arbitrary [rBegin, rEnd)), so that every leaf gathers ALL its
             * ties and the ascending, capped list is the lowest `want` states whatever the degeneracy (W = 0 included) */
            unsigned long long mid = b + (e - b) / 2;
            if ((e - b) > spanSize) {
                const unsigned long long midSpan = ((b >> (L_ + M_)) + ((e - 1) >> (L_ + M_)) + 1) / 2;
                if ((midSpan << (L_ + M_)) > b && (midSpan << (L_ + M_)) < e) mid = midSpan << (L_ + M_);
            }
            collectRec(fullBegin, fullEnd, b, mid, target, want, found);
            collectRec(fullBegin, fullEnd, mid, e, target, want, found);
            return;
        }
        unsigned int n = std::min(count, (unsigned int)BF_COLLECT_CAP);
        size_t at = found->size();
        found->resize(at + n);
        dev_->d2h(found->data() + at, dOut_.p, sizeof(unsigned long long) * n);
        dev_->synchronize();
    }

    B200Device *dev_;
    int N_, L_, M_, threads_;
    size_t smem_;
    unsigned int maxSpans_;
    DevBuf<double> dW_, dEL_, dCtaMin_;
    DevBuf<unsigned long long> dGlobalMin_, dOut_;
    DevBuf<unsigned int> dCount_;
};

/* =====================================================================================
 * dense-graph searcher
 * ===================================================================================== */
template <class real>
class B200DenseGraphBFSearcher : public sq::cuda::DenseGraphBFSearcher<real>, public DenseBFExtras {
    typedef sq::MatrixType<real> Matrix;
    typedef sq::VectorType<real> Vector;
    typedef sq::DenseGraphBFSearcher<real> Base;
    typedef B200DenseGraphBFSearcher<real> This;
    using Base::N_; using Base::om_; using Base::tileSize_; using Base::x_; using Base::xMax_;

public:
    B200DenseGraphBFSearcher() : Emin_(DBL_MAX), rangeBegin_(0), rangeEnd_(0), rangeSet_(false) { tileSize_ = 1 << 30; }
    void assignDevice(sq::cuda::Device &device) {
        sqb_throwErrorIf(engine_.device() != NULL, "Device assigned more than once.");
        engine_.assign(&asB200(device));
    }
    void setQUBO(const Matrix &W, sq::OptimizeMethod om = sq::optMinimize) {
        sqb_throwErrorIf(W.rows != W.cols, "%s, W is not a sqare matrix.", __func__);
        sqb_throwErrorIf(!sq::isSymmetric(W), "%s, Matrix is not symmetric.", __func__);
        sqb_throwErrorIf(63 < W.rows, "N must be smaller than 64, N=%d.", W.rows);
        sqb_throwErrorIf(engine_.device() == NULL, "Device not set.");
        this->clearState(Base::solProblemSet);
        N_ = W.rows;
        om_ = om;
        const double sign = (om == sq::optMaximize) ? -1. : 1.; /* CUDADenseGraphBFSearcher.cpp:60-62: W is negated */
        Wk_.assign((size_t)N_ * N_, 0.);
        for (int a = 0; a < N_; ++a)      /* kernel bit a <-> variable N-1-a (Common.cpp:78-93, MSB first) */
            for (int b = 0; b < N_; ++b) Wk_[(size_t)a * N_ + b] = sign * (double)W(N_ - 1 - a, N_ - 1 - b);
        rangeSet_ = false;
        this->setState(Base::solProblemSet);
    }
    sq::Preferences getPreferences() const {
        sq::Preferences prefs = Base::getPreferences();
        prefs.pushBack(sq::Preference(sq::pnDevice, "cuda"));
        return prefs;
    }
    const Vector &get_E() const {
        if (!this->isEAvailable()) const_cast<This *>(this)->calculate_E();
        return E_;
    }
    const sq::BitSetArray &get_x() const {
        if (!this->isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
        return xList_;
    }
    void prepare() {
        this->throwErrorIfProblemNotSet();
        engine_.setProblem(Wk_, N_);
        Emin_ = DBL_MAX;
        packed_.clear();
        xList_.clear();
        xMax_ = 1ull << N_;
        if (!rangeSet_) { rangeBegin_ = 0; rangeEnd_ = xMax_; }
        rangeEnd_ = std::min(rangeEnd_, xMax_);
        x_ = rangeBegin_;
        if (xMax_ < (sq::PackedBitSet)tileSize_) {
            tileSize_ = (sq::SizeType)xMax_;
            sq::log("Tile size is adjusted to %d for N=%d", tileSize_, N_);
        }
        this->setState(Base::solPrepared);
    }
    size_t solutionCap() const { return (size_t)std::min((long long)tileSize_, 1ll << 16); }
    bool searchRange(sq::PackedBitSet *curXEnd) {
        this->throwErrorIfNotPrepared();
        this->clearState(Base::solSolutionAvailable);
        const sq::PackedBitSet b = x_, e = std::min(x_ + (sq::PackedBitSet)tileSize_, rangeEnd_);
        if (b < e) {
            const double tileMin = engine_.rangeMin(b, e);
            if (tileMin < Emin_) {
                Emin_ = tileMin;
                packed_.clear();
            }
            if (tileMin == Emin_ && packed_.size() < solutionCap()) engine_.collect(b, e, Emin_, solutionCap(), &packed_);
        }
        x_ = e;
        if (curXEnd != NULL) *curXEnd = x_;
        return x_ == rangeEnd_;
    }
    void calculate_E() {
        this->throwErrorIfNotPrepared();
        E_.resize(packed_.empty() ? 1 : (int)packed_.size());
        real v = (real)((om_ == sq::optMaximize) ? -Emin_ : Emin_);
        E_ = v;
        this->setState(Base::solEAvailable);
    }
    void makeSolution() {
        this->throwErrorIfNotPrepared();
        xList_.clear();
        std::sort(packed_.begin(), packed_.end());
        for (size_t i = 0; i < packed_.size() && i < solutionCap(); ++i) {
            sq::BitSet bits;
            sq::unpackBitSet(&bits, packed_[i], N_);
            xList_.pushBack(bits);
        }
        calculate_E();
        this->setState(Base::solSolutionAvailable);
    }
    /* ---- DenseBFExtras ---- */
    void setRange(sq::PackedBitSet xBegin, sq::PackedBitSet xEnd) {
        sqb_throwErrorIf(xEnd < xBegin, "invalid range.");
        rangeBegin_ = xBegin; rangeEnd_ = xEnd; rangeSet_ = true;
        this->clearState(Base::solPrepared);
    }
    double getEmin() const { return (om_ == sq::optMaximize) ? -Emin_ : Emin_; }
    const sq::PackedBitSetArray &packedSolutions() const {
        packedArr_.clear();
        for (size_t i = 0; i < packed_.size(); ++i) packedArr_.pushBack(packed_[i]);
        return packedArr_;
    }
    void setPackedSolutions(double Emin, const sq::PackedBitSet *x, int n) {
        this->throwErrorIfNotPrepared();
        Emin_ = (om_ == sq::optMaximize) ? -Emin : Emin;
        packed_.assign(x, x + n);
        this->clearState(Base::solSolutionAvailable);
    }

private:
    BFEngine engine_;
    std::vector<double> Wk_;
    double Emin_;
    std::vector<unsigned long long> packed_;
    mutable sq::PackedBitSetArray packedArr_;
    sq::PackedBitSet rangeBegin_, rangeEnd_;
    bool rangeSet_;
    Vector E_;
    sq::BitSetArray xList_;
};

/* =====================================================================================
 * bipartite-graph searcher
 * ===================================================================================== */
template <class real>
class B200BipartiteGraphBFSearcher : public sq::cuda::BipartiteGraphBFSearcher<real> {
    typedef sq::MatrixType<real> Matrix;
    typedef sq::VectorType<real> Vector;
    typedef sq::BipartiteGraphBFSearcher<real> Base;
    typedef B200BipartiteGraphBFSearcher<real> This;
    using Base::N0_; using Base::N1_; using Base::om_; using Base::tileSize0_; using Base::tileSize1_;
    using Base::x0_; using Base::x1_; using Base::x0max_; using Base::x1max_;

public:
    B200BipartiteGraphBFSearcher() : Emin_(DBL_MAX), k0_(0), k1_(0), r_(0), rMax_(0) {
        tileSize0_ = 1 << 15;
        tileSize1_ = 1 << 15;
    }
    void assignDevice(sq::cuda::Device &device) {
        sqb_throwErrorIf(engine_.device() != NULL, "Device assigned more than once.");
        engine_.assign(&asB200(device));
    }
    void setQUBO(const Vector &b0, const Vector &b1, const Matrix &W, sq::OptimizeMethod om = sq::optMinimize) {
        sqb_throwErrorIf(W.cols != b0.size || W.rows != b1.size, "%s, shape mismatch between b0, b1 and W.", __func__);
        sqb_throwErrorIf(b0.size > 63 || b1.size > 63, "N0 and N1 must be smaller than 64.");
        sqb_throwErrorIf(b0.size + b1.size > 63, "N0 + N1 = %d: a search over more than 2^63 pairs is not supported.", b0.size + b1.size);
        sqb_throwErrorIf(engine_.device() == NULL, "Device not set.");
        this->clearState(Base::solProblemSet);
        N0_ = b0.size; N1_ = b1.size; om_ = om;
        const double sign = (om == sq::optMaximize) ? -1. : 1.;
        b0_.assign(N0_, 0.); b1_.assign(N1_, 0.); W_.assign((size_t)N0_ * N1_, 0.);
        for (int j = 0; j < N0_; ++j) b0_[j] = sign * (double)b0(j);
        for (int i = 0; i < N1_; ++i) b1_[i] = sign * (double)b1(i);
        for (int i = 0; i < N1_; ++i) for (int j = 0; j < N0_; ++j) W_[(size_t)i * N0_ + j] = sign * (double)W(i, j);
        this->setState(Base::solProblemSet);
    }
    sq::Preferences getPreferences() const {
        sq::Preferences prefs = Base::getPreferences();
        prefs.pushBack(sq::Preference(sq::pnDevice, "cuda"));
        return prefs;
    }
    const Vector &get_E() const {
        if (!this->isEAvailable()) const_cast<This *>(this)->calculate_E();
        return E_;
    }
    const sq::BitSetPairArray &get_x() const {
        if (!this->isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
        return xPairList_;
    }
    void prepare() {
        this->throwErrorIfProblemNotSet();
        x0max_ = 1ull << N0_;
        x1max_ = 1ull << N1_;
        /* tiles are powers of two so that a tile0 x tile1 rectangle is a contiguous range of kernel indices */
        k0_ = floorLog2(std::min((unsigned long long)tileSize0_, x0max_));
        k1_ = floorLog2(std::min((unsigned long long)tileSize1_, x1max_));
        while (k0_ + k1_ > 30) { if (k0_ >= k1_) --k0_; else --k1_; }
        if ((1 << k0_) != tileSize0_ || (1 << k1_) != tileSize1_)
            sq::log("Tile sizes are adjusted to %d x %d.", 1 << k0_, 1 << k1_);
        tileSize0_ = 1 << k0_;
        tileSize1_ = 1 << k1_;
        /* kernel bit order: [x0 low k0][x1 low k1][x1 high][x0 high]; bit p of x0 is variable N0-1-p (MSB first) */
        const int N = N0_ + N1_;
        side_.assign(N, 0); pos_.assign(N, 0);
        int kb = 0;
        for (int p = 0; p < k0_; ++p, ++kb) { side_[kb] = 0; pos_[kb] = p; }
        for (int p = 0; p < k1_; ++p, ++kb) { side_[kb] = 1; pos_[kb] = p; }
        for (int p = k1_; p < N1_; ++p, ++kb) { side_[kb] = 1; pos_[kb] = p; }
        for (int p = k0_; p < N0_; ++p, ++kb) { side_[kb] = 0; pos_[kb] = p; }
        std::vector<double> Wk((size_t)N * N, 0.);
        for (int a = 0; a < N; ++a) {
            const int va = (side_[a] == 0 ? N0_ : N1_) - 1 - pos_[a];
            Wk[(size_t)a * N + a] = (side_[a] == 0) ? b0_[va] : b1_[va];
            for (int b = 0; b < N; ++b) {
                if (side_[a] == side_[b]) continue;
                const int vb = (side_[b] == 0 ? N0_ : N1_) - 1 - pos_[b];
                const double w = (side_[a] == 1) ? W_[(size_t)va * N0_ + vb] : W_[(size_t)vb * N0_ + va];
                Wk[(size_t)a * N + b] = 0.5 * w;
            }
        }
        engine_.setProblem(Wk, N);
        Emin_ = DBL_MAX;
        packed_.clear();
        xPairList_.clear();
        r_ = 0;
        rMax_ = 1ull << N;
        x0_ = x1_ = 0;
        this->setState(Base::solPrepared);
    }
    size_t solutionCap() const { return (size_t)std::min((long long)tileSize0_ + (long long)tileSize1_, 1ll << 16); }
    bool searchRange(sq::PackedBitSet *curX0End, sq::PackedBitSet *curX1End) {
        this->throwErrorIfNotPrepared();
        this->clearState(Base::solSolutionAvailable);
        const unsigned long long tile = 1ull << (k0_ + k1_);
        const unsigned long long b = r_, e = std::min(r_ + tile, rMax_);
        if (b < e) {
            const double tileMin = engine_.rangeMin(b, e);
            if (tileMin < Emin_) { Emin_ = tileMin; packed_.clear(); }
            if (tileMin == Emin_ && packed_.size() < solutionCap()) engine_.collect(b, e, Emin_, solutionCap(), &packed_);
        }
        r_ = e;
        /* cursor in the reference's terms (CPUBipartiteGraphBFSearcher.cpp:155-185): x1 advances first, then x0 */
        if (r_ == rMax_) { x0_ = x0max_; x1_ = 0; }
        else {
            sq::PackedBitSet a0, a1;
            split(r_, &a0, &a1);
            x0_ = a0; x1_ = a1;
        }
        if (curX0End != NULL) *curX0End = x0_;
        if (curX1End != NULL) *curX1End = x1_;
        return r_ == rMax_;
    }
    void calculate_E() {
        this->throwErrorIfNotPrepared();
        E_.resize(packed_.empty() ? 1 : (int)std::min(packed_.size(), solutionCap()));
        real v = (real)((om_ == sq::optMaximize) ? -Emin_ : Emin_);
        E_ = v;
        this->setState(Base::solEAvailable);
    }
    void makeSolution() {
        this->throwErrorIfNotPrepared();
        xPairList_.clear();
        std::vector<std::pair<unsigned long long, unsigned long long> > pairs;
        for (size_t i = 0; i < packed_.size(); ++i) {
            sq::PackedBitSet a0, a1;
            split(packed_[i], &a0, &a1);
            pairs.push_back(std::make_pair(a0, a1));
        }
        std::sort(pairs.begin(), pairs.end());
        for (size_t i = 0; i < pairs.size() && i < solutionCap(); ++i) {
            sq::BitSet x0, x1;
            sq::unpackBitSet(&x0, pairs[i].first, N0_);
            sq::unpackBitSet(&x1, pairs[i].second, N1_);
            xPairList_.pushBack(sq::BitSetPair(x0, x1));
        }
        calculate_E();
        this->setState(Base::solSolutionAvailable);
    }

private:
    static int floorLog2(unsigned long long v) { int k = 0; while ((2ull << k) <= v) ++k; return k; }
    void split(unsigned long long r, sq::PackedBitSet *x0, sq::PackedBitSet *x1) const {
        sq::PackedBitSet a0 = 0, a1 = 0;
        for (size_t kb = 0; kb < side_.size(); ++kb) {
            if (!((r >> kb) & 1ull)) continue;
            if (side_[kb] == 0) a0 |= 1ull << pos_[kb]; else a1 |= 1ull << pos_[kb];
        }
        *x0 = a0; *x1 = a1;
    }
    BFEngine engine_;
    std::vector<double> b0_, b1_, W_;
    std::vector<int> side_, pos_;
    double Emin_;
    int k0_, k1_;
    unsigned long long r_, rMax_;
    std::vector<unsigned long long> packed_;
    Vector E_;
    sq::BitSetPairArray xPairList_;
};

} // namespace sqb

namespace sqaod { namespace cuda {
template <> DenseGraphBFSearcher<float> *newDenseGraphBFSearcher<float>() { return new sqb::B200DenseGraphBFSearcher<float>(); }
template <> DenseGraphBFSearcher<double> *newDenseGraphBFSearcher<double>() { return new sqb::B200DenseGraphBFSearcher<double>(); }
template <> BipartiteGraphBFSearcher<float> *newBipartiteGraphBFSearcher<float>() { return new sqb::B200BipartiteGraphBFSearcher<float>(); }
template <> BipartiteGraphBFSearcher<double> *newBipartiteGraphBFSearcher<double>() { return new sqb::B200BipartiteGraphBFSearcher<double>(); }
}} // namespace sqaod::cuda
