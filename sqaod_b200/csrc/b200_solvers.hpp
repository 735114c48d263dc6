/* b200_solvers.hpp -- concrete solver classes of the B200 back end (internal header). */
#pragma once
#include "device.hpp"
#include "tc_gemm.hpp"
#include <vector>

namespace sqb {

/* ---- device-side formulas (formulas.cu) ---- */
/* h = -1/2 colsum(sW), J = -1/4 sW with zero diagonal, c = sum(J before zeroing) + sum(diag)   (s = +-1) */
template <class real>
void devDenseHamiltonian(const B200Device &dev, real *d_h, real *d_J, int ldJ, real *d_c, const real *d_W, int ldW, int N, real sign);
template <class real>
void devBipartiteHamiltonian(const B200Device &dev, real *d_h0, real *d_h1, real *d_J, int ldJ, real *d_c, const real *d_b0,
                             const real *d_b1, const real *d_W, int ldW, int N0, int N1, real sign);
/* E_b = alpha * ( sum_i v_bi (g_i + sum_j A_ij u_bj) + sum_j f_j u_bj ) + beta0 ; A is R x C, u is nBatch x C, v is nBatch x R
 * (g, f may be NULL).  Dense Ising: u = v = q, g = h.  QUBO: u = v = x.  Bipartite: A = J (N1 x N0), u = q0, v = q1. */
template <class real>
void devBatchedEnergy(const B200Device &dev, real *d_E, const real *d_A, int ldA, int R, int C, const signed char *d_u, int ldu,
                      const signed char *d_v, int ldv, const real *d_g, const real *d_f, int nBatch, real alpha, real beta0);
/* E[i1][i0] = b0.x0_i0 + b1.x1_i1 + x1_i1^T W x0_i0 */
template <class real>
void devBipartiteEnergy2D(const B200Device &dev, real *d_E, int ldE, const real *d_b0, const real *d_b1, const real *d_W, int ldW,
                          int N0, int N1, const signed char *d_x0, int ldx0, int n0, const signed char *d_x1, int ldx1, int n1);

/* C[y][i] = sum_k Q[y][k] A[i][k] on CUDA cores (bipartite_annealer.cu) */
template <class real>
void devSpinGemm(const B200Device &dev, real *C, int ldc, const real *A, int ldA, const signed char *Q, int ldq, int m, int NA, int NF);

void launchRandomizeSpin(const B200Device &dev, signed char *q, int ldq, int N, int m, unsigned long long seed,
                         unsigned long long count, unsigned domain, int yOff = 0, int mPerReplica = 0);
long long ringSpinDot(const B200Device &dev, const signed char *q, int ldq, int N, int m);

/* extras of the dense brute-force searcher reachable through the C ABI (sharded search) */
struct DenseBFExtras {
    virtual ~DenseBFExtras() {}
    virtual void setRange(sq::PackedBitSet xBegin, sq::PackedBitSet xEnd) = 0;
    virtual double getEmin() const = 0;
    virtual const sq::PackedBitSetArray &packedSolutions() const = 0;
    virtual void setPackedSolutions(double Emin, const sq::PackedBitSet *x, int n) = 0;
};

/* ---- dense-graph annealer ---- */
template <class real> class B200DenseGraphAnnealer : public sq::cuda::DenseGraphAnnealer<real> {
    typedef sq::MatrixType<real> HostMatrix;
    typedef sq::VectorType<real> HostVector;
    typedef B200DenseGraphAnnealer<real> This;
    typedef sq::DenseGraphAnnealer<real> Base;
public:
    B200DenseGraphAnnealer();
    ~B200DenseGraphAnnealer();
    void assignDevice(sq::cuda::Device &device);
    sq::Algorithm selectAlgorithm(sq::Algorithm algo);
    void seed(unsigned long long seed);
    void setQUBO(const HostMatrix &W, sq::OptimizeMethod om = sq::optMinimize);
    void setHamiltonian(const HostVector &h, const HostMatrix &J, real c = real(0.));
    void getHamiltonian(HostVector *h, HostMatrix *J, real *c) const;
    sq::Preferences getPreferences() const;
    const HostVector &get_E() const;
    const sq::BitSetArray &get_x() const;
    void set_q(const sq::BitSet &q);
    void set_qset(const sq::BitSetArray &q);
    const sq::BitSetArray &get_q() const;
    void randomizeSpin();
    void prepare();
    void calculate_E();
    void makeSolution();
    real getSystemE(real G, real beta) const;
    void annealOneStep(real G, real beta);

    /* extras used by the C ABI */
    void setSpinsRaw(const signed char *q, int m);
    void getSpinsRaw(signed char *q) const;
    void getStats(unsigned long long *accepted, unsigned long long *waits) const;
    void getBarrierStats(unsigned long long *dot, unsigned long long *chain) const { *dot = lastBarrierWaitDot_; *chain = lastBarrierWaitChain_; }
    void getCounters(unsigned long long out[8]) const; /* raw sweep counters, see SweepParams::stats */
    void getProfile(unsigned long long out[16]) const;
    int getCtaProfile(unsigned long long *out, int maxCtas) const; /* per CTA, last launch: [T, wait fields, wait neighbours, chain work, loop cycles, end time (ns), accepted, passes] */ /* counters + the field-mode chain profile (stats[8..15]) */
    int numTrotters() const { return m_; }
    /* replica batch: R independent replicas of the problem (seed + r) annealed side by side; spin / energy rows are [r][y] */
    void setNumReplicas(int n);
    void setQUBOBatch(const real *W, int nProblems, int N, int ldW, sq::OptimizeMethod om);
    /* synthetic symmetric W ~ U(-0.5, 0.5) generated on the device (benchmarks / multi-GPU tests); getQUBORandom returns the same matrix */
    void setQUBORandom(int N, unsigned long long seed, bool quantize, sq::OptimizeMethod om);
    void getQUBORandom(real *W, int N, int ldW, unsigned long long seed, bool quantize) const;
    int numProblems() const { return nProblems_; }
    int numReplicas() const { return nReplicas_; }
    /* ring sharding over several GPUs: this solver anneals trotters [rank*m/world, (rank+1)*m/world) of one ring */
    void ringConfigure(int rank, int world, int mGlobal);
    void ringExport(unsigned char handle[64]) const;
    void ringAttach(const unsigned char left[64], const unsigned char right[64]);
    void ringPushHalos();
    B200Device *device() const { return dev_; }

private:
    void uploadProblem(const real *h, const real *J, int strideJ);
    void hamiltonianFromDeviceQUBO(const real *dW);
    void prepareTensorCoreOperand();
    void syncBits();

    B200Device *dev_;
    DevBuf<real> dJ_, dh_, dE_;
    DevBuf<signed char> dq_, dq2_;  /* current spins; the buffer the next sweep writes (swapped after every step) */
    DevBuf<unsigned long long> dStats_;
    void *handoff_;            /* accept flags, snapshots, halo rows (HandoffLayout in dense_annealer.cu) */
    bool handoffIpc_;
    void *peerBase_[2];        /* the left / right peer's hand-off block, opened through CUDA IPC */
    int ringRank_, ringWorld_, mRing_, yOff_;
    unsigned long long ringEpoch_;
    int nReplicas_, replicasPerLaunch_;
    int nProblems_ = 1;            /* > 1: replica r anneals problem r (setQUBOBatch) */
    std::vector<real> cBatch_;
    void allocHandoff();
    void freeHandoff();
    void closePeer(int side);
    TcOperand tcJ_;
    TcWorkspace tcWs_;
    int ldJ_, ldq_;
    real c_;
    unsigned long long seed_, step_, randomizeCount_, launchCount_;
    int grid_, chunkElems_, chunksPerRow_, stages_, nw64_, nWindows_, K_;
    int dotWarps_ = 12;
    /* field mode of the sweep (dense_annealer.cu): local fields J.q kept in shared memory and updated per accepted flip */
    bool fieldMode_ = false;
    bool specChain_ = true;        /* window-parallel accept chain (SQAOD_B200_SWEEP_SPEC=0: the sequential per-round chain) */
    DevBuf<real> dF_;              /* [m * replicas][ldJ] fields at step start */
    AsyncReadback<real> eBack_;     /* energies: pinned landing buffer + completion event (calculate_E never synchronises the stream) */
    DevBuf<unsigned char> dTables_; /* field mode: per-step tables of the sweep (sweepTablesKernel) */
    DevBuf<real> dRowMax_;         /* scratch of prepare(): max_j |J[i][j]| per row */
    real jAbsMax_ = real(0);       /* max |J| (field mode: bound of a cross term in flight) */
    bool fieldsHaveH_ = false;     /* dF_ holds h + 2 J.q (written back by a sweep) instead of the spin GEMM's J.q */
    bool fieldsValid_ = false;     /* dF_ matches dq_ (cleared by everything that writes spins or the problem) */
    int fieldRefresh_ = 1, stepsSinceRefresh_ = 0; /* recompute F = J.q with the spin GEMM every fieldRefresh_ steps */
    int sweepModeWanted_ = -1, fieldRefreshWanted_ = 0; /* setSweepMode(); -1 / 0: automatic */
    void refreshFields();
public:
    bool fieldMode() const { return fieldMode_; }
    /* mode -1: automatic (field mode whenever the field rows fit in shared memory), 0: classic (one J row per attempt),
     * 1: field mode (error in prepare() when it does not fit); fieldRefresh > 0: steps between two J.q recomputations */
    void setSweepMode(int mode, int fieldRefresh);
    bool getFields(real *H, int ldH) const; /* carried local fields h + 2 J.q (field mode with write-back), [rows][ldH] */
private:
    size_t smemBytes_;
    mutable unsigned long long lastBarrierWaitDot_ = 0, lastBarrierWaitChain_ = 0;
    HostVector E_;
    std::vector<signed char> hq_;
    sq::BitSetArray xlist_, qlist_;

    using Base::selectDefaultAlgorithm;
    using Base::selectDefaultSAAlgorithm;
    using Base::N_;
    using Base::m_;
    using Base::om_;
    using Base::algo_;
    using Base::solRandSeedGiven;
    using Base::solPrepared;
    using Base::solProblemSet;
    using Base::solQSet;
    using Base::solEAvailable;
    using Base::solSolutionAvailable;
    using Base::setState;
    using Base::clearState;
    using Base::isRandSeedGiven;
    using Base::isPrepared;
    using Base::isEAvailable;
    using Base::isSolutionAvailable;
    using Base::throwErrorIfProblemNotSet;
    using Base::throwErrorIfNotPrepared;
    using Base::throwErrorIfQNotSet;
};

} // namespace sqb
