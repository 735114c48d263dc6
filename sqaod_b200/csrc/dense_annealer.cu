/* dense_annealer.cu -- dense-graph SQA / SA annealer for B200 (sm_100a).
 *
 * Replaces the reference's CUDADenseGraphAnnealer (sqaodc/cuda/CUDADenseGraphAnnealer.cu:128-602) together with its
 * J.q reduction kernels (DeviceSegmentedSum.cuh, DeviceBatchedDot.cuh:211-262), flip kernels (:428-480, :543-561) and
 * random-number pool (DeviceRandomBuffer.cu).  The Markov chain is the reference CPU solver's
 * (sqaodc/cpu/CPUDenseGraphAnnealer.cpp:250-338): per annealOneStep N rounds; in each round every trotter y draws a
 * spin x and a uniform u and does one Metropolis attempt with
 *     dE = (2/m) q_yx (h_x + 2 sum_j J_xj q_yj) - q_yx (q_{y-1,x} + q_{y+1,x}) coef,   coef = ln tanh(G beta/m)/beta
 * even trotters first, then (m odd) trotter m-1, then odd trotters.  (x, u) come from Philox keyed by
 * (seed, step, round, y) instead of per-thread MT19937 streams, so the trajectory is a pure function of the seed and
 * is reproduced bit for bit by the CPU oracle in "philox" mode (tests/test_dense_annealer_gpu.py).
 *
 * ONE persistent cooperative kernel per annealOneStep (the reference needs 2N dependent launches):
 *   - CTA c owns T contiguous trotters (m spread over min(#SM, m) CTAs); their spins live bit-packed in shared memory.
 *   - 12 or 14 "dot" warps stream the J rows the owned
 *     trotters will need, one window of K <= 16 rounds ahead of the accept chain, through per-warp rings of TMA bulk
 *     copies (cp.async.bulk + mbarrier), and reduce sum_j J_xj q_yj against a SNAPSHOT of q_y with warp shuffles.
 *     Flip positions are state-independent (Philox), so rows are known arbitrarily far ahead; rows are claimed from a
 *     shared counter.
 *   - because at most one spin per trotter changes per round, the dot product against the stale snapshot is repaired
 *     exactly with one term per accepted flip since the snapshot: -2 q_old[x'] J[x][x'].  The J[x][x'] cross terms are
 *     picked out of the row while it sits in shared memory (one per lane, <= 2K-1 = 31 of them).
 *   - 1 "chain" warp (lane = trotter) replays the K rounds in the
 *     reference's order using the finished dot products, the cross terms and the neighbours' spins.  Neighbours inside
 *     the CTA are read from shared memory; for the two trotters owned by other CTAs the chain uses a published
 *     snapshot plus the accept bits of exactly those neighbour attempts that hit the same spin index (probability
 *     ~K/N per attempt), so there is no per-round grid barrier: CTAs meet only through per-window snapshots and rare
 *     per-attempt flag waits (L2, release/acquire).
 *   - helper roles: "snapshot" (builds S_w from the window's accept bits, publishes the edge trotters), "neighbour"
 *     (fetches the neighbours' snapshots, builds the conflict masks), "prep" (Philox tables two windows ahead).
 *   - two warp layouts (prepare() picks one from the trotters per CTA): up to 2 trotters per CTA the chain bounds the
 *     step, so warp 0 (chain) shares its scheduler only with the three helper warps 4 / 8 / 12 and the 12 dot warps are
 *     those of the other three schedulers; from 3 trotters per CTA on, warps 4 and 8 stream rows too (14 dot warps) and
 *     warp 12 does the three helper jobs in turn.
 *   - inside the CTA there is no barrier in the sweep either: the warps hand work over through five release/acquire
 *     counters in shared memory (see the kernel body).
 * HBM traffic of this "classic" form is one J row per attempt (N*sizeof(real) bytes), the figure SURVEY.md section 8(d) uses.
 *
 * FIELD mode (template flag FIELD; the north star's "local-field update kept incrementally in shared memory"): the local fields
 * F[y][j] = sum_i J[j][i] q[y][i] of every trotter are computed once per step by the tcgen05 spin GEMM (energy_tc.cu; CUDA-core
 * GEMM for fp64), loaded into shared memory (T rows of N reals per CTA) and updated with one J row per ACCEPTED flip, two
 * windows behind the chain; the look-ahead machinery (cross terms, fold, repair) bridges those two windows exactly as it
 * bridges the snapshot lag of the classic form, so both forms run the same Markov chain bit for bit.  The dot warps turn into
 * "field warps": they own column groups of the field rows, stream the rows of accepted flips (L2 prefetch + 128-bit loads),
 * gather the <= 2K-1 cross terms J[x][x'] of every attempt with 4-byte loads a window ahead, and hand the chain
 * scaleA (h[x] + 2 F[y][x]).  Traffic drops from one row per attempt to (acceptance rate) rows + 31 sectors per attempt.
 *
 * CTAs meet through (a) one 64-bit word per edge trotter and window (tag << 16 | accept bits) from which the neighbouring CTA
 * rebuilds the trotter's spins with its own copy of the Philox draws, and (b) per-attempt accept flags, published only for
 * the rounds in which the neighbour can draw the same spin index.  Both are mirrored to the neighbouring GPU when the ring is
 * sharded.
 */
#include "device.hpp"
#include "kernels_common.cuh"
#include "philox.cuh"
#include "b200_solvers.hpp"
#include <math.h>
#include <type_traits>
#include <time.h>
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>

namespace sqb {

enum { SW_MAX_K = 16, SW_DOT_WARPS = 12 /* chain-critical layout */, SW_DOT_WARPS_WIDE = 14 /* many trotters per CTA */, SW_THREADS = 512, SW_FLAG_RING = 128, SW_SNAP_SLOTS = 4, SW_TAB_SLOTS = 4, SW_TAB_SLOTS_FIELD = 8 /* field mode prepares tables three windows ahead */,
       SW_CHAIN_WARP = 0, SW_SNAP_WARP = 4, SW_PREP_WARP = 8, SW_NB_WARP = 12 /* the dot warps are those with warp & 3 != 0 */,
       /* field mode: warps 0-3 are accept-chain warps (one per scheduler, trotter t -> warp t % 4), warp 4 does all the helper
        * work, warps 5-15 own the column groups of the field rows */
       SW_FIELD_CHAIN_WARPS = 4, SW_FIELD_NB_WARP = 4, SW_FIELD_PREP_WARP = 5, SW_FIELD_WARPS = 10 };

template <class real> struct SweepParams {
    const real *J;
    const real *h;
    signed char *q;     /* spins at step start (read) */
    signed char *qOut;  /* spins at step end (written): a second buffer, swapped by the host after the launch -- a CTA may still be
                         * packing a neighbour's step-start row while the owner of that row has finished the sweep */
    int ldJ, ldq, N, m;
    unsigned long long seed, step;
    real twoDivM, coef, beta;
    real scaleA, scaleNb; /* accept iff q (scaleA (h + 2 sum) - scaleNb (ql + qr)) < -ln u : scaleA = beta 2/m (SA: 2/kT), scaleNb = beta coef */
    int chunkElems, chunksPerRow, stages, nw64, K;
    int dotWarps; /* SW_DOT_WARPS: chain alone on scheduler 0; SW_DOT_WARPS_WIDE: warps 4 and 8 stream too, warp 12 does all the helper work */
    /* hand-off arrays have m + 2 slots: local trotter l -> slot l; slot m / m + 1 = the trotter left of local 0 / right of
     * local m-1 when it lives on another GPU (ring sharding); the owning GPU mirrors its publications into them */
    unsigned long long *acceptFlags; /* [m+2][SW_FLAG_RING] */
    unsigned long long *snapFlags;   /* [m+2] */
    unsigned long long *snapBits;    /* [m+2][SW_SNAP_SLOTS][nw64] */
    /* ring sharding: this launch owns trotters yOff .. yOff+m-1 of a ring of mRing; mRing == m: unsharded */
    int mRing, yOff;
    /* replica batch: blockIdx.y selects replica replicaBase + blockIdx.y (seed + replica, own spins and hand-off block) */
    int replicaBase;
    size_t qReplicaStride, handoffReplicaStride; /* in bytes */
    size_t jReplicaStride, hReplicaStride;       /* in elements; 0: every replica anneals the same problem, else replica r has its own J and h */
    const signed char *haloQ[2];         /* spins of the left / right foreign neighbour at step start (pushed by the peers) */
    const unsigned long long *stepFlags; /* [2]: epoch of the last halo push received from the left / right peer */
    unsigned long long stepEpoch;
    unsigned long long *peerFlags[2], *peerSnapFlags[2], *peerSnapBits[2]; /* the peers' arrays (NVLink P2P), or NULL */
    unsigned long long roundBase, snapBase;
    /* field mode (FIELD kernels): local fields F[y][j] = sum_i J[j][i] q[y][i] of every trotter, [rows][ldF] in global memory at
     * step start (spin GEMM or the previous step's write-back); kept in shared memory and updated incrementally by the sweep */
    real *F;
    int ldF, writeBackF;
    /* field mode: per (replica of the launch, CTA, window) record of Philox draws and conflict masks, written by
     * sweepTablesKernel before the sweep (SweepTabRec layout) */
    const unsigned char *tables;
    real uncBound;   /* field mode: 4 scaleA max|J| (with slack) -- bounds the cross term of a gather that is still in flight */
    int fieldHasH;   /* field mode: the rows in F already hold h + 2 J.q (written back by the previous step) instead of J.q */
    int specChain; /* accept chain evaluates a whole window in parallel and commits flips in order (see the chain warp) */
    unsigned long long *stats;       /* [0] accepted flips, [1] remote wait polls, [2]/[3] busy cycles of dot warp 0 / the chain warp, summed over CTAs */
};

/* shared-memory carve-up, identical on host and device */
template <class real> struct SweepSmem {
    size_t field, ring, bars, qcur, qsnap, nbsnap, dots, cross, xs, xb, us, hs, need, xn, conf, confAny, accLog, sgnLog, spec, carry, pend, cstate, flipQ, counter, total;
    int flipQLen;
    /* fieldElems > 0: field mode -- T rows of fieldElems local fields instead of the TMA ring (stages == 0) */
    __host__ __device__ SweepSmem(int T, int nw64, int chunkElems, int stages, int K, int dotWarps, int fieldElems = 0) {
        size_t o = 0;
        field = o; o += (size_t)T * fieldElems * sizeof(real);
        ring = o; o += (size_t)dotWarps * stages * chunkElems * sizeof(real);
        bars = o; o += (size_t)dotWarps * stages * 8;
        qcur = o; o += (size_t)T * nw64 * 8;
        qsnap = o; o += (size_t)(fieldElems ? 0 : 2) * T * nw64 * 8; /* field mode has no use for snapshots */
        nbsnap = o; o += (size_t)2 * 2 * nw64 * 8;
        dots = o; o += (size_t)2 * T * K * sizeof(real);
        o = (o + 15) & ~(size_t)15;
        cross = o; o += (size_t)(fieldElems ? 0 : 2) * T * K * (2 * K) * sizeof(real); /* field mode gathers cross terms per ACCEPTED flip */
        const size_t tab = fieldElems ? SW_TAB_SLOTS_FIELD : SW_TAB_SLOTS;
        xs = o; o += tab * T * K * 4;
        xb = o; o += tab * T * K * 4;
        o = (o + 15) & ~(size_t)15;
        us = o; o += tab * T * K * sizeof(real);
        hs = o; o += (fieldElems ? 0 : tab) * T * K * sizeof(real); /* field mode keeps h inside its field rows */
        need = o; o += (fieldElems ? tab : 0) * T * K * 4;          /* field mode: local-neighbour conflict masks from the table pre-pass */
        xn = o; o += (size_t)2 * tab * K * 4;
        conf = o; o += tab * 2 * K * 4;      /* [tab][2 sides][K] */
        confAny = o; o += tab * 2 * 4 * 2;   /* [tab][2 sides] + pubMask[tab][2 sides] */
        accLog = o; o += (size_t)2 * T * 4;
        sgnLog = o; o += (size_t)2 * T * 4;
        spec = o; o += (size_t)(2 * T + 2 * T * K) * 4; /* window-parallel chain: trotter info, frontiers, local conflict masks */
        o = (o + 15) & ~(size_t)15;
        /* field mode: carry[T][K] = corrections already known for the NEXT window's attempts; pend[T][32] = landing slots of the
         * cross-term gathers in flight; cstate[T][8] = per-trotter chain state (pending generation, accept / sign logs) */
        carry = o; o += (size_t)(fieldElems ? 1 : 0) * T * K * sizeof(real);
        pend = o; o += (size_t)(fieldElems ? 1 : 0) * T * 32 * sizeof(real);
        cstate = o; o += (size_t)(fieldElems ? 1 : 0) * T * 8 * sizeof(real);
        o = (o + 15) & ~(size_t)15;
        /* field mode: queue of committed flips (chain warps -> field warps), 64-bit entries (sequence number << 32 | payload).  The
         * chain cannot be more than two windows ahead of the slowest field warp: 2 (T K flips + 1 marker) entries at most */
        flipQLen = 0;
        if (fieldElems) { flipQLen = 64; while (flipQLen < 2 * (T * K + 1) + 2) flipQLen <<= 1; }
        flipQ = o; o += (size_t)flipQLen * 8;
        counter = o; o += 32;
        total = (o + 127) & ~(size_t)127;
    }
};

/* -ln(u) for the accept test exp(-dE beta) > u  <=>  dE beta < -ln(u).  u == 1 (possible after rounding to fp32) never
 * accepts, like `1 > u` in the reference; u == 0 accepts whenever exp() would not underflow to zero. */
template <class real> __device__ __forceinline__ real negLogUniform(const Philox4 &p);
template <> __device__ __forceinline__ float negLogUniform<float>(const Philox4 &p) {
    const float u = philoxUniform<float>(p);
    if (u >= 1.f) return -INFINITY;
    if (u <= 0.f) return 103.f;
    return -logf(u);
}
template <> __device__ __forceinline__ double negLogUniform<double>(const Philox4 &p) {
    const double u = philoxUniform<double>(p);
    if (u >= 1.) return -INFINITY;
    if (u <= 0.) return 745.;
    return -log(u);
}
template <class real> __device__ __forceinline__ real expReal(real v);
template <> __device__ __forceinline__ float expReal<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double expReal<double>(double v) { return exp(v); }

/* accumulate the signed sum of one 128-spin group: lane owns 4 consecutive elements.  The four fp32 accumulators are kept
 * as two packed pairs so that the adds are `add.rn.f32x2` (SASS FADD2): same IEEE result per lane, half the add issue slots. */
__device__ __forceinline__ unsigned long long packF32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void addF32x2(unsigned long long &acc, unsigned long long v) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(v)); }
struct Acc4f { /* a0..a3 of the fp32 dot product */
    unsigned long long a01, a23;
    __device__ __forceinline__ Acc4f() : a01(0ull), a23(0ull) {}
    __device__ __forceinline__ float sum() const {
        float a0, a1, a2, a3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(a23));
        return (a0 + a1) + (a2 + a3);
    }
};
struct Acc4d {
    double a0, a1, a2, a3;
    __device__ __forceinline__ Acc4d() : a0(0.), a1(0.), a2(0.), a3(0.) {}
    __device__ __forceinline__ double sum() const { return (a0 + a1) + (a2 + a3); }
};
template <class real> struct Acc4;
template <> struct Acc4<float> { typedef Acc4f type; };
template <> struct Acc4<double> { typedef Acc4d type; };

__device__ __forceinline__ void accumGroup(const float *buf, uint32_t nib, Acc4f &a) {
    float4 v = *reinterpret_cast<const float4 *>(buf);
    uint32_t neg = ~nib; /* bit set -> spin +1 -> keep sign */
    addF32x2(a.a01, packF32x2(signFlip(v.x, neg << 31), signFlip(v.y, neg << 30)));
    addF32x2(a.a23, packF32x2(signFlip(v.z, neg << 29), signFlip(v.w, neg << 28)));
}
__device__ __forceinline__ void accumGroup(const double *buf, uint32_t nib, Acc4d &a) {
    double2 v0 = *reinterpret_cast<const double2 *>(buf);
    double2 v1 = *reinterpret_cast<const double2 *>(buf + 2);
    uint32_t neg = ~nib;
    a.a0 += signFlip(v0.x, neg << 31);
    a.a1 += signFlip(v0.y, neg << 30);
    a.a2 += signFlip(v1.x, neg << 29);
    a.a3 += signFlip(v1.y, neg << 28);
}

/* field mode: F[0..3] += c * J[0..3] (c = +-2, so the products are exact); 4 consecutive elements per lane */
struct RowVec4f { float4 v; };
struct RowVec4d { double2 v0, v1; };
template <class real> struct RowVec4;
template <> struct RowVec4<float> { typedef RowVec4f type; };
template <> struct RowVec4<double> { typedef RowVec4d type; };
__device__ __forceinline__ void loadRow4(const float *p, RowVec4f &r) { r.v = __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void loadRow4(const double *p, RowVec4d &r) {
    r.v0 = __ldg(reinterpret_cast<const double2 *>(p));
    r.v1 = __ldg(reinterpret_cast<const double2 *>(p) + 1);
}
__device__ __forceinline__ void axpyRow4(float *F, float c, const RowVec4f &r) {
    float4 f = *reinterpret_cast<float4 *>(F);
    f.x = fmaf(c, r.v.x, f.x); f.y = fmaf(c, r.v.y, f.y); f.z = fmaf(c, r.v.z, f.z); f.w = fmaf(c, r.v.w, f.w);
    *reinterpret_cast<float4 *>(F) = f;
}
__device__ __forceinline__ void axpyRow4(double *F, double c, const RowVec4d &r) {
    double2 f0 = *reinterpret_cast<double2 *>(F), f1 = *(reinterpret_cast<double2 *>(F) + 1);
    f0.x = fma(c, r.v0.x, f0.x); f0.y = fma(c, r.v0.y, f0.y); f1.x = fma(c, r.v1.x, f1.x); f1.y = fma(c, r.v1.y, f1.y);
    *reinterpret_cast<double2 *>(F) = f0;
    *(reinterpret_cast<double2 *>(F) + 1) = f1;
}

/* Rare path of the accept chain, kept out of line so that the hot loop stays small enough for the instruction cache: the
 * spin of a trotter owned by another CTA -- published snapshot (one window old), corrected by the accept bits of exactly
 * those of its attempts (bits of `mask`: j < K previous window, j >= K this window) that drew the same spin index. */
__device__ __noinline__ int remoteSpinSlow(const unsigned long long *row, int x, uint32_t mask, long long rrBase, unsigned long long roundBase,
                                           const unsigned long long *flags, int sys, unsigned long long *nWaits) {
    int v = spinAt(row, x);
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const long long rr = rrBase + j;
        const unsigned long long want = roundBase + (unsigned long long)rr + 1ull;
        const unsigned long long *f = flags + (rr % SW_FLAG_RING);
        /* the flag word carries its own payload (tag, accept bit): relaxed accesses are enough */
        unsigned long long got = sys ? ldRelaxedSys(f) : ldRelaxed(f);
        while ((got >> 1) != want) { ++*nWaits; __nanosleep(20); got = sys ? ldRelaxedSys(f) : ldRelaxed(f); }
        if (got & 1ull) v = -v;
    }
    return v;
}

/* non-blocking variant for the window-parallel chain: spin bit (0 / 1), or -1 when one of the accept flags is not there yet */
__device__ __noinline__ int remoteBitTry(const unsigned long long *row, int x, uint32_t mask, long long rrBase, unsigned long long roundBase,
                                         const unsigned long long *flags, int sys) {
    int v = spinAt(row, x) > 0 ? 1 : 0;
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        const long long rr = rrBase + j;
        const unsigned long long *f = flags + (rr % SW_FLAG_RING);
        const unsigned long long got = sys ? ldRelaxedSys(f) : ldRelaxed(f);
        if ((got >> 1) != roundBase + (unsigned long long)rr + 1ull) return -1;
        v ^= (int)(got & 1ull);
    }
    return v;
}

__device__ __forceinline__ int sweepPhase(int y, int m) { /* 0: even, 1: trotter m-1 of an odd ring, 2: odd */
    if (y & 1) return 2;
    return ((m & 1) && y == m - 1) ? 1 : 0;
}

/* ---------------- field mode: table pre-pass ----------------
 * Everything about a step that depends on the Philox stream only -- which spin each attempt draws, -ln(u), and which attempts
 * of neighbouring trotters draw the same spin index within the look-back range -- is computed for ALL windows of the step by
 * one throughput kernel before the sweep (one warp per CTA-window, every SM busy), instead of by one latency-bound helper warp
 * per sweep CTA that had to keep three windows ahead of the accept chain (it took 3.7 us per 16-round window and paced the
 * whole sweep).  The sweep kernel's table warp only copies one ~1 KB record per window into shared memory. */
template <class real> struct SweepTabRec { /* byte offsets inside the record of one (CTA, window); arrays are [maxT][K] / [2 sides][K] */
    size_t xs, xb, us, need, xn, conf, misc, bytes;
    __host__ __device__ SweepTabRec(int maxT, int K) {
        size_t o = 0;
        xs = o; o += (size_t)maxT * K * 4;
        xb = o; o += (size_t)maxT * K * 4;
        need = o; o += (size_t)maxT * K * 4;
        xn = o; o += (size_t)2 * K * 4;
        conf = o; o += (size_t)2 * K * 4;
        misc = o; o += 16; /* confAny[2], pubMask[2] */
        us = o; o += (size_t)maxT * K * sizeof(real);
        bytes = (o + 15) & ~(size_t)15;
    }
};

struct SweepTabParams {
    unsigned char *tables;
    int N, m, G, K, nW, replicaBase;
    unsigned long long seed, step;
    int sqa;
};

template <class real> __global__ void __launch_bounds__(128) sweepTablesKernel(const SweepTabParams P) {
    extern __shared__ int tabScratch[]; /* per warp: own draws [maxT][K], then the foreign neighbours' draws [2 sides][3 windows][K] */
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * 4 + wib, cta = blockIdx.y, rp = blockIdx.z;
    if (w >= P.nW) return;
    const int N = P.N, m = P.m, G = P.G, K = P.K;
    const int baseT = m / G, remT = m % G;
    const int T = baseT + (cta < remT ? 1 : 0), maxT = baseT + (remT ? 1 : 0);
    const int y0 = cta * baseT + min(cta, remT);
    const unsigned long long seedR = P.seed + (unsigned long long)(P.replicaBase + rp);
    const SweepTabRec<real> R(maxT, K);
    unsigned char *rec = P.tables + (((size_t)rp * G + cta) * P.nW + w) * R.bytes;
    int *rxs = reinterpret_cast<int *>(rec + R.xs), *rxb = reinterpret_cast<int *>(rec + R.xb), *rxn = reinterpret_cast<int *>(rec + R.xn);
    uint32_t *rneed = reinterpret_cast<uint32_t *>(rec + R.need), *rconf = reinterpret_cast<uint32_t *>(rec + R.conf), *rmisc = reinterpret_cast<uint32_t *>(rec + R.misc);
    real *rus = reinterpret_cast<real *>(rec + R.us);
    int *sx = tabScratch + (size_t)wib * (maxT * K + 6 * K);
    int *snb = sx + maxT * K;
    const int Kw = min(K, N - w * K);
    auto roundsIn = [&](int ww) { return min(K, N - ww * K); };
    const uint32_t kM = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);

    for (int idx = lane; idx < T * K; idx += 32) { /* own draws */
        const int t = idx / K, r = idx % K;
        int x = -1 - idx;
        if (r < Kw) {
            const Philox4 p = sqbPhilox(seedR, P.step, DOM_DENSE_SWEEP, (uint32_t)(w * K + r), (uint32_t)(y0 + t));
            x = (int)(p.w[0] % (uint32_t)N);
            int w64, bit;
            spinBitPos(x, w64, bit);
            rxs[t * K + r] = x;
            rxb[t * K + r] = ((2 * w64 + (bit >> 5)) << 5) | (bit & 31);
            rus[t * K + r] = negLogUniform<real>(p);
        }
        sx[idx] = x;
    }
    const bool remote = P.sqa && G > 1;
    const int yLeft = (y0 == 0) ? m - 1 : y0 - 1, yRight = (y0 + T - 1 == m - 1) ? 0 : y0 + T;
    if (remote) { /* the foreign neighbours' draws of windows w-1, w, w+1 */
        for (int idx = lane; idx < 6 * K; idx += 32) {
            const int side = idx / (3 * K), u = (idx / K) % 3, j = idx % K;
            const int wu = w - 1 + u;
            int x = -100 - idx;
            if (wu >= 0 && wu < P.nW && j < roundsIn(wu)) {
                const Philox4 p = sqbPhilox(seedR, P.step, DOM_DENSE_SWEEP, (uint32_t)(wu * K + j), (uint32_t)(side ? yRight : yLeft));
                x = (int)(p.w[0] % (uint32_t)N);
            }
            snb[idx] = x;
            if (u == 1 && j < Kw) rxn[side * K + j] = x;
        }
    }
    __syncwarp();
    /* local neighbours: rounds of trotter t-1 / t+1 (same CTA) that must be final before round r of trotter t -- same spin
     * index, earlier in the reference order (earlier round, or same round and earlier phase) */
    for (int idx = lane; idx < T * K; idx += 32) {
        const int t = idx / K, r = idx % K;
        uint32_t nd = 0u;
        if (P.sqa && r < Kw) {
            const int x = sx[idx];
            const int g = y0 + t, ph = sweepPhase(g, m);
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int gn = side ? (g == m - 1 ? 0 : g + 1) : (g == 0 ? m - 1 : g - 1);
                if (gn < y0 || gn >= y0 + T) continue;
                const int tn = gn - y0;
                uint32_t msk = 0u;
                for (int j = 0; j < Kw; ++j) msk |= (sx[tn * K + j] == x ? 1u : 0u) << j;
                msk &= ((1u << r) - 1u) | ((sweepPhase(gn, m) < ph ? 1u : 0u) << r);
                nd |= msk << (16 * side);
            }
        }
        if (r < Kw) rneed[idx] = nd;
    }
    if (remote) { /* foreign neighbours: conf = their attempts of windows w-1 / w on my edge trotter's spin index; pubMask = rounds
                   * whose accept flag they can ever read (their attempts of windows w / w+1 draw the same index) */
        uint32_t anyMine = 0u, pubMine = 0u; /* K <= 16: lane = (side, r) */
        const int side = lane / K, r = lane % K;
        if (side < 2 && r < Kw) {
            uint32_t mP = 0u, mC = 0u, mQ = 0u;
            const int xe = sx[(side ? T - 1 : 0) * K + r];
            const int *nb = snb + side * 3 * K;
            for (int j = 0; j < K; ++j) {
                mP |= (nb[j] == xe ? 1u : 0u) << j;
                mC |= (nb[K + j] == xe ? 1u : 0u) << j;
                mQ |= (nb[2 * K + j] == xe ? 1u : 0u) << j;
            }
            rconf[side * K + r] = mP | (mC << K);
            anyMine = (mP | mC) ? 1u : 0u;
            pubMine = (mC | mQ) ? 1u : 0u;
        }
        const uint32_t nz = __ballot_sync(0xffffffffu, anyMine != 0u), pb = __ballot_sync(0xffffffffu, pubMine != 0u);
        if (lane < 2) { rmisc[lane] = (nz >> (lane * K)) & kM; rmisc[2 + lane] = (pb >> (lane * K)) & kM; }
    } else if (lane < 4) rmisc[lane] = 0u;
}

template <class real, bool SQA, int K, bool FIELD>
__global__ void __launch_bounds__(SW_THREADS, 1) denseSweepKernel(const SweepParams<real> P) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TAB = FIELD ? SW_TAB_SLOTS_FIELD : SW_TAB_SLOTS; /* slots of the per-window Philox tables */
    /* independent replicas of the same problem share J and h; seed, spins and hand-off block are per replica */
    const int replica = P.replicaBase + (int)blockIdx.y;
    const unsigned long long seedR = P.seed + (unsigned long long)replica;
    signed char *const qBase = P.q + (size_t)replica * P.qReplicaStride;
    const real *const Jr = P.J + (size_t)replica * P.jReplicaStride;
    const real *const hr = P.h + (size_t)replica * P.hReplicaStride;
    const size_t handoffOff = (size_t)replica * P.handoffReplicaStride / 8;
    unsigned long long *const aFlags = P.acceptFlags + handoffOff;
    unsigned long long *const sFlags = P.snapFlags + handoffOff;
    unsigned long long *const sBits = P.snapBits + handoffOff;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    /* warp roles.  Warps are spread over the SM's four schedulers by (warp & 3): scheduler 0 is kept for the latency-bound
     * accept chain and its helpers, the twelve streaming dot warps share the other three. */
    const bool wide = !FIELD && (P.dotWarps == SW_DOT_WARPS_WIDE);
    const bool dotWarp = FIELD ? (warp > SW_FIELD_PREP_WARP) : ((warp & 3) != 0 || (wide && (warp == SW_SNAP_WARP || warp == SW_PREP_WARP)));
    /* dot warp index: 0..11 for the warps of schedulers 1-3, 12 / 13 for warps 4 / 8 in the wide layout; field mode: warps 6..15 */
    const int dw = FIELD ? warp - (SW_FIELD_PREP_WARP + 1) : ((warp & 3) ? (warp >> 2) * 3 + (warp & 3) - 1 : SW_DOT_WARPS + (warp >> 2) - 1);
    const bool chainWarp = FIELD ? (warp < SW_FIELD_CHAIN_WARPS) : (warp == SW_CHAIN_WARP);
    const bool snapWarp = !FIELD && !wide && (warp == SW_SNAP_WARP);
    const bool prepWarp = FIELD ? (warp == SW_FIELD_PREP_WARP) : (!wide && (warp == SW_PREP_WARP));
    const bool nbWarp = FIELD ? (warp == SW_FIELD_NB_WARP) : (!wide && (warp == SW_NB_WARP));
    const bool allHelperWarp = !FIELD && wide && (warp == SW_NB_WARP); /* wide layout: warp 12 builds snapshots, tables and neighbour data in turn */
    const int N = P.N, m = P.m;
    const int G = gridDim.x, cta = blockIdx.x;
    const int baseT = m / G, remT = m % G;
    const int T = baseT + (cta < remT ? 1 : 0);
    const int y0 = cta * baseT + min(cta, remT);
    const int maxT = baseT + (remT ? 1 : 0);
    const int nW = (N + K - 1) / K;
    const int CH = P.chunkElems, CPR = P.chunksPerRow, S = P.stages, NW = P.nw64;
    const int GPC = CH >> 7;

    const SweepSmem<real> L(maxT, NW, CH, S, K, P.dotWarps, FIELD ? P.ldF : 0);
    real *field = reinterpret_cast<real *>(smem + L.field);  /* FIELD: [maxT][ldF] local fields sum_i J[j][i] q[t][i] */
    real *ring = reinterpret_cast<real *>(smem + L.ring);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    unsigned long long *qcur = reinterpret_cast<unsigned long long *>(smem + L.qcur);
    unsigned long long *qsnap = reinterpret_cast<unsigned long long *>(smem + L.qsnap);
    unsigned long long *nbsnap = reinterpret_cast<unsigned long long *>(smem + L.nbsnap); /* [2 buffers][2 sides][NW] */
    real *dots = reinterpret_cast<real *>(smem + L.dots);    /* [2][maxT][K] */
    real *cross = reinterpret_cast<real *>(smem + L.cross);  /* [2][maxT][K][2K] */
    int *xs = reinterpret_cast<int *>(smem + L.xs);          /* [3][maxT][K] */
    int *xb = reinterpret_cast<int *>(smem + L.xb);          /* [3][maxT][K]: (32-bit word index << 5) | bit of spin x in a packed row */
    real *us = reinterpret_cast<real *>(smem + L.us);
    real *hs = reinterpret_cast<real *>(smem + L.hs);
    uint32_t *need = reinterpret_cast<uint32_t *>(smem + L.need); /* FIELD: [TAB][maxT][K]: lo 16 bits left neighbour, hi 16 right */
    real *carry = reinterpret_cast<real *>(smem + L.carry);  /* FIELD: [maxT][K] */
    real *pend = reinterpret_cast<real *>(smem + L.pend);    /* FIELD: [maxT][32] */
    unsigned char *cstate = smem + L.cstate;                 /* FIELD: [maxT][8 * sizeof(real)] */
    unsigned long long *flipQ = reinterpret_cast<unsigned long long *>(smem + L.flipQ); /* FIELD: [L.flipQLen] */
    int *xn = reinterpret_cast<int *>(smem + L.xn);          /* [2 sides][3][K] */
    uint32_t *conf = reinterpret_cast<uint32_t *>(smem + L.conf);       /* [TAB][2 sides][K], by window slot */
    uint32_t *confAny = reinterpret_cast<uint32_t *>(smem + L.confAny); /* [TAB][2 sides]: rounds with a non-empty mask */
    uint32_t *pubMask = confAny + 2 * TAB; /* [2 buffers][2 sides]: rounds of my edge trotter whose accept flag the neighbouring CTA may read */
    uint32_t *accLog = reinterpret_cast<uint32_t *>(smem + L.accLog);   /* [2 buffers][maxT]: accept bits of a window */
    uint32_t *sgnLog = reinterpret_cast<uint32_t *>(smem + L.sgnLog);   /* [2 buffers][maxT]: spin (1 = up) before each attempt of a window */
    uint32_t *tinfo = reinterpret_cast<uint32_t *>(smem + L.spec);      /* [maxT]: phase | (left local index + 1) << 2 | (right local index + 1) << 8 */
    uint32_t *front = tinfo + maxT;                                     /* [maxT]: rounds of the current window that are final */
    uint32_t *lconf = front + maxT;                                     /* [maxT][2 sides][K]: rounds of the local neighbour drawing the same spin */
    unsigned int *taskCounter = reinterpret_cast<unsigned int *>(smem + L.counter);

    /* ring topology: local index l <-> global trotter (yOff + l) mod mRing */
    const int mRing = P.mRing, yOff = P.yOff;
    const bool ringSharded = (mRing != m);
    auto gOf = [&](int l) { int g = yOff + l; return g >= mRing ? g - mRing : g; };
    auto slotOf = [&](int g) { /* hand-off slot of global trotter g: local index, or m / m+1 for the foreign neighbours */
        int dd = g - yOff;
        if (dd < 0) dd += mRing;
        if (dd < m) return dd;
        return (dd == mRing - 1) ? m : m + 1;
    };
    /* trotters of other CTAs (or GPUs) adjacent to this CTA's range (SQA only); global indices */
    const int gFirst = gOf(y0), gLast = gOf(y0 + T - 1);
    const int yLeft = (gFirst == 0) ? mRing - 1 : gFirst - 1;
    const int yRight = (gLast == mRing - 1) ? 0 : gLast + 1;
    const int slotL = slotOf(yLeft), slotR = slotOf(yRight);
    const bool remote = SQA && (G > 1 || ringSharded);

    auto roundsIn = [&](int w) { return min(K, N - w * K); };

    /* (x, -ln u, h[x]) of every attempt of window w for the owned trotters, plus the remote neighbours' x */
    auto prepWindow = [&](int w, int t0, int nthr) {
        if (w >= nW) return;
        const int Kw = roundsIn(w), slot = w & (TAB - 1);
        for (int idx = t0; idx < Kw * T; idx += nthr) {
            int t = idx % T, rl = idx / T;
            Philox4 p = sqbPhilox(seedR, P.step, DOM_DENSE_SWEEP, (uint32_t)(w * K + rl), (uint32_t)gOf(y0 + t));
            int x = (int)(p.w[0] % (uint32_t)N);
            int o = (slot * maxT + t) * K + rl;
            xs[o] = x;
            {
                int w64, bit;
                spinBitPos(x, w64, bit);
                xb[o] = ((2 * w64 + (bit >> 5)) << 5) | (bit & 31);
            }
            us[o] = negLogUniform<real>(p); /* accept iff dE*beta < -ln(u): no exp on the chain's critical path */
            if (!FIELD) hs[o] = hr[x]; /* field mode keeps h inside its field rows */
        }
        if (remote) {
            for (int idx = t0; idx < 2 * Kw; idx += nthr) {
                int side = idx / Kw, rl = idx % Kw;
                Philox4 p = sqbPhilox(seedR, P.step, DOM_DENSE_SWEEP, (uint32_t)(w * K + rl), (uint32_t)(side ? yRight : yLeft));
                xn[(side * TAB + slot) * K + rl] = (int)(p.w[0] % (uint32_t)N);
            }
        }
    };

    /* field mode: the tables of window w come from the pre-pass (sweepTablesKernel); copy the record into table slot w & (TAB-1) */
    auto loadWindow = [&](int w, int t0, int nthr) {
        if (w >= nW) return;
        const SweepTabRec<real> R(maxT, K);
        const unsigned char *rec = P.tables + (((size_t)blockIdx.y * G + cta) * nW + w) * R.bytes;
        const int slot = w & (TAB - 1);
        const int nTK = maxT * K;
        const int *gxs = reinterpret_cast<const int *>(rec + R.xs), *gxb = reinterpret_cast<const int *>(rec + R.xb), *gxn = reinterpret_cast<const int *>(rec + R.xn);
        const uint32_t *gneed = reinterpret_cast<const uint32_t *>(rec + R.need), *gconf = reinterpret_cast<const uint32_t *>(rec + R.conf);
        const uint32_t *gmisc = reinterpret_cast<const uint32_t *>(rec + R.misc);
        const real *gus = reinterpret_cast<const real *>(rec + R.us);
        for (int i = t0; i < nTK; i += nthr) {
            xs[slot * nTK + i] = __ldcg(gxs + i);
            xb[slot * nTK + i] = __ldcg(gxb + i);
            need[slot * nTK + i] = __ldcg(gneed + i);
            us[slot * nTK + i] = __ldcg(gus + i);
        }
        for (int i = t0; i < 2 * K; i += nthr) {
            xn[((i / K) * TAB + slot) * K + (i % K)] = __ldcg(gxn + i);
            conf[slot * 2 * K + i] = __ldcg(gconf + i);
        }
        if (t0 < 2) { confAny[slot * 2 + t0] = __ldcg(gmisc + t0); pubMask[slot * 2 + t0] = __ldcg(gmisc + 2 + t0); }
    };

    unsigned long long nWaits = 0;
    long long waited = 0; /* cycles lane 0 of this warp spent waiting for another warp or CTA */

    /* helper warp: what chain window wn needs from the neighbouring CTAs -- their snapshot S_{wn-1} (spins of the two foreign
     * neighbour trotters before window wn-1), into buffer wn & 1 while the chain replays window wn - 1 out of the other one.
     * A neighbour hands the flips of a window over as ONE 64-bit word -- tag << 16 | accept bits, written by its chain warp when the
     * window ends; the spin indices are its Philox draws, which this CTA tabulates itself.  Buffer wn & 1 still holds S_{wn-3}
     * (built for window wn-2): the flips of windows wn-3 and wn-2 bring it up to date, so nothing is copied. */
    unsigned long long nbPrevWord = 0ull; /* lanes 0 / 1: the left / right neighbour's word of the window before the newest one */
    int nbPrevX = 0;                      /* lane (side, j): that window's draw j of the neighbour (its table slot may be gone by now) */
    auto neighbourWindow = [&](int wn) {
        if (!remote || wn < 2) return;
        const int bN = wn & 1;
        unsigned long long *dstB = nbsnap + (size_t)(bN * 2) * NW;
        unsigned long long word = 0ull;
        if (lane < 2) { /* lane 0 / 1 wait for the left / right neighbour's word of window wn-2 */
            const int sl = lane ? slotR : slotL;
            const unsigned long long want = P.snapBase + (unsigned long long)(wn - 1);
            const unsigned long long *f = sBits + ((size_t)sl * SW_SNAP_SLOTS + ((wn - 2) % SW_SNAP_SLOTS)) * NW;
            const long long t0 = clock64();
            unsigned ns = 20;
            word = ringSharded ? ldRelaxedSys(f) : ldRelaxed(f);
            while ((word >> 16) != want) {
                ++nWaits;
                __nanosleep(ns);
                if (ns < 160u) ns <<= 1;
                word = ringSharded ? ldRelaxedSys(f) : ldRelaxed(f);
            }
            if (lane == 0) waited += clock64() - t0;
        }
        __syncwarp();
        const int sideF = lane / K, jF = lane % K;
        const int xNew = (sideF < 2) ? xn[(sideF * TAB + ((wn - 2) & (TAB - 1))) * K + jF] : 0;
#pragma unroll
        for (int age = 0; age < 2; ++age) { /* age 0: window wn-2 (the new word), age 1: window wn-3 (last call's word and draws) */
            if (wn - 2 - age < 0) break;
            const unsigned long long wd = age ? nbPrevWord : word;
            const uint32_t mL = (uint32_t)__shfl_sync(0xffffffffu, wd, 0) & 0xffffu, mR = (uint32_t)__shfl_sync(0xffffffffu, wd, 1) & 0xffffu;
            if (sideF < 2 && (((sideF ? mR : mL) >> jF) & 1u)) {
                int w64, bit;
                spinBitPos(age ? nbPrevX : xNew, w64, bit);
                atomicXor(reinterpret_cast<unsigned int *>(dstB + (size_t)sideF * NW + w64) + (bit >> 5), 1u << (bit & 31));
            }
        }
        nbPrevX = xNew;
        nbPrevWord = word;
        __syncwarp();
    };
    /* Conflict masks of window v for my two edge trotters (they depend on Philox draws only, so they are built ahead of time, right
     * after the tables of window v+1): per round r, which of the foreign neighbour's attempts of windows v-1 and v drew the same
     * spin index (conf), and whether its attempts of windows v / v+1 did -- only then will it ever read my accept flag (pubMask).
     * Lane (side, r) holds my draw and the neighbour's draw r of the three windows; K lane-addressed shuffles compare all pairs. */
    auto maskWindow = [&](int v) {
        if (!remote || v >= nW) return;
        const int Kv = roundsIn(v), sv = v & (TAB - 1);
        const int Kq = (v + 1 < nW) ? roundsIn(v + 1) : 0;
        const int side = lane / K, r = lane % K;
        const bool act = side < 2;
        const uint32_t kM = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        const int xe = (act && r < Kv) ? xs[(sv * maxT + (side ? T - 1 : 0)) * K + r] : -1 - lane;
        const int nP = (act && v > 0) ? xn[(side * TAB + ((v - 1) & (TAB - 1))) * K + r] : -100 - lane;
        const int nC = (act && r < Kv) ? xn[(side * TAB + sv) * K + r] : -100 - lane;
        const int nQ = (act && r < Kq) ? xn[(side * TAB + ((v + 1) & (TAB - 1))) * K + r] : -100 - lane;
        uint32_t mP = 0u, mC = 0u, mQ = 0u;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int src = (lane - r) + j;
            mP |= (__shfl_sync(0xffffffffu, nP, src) == xe ? 1u : 0u) << j;
            mC |= (__shfl_sync(0xffffffffu, nC, src) == xe ? 1u : 0u) << j;
            mQ |= (__shfl_sync(0xffffffffu, nQ, src) == xe ? 1u : 0u) << j;
        }
        const bool mine = act && r < Kv;
        const uint32_t mask = mine ? (mP | (mC << K)) : 0u;
        if (mine) conf[(sv * 2 + side) * K + r] = mask;
        const uint32_t nz = __ballot_sync(0xffffffffu, mask != 0u);
        const uint32_t pb = __ballot_sync(0xffffffffu, mine && (mC | mQ) != 0u);
        if (lane < 2) {
            confAny[sv * 2 + lane] = (nz >> (lane * K)) & kM;
            pubMask[sv * 2 + lane] = (pb >> (lane * K)) & kM;
        }
        __syncwarp();
    };

    /* ---------------- setup (all warps) ---------------- */
    for (int i = tid; i < maxT * NW; i += SW_THREADS) qcur[i] = 0ull;
    for (int i = tid; i < 4 * NW; i += SW_THREADS) nbsnap[i] = 0ull;
    if (tid == 0) {
        *taskCounter = 0u;
        for (int i = 0; i < P.dotWarps * S; ++i) mbarInit(&bars[i], 1);
        mbarInitFence();
    }
    __syncthreads();
    if (ringSharded && remote && (slotL >= m || slotR >= m)) {
        /* a neighbour lives on another GPU: its spins at step start arrive through the halo push of that GPU */
        if (tid == 0) {
            if (slotL >= m) while (ldAcquireSys(P.stepFlags + 0) < P.stepEpoch) __nanosleep(100);
            if (slotR >= m) while (ldAcquireSys(P.stepFlags + 1) < P.stepEpoch) __nanosleep(100);
        }
        __syncthreads();
    }
    {   /* pack int8 spins -> bits; 4 spins (one nibble) per thread step */
        const int n4 = (N + 3) >> 2;
        const int rows = T + (remote ? 2 : 0);
        for (int idx = tid; idx < rows * n4; idx += SW_THREADS) {
            int r = idx / n4, j = (idx % n4) << 2;
            const signed char *rowp;
            if (r < T) rowp = qBase + (size_t)(y0 + r) * P.ldq;
            else {
                const int sl = (r == T) ? slotL : slotR;
                rowp = (sl < m) ? qBase + (size_t)sl * P.ldq : P.haloQ[sl - m];
            }
            const signed char *src = rowp + j;
            unsigned nib = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j + e < N && src[e] > 0) nib |= 1u << e;
            int w64, bit;
            spinBitPos(j, w64, bit);
            unsigned long long *dst = (r < T) ? (qcur + (size_t)r * NW) : (nbsnap + (size_t)(r - T) * NW);
            atomicOr(&dst[w64], (unsigned long long)nib << bit);
        }
    }
    if (FIELD) { for (int w0 = 0; w0 < 4; ++w0) loadWindow(w0, tid, SW_THREADS); }
    else { prepWindow(0, tid, SW_THREADS); prepWindow(1, tid, SW_THREADS); prepWindow(2, tid, SW_THREADS); }
    real *const Fg = FIELD ? P.F + ((size_t)replica * m + y0) * P.ldF : NULL; /* this CTA's rows of the field matrix */
    if (FIELD) { /* the field rows in shared memory hold H[t][j] = h[j] + 2 sum_i J[j][i] q_t[i]; ldF is a multiple of 128 elements */
        if (P.fieldHasH) {
            const int n16 = (int)((size_t)T * P.ldF * sizeof(real) / 16);
            const int4 *src = reinterpret_cast<const int4 *>(Fg);
            int4 *dst = reinterpret_cast<int4 *>(field);
            for (int i = tid; i < n16; i += SW_THREADS) dst[i] = __ldcg(src + i);
        } else {
            for (int i = tid; i < T * P.ldF; i += SW_THREADS) {
                const int j = i % P.ldF;
                field[i] = (j < N ? hr[j] : real(0)) + real(2) * __ldcg(Fg + i);
            }
        }
    }
    __syncthreads();
    if (!FIELD) for (int i = tid; i < T * NW; i += SW_THREADS) qsnap[i] = qcur[i];
    for (int i = tid; i < 2 * NW; i += SW_THREADS) nbsnap[2 * NW + i] = nbsnap[i]; /* windows 0 and 1 both start from S_0 */
    if (!FIELD && warp == SW_NB_WARP) { /* conflict masks of the windows whose successor's tables exist (field mode: from the pre-pass) */
        constexpr int PW = 3;
        for (int v = 0; v < PW - 1; ++v) maskWindow(v);
        if (nW <= PW) maskWindow(PW - 1);
    }
    if (FIELD) { /* per-trotter chain state: phase and local neighbours, frontier (rounds of the step that are final), carries */
        if (tid < T) {
            const int gy = gOf(y0 + tid);
            const int yl = slotOf(gy == 0 ? mRing - 1 : gy - 1), yr = slotOf(gy == mRing - 1 ? 0 : gy + 1);
            const bool lLocal = (yl >= y0 && yl < y0 + T), rLocal = (yr >= y0 && yr < y0 + T);
            tinfo[tid] = (uint32_t)sweepPhase(gy, mRing) | ((lLocal ? (uint32_t)(yl - y0 + 1) : 0u) << 2) | ((rLocal ? (uint32_t)(yr - y0 + 1) : 0u) << 8);
            front[tid] = 0u;
            uint32_t *cs = reinterpret_cast<uint32_t *>(cstate + (size_t)tid * 8 * sizeof(real));
            cs[0] = 0u; cs[1] = 0u; cs[2] = 0u; cs[3] = 0u;
        }
        for (int i = tid; i < T * K; i += SW_THREADS) carry[i] = real(0);
        for (int i = tid; i < L.flipQLen; i += SW_THREADS) flipQ[i] = 0ull;
    }

    /* ---------------- hand-off counters between the warps of this CTA (shared memory, release/acquire at CTA scope) ------
     * There is no CTA-wide barrier inside the sweep: every warp runs its own loop over the windows and waits only for what it
     * consumes.
     *   rowsDone[b]  rows of the windows of parity b whose dot product + cross terms are stored      (dot warps -> chain)
     *   replayDone   windows replayed                                                                  (chain -> helper)
     *   snapCount    snapshots S_0 .. S_{c-1} built; S_k = spins before window k, kept in qsnap[k & 1]  (helper -> dot, prep)
     *   nbCount      windows whose neighbour snapshot + conflict masks are in place                     (helper -> chain)
     *   prepCount    windows whose (x, -ln u, h) tables are in place (4 slots)                          (prep -> dot, helper)
     * Rows of window w are reduced against S_{w-1} (S_0 for w = 0), i.e. they can start as soon as window w-2 has been
     * replayed, one full window before the chain needs them. */
    const uint32_t aSync = smemAddr(taskCounter);
    /* rowsDone[2], nbCount, prepCount share one aligned 16-byte line: the accept chain polls all of them with ONE 128-bit load */
    const uint32_t aRowsDone = aSync + 16, aReplayDone = aSync + 4, aSnapCount = aSync + 8, aNbCount = aSync + 24, aPrepCount = aSync + 28;
    auto waitCount = [&](uint32_t addr, uint32_t want, unsigned ns) { /* whole warp; lane 0 polls.  ns == 0: latency-critical (accept chain) */
        const bool chainPoll = (ns == 0u); /* fence-free polls on the accept chain, see ldVolatileCta */
        if (lane == 0 && (chainPoll ? ldVolatileCta(addr) : ldAcquireCta(addr)) < want) {
            const long long t0 = clock64();
            /* a waiting warp must not eat the issue slots of the warps it waits for (they share its scheduler): sleep between
             * polls, 32..128 ns for the chain, ns..8 ns for everybody else */
            const unsigned nsMax = ns ? ns * 8u : 128u;
            if (!ns) ns = 32u;
            while ((chainPoll ? ldVolatileCta(addr) : ldAcquireCta(addr)) < want) {
                __nanosleep(ns);
                if (ns < nsMax) ns <<= 1;
            }
            waited += clock64() - t0;
        }
        __syncwarp();
    };
    auto signalCount = [&](uint32_t addr, uint32_t v) { /* whole warp: everything the warp wrote is visible before the count */
        __syncwarp();
        if (lane == 0) stReleaseCta(addr, v);
    };

    /* ---------------- dot warps: per-warp TMA ring state ---------------- */
    /* task g = w * (K*T) + id, id = rl * T + t.  Tasks are CLAIMED dynamically (shared counter) by whichever dot warp is about to
     * issue a new row, so a warp slowed down by HBM/L2 queueing simply takes fewer rows.  Claims are monotone, hence in
     * window order; the TMA ring keeps streaming across window boundaries (rows are known from Philox alone). */
    const int TPW = K * T;                                   /* tasks per full window */
    const int totalTasks = (nW - 1) * TPW + roundsIn(nW - 1) * T;
    int ic = 0, ix = 0;                     /* issue cursor (lane 0): chunk within the row being issued, its row index */
    int fifo0 = -1, fifo1 = -1;             /* lane 0: claimed tasks not yet consumed (oldest first) */
    int iStage = 0;                         /* ring slot the next issue goes to */
    int cStage = 0;                         /* ring slot the next consume reads, and its mbarrier phase parity */
    uint32_t cParity = 0;
    bool issueDone = false;
    uint64_t *myBars = bars + (dotWarp ? dw : 0) * S;
    real *myRing = ring + (size_t)(dotWarp ? dw : 0) * S * CH;

    auto issueNext = [&]() { /* lane 0 of a dot warp */
        if (issueDone) return;
        if (ic == 0) {
            if (fifo1 >= 0) return; /* two rows pending already (rows shorter than the ring): claim again after the next pop */
            const int g = (int)atomicAdd(taskCounter, 1u);
            if (g >= totalTasks) { issueDone = true; return; }
            if (fifo0 < 0) fifo0 = g; else fifo1 = g;
            const int iw = g / TPW, iid = g - iw * TPW;
            const int t = iid % T, rl = iid / T;
            if (ldAcquireCta(aPrepCount) > (uint32_t)iw) { /* the window's tables are already in place */
                ix = xs[((iw & (TAB - 1)) * maxT + t) * K + rl];
            } else {
                Philox4 p = sqbPhilox(seedR, P.step, DOM_DENSE_SWEEP, (uint32_t)(iw * K + rl), (uint32_t)gOf(y0 + t));
                ix = (int)(p.w[0] % (uint32_t)N);
            }
        }
        const int elems = min(CH, P.ldJ - ic * CH);
        const uint32_t bytes = (uint32_t)(elems * sizeof(real));
        mbarArriveExpectTx(&myBars[iStage], bytes);
        tmaLoad1D(myRing + (size_t)iStage * CH, Jr + (size_t)ix * P.ldJ + (size_t)ic * CH, bytes, &myBars[iStage]);
        if (++iStage == S) iStage = 0;
        if (++ic == CPR) ic = 0;
    };

    /* reduce the row of task g (window w) against the snapshot the window is defined on; results go to buffer w & 1 */
    auto dotRow = [&](int g, int w) {
        const int buf = w & 1, slot = w & (TAB - 1);
        const int id = g - w * TPW;
        const int t = id % T, rl = id / T;
        /* column whose J[x][col] this lane must pick up: lane j < K -> round j of window w-1, else round j-K of w */
        int px = -1;
        if (lane < K) { if (w > 0) px = xs[(((w - 1) & (TAB - 1)) * maxT + t) * K + lane]; }
        else if (lane < 2 * K && lane - K < rl) px = xs[(slot * maxT + t) * K + (lane - K)];
        real crossv = real(0);
        typename Acc4<real>::type acc;
        const unsigned long long *qrow = qsnap + ((size_t)((w > 0) ? ((w - 1) & 1) : 0) * maxT + t) * NW;
        for (int c = 0; c < CPR; ++c) {
            mbarWait(&myBars[cStage], cParity);
            const real *buf_ = myRing + (size_t)cStage * CH;
            const int c0 = c * CH;
            const int groups = min(GPC, (P.ldJ - c0) >> 7);
            const int g0 = c * GPC;
            unsigned long long bits = qrow[((g0 >> 4) << 5) + lane] >> ((g0 & 15) << 2);
            const real *src = buf_ + lane * 4;
            if (groups == 16) {
                const uint32_t blo = (uint32_t)bits, bhi = (uint32_t)(bits >> 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) accumGroup(src + i * 128, (blo >> (4 * i)) & 0xfu, acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) accumGroup(src + (i + 8) * 128, (bhi >> (4 * i)) & 0xfu, acc);
            } else if (groups == 8) {
                const uint32_t blo = (uint32_t)bits;
#pragma unroll
                for (int i = 0; i < 8; ++i) accumGroup(src + i * 128, (blo >> (4 * i)) & 0xfu, acc);
            } else {
                for (int i = 0; i < groups; ++i) accumGroup(src + i * 128, (uint32_t)(bits >> (4 * i)) & 0xfu, acc);
            }
            if (px >= c0 && px < c0 + (groups << 7)) crossv = buf_[px - c0];
            __syncwarp();
            if (lane == 0) {
                if (c == CPR - 1) { fifo0 = fifo1; fifo1 = -1; } /* this row is done: make room before claiming */
                issueNext();
            }
            if (++cStage == S) { cStage = 0; cParity ^= 1u; }
        }
        real s = warpSum(acc.sum());
        if (lane == 0) dots[(buf * maxT + t) * K + rl] = P.scaleA * (hs[(slot * maxT + t) * K + rl] + real(2) * s);
        if (lane < 2 * K) cross[((buf * maxT + t) * K + rl) * (2 * K) + lane] = crossv;
        __syncwarp();
        if (lane == 0) redAddReleaseCta(aRowsDone + 4u * (uint32_t)buf, 1u);
    };

    /* S_w for the dot warps (shared memory, double buffered) and for the neighbouring CTAs (global memory) */
    auto snapshotWindow = [&](int w) {
        if (!FIELD) { /* S_w = S_{w-1} with the accepted flips of window w-1 (field mode has no use for snapshots) */
            const unsigned long long *src = qsnap + (size_t)((w - 1) & 1) * maxT * NW;
            unsigned long long *dst = qsnap + (size_t)(w & 1) * maxT * NW;
            for (int i = lane; i < T * NW; i += 32) dst[i] = src[i];
            __syncwarp();
            if (lane < T) {
                uint32_t bitsAcc = accLog[((w - 1) & 1) * maxT + lane];
                const int *xbRow = xb + (((w - 1) & (TAB - 1)) * maxT + lane) * K;
                uint32_t *row = reinterpret_cast<uint32_t *>(dst + (size_t)lane * NW);
                while (bitsAcc) {
                    const int rl = __ffs(bitsAcc) - 1;
                    bitsAcc &= bitsAcc - 1;
                    const int xbv = xbRow[rl];
                    row[xbv >> 5] ^= 1u << (xbv & 31);
                }
            }
        }
        signalCount(aSnapCount, (uint32_t)w + 1u);
    };

    /* ---------------- field mode: what the dot warps do instead of streaming one J row per attempt ----------------
     * The local fields F[t][j] = sum_i J[j][i] q_t[i] of the owned trotters live in shared memory.  Field warp d owns the
     * 128-column groups g = d (mod #field warps) of every row.  For chain window w it
     *   1. folds the flips ACCEPTED in window w-2 into its columns: F[t][.] -= 2 q_old J[x][.] (J symmetric) -- the only
     *      full-row traffic, acceptance-rate x one row per attempt; the chain warp that accepted the flip has already asked
     *      L2 for the row (prefetch at commit time, one to two windows earlier),
     *   2. hands the chain scaleA (h[x] + 2 F[t][x]) for the attempts of window w whose column it owns.
     * F then holds exactly the flips of windows <= w-2; the chain adds the cross terms J[x'][x] of the flips accepted since
     * (windows w-1 and w), which it gathers itself when it commits a flip -- 2K four-byte loads per ACCEPTED flip instead
     * of 2K-1 per attempt. */
    const int nDot = P.dotWarps;
    const int nGroups = FIELD ? (P.ldF >> 7) : 0;
    /* queue of committed flips: entry = (index + 1) << 32 | payload; payload = marker << 31 | sign << 30 | trotter << 24 | x
     * (marker entries: window index in the low bits).  Producers take a slot with an atomic add and publish the entry with one
     * 64-bit store; every field warp reads the queue in order at its own pace. */
    const uint32_t aFlipQ = smemAddr(flipQ);
    const uint32_t qMask = (uint32_t)L.flipQLen - 1u;
    auto pushFlip = [&](uint32_t payload) { /* one lane */
        const uint32_t slot = atomicAdd(reinterpret_cast<unsigned int *>(taskCounter) + 3, 1u);
        const unsigned long long e = ((unsigned long long)(slot + 1u) << 32) | payload;
        asm volatile("st.relaxed.cta.shared.u64 [%0], %1;" ::"r"(aFlipQ + ((slot & qMask) << 3)), "l"(e) : "memory");
    };
    /* H[t][.] += c J[x][.] on this warp's column groups (J symmetric), split into the loads and the shared-memory update so
     * that the loads of the next queued flip are in flight while this one is applied */
    typedef typename RowVec4<real>::type RowVec;
    auto loadFlip = [&](int x, RowVec (&v)[8], int g0) {
        const real *Jrow = Jr + (size_t)x * P.ldJ + lane * 4;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = g0 + u * nDot;
            if (g < nGroups) loadRow4(Jrow + (size_t)g * 128, v[u]);
        }
    };
    auto storeFlip = [&](int t, real c, const RowVec (&v)[8], int g0) {
        real *Frow = field + (size_t)t * P.ldF + lane * 4;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = g0 + u * nDot;
            if (g < nGroups) axpyRow4(Frow + (size_t)g * 128, c, v[u]);
        }
    };
    auto applyFlip = [&](int t, int x, real c) {
        for (int g0 = dw; g0 < nGroups; g0 += nDot * 8) {
            RowVec v[8];
            loadFlip(x, v, g0);
            storeFlip(t, c, v, g0);
        }
    };
    auto fieldDots = [&](int w) { /* scaleA H[t][x] for the attempts of window w whose column this warp owns */
        const int Kw = roundsIn(w), buf = w & 1, slot = w & (TAB - 1);
        const int nEnt = T * Kw;
        __syncwarp();
        for (int e = lane; e < nEnt; e += 32) {
            const int t = e % T, rl = e / T;
            const int x = xs[(slot * maxT + t) * K + rl];
            if ((x >> 7) % nDot == dw) dots[(buf * maxT + t) * K + rl] = P.scaleA * field[(size_t)t * P.ldF + x];
        }
        /* one count per field warp and window (parity buffers): the chain starts window w at nDot * (w / 2 + 1) */
        __syncwarp();
        if (lane == 0) redAddReleaseCta(aRowsDone + 4u * (uint32_t)buf, 1u);
    };

    if (tid == 0) {
        stReleaseCta(aRowsDone, 0u); stReleaseCta(aRowsDone + 4, 0u); stReleaseCta(aReplayDone, 0u);
        stReleaseCta(aSnapCount, 1u); stReleaseCta(aNbCount, 1u); stReleaseCta(aPrepCount, FIELD ? 4u : 3u);
        stReleaseCta(aSync + 12, 0u); /* tail of the flip queue (field mode) */
    }
    __syncthreads();
    const long long tLoop0 = clock64();

    /* CTAs with fewer trotters than the heaviest ones keep proportionally fewer rows in flight: when the sweep is bound by
     * HBM, bandwidth is shared in proportion to the bytes each SM has outstanding, and all CTAs should finish a window at
     * the same time (512 trotters on 148 SMs: 4 or 3 per CTA -> 12 or 9 streaming warps). */
    const int activeDotWarps = max(1, (P.dotWarps * T + maxT - 1) / maxT);
    if (FIELD && dotWarp) {
        /* The field rows hold every flip of the windows <= w when the local fields of window w + 2 are read: the chain warps
         * queue each flip the moment they commit it (so its row streams in while the chain goes on) and a marker when a window
         * is complete; a field warp applies the flips in queue order -- per trotter that is commit order, so every element sees
         * the same sequence of additions whatever the timing -- and on the marker of window w hands over the fields of w + 2. */
        fieldDots(0);
        if (nW > 1) fieldDots(1);
        uint32_t pos = 0;
        for (;;) {
            unsigned long long e = 0ull;
            if (lane == 0) {
                const uint32_t a = aFlipQ + ((pos & qMask) << 3);
                unsigned ns = 16u;
                const long long t0 = clock64();
                bool slept = false;
                for (;;) {
                    asm volatile("ld.relaxed.cta.shared.u64 %0, [%1];" : "=l"(e) : "r"(a) : "memory");
                    if ((uint32_t)(e >> 32) == pos + 1u) break;
                    slept = true;
                    __nanosleep(ns);
                    if (ns < 128u) ns <<= 1;
                }
                if (slept) waited += clock64() - t0;
            }
            const uint32_t pl = (uint32_t)__shfl_sync(0xffffffffu, e, 0);
            ++pos;
            if (pl >> 31) { /* marker: window wm is complete */
                const int wm = (int)(pl & 0x7fffffffu);
                if (wm + 2 < nW) {
                    waitCount(aPrepCount, (uint32_t)wm + 3u, 20); /* tables of window wm + 2 */
                    fieldDots(wm + 2);
                }
                if (wm == nW - 1) break;
            } else {
                const int t0 = (int)((pl >> 24) & 63u), x0 = (int)(pl & 0xffffffu);
                const real c0 = ((pl >> 30) & 1u) ? real(-4) : real(4); /* q_old = +1: h + 2 sum loses 4 J */
                bool done2 = false;
                if constexpr (sizeof(real) == 4) if (nGroups <= nDot * 8) { /* one batch of loads per flip: overlap it with the next queued flip, if there is one already (fp32: 64 registers) */
                    done2 = true;
                    RowVec va[8], vb[8];
                    loadFlip(x0, va, dw);
                    unsigned long long e2 = 0ull;
                    if (lane == 0) asm volatile("ld.relaxed.cta.shared.u64 %0, [%1];" : "=l"(e2) : "r"(aFlipQ + ((pos & qMask) << 3)) : "memory");
                    e2 = __shfl_sync(0xffffffffu, e2, 0);
                    const uint32_t pl2 = (uint32_t)e2;
                    const bool two = ((uint32_t)(e2 >> 32) == pos + 1u) && !(pl2 >> 31);
                    if (two) {
                        ++pos;
                        loadFlip((int)(pl2 & 0xffffffu), vb, dw);
                    }
                    storeFlip(t0, c0, va, dw);
                    if (two) storeFlip((int)((pl2 >> 24) & 63u), ((pl2 >> 30) & 1u) ? real(-4) : real(4), vb, dw);
                }
                if (!done2) applyFlip(t0, x0, c0);
            }
        }
        if (P.writeBackF) { /* H matches the final spins: back to global memory for the next step */
            __syncwarp();
            for (int t = 0; t < T; ++t)
                for (int g = dw; g < nGroups; g += nDot) {
                    const size_t o = (size_t)t * P.ldF + (size_t)g * 128 + lane * 4;
                    if (sizeof(real) == 4) *reinterpret_cast<int4 *>(Fg + o) = *reinterpret_cast<const int4 *>(field + o);
                    else {
                        *reinterpret_cast<int4 *>(Fg + o) = *reinterpret_cast<const int4 *>(field + o);
                        *(reinterpret_cast<int4 *>(Fg + o) + 1) = *(reinterpret_cast<const int4 *>(field + o) + 1);
                    }
                }
        }
    } else if (!FIELD && dotWarp && dw < activeDotWarps) {
        if (lane == 0)
            for (int s = 0; s < S; ++s) issueNext();
        int curW = -1;
        for (;;) {
            const int g = __shfl_sync(0xffffffffu, fifo0, 0); /* oldest row this warp has claimed and not yet reduced */
            if (g < 0) break;
            const int w = g / TPW;
            if (w != curW) { /* first row of a later window: its snapshot S_{w-1} and its tables must be in place */
                waitCount(aSnapCount, (uint32_t)w, 20);
                waitCount(aPrepCount, (uint32_t)w + 1u, 20);
                curW = w;
            }
            dotRow(g, w);
        }
    } else if (snapWarp) {
        for (int w = 1; w < nW; ++w) { /* as soon as window w-1 has been replayed */
            waitCount(aReplayDone, (uint32_t)w, 20);
            snapshotWindow(w);
        }
    } else if (allHelperWarp) {
        /* wide layout: after window w-1 has been replayed -- S_w, then the tables of window w+2 (slot of window w-2, dead
         * once S_w exists), then the neighbour data of window w+1 */
        if (1 < nW) { neighbourWindow(1); signalCount(aNbCount, 2u); }
        for (int w = 1; w < nW; ++w) {
            waitCount(aReplayDone, (uint32_t)w, 20);
            snapshotWindow(w);
            if (w + 2 < nW) { /* tables of window w+2, then the conflict masks of window w+1 (they look one window ahead) */
                prepWindow(w + 2, lane, 32);
                __syncwarp();
                maskWindow(w + 1);
                if (w + 2 == nW - 1) maskWindow(w + 2);
                signalCount(aPrepCount, (uint32_t)w + 3u);
            }
            if (w + 1 < nW) { neighbourWindow(w + 1); signalCount(aNbCount, (uint32_t)w + 2u); }
        }
    } else if (nbWarp) {
        /* the neighbours' S_{wn-1} and the conflict masks of window wn, into the buffers window wn-2 has finished with */
        long long nbW0 = 0, nbW1 = 0, nbW2 = 0;
        for (int wn = 1; wn < nW; ++wn) {
            const long long a0 = waited;
            waitCount(aReplayDone, (uint32_t)(wn - 1), 20);
            const long long a1 = waited;
            waitCount(aPrepCount, (uint32_t)min(wn + 2, nW), 20); /* the conflict masks of window wn are in place (the chain reads them) */
            const long long a2 = waited;
            neighbourWindow(wn);
            nbW0 += a1 - a0; nbW1 += a2 - a1; nbW2 += waited - a2;
            signalCount(aNbCount, (uint32_t)wn + 1u);
        }
        if (FIELD && P.stats && lane == 0 && blockIdx.y == 0) { /* per-CTA profile: the neighbour warp's waits (own chain, tables, remote words) */
            unsigned long long *pc = P.stats + 16 + 16 * (size_t)cta;
            pc[8] = (unsigned long long)nbW0; pc[9] = (unsigned long long)nbW1; pc[10] = (unsigned long long)nbW2;
            pc[11] = (unsigned long long)(clock64() - tLoop0);
        }
    } else if (prepWarp) {
        /* tables of window wp go to the slot of window wp-4, dead once S_{wp-2} is built (window wp-3 replayed) */
        long long prepT0 = 0, prepT1 = 0;
        for (int wp = FIELD ? 4 : 3; wp < nW; ++wp) {
            /* field mode: eight slots, one more window of look-ahead (the dot warps still read window wp-4's slot then) */
            if (FIELD) waitCount(aReplayDone, (uint32_t)wp - 3u, 20); /* window wp-4 replayed: the slot of window wp-8 is dead */
            else waitCount(aSnapCount, (uint32_t)wp - 1u, 20);
            if (!FIELD && remote) waitCount(aNbCount, (uint32_t)wp - 1u, 20); /* the neighbour warp reads window wp-4's draws for window wp-2 */
            const long long c0 = clock64();
            if (FIELD) loadWindow(wp, lane, 32); /* one record of the table pre-pass */
            else prepWindow(wp, lane, 32);
            __syncwarp();
            const long long c1 = clock64();
            if (!FIELD) {
                maskWindow(wp - 1); /* the conflict masks look one window ahead */
                if (wp == nW - 1) maskWindow(wp);
            }
            prepT0 += c1 - c0; prepT1 += clock64() - c1;
            signalCount(aPrepCount, (uint32_t)wp + 1u);
        }
        if (FIELD && P.stats && lane == 0 && blockIdx.y == 0) { /* per-CTA profile: cycles in the Philox tables / the conflict masks / waiting */
            unsigned long long *pc = P.stats + 16 + 16 * (size_t)cta;
            pc[12] = (unsigned long long)prepT0; pc[13] = (unsigned long long)prepT1; pc[14] = (unsigned long long)waited;
        }
    } else if (FIELD && chainWarp) {
        /* ---------------- field mode: the accept chain, one warp per trotter, a whole window evaluated at once ----------------
         * Chain warp cw owns the trotters t = cw (mod 4).  Lanes 0..K-1 hold the K attempts of the current window of one
         * trotter, lanes K..2K-1 the attempts of the NEXT window (they only take part in the cross-term gathers).
         *   evaluate   every attempt at or behind the trotter's frontier is tested against the current spins and fields.  The
         *              attempts in front of the first "stop" are rejections whose inputs were exact: they are final.
         *   stop       (a) an accept: committed -- spin flipped, frontier behind it, cross terms J[x'][x_r] of the trotter's later
         *              attempts of this window and of all attempts of the next one gathered with one asynchronous 4-byte copy
         *              per lane (cp.async), the row prefetched into L2 for the field warps;
         *              (b) an attempt that has to wait because an EARLIER attempt of a neighbouring trotter on the same spin
         *              index is not final yet (rare: same index within a window);
         *              (c) an attempt whose sign is not certain while the gathers of the last commit are in flight: every
         *              affected attempt carries the bound 4 scaleA max|J[x'][.]| and is decided only if its margin exceeds it;
         *              otherwise the warp waits for the copies, applies them and evaluates again.
         * Only accepted flips serialise, and the HBM latency of their cross terms is hidden behind the attempts that are certain
         * anyway.  Commit order per trotter is the reference order; attempts of neighbouring trotters on the same spin index are
         * ordered by the frontier protocol (the later one waits for the earlier one), so the chain is the reference chain exactly.
         * Trotters of one CTA run on different warps and meet at the end of every window (named barrier 1). */
        if constexpr (FIELD) {
        constexpr int CW = SW_FIELD_CHAIN_WARPS;
        const int cw = warp;
        const int rI = lane % K, hI = lane / K;
        const uint32_t rowBytes = (uint32_t)NW * 8u;
        const uint32_t aMy0 = smemAddr(qcur), aNb = smemAddr(nbsnap), aFront = smemAddr(front), aDots = smemAddr(dots);
        const uint32_t aPend = smemAddr(pend), aCarry = smemAddr(carry);
        const int nbPhaseL = sweepPhase(yLeft, mRing), nbPhaseR = sweepPhase(yRight, mRing);
        const real corrScale = real(-4) * P.scaleA; /* a flip of spin x' changes scaleA (h + 2 sum) of a later attempt on x by -4 scaleA q_old J[x'][x] */
        const real nbScale2 = real(2) * P.scaleNb;
        const int rowLines = (int)((size_t)P.ldJ * sizeof(real) / 128);
        const uint32_t kMask = (K == 32) ? 0xffffffffu : ((1u << K) - 1u);
        unsigned long long nAccepted = 0;
        /* profile of chain warp 0 (stats[8..15]): cycles in gather waits / idle polls / the window barrier, evaluation passes,
         * resolves forced by an uncertain attempt / by a second commit, passes that ended on a blocked attempt */
        long long cycGather = 0, cycIdle = 0, cycBar = 0, cycStart = 0, cycEval = 0, cycEnd = 0, waitedRows = 0, waitedNbF = 0;
        unsigned long long nEval = 0, nUncRes = 0, nCommitRes = 0, nBlkStop = 0;
        auto cst = [&](int t) { return reinterpret_cast<uint32_t *>(cstate + (size_t)t * 8 * sizeof(real)); }; /* [0] pending window + 1, [1] its round, [2] accept bits, [3] sign bits */
        auto cstR = [&](int t) { return reinterpret_cast<real *>(cstate + (size_t)t * 8 * sizeof(real) + 16); }; /* [0] bound, [1] signed scale of the pending generation */

        /* ---- fast path: at most one trotter per chain warp (T <= 4, e.g. 512 trotters on 148 SMs).  Everything a window needs
         * per attempt is held in registers -- lane r < K: round r of the current window, lane K + r: round r of the next window
         * (cross-term gathers, local conflict masks of the right-hand neighbour) -- so that an evaluation pass is three
         * state-dependent shared-memory loads (own spin word, the two neighbours' words), a handful of ALU instructions and a vote. */
        const bool fast = (T <= CW);
        const int tF = cw;
        const bool haveT = fast && (cw < T);
        const uint32_t infoF = haveT ? tinfo[tF] : 0u;
        const int phF = (int)(infoF & 3u), tnLF = (int)((infoF >> 2) & 63u) - 1, tnRF = (int)((infoF >> 8) & 63u) - 1;
        const bool edgeLF = haveT && remote && tnLF < 0, edgeRF = haveT && remote && tnRF < 0;
        uint32_t pwF = 0u, pr0F = 0u;   /* pending generation of gathers: window it was issued in + 1 (0: none), its round */
        real pSF = real(0), carryN = real(0); /* its signed scale; lanes K..2K-1: corrections known so far for the next window's rounds */

        for (int w = 0; w < nW; ++w) {
            const int Kw = roundsIn(w), KwN = (w + 1 < nW) ? roundsIn(w + 1) : 0;
            const int buf = w & 1, slot = w & (TAB - 1), slotN = (w + 1) & (TAB - 1);
            const uint32_t wBase = (uint32_t)(w * K);
            {   /* what the window needs: its local fields (field warps), the foreign neighbours' snapshot, the tables of windows w and
                 * w+1 (the gathers of a commit look one window ahead) -- one 128-bit poll of the four counters */
                const uint32_t wantRows = (uint32_t)(P.dotWarps * ((w >> 1) + 1)), wantNb = remote ? (uint32_t)w + 1u : 0u, wantPrep = (uint32_t)min(w + 2, nW);
                if (lane == 0) {
                    uint32_t c0, c1, c2, c3;
                    unsigned ns = 32u;
                    long long tFirst = 0, tRows = 0;
                    for (;;) {
                        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3) : "r"(aRowsDone) : "memory");
                        const bool okRows = (buf ? c1 : c0) >= wantRows;
                        if (okRows && tFirst && !tRows) tRows = clock64();
                        if (okRows && c2 >= wantNb && c3 >= wantPrep) break;
                        if (!tFirst) tFirst = clock64();
                        __nanosleep(ns);
                        if (ns < 128u) ns <<= 1;
                    }
                    if (tFirst) { /* split the wait: until the fields were there / the rest (neighbour data, tables) */
                        const long long tEnd = clock64();
                        if (!tRows) tRows = tEnd;
                        waited += tEnd - tFirst; waitedRows += tRows - tFirst; waitedNbF += tEnd - tRows;
                    }
                }
                __syncwarp();
            }
            const unsigned long long flagBase = (P.roundBase + (unsigned long long)w * K + 1ull) << 1;
            const long long rrBase = (long long)w * K - K; /* round index of bit 0 of a remote conflict mask */
            const int fs0 = (w * K) % SW_FLAG_RING;
            const unsigned long long *nbRowsW = nbsnap + (size_t)buf * 2 * NW;

            if (fast) {
            const long long tw0 = clock64();
            if (haveT) {
                const int t = tF;
                const int o = (slot * maxT + t) * K + rI;
                /* per-window constants of my attempt */
                const uint32_t xbv = (uint32_t)xb[o];
                const uint32_t aw = (xbv >> 5) << 2, bit = xbv & 31u;
                const uint32_t aOwn = aMy0 + (uint32_t)t * rowBytes + aw;
                const uint32_t aL = (tnLF >= 0 ? aMy0 + (uint32_t)tnLF * rowBytes : aNb + (uint32_t)(buf * 2) * rowBytes) + aw;
                const uint32_t aR = (tnRF >= 0 ? aMy0 + (uint32_t)tnRF * rowBytes : aNb + (uint32_t)(buf * 2 + 1) * rowBytes) + aw;
                const real lnu = us[o];
                const int xMine = xs[o];
                const int xGather = (hI == 0) ? xMine : ((hI == 1 && rI < KwN) ? xs[(slotN * maxT + t) * K + rI] : -1);
                uint32_t needL = 0u, needR = 0u, cmL = 0u, cmR = 0u;
                if (SQA) {
                    if (hI == 0 && rI < Kw) { /* local neighbours (table pre-pass): their earlier attempts on my spin index */
                        const uint32_t nd = need[o];
                        needL = nd & 0xffffu; needR = nd >> 16;
                    }
                    if (hI == 0 && rI < Kw) { /* neighbours owned by other CTAs: attempts of theirs that precede mine on the same spin index */
                        const uint32_t precBase = (w > 0 ? kMask : 0u) | (((1u << rI) - 1u) << K);
                        if (edgeLF) cmL = conf[(slot * 2 + 0) * K + rI] & (precBase | (((nbPhaseL < phF) ? 1u : 0u) << (K + rI)));
                        if (edgeRF) cmR = conf[(slot * 2 + 1) * K + rI] & (precBase | (((nbPhaseR < phF) ? 1u : 0u) << (K + rI)));
                    }
                }
                const bool rare = (needL | needR | cmL | cmR) != 0u;
                const uint32_t pmaskW = ((edgeLF && t == 0) ? pubMask[slot * 2] : 0u) | ((edgeRF && t == T - 1) ? pubMask[slot * 2 + 1] : 0u);
                if (pwF != 0u && pwF != (uint32_t)w) { cpAsyncWaitAll(); pwF = 0u; } /* issued two windows ago: the field rows have the flip by now */
                real v;
                ldsReal(aDots + (uint32_t)(((buf * maxT + t) * K + rI) * sizeof(real)), v);
                v += __shfl_down_sync(0xffffffffu, carryN, K); /* corrections gathered during the previous window */
                carryN = real(0);
                uint32_t accC = 0u, sgnC = 0u, fr = 0u;
                const long long tw1 = clock64();
                cycStart += tw1 - tw0;

                while (fr < (uint32_t)Kw) {
                    const bool valid = (hI == 0) && (rI < Kw) && ((uint32_t)rI >= fr);
                    uint32_t code = 0u; /* 0: final rejection, 1: blocked, 2: uncertain, 3: accept */
                    uint32_t up = 0u, wv = 0u;
                    if (valid) {
                        wv = ldsU32(aOwn);
                        up = (wv >> bit) & 1u;
                        real vv = v;
                        bool blk = false;
                        if (SQA) {
                            uint32_t lb = (ldsU32(aL) >> bit) & 1u, rb = (ldsU32(aR) >> bit) & 1u;
                            if (rare) { /* a neighbour draws this spin index within the look-back range: its earlier attempts must be final */
                                if (needL) {
                                    const uint32_t fn = ldAcquireCta(aFront + 4u * (uint32_t)tnLF) - wBase;
                                    if (needL & ~((fn >= 32u) ? 0xffffffffu : ((1u << fn) - 1u))) blk = true;
                                    lb = (ldsU32(aL) >> bit) & 1u;
                                }
                                if (needR) {
                                    const uint32_t fn = ldAcquireCta(aFront + 4u * (uint32_t)tnRF) - wBase;
                                    if (needR & ~((fn >= 32u) ? 0xffffffffu : ((1u << fn) - 1u))) blk = true;
                                    rb = (ldsU32(aR) >> bit) & 1u;
                                }
                                if (cmL) {
                                    const int q = remoteBitTry(nbRowsW, xMine, cmL, rrBase, P.roundBase, aFlags + (size_t)slotL * SW_FLAG_RING, 0);
                                    if (q < 0) blk = true; else lb = (uint32_t)q;
                                }
                                if (cmR) {
                                    const int q = remoteBitTry(nbRowsW + NW, xMine, cmR, rrBase, P.roundBase, aFlags + (size_t)slotR * SW_FLAG_RING, 0);
                                    if (q < 0) blk = true; else rb = (uint32_t)q;
                                }
                            }
                            vv -= nbScale2 * real((int)(lb + rb) - 1);
                        }
                        if (blk) code = 1u;
                        else {
                            const real sv = up ? vv : -vv;
                            const bool affected = (pwF != 0u) && (pwF != (uint32_t)w + 1u || (uint32_t)rI > pr0F);
                            if (affected && fabs(sv - lnu) <= P.uncBound + real(1e-5) * fabs(sv)) code = 2u;
                            else if (sv < lnu) code = 3u; /* exp(-dE beta) > u */
                        }
                    }
                    const uint32_t stopBits = __ballot_sync(0xffffffffu, code != 0u) & kMask;
                    ++nEval;
                    uint32_t newFront = (uint32_t)Kw;
                    uint32_t kind = 0u;
                    int rs = Kw;
                    if (stopBits) {
                        rs = __ffs(stopBits) - 1;
                        kind = __shfl_sync(0xffffffffu, code, rs);
                        newFront = (uint32_t)rs;
                        if (pwF != 0u && kind >= 2u) { /* the gathers in flight must land before an uncertain attempt is decided / before the next commit */
                            const long long tg = clock64();
                            cpAsyncWaitAll();
                            cycGather += clock64() - tg;
                            real val;
                            ldsReal(aPend + (uint32_t)((t * 32 + lane) * sizeof(real)), val);
                            val *= pSF;
                            const real fromNext = __shfl_down_sync(0xffffffffu, (hI == 1) ? val : real(0), K);
                            if (pwF == (uint32_t)w + 1u) { /* issued in this window: lanes < K later rounds of it, lanes K.. the next window */
                                if (hI == 0) v += val; else if (hI == 1) carryN += val;
                            } else if (hI == 0) v += fromNext; /* issued in the previous window: its "next window" is this one */
                            pwF = 0u;
                        }
                        if (kind == 3u) { /* commit */
                            const uint32_t upj = __shfl_sync(0xffffffffu, up, rs);
                            if (lane == rs) stsU32(aOwn, wv ^ (1u << bit));
                            const int xF = __shfl_sync(0xffffffffu, xMine, rs);
                            const real *Jrow = Jr + (size_t)xF * P.ldJ;
                            const bool want = (hI == 0) ? (rI > rs && rI < Kw) : (xGather >= 0);
                            const uint32_t aP = aPend + (uint32_t)((t * 32 + lane) * sizeof(real));
                            if (want) cpAsyncReal(aP, Jrow + xGather); else stsReal(aP, real(0));
                            cpAsyncCommit();
                            if (lane == 1) prefetchL2Bulk(Jrow, (uint32_t)((size_t)P.ldJ * sizeof(real))); /* the whole row, for the field warps */
                            pwF = (uint32_t)w + 1u; pr0F = (uint32_t)rs;
                            pSF = upj ? corrScale : -corrScale;
                            accC |= 1u << rs; sgnC |= upj << rs;
                            if (lane == 0) { pushFlip((upj << 30) | ((uint32_t)t << 24) | (uint32_t)xF); ++nAccepted; }
                            newFront = (uint32_t)rs + 1u;
                        }
                    }
                    if (pmaskW) { /* rounds that became final and whose accept flag a neighbouring CTA may read */
                        uint32_t pm = pmaskW & ((newFront >= 32u) ? 0xffffffffu : ((1u << newFront) - 1u)) & ~((1u << fr) - 1u);
                        if (lane == 0) {
                            unsigned long long *myFlags = aFlags + (size_t)(y0 + t) * SW_FLAG_RING;
                            while (pm) {
                                const int r = __ffs(pm) - 1;
                                pm &= pm - 1;
                                stRelaxed(myFlags + (fs0 + r) % SW_FLAG_RING, flagBase + (unsigned long long)(2 * r) + ((kind == 3u && r == rs) ? 1ull : 0ull));
                            }
                        }
                    }
                    if (newFront != fr) {
                        __syncwarp(); /* the flipped spin word is written before the frontier moves */
                        if (lane == 0) stReleaseCta(aFront + 4u * (uint32_t)t, wBase + newFront);
                        fr = newFront;
                    } else if (kind == 1u) { /* blocked on another warp or CTA */
                        const long long ti = clock64();
                        ++nWaits;
                        __nanosleep(20);
                        cycIdle += clock64() - ti;
                    }
                }
                const long long tw2 = clock64();
                cycEval += tw2 - tw1;
                if (lane == 0 && pmaskW == 0u) {} /* (flags are only published where readable) */
                if (lane == 0 && remote && (t == 0 || t == T - 1)) /* the neighbouring CTAs rebuild this trotter's spins from the accept bits of the window */
                    stRelaxed(sBits + ((size_t)(y0 + t) * SW_SNAP_SLOTS + (size_t)(w % SW_SNAP_SLOTS)) * NW,
                              ((P.snapBase + (unsigned long long)w + 1ull) << 16) | (unsigned long long)accC);
                (void)sgnC;
                cycEnd += clock64() - tw2;
            }
            } else {
            /* window start: the corrections carried over from the previous window, the local conflict masks, the logs */
            const long long tw0 = clock64();
            for (int t = cw; t < T; t += CW) {
                uint32_t *cs = cst(t);
                if (cs[0] != 0u && cs[0] != (uint32_t)w) { /* a generation issued two windows ago: F has the flip by now */
                    cpAsyncWaitAll();
                    __syncwarp();
                    if (lane == 0) cs[0] = 0u;
                }
                if (hI == 0 && rI < Kw) {
                    const int o = (buf * maxT + t) * K + rI;
                    dots[o] += carry[t * K + rI];
                    carry[t * K + rI] = real(0);
                }
                if (lane == 0) { cs[2] = 0u; cs[3] = 0u; }
            }
            __syncwarp();
            const long long tw1 = clock64();
            cycStart += tw1 - tw0;

            for (;;) {
                bool allDone = true, progress = false;
                for (int t = cw; t < T; t += CW) {
                    const uint32_t fr = front[t] - wBase; /* rounds of this window that are final (only this warp writes it) */
                    if (fr >= (uint32_t)Kw) continue;
                    allDone = false;
                    uint32_t *cs = cst(t);
                    const uint32_t pw = cs[0], pr0 = cs[1];
                    const real pB = cstR(t)[0];
                    const uint32_t info = tinfo[t];
                    const int o = (slot * maxT + t) * K + rI;
                    const bool valid = (hI == 0) && (rI < Kw) && ((uint32_t)rI >= fr);
                    bool stop = false, blk = false, unc = false;
                    uint32_t up = 0, wv = 0, aw = 0, bit = 0;
                    if (valid) {
                        const uint32_t xbv = (uint32_t)xb[o];
                        aw = (xbv >> 5) << 2; bit = xbv & 31u;
                        wv = ldsU32(aMy0 + (uint32_t)t * rowBytes + aw);
                        up = (wv >> bit) & 1u;
                        real vv;
                        ldsReal(aDots + (uint32_t)(((buf * maxT + t) * K + rI) * sizeof(real)), vv);
                        if (SQA) {
                            const int ph = (int)(info & 3u);
                            int nb = 0;
#pragma unroll
                            for (int side = 0; side < 2; ++side) {
                                const int tn = (int)((info >> (2 + 6 * side)) & 63u) - 1;
                                uint32_t nbit;
                                if (tn >= 0) {
                                    const uint32_t ndm = (need[o] >> (16 * side)) & 0xffffu; /* table pre-pass: its earlier attempts on this spin index */
                                    if (ndm) { /* rare: they must be final -- frontier first, spin word after */
                                        const uint32_t fn = ldAcquireCta(aFront + 4u * (uint32_t)tn) - wBase;
                                        const uint32_t fin = (fn >= 32u) ? 0xffffffffu : ((1u << fn) - 1u);
                                        if (ndm & ~fin) blk = true;
                                    }
                                    nbit = (ldsU32(aMy0 + (uint32_t)tn * rowBytes + aw) >> bit) & 1u;
                                } else {
                                    nbit = (ldsU32(aNb + (uint32_t)(buf * 2 + side) * rowBytes + aw) >> bit) & 1u;
                                    const uint32_t cm = conf[(slot * 2 + side) * K + rI];
                                    if (cm) { /* rare: so does the neighbour owned by another CTA -- its accept flags decide */
                                        const uint32_t prec = (w > 0 ? kMask : 0u) | (((1u << rI) - 1u) << K) |
                                                              ((((side ? nbPhaseR : nbPhaseL) < ph) ? 1u : 0u) << (K + rI));
                                        if (cm & prec) {
                                            const int rb = remoteBitTry(nbRowsW + (size_t)side * NW, xs[o], cm & prec, rrBase, P.roundBase,
                                                                        aFlags + (size_t)(side ? slotR : slotL) * SW_FLAG_RING, 0);
                                            if (rb < 0) blk = true; else nbit = (uint32_t)rb;
                                        }
                                    }
                                }
                                nb += (int)nbit;
                            }
                            vv -= nbScale2 * real(nb - 1);
                        }
                        if (!blk) {
                            const real sv = up ? vv : -vv, lnu = us[o];
                            /* gathers in flight change vv by at most pB: issued this window -> the rounds after the commit,
                             * issued in the previous window -> every round */
                            const bool affected = (pw != 0u) && (pw != (uint32_t)w + 1u || (uint32_t)rI > pr0);
                            if (affected && fabs(sv - lnu) <= pB + real(1e-5) * fabs(sv)) unc = true;
                            else stop = sv < lnu; /* exp(-dE beta) > u */
                        }
                    }
                    const uint32_t stopBits = __ballot_sync(0xffffffffu, stop || blk || unc) & kMask;
                    const uint32_t blkBits = __ballot_sync(0xffffffffu, blk), uncBits = __ballot_sync(0xffffffffu, unc);
                    const uint32_t upBits = __ballot_sync(0xffffffffu, up != 0u);
                    const int rs = stopBits ? (__ffs(stopBits) - 1) : Kw;
                    const bool isBlk = stopBits && ((blkBits >> rs) & 1u), isUnc = stopBits && !isBlk && ((uncBits >> rs) & 1u);
                    const bool commit = stopBits && !isBlk && !isUnc;
                    ++nEval;
                    if (isBlk) ++nBlkStop;
                    if (pw != 0u && (isUnc || commit)) {
                        /* the copies of the pending generation must land: lane < K -> later rounds of the window it was issued in,
                         * lanes K..2K-1 -> the rounds of the window after that one */
                        if (isUnc) ++nUncRes; else ++nCommitRes;
                        const long long tg = clock64();
                        cpAsyncWaitAll();
                        cycGather += clock64() - tg;
                        real val;
                        ldsReal(aPend + (uint32_t)((t * 32 + lane) * sizeof(real)), val);
                        val *= cstR(t)[1];
                        if (pw == (uint32_t)w + 1u) {
                            if (hI == 0 && rI < Kw) dots[(buf * maxT + t) * K + rI] += val;
                            else if (hI == 1 && rI < KwN) carry[t * K + rI] += val;
                        } else if (hI == 1 && rI < Kw) dots[(buf * maxT + t) * K + rI] += val;
                        __syncwarp();
                        if (lane == 0) cs[0] = 0u;
                        progress = true;
                    }
                    uint32_t newFront = (uint32_t)rs;
                    if (commit) {
                        const uint32_t upj = (upBits >> rs) & 1u;
                        if (lane == rs) stsU32(aMy0 + (uint32_t)t * rowBytes + aw, wv ^ (1u << bit));
                        const int oS = (slot * maxT + t) * K + rs;
                        const real *Jrow = Jr + (size_t)xs[oS] * P.ldJ;
                        int xt = -1;
                        if (hI == 0) { if (rI > rs && rI < Kw) xt = xs[o]; }
                        else if (hI == 1) { if (rI < KwN) xt = xs[(slotN * maxT + t) * K + rI]; }
                        const uint32_t aP = aPend + (uint32_t)((t * 32 + lane) * sizeof(real));
                        if (xt >= 0) cpAsyncReal(aP, Jrow + xt); else stsReal(aP, real(0));
                        cpAsyncCommit();
                        if (lane == 1) prefetchL2Bulk(Jrow, (uint32_t)((size_t)P.ldJ * sizeof(real)));
                        if (lane == 0) {
                            cs[0] = (uint32_t)w + 1u; cs[1] = (uint32_t)rs;
                            cs[2] |= 1u << rs; cs[3] |= upj << rs;
                            cstR(t)[0] = P.uncBound;
                            pushFlip((upj << 30) | ((uint32_t)t << 24) | (uint32_t)xs[oS]);
                            cstR(t)[1] = upj ? corrScale : -corrScale;
                            ++nAccepted;
                        }
                        newFront = (uint32_t)rs + 1u;
                    }
                    if (lane == 0 && remote && (t == 0 || t == T - 1)) { /* rounds that became final and whose accept flag a neighbouring CTA may read */
                        uint32_t pmask = ((t == 0 && !((info >> 2) & 63u)) ? pubMask[slot * 2] : 0u) | ((t == T - 1 && !((info >> 8) & 63u)) ? pubMask[slot * 2 + 1] : 0u);
                        pmask &= ((newFront >= 32u) ? 0xffffffffu : ((1u << newFront) - 1u)) & ~((1u << fr) - 1u);
                        unsigned long long *myFlags = aFlags + (size_t)(y0 + t) * SW_FLAG_RING;
                        while (pmask) {
                            const int r = __ffs(pmask) - 1;
                            pmask &= pmask - 1;
                            stRelaxed(myFlags + (fs0 + r) % SW_FLAG_RING, flagBase + (unsigned long long)(2 * r) + ((commit && r == rs) ? 1ull : 0ull));
                        }
                    }
                    __syncwarp(); /* the flipped spin word and the state words are written before the frontier moves */
                    if (newFront != fr) {
                        progress = true;
                        if (lane == 0) stReleaseCta(aFront + 4u * (uint32_t)t, wBase + newFront);
                    }
                    __syncwarp();
                }
                if (allDone) break;
                if (!progress) { /* everything left waits for another warp or CTA */
                    const long long ti = clock64();
                    ++nWaits;
                    __nanosleep(20);
                    cycIdle += clock64() - ti;
                }
            }

            const long long tw2 = clock64();
            cycEval += tw2 - tw1;
            for (int t = cw; t < T; t += CW) {
                if (lane == 0) {
                    const uint32_t accC = cst(t)[2];
                    if (remote && (t == 0 || t == T - 1)) /* the neighbouring CTAs rebuild this trotter's spins from the accept bits of the window */
                        stRelaxed(sBits + ((size_t)(y0 + t) * SW_SNAP_SLOTS + (size_t)(w % SW_SNAP_SLOTS)) * NW,
                                  ((P.snapBase + (unsigned long long)w + 1ull) << 16) | (unsigned long long)accC);
                }
            }
            }
            __syncwarp();
            const long long tb = clock64();
            namedBarSync(1, 32 * CW); /* every trotter of the CTA has finished the window */
            cycBar += clock64() - tb;
            if (cw == 0) {
                if (lane == 0) pushFlip(0x80000000u | (uint32_t)w); /* after every flip of the window */
                signalCount(aReplayDone, (uint32_t)w + 1u);
            }
        }
        cpAsyncWaitAll();
        if (P.stats) {
            if (lane == 0 && nAccepted) atomicAdd(P.stats, nAccepted);
            if (lane == 0 && cw == 0) {
                atomicAdd(P.stats + 4, (unsigned long long)waitedRows);                      /* waiting for the field warps */
                atomicAdd(P.stats + 7, (unsigned long long)waitedNbF);                       /* waiting for neighbour data (other CTAs) */
                atomicAdd(P.stats + 6, (unsigned long long)(waited - waitedRows - waitedNbF)); /* waiting for the Philox tables */
                atomicAdd(P.stats + 8, (unsigned long long)cycGather);
                atomicAdd(P.stats + 9, (unsigned long long)cycIdle);
                atomicAdd(P.stats + 10, (unsigned long long)cycBar);
                atomicAdd(P.stats + 11, nEval);
                atomicAdd(P.stats + 12, (unsigned long long)cycStart);
                atomicAdd(P.stats + 13, (unsigned long long)cycEval);
                atomicAdd(P.stats + 14, (unsigned long long)cycEnd);
            }
            if (lane == 0 && cw != 0) atomicAdd(P.stats + 15, (unsigned long long)cycBar); /* chain warps 1..3: cycles waiting in the window barrier */
            if (lane == 0 && cw == 0 && blockIdx.y == 0) { /* per-CTA profile of the last launch (chain warp 0) */
                unsigned long long *pc = P.stats + 16 + 16 * (size_t)cta;
                unsigned long long gt;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
                pc[0] = (unsigned long long)T; pc[1] = (unsigned long long)waitedRows; pc[2] = (unsigned long long)waitedNbF;
                pc[3] = (unsigned long long)(cycStart + cycEval + cycEnd); pc[4] = (unsigned long long)((clock64() - tLoop0));
                pc[5] = gt; pc[6] = nAccepted; pc[7] = nEval;
            }
        }
        } /* if constexpr (FIELD) */
    } else if (!FIELD && chainWarp) {
        /* ---------------- the accept chain: lane = trotter ---------------- */
        const bool active = (lane < T);
        const int tl = active ? lane : 0;
        const int gy = gOf(y0 + tl);             /* global trotter */
        const int myPhase = active ? sweepPhase(gy, mRing) : -1;
        /* phases this CTA has to run per round: 0 and 2, plus 1 where trotter m-1 of an odd ring lives (a vote result, so
         * that it stays in a register instead of being re-derived from the kernel parameters every round) */
        const int phStep = __any_sync(0xffffffffu, myPhase == 1) ? 1 : 2;
        const int yl = slotOf(gy == 0 ? mRing - 1 : gy - 1), yr = slotOf(gy == mRing - 1 ? 0 : gy + 1); /* slots of the neighbours */
        const bool lLocal = (yl >= y0 && yl < y0 + T), rLocal = (yr >= y0 && yr < y0 + T);
        const bool remoteLane = remote && active && (!lLocal || !rLocal);
        const bool publishes = remote && active && (lane == 0 || lane == T - 1);
        const uint32_t rowBytes = (uint32_t)NW * 8u;
        const uint32_t aMy0 = smemAddr(qcur);
        const uint32_t aMy = aMy0 + (uint32_t)tl * rowBytes;
        const uint32_t aLeftLocal = smemAddr(qcur) + (uint32_t)(lLocal ? yl - y0 : 0) * rowBytes;
        const uint32_t aRightLocal = smemAddr(qcur) + (uint32_t)(rLocal ? yr - y0 : 0) * rowBytes;
        const uint32_t aNb = smemAddr(nbsnap), aDots = smemAddr(dots), aCross = smemAddr(cross), aXb = smemAddr(xb), aUs = smemAddr(us);
        const uint32_t aXs = smemAddr(xs), aConf = smemAddr(conf);
        const int nbPhaseL = sweepPhase(yLeft, mRing), nbPhaseR = sweepPhase(yRight, mRing);
        unsigned long long *myFlags = aFlags + (size_t)(y0 + tl) * SW_FLAG_RING;
        /* the first / last trotter of a sharded ring also publishes into the neighbouring GPU's arrays */
        unsigned long long *mirror0 = (ringSharded && active && y0 + lane == 0 && P.peerFlags[0]) ? P.peerFlags[0] + (size_t)(m + 1) * SW_FLAG_RING : NULL;
        unsigned long long *mirror1 = (ringSharded && active && y0 + lane == m - 1 && P.peerFlags[1]) ? P.peerFlags[1] + (size_t)m * SW_FLAG_RING : NULL;
        unsigned long long *const snapWord = sBits + (size_t)(y0 + tl) * SW_SNAP_SLOTS * NW; /* + (w & 3) * NW: this trotter's accept word of window w */
        unsigned long long *snapMirror0 = (ringSharded && active && y0 + lane == 0 && P.peerSnapBits[0]) ? P.peerSnapBits[0] + (size_t)(m + 1) * SW_SNAP_SLOTS * NW : NULL;
        unsigned long long *snapMirror1 = (ringSharded && active && y0 + lane == m - 1 && P.peerSnapBits[1]) ? P.peerSnapBits[1] + (size_t)m * SW_SNAP_SLOTS * NW : NULL;
        const real corrScale = real(-4) * P.scaleA; /* a flip of spin x' accepted since the snapshot changes sum by -2 q_old J[x][x'] */
        const real nbScale2 = real(2) * P.scaleNb;
        if (active) tinfo[lane] = (uint32_t)myPhase | ((lLocal ? (uint32_t)(yl - y0 + 1) : 0u) << 2) | ((rLocal ? (uint32_t)(yr - y0 + 1) : 0u) << 8);
        __syncwarp();
        uint32_t accP = 0, sgnP = 0;
        unsigned long long nAccepted = 0;
        unsigned long long slowWaits = 0; /* flag polls of the out-of-line conflict path (its address is taken: lives in local memory) */
        long long waitedNb = 0;

        for (int w = 0; w < nW; ++w) {
            const int Kw = roundsIn(w), buf = w & 1, slot = w & (TAB - 1);
            const unsigned long long *nbRows = nbsnap + (size_t)buf * 2 * NW;
            /* every row of this window reduced (and, through it, the window's tables in place); neighbour data in place */
            waitCount(aRowsDone + 4u * (uint32_t)buf, FIELD ? (uint32_t)(P.dotWarps * ((w >> 1) + 1)) : (uint32_t)((w >> 1) * TPW + Kw * T), 0);
            const long long waitedRows = waited;
            if (remote) waitCount(aNbCount, (uint32_t)w + 1u, 0);
            waitedNb += waited - waitedRows;

            /* 1. fold the flips of window w-1 into the snapshot dot products of window w: one (trotter, round) item per
             *    lane, so this is off the serial path */
            if (__any_sync(0xffffffffu, accP != 0u)) {
                for (int i0 = 0; i0 < T * K; i0 += 32) {
                    const int i = i0 + lane;
                    const int t = min(i / K, T - 1), rl = i % K;
                    uint32_t ev = __shfl_sync(0xffffffffu, accP, t);
                    const uint32_t sg = __shfl_sync(0xffffffffu, sgnP, t);
                    if (i < T * K && rl < Kw && ev) {
                        const uint32_t aV = aDots + (uint32_t)(((buf * maxT + t) * K + rl) * sizeof(real));
                        const uint32_t aC = aCross + (uint32_t)(((((FIELD ? (w & 3) : buf) * maxT + t) * K + rl) * (2 * K)) * sizeof(real));
                        real v, c;
                        ldsReal(aV, v);
                        do {
                            const int j = __ffs(ev) - 1;
                            ev &= ev - 1;
                            ldsReal(aC + (uint32_t)(j * sizeof(real)), c);
                            v += (((sg >> j) & 1u) ? corrScale : -corrScale) * c;
                        } while (ev);
                        stsReal(aV, v);
                    }
                }
                __syncwarp();
            }

            /* 2. replay.  Per round: the state-independent part (tables) for all lanes, then the state-dependent core once per
             *    phase of the reference order (even y, [y = m-1 of an odd ring], odd y) -- ONE copy of the core in a loop over
             *    the phases, rare paths out of line, so that the loop fits the instruction cache -- then, if a flip was
             *    accepted, the local fields of the trotter's later attempts in this window are repaired by all 32 lanes. */
            const uint32_t aLeft = lLocal ? aLeftLocal : aNb + (uint32_t)(buf * 2) * rowBytes;
            const uint32_t aRight = rLocal ? aRightLocal : aNb + (uint32_t)(buf * 2 + 1) * rowBytes;
            uint32_t cmask = 0; /* rounds in which a neighbour owned by another CTA attempts the same spin index */
            if (remoteLane) cmask = (lLocal ? 0u : confAny[slot * 2]) | (rLocal ? 0u : confAny[slot * 2 + 1]);
            uint32_t pmask = 0; /* rounds whose accept flag a neighbouring CTA may read */
            if (publishes) pmask = ((lane == 0 && !lLocal) ? pubMask[slot * 2] : 0u) | ((lane == T - 1 && !rLocal) ? pubMask[slot * 2 + 1] : 0u);
            uint32_t pXb = aXb + (uint32_t)(((slot * maxT + tl) * K) * 4);
            uint32_t pUs = aUs + (uint32_t)(((slot * maxT + tl) * K) * sizeof(real));
            uint32_t pDot = aDots + (uint32_t)(((buf * maxT + tl) * K) * sizeof(real));
            const uint32_t aDotsW = aDots + (uint32_t)((buf * maxT * K) * sizeof(real));                             /* dots of this window, [t][round] */
            const uint32_t aCrossW = aCross + (uint32_t)(((FIELD ? (w & 3) : buf) * maxT * K * 2 * K + K) * sizeof(real)); /* this window's columns of [t][round][2K] */
            const uint32_t aXsW = aXs + (uint32_t)(((slot * maxT + tl) * K) * 4);
            const uint32_t aConfW = aConf + (uint32_t)((slot * 2 * K) * 4);
            const unsigned long long flagBase = (P.roundBase + (unsigned long long)w * K + 1ull) << 1;
            const long long rrBase = (long long)w * K - K; /* round index of bit 0 of a conflict mask */
            int fs = (w * K) % SW_FLAG_RING;
            uint32_t accC = 0, sgnC = 0;
            if (P.specChain) {
                /* Window-parallel accept chain.  Every attempt (t, r) of the window is evaluated at once -- lane = (trotter, round)
                 * -- against the current spins and fields.  Per trotter, the attempts before its first "stop" are rejections whose
                 * inputs were exact, so they are final; a stop is either an accept (committed: spin flipped, the trotter's later
                 * fields repaired, frontier behind it) or an attempt that must wait because an EARLIER attempt of a neighbouring
                 * trotter on the same spin index is not final yet (frontier stays in front of it).  Only accepted flips serialise:
                 * the loop runs (accepted flips per trotter and window) + 1 times instead of once per round, and reproduces the
                 * sequential reference order exactly. */
                constexpr int TPP = 32 / K; /* trotters per pass */
                const int nPass = (T + TPP - 1) / TPP;
                const int rI = lane % K, jI = lane / K;
                const int sysFlag = ringSharded ? 1 : 0;
                const unsigned long long *nbRowsW = nbsnap + (size_t)buf * 2 * NW;
                if (active) front[lane] = 0u;
                if (SQA) {
                    for (int pass = 0; pass < nPass; ++pass) { /* rounds of the local neighbours that draw the same spin index */
                        const int t = pass * TPP + jI;
                        if (t < T && rI < Kw) {
                            const uint32_t info = tinfo[t];
                            const int x = xs[(slot * maxT + t) * K + rI];
#pragma unroll
                            for (int side = 0; side < 2; ++side) {
                                const int tn = (int)((info >> (2 + 6 * side)) & 63u) - 1;
                                uint32_t msk = 0;
                                if (tn >= 0) {
                                    const int *xo = xs + (slot * maxT + tn) * K;
#pragma unroll
                                    for (int j = 0; j < K; ++j) msk |= ((j < Kw && xo[j] == x) ? 1u : 0u) << j;
                                }
                                lconf[(t * 2 + side) * K + rI] = msk;
                            }
                        }
                    }
                }
                __syncwarp();
                for (;;) {
                    bool progress = false;
                    for (int pass = 0; pass < nPass; ++pass) {
                        const int t = pass * TPP + jI;
                        const bool valid = (t < T) && (rI < Kw);
                        bool stop = false, blk = false;
                        uint32_t up = 0, wv = 0, aw = 0, bit = 0;
                        if (valid && (uint32_t)rI >= front[t]) {
                            const int o = (slot * maxT + t) * K + rI;
                            const uint32_t xbv = (uint32_t)xb[o];
                            aw = (xbv >> 5) << 2; bit = xbv & 31u;
                            wv = ldsU32(aMy0 + (uint32_t)t * rowBytes + aw);
                            up = (wv >> bit) & 1u;
                            real vv = dots[(buf * maxT + t) * K + rI];
                            if (SQA) {
                                const uint32_t info = tinfo[t];
                                const int ph = (int)(info & 3u);
                                int nb = 0;
#pragma unroll
                                for (int side = 0; side < 2; ++side) {
                                    const int tn = (int)((info >> (2 + 6 * side)) & 63u) - 1;
                                    uint32_t nbit;
                                    if (tn >= 0) {
                                        nbit = (ldsU32(aMy0 + (uint32_t)tn * rowBytes + aw) >> bit) & 1u;
                                        const uint32_t lm = lconf[(t * 2 + side) * K + rI];
                                        if (lm) { /* rare: the neighbour draws this spin index in this window too */
                                            const uint32_t prec = ((1u << rI) - 1u) | ((((int)(tinfo[tn] & 3u) < ph) ? 1u : 0u) << rI);
                                            if (lm & prec & ~((1u << front[tn]) - 1u)) blk = true; /* an earlier one is not final yet */
                                        }
                                    } else {
                                        nbit = (ldsU32(aNb + (uint32_t)(buf * 2 + side) * rowBytes + aw) >> bit) & 1u;
                                        const uint32_t cm = conf[(slot * 2 + side) * K + rI];
                                        if (cm) { /* rare: so does the neighbour owned by another CTA -- its accept flags decide */
                                            const uint32_t prec = (w > 0 ? ((1u << K) - 1u) : 0u) | (((1u << rI) - 1u) << K) |
                                                                  ((((side ? nbPhaseR : nbPhaseL) < ph) ? 1u : 0u) << (K + rI));
                                            if (cm & prec) {
                                                const int rb = remoteBitTry(nbRowsW + (size_t)side * NW, xs[o], cm & prec, rrBase, P.roundBase,
                                                                            aFlags + (size_t)(side ? slotR : slotL) * SW_FLAG_RING, sysFlag);
                                                if (rb < 0) blk = true; else nbit = (uint32_t)rb;
                                            }
                                        }
                                    }
                                    nb += (int)nbit;
                                }
                                vv -= nbScale2 * real(nb - 1);
                            }
                            stop = blk || ((up ? vv : -vv) < us[o]); /* exp(-dE beta) > u */
                        }
                        const uint32_t stopBits = __ballot_sync(0xffffffffu, stop), blkBits = __ballot_sync(0xffffffffu, blk);
                        const uint32_t upBits = __ballot_sync(0xffffffffu, up != 0u);
#pragma unroll
                        for (int j = 0; j < TPP; ++j) {
                            const int tj = pass * TPP + j;
                            if (tj >= T) break;
                            const uint32_t h = (stopBits >> (j * K)) & ((K == 32) ? 0xffffffffu : ((1u << K) - 1u));
                            const int rs = h ? (__ffs(h) - 1) : Kw;
                            const bool commit = h && !((blkBits >> (j * K + rs)) & 1u);
                            const uint32_t oldFront = front[tj];
                            const uint32_t newFront = (uint32_t)(commit ? rs + 1 : rs);
                            progress |= (newFront != oldFront);
                            __syncwarp(); /* everybody has read front[tj] */
                            uint32_t upj = 0;
                            if (commit) {
                                upj = (upBits >> (j * K + rs)) & 1u;
                                if (lane == j * K + rs) stsU32(aMy0 + (uint32_t)tj * rowBytes + aw, wv ^ (1u << bit));
                                if (lane > rs && lane < Kw) { /* the trotter's later local fields of this window */
                                    const uint32_t aV = aDotsW + (uint32_t)((tj * K + lane) * sizeof(real));
                                    real dv, c;
                                    ldsReal(aV, dv);
                                    ldsReal(aCrossW + (uint32_t)(((tj * K + lane) * 2 * K + rs) * sizeof(real)), c);
                                    stsReal(aV, dv + (upj ? corrScale : -corrScale) * c);
                                }
                            }
                            if (lane == tj) { /* the trotter's own lane keeps its logs, frontier and published flags */
                                if (commit) { accC |= 1u << rs; sgnC |= upj << rs; }
                                uint32_t pend = pmask & ((1u << newFront) - 1u) & ~((1u << oldFront) - 1u);
                                while (pend) { /* rounds that became final and whose accept flag a neighbouring CTA may read */
                                    const int r = __ffs(pend) - 1;
                                    pend &= pend - 1;
                                    const unsigned long long fv = flagBase + (unsigned long long)(2 * r) + ((commit && r == rs) ? 1ull : 0ull);
                                    const int fsr = (fs + r) % SW_FLAG_RING;
                                    stRelaxed(myFlags + fsr, fv);
                                    if (mirror0) stRelaxedSys(mirror0 + fsr, fv);
                                    if (mirror1) stRelaxedSys(mirror1 + fsr, fv);
                                }
                                front[tj] = newFront;
                            }
                        }
                        __syncwarp();
                    }
                    const bool mineDone = !active || front[lane] >= (uint32_t)Kw;
                    if (__all_sync(0xffffffffu, mineDone)) break;
                    if (!progress) { ++nWaits; __nanosleep(20); } /* everything left waits for another CTA's accept flag */
                }
            } else {
            uint32_t xbN = ldsU32(pXb);
            real lnuN, vN;
            ldsReal(pUs, lnuN);
            ldsReal(pDot, vN);

            for (int rl = 0; rl < Kw; ++rl) {
                const uint32_t aw = (xbN >> 5) << 2, bit = xbN & 31u;
                const real lnu = lnuN;
                const real v = vN; /* scaleA (h + 2 sum) with every flip accepted so far folded in */
                if (rl + 1 < Kw) { /* next round's table entries (state independent) */
                    pXb += 4; pUs += (uint32_t)sizeof(real); pDot += (uint32_t)sizeof(real);
                    xbN = ldsU32(pXb);
                    ldsReal(pUs, lnuN);
                    ldsReal(pDot, vN);
                }
                const bool conflict = (cmask >> rl) & 1u;
                bool accNow = false;
                uint32_t upNow = 0;
#pragma unroll 1
                for (int ph = 0; ph < 3; ph += phStep) {
                    if (myPhase == ph) {
                        /* state-dependent shared-memory reads: own word and both neighbours' words */
                        const uint32_t wv = ldsU32(aMy + aw);
                        upNow = (wv >> bit) & 1u;
                        real vv = v;
                        if (SQA) {
                            const uint32_t lv = ldsU32(aLeft + aw), rv = ldsU32(aRight + aw);
                            int nb = (int)((lv >> bit) & 1u) + (int)((rv >> bit) & 1u); /* number of up neighbours */
                            if (conflict) { /* rare: a neighbour owned by another CTA attempted this very spin */
                                const int x = (int)ldsU32(aXsW + (uint32_t)rl * 4u);
                                int ql = ((lv >> bit) & 1u) ? 1 : -1, qr = ((rv >> bit) & 1u) ? 1 : -1;
                                /* visible: the previous window, earlier rounds of this one, this round if the neighbour's phase is earlier */
                                const uint32_t visBase = (w > 0 ? ((1u << K) - 1u) : 0u) | (((1u << rl) - 1u) << K);
                                const uint32_t mL = lLocal ? 0u : ldsU32(aConfW + (uint32_t)rl * 4u);
                                const uint32_t mR = rLocal ? 0u : ldsU32(aConfW + (uint32_t)(K + rl) * 4u);
                                if (mL) ql = remoteSpinSlow(nbRows, x, mL & (visBase | ((nbPhaseL < myPhase) ? (1u << (K + rl)) : 0u)), rrBase, P.roundBase,
                                                            aFlags + (size_t)slotL * SW_FLAG_RING, ringSharded ? 1 : 0, &slowWaits);
                                if (mR) qr = remoteSpinSlow(nbRows + NW, x, mR & (visBase | ((nbPhaseR < myPhase) ? (1u << (K + rl)) : 0u)), rrBase, P.roundBase,
                                                            aFlags + (size_t)slotR * SW_FLAG_RING, ringSharded ? 1 : 0, &slowWaits);
                                nb = (ql + qr + 2) >> 1;
                            }
                            vv -= nbScale2 * real(nb - 1);
                        }
                        accNow = (upNow ? vv : -vv) < lnu; /* exp(-dE beta) > u */
                        if (accNow) stsU32(aMy + aw, wv ^ (1u << bit));
                        if ((pmask >> rl) & 1u) {
                            const unsigned long long fv = flagBase + (unsigned long long)(2 * rl) + (accNow ? 1ull : 0ull);
                            stRelaxed(myFlags + fs, fv);
                            if (mirror0) stRelaxedSys(mirror0 + fs, fv);
                            if (mirror1) stRelaxedSys(mirror1 + fs, fv);
                        }
                    }
                    __syncwarp();
                }
                accC |= (accNow ? 1u : 0u) << rl;
                sgnC |= upNow << rl;
                /* an accepted flip of spin x changes the trotter's later local fields of this window by -4 scaleA q_old J[x'][x]:
                 * lane r repairs round r (in the classic order of additions: flips in the order they were accepted) */
                uint32_t accLanes = __ballot_sync(0xffffffffu, accNow);
                if (accLanes) {
                    do {
                        const int t = __ffs(accLanes) - 1;
                        accLanes &= accLanes - 1;
                        const uint32_t upT = __shfl_sync(0xffffffffu, upNow, t);
                        if (lane > rl && lane < Kw) {
                            const uint32_t aV = aDotsW + (uint32_t)((t * K + lane) * sizeof(real));
                            real dv, c;
                            ldsReal(aV, dv);
                            ldsReal(aCrossW + (uint32_t)(((t * K + lane) * 2 * K + rl) * sizeof(real)), c);
                            stsReal(aV, dv + (upT ? corrScale : -corrScale) * c);
                        }
                    } while (accLanes);
                    __syncwarp();
                    if (rl + 1 < Kw) ldsReal(pDot, vN);
                }
                if (++fs == SW_FLAG_RING) fs = 0;
            }
            } /* sequential chain */
            if (active) {
                accLog[buf * maxT + lane] = accC;
                if (FIELD) sgnLog[buf * maxT + lane] = sgnC;
            }
            if (publishes) { /* the neighbouring CTAs (GPUs) rebuild this trotter's spins from the accept bits of the window */
                const unsigned long long sv = ((P.snapBase + (unsigned long long)w + 1ull) << 16) | (unsigned long long)accC;
                const size_t so = (size_t)(w % SW_SNAP_SLOTS) * NW;
                stRelaxed(snapWord + so, sv);
                if (snapMirror0) stRelaxedSys(snapMirror0 + so, sv);
                if (snapMirror1) stRelaxedSys(snapMirror1 + so, sv);
            }
            nAccepted += (unsigned long long)__popc(accC);
            accP = accC; sgnP = sgnC;
            signalCount(aReplayDone, (uint32_t)w + 1u);
        }
        nWaits += slowWaits;
        if (P.stats) {
            nAccepted = warpSum(nAccepted);
            if (lane == 0) {
                atomicAdd(P.stats, nAccepted);
                atomicAdd(P.stats + 4, (unsigned long long)(waited - waitedNb)); /* chain: cycles waiting for dot products */
                atomicAdd(P.stats + 7, (unsigned long long)waitedNb);            /* chain: cycles waiting for neighbour data */
            }
        }
    }
    const long long busy = (clock64() - tLoop0) - waited; /* cycles lane 0 of this warp spent working */
    __syncthreads();

    /* ---------------- write the spins back ---------------- */
    {
        const int n4 = (N + 3) >> 2;
        for (int idx = tid; idx < T * n4; idx += SW_THREADS) {
            int r = idx / n4, j = (idx % n4) << 2;
            int w64, bit;
            spinBitPos(j, w64, bit);
            unsigned nib = (unsigned)(qcur[(size_t)r * NW + w64] >> bit) & 0xfu;
            signed char *dst = P.qOut + (size_t)replica * P.qReplicaStride + (size_t)(y0 + r) * P.ldq + j;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j + e < N) dst[e] = ((nib >> e) & 1u) ? 1 : -1;
        }
    }
    if (P.stats) {
        nWaits = warpSum(nWaits);
        if (lane == 0) {
            if (nWaits) atomicAdd(P.stats + 1, nWaits);
            if (dotWarp && dw == 0) atomicAdd(P.stats + 2, (unsigned long long)busy); /* dot warp 0: cycles spent on dot products */
            if (chainWarp && warp == 0) atomicAdd(P.stats + 3, (unsigned long long)busy); /* chain warp (field mode: the first of four): cycles spent replaying */
            if (snapWarp || nbWarp || allHelperWarp) atomicAdd(P.stats + 5, (unsigned long long)busy); /* snapshot + neighbour (or all-helper) warps */
            if (prepWarp && !FIELD) atomicAdd(P.stats + 6, (unsigned long long)busy);  /* prep warp: Philox tables (field mode: chain cycles waiting for them) */
        }
    }
}

/* ---------------- small element-wise kernels ---------------- */
__global__ void randomizeSpinKernel(signed char *q, int ldq, int N, int m, unsigned long long seed,
                                    unsigned long long count, unsigned domain, int yOff, int mPerReplica) {
    /* one Philox call per 128 spins (reference: DeviceKernels.cu:549-572 takes the LSB of a pool word per spin) */
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int groups = (N + 127) >> 7;
    if (g >= groups * m) return;
    int y = g / groups, grp = g % groups;
    /* rows are [replica][trotter]; replica r draws from seed + r, exactly like a separate solver seeded seed + r */
    const int rp = mPerReplica > 0 ? y / mPerReplica : 0, yy = mPerReplica > 0 ? y % mPerReplica : y;
    Philox4 p = sqbPhilox(seed + (unsigned long long)rp, count, domain, (uint32_t)grp, (uint32_t)(yy + yOff));
    signed char *row = q + (size_t)y * ldq;
    int x0 = grp << 7;
    for (int k = 0; k < 128 && x0 + k < N; ++k) row[x0 + k] = ((p.w[(k >> 5) & 3] >> (k & 31)) & 1u) ? 1 : -1;
}

__global__ void broadcastSpinRowKernel(signed char *q, int ldq, int N, int m, const signed char *src) {
    int x = blockIdx.y * blockDim.x + threadIdx.x, y = blockIdx.x; /* rows on grid.x */
    if (x < N && y < m) q[(size_t)y * ldq + x] = src[x];
}

/* sum_y q_y . q_{(y+1) mod m}  (reference: DeviceBatchedDot.cuh:62-74, 144-160) */
__global__ void ringSpinDotKernel(const signed char *q, int ldq, int N, int m, long long *out) {
    int y = blockIdx.x;
    const signed char *a = q + (size_t)y * ldq, *b = q + (size_t)((y + 1) % m) * ldq;
    int s = 0;
    for (int x = threadIdx.x; x < N; x += blockDim.x) s += (int)a[x] * (int)b[x];
    s = warpSum(s);
    if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)out, (unsigned long long)(long long)s);
}

/* synthetic problems generated on the device (benchmarks, multi-GPU tests): W symmetric, W[i][j] = W[j][i] ~ U(-0.5, 0.5) from
 * Philox(seed, 0, DOM_PROBLEM, min(i,j), max(i,j)); optionally rounded to the 2^-14 grid the reference tests use
 * (sqaodpy/tests/example_problems.py:16-22), which keeps every sum exact */
template <class real> __global__ void randomSymmetricKernel(real *W, int ldW, int N, unsigned long long seed, int quantize) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N) return;
    const Philox4 p = sqbPhilox(seed, 0ull, DOM_PROBLEM, (uint32_t)min(i, j), (uint32_t)max(i, j));
    double u = (double)p.w[0] * (1.0 / 4294967296.0) - 0.5;
    if (quantize) u = rint(u * 16384.0) * (1.0 / 16384.0);
    W[(size_t)i * ldW + j] = (real)u;
}

/* out[i] = max_j |A[i][j]| -- one warp per row (field mode: bound of a cross term whose gather is still in flight) */
template <class real> __global__ void rowAbsMaxKernel(real *out, const real *A, int ldA, int rows, int cols) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    real mx = real(0);
    for (int j = lane; j < cols; j += 32) mx = max(mx, fabs(A[(size_t)row * ldA + j]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) out[row] = mx;
}
template <class real> void devRowAbsMax(const B200Device &dev, real *out, const real *A, int ldA, int rows, int cols) {
    dev.makeCurrent();
    rowAbsMaxKernel<real><<<(rows + 7) / 8, 256, 0, dev.stream()>>>(out, A, ldA, rows, cols);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}

void launchRandomizeSpin(const B200Device &dev, signed char *q, int ldq, int N, int m, unsigned long long seed,
                         unsigned long long count, unsigned domain, int yOff, int mPerReplica) {
    int groups = ((N + 127) >> 7) * m;
    randomizeSpinKernel<<<(groups + 127) / 128, 128, 0, dev.stream()>>>(q, ldq, N, m, seed, count, domain, yOff, mPerReplica);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
}

long long ringSpinDot(const B200Device &dev, const signed char *q, int ldq, int N, int m) {
    DevBuf<long long> d;
    d.alloc(&dev, 1);
    ringSpinDotKernel<<<m, 128, 0, dev.stream()>>>(q, ldq, N, m, d.p);
    CUDA_CHECK(cudaGetLastError());
    ++dev.launchCount;
    long long h = 0;
    dev.d2h(&h, d.p, sizeof(h));
    dev.synchronize();
    return h;
}

struct HandoffLayout { /* one block per solver so that a single IPC handle exposes everything a peer writes */
    size_t flags, snapFlags, snapBits, stepFlags, haloQ, total;
    HandoffLayout(int m, int nw64, int ldq) {
        size_t o = 0;
        flags = o; o += (size_t)(m + 2) * SW_FLAG_RING * 8;
        snapFlags = o; o += (size_t)(m + 2) * 8;
        snapBits = o; o += (size_t)(m + 2) * SW_SNAP_SLOTS * nw64 * 8;
        stepFlags = o; o += 16;
        haloQ = o; o += (size_t)4 * ldq; /* [side][epoch parity][ldq] */
        total = (o + 255) & ~(size_t)255;
    }
};

template <class real, bool FIELD> static const void *sweepKernelForMode(bool sqa, int K) {
    switch (K) {
    case 16: return sqa ? (const void *)denseSweepKernel<real, true, 16, FIELD> : (const void *)denseSweepKernel<real, false, 16, FIELD>;
    case 8: return sqa ? (const void *)denseSweepKernel<real, true, 8, FIELD> : (const void *)denseSweepKernel<real, false, 8, FIELD>;
    default: return sqa ? (const void *)denseSweepKernel<real, true, 4, FIELD> : (const void *)denseSweepKernel<real, false, 4, FIELD>;
    }
}
template <class real> static const void *sweepKernelFor(bool sqa, int K, bool field) {
    return field ? sweepKernelForMode<real, true>(sqa, K) : sweepKernelForMode<real, false>(sqa, K);
}

/* =====================================================================================
 * host class
 * ===================================================================================== */
template <class real> B200DenseGraphAnnealer<real>::B200DenseGraphAnnealer()
    : dev_(NULL), ldJ_(0), ldq_(0), c_(0), seed_(0), step_(0), randomizeCount_(0), launchCount_(0), nWindows_(0) {
    handoff_ = NULL; handoffIpc_ = false; peerBase_[0] = peerBase_[1] = NULL;
    ringRank_ = 0; ringWorld_ = 1; mRing_ = 0; yOff_ = 0; ringEpoch_ = 0;
    nReplicas_ = 1; replicasPerLaunch_ = 1;
    m_ = -1;
    selectAlgorithm(sq::algoDefault);
}
template <class real> B200DenseGraphAnnealer<real>::~B200DenseGraphAnnealer() {
    for (int side = 0; side < 2; ++side) closePeer(side);
    freeHandoff();
}

template <class real> void B200DenseGraphAnnealer<real>::assignDevice(sq::cuda::Device &device) {
    sqb_throwErrorIf(dev_ != NULL, "Device assigned more than once.");
    dev_ = &asB200(device);
}

template <class real> sq::Algorithm B200DenseGraphAnnealer<real>::selectAlgorithm(sq::Algorithm algo) {
    /* same table as CUDADenseGraphAnnealer.cu:114-126: coloring and sa_naive are native, the rest fall back */
    switch (algo) {
    case sq::algoColoring:
    case sq::algoSANaive:
        algo_ = algo;
        break;
    default:
        selectDefaultAlgorithm(algo, sq::algoColoring, sq::algoSANaive);
        break;
    }
    return algo_;
}

template <class real> void B200DenseGraphAnnealer<real>::seed(unsigned long long seed) {
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    seed_ = seed;
    step_ = 0;
    randomizeCount_ = 0;
    setState(solRandSeedGiven);
}

template <class real> void B200DenseGraphAnnealer<real>::uploadProblem(const real *h, const real *J, int strideJ) {
    ldJ_ = sq::roundUp(N_, 128);
    dJ_.alloc(dev_, (size_t)N_ * ldJ_);
    dh_.alloc(dev_, N_);
    dev_->h2d2D(dJ_.p, sizeof(real) * ldJ_, J, sizeof(real) * strideJ, sizeof(real) * N_, N_);
    dev_->h2d(dh_.p, h, sizeof(real) * N_);
    prepareTensorCoreOperand();
    dev_->synchronize();
}

template <class real> void B200DenseGraphAnnealer<real>::prepareTensorCoreOperand() {
    /* fp32 only: J split into bf16 hi/mid/lo once per problem for the tcgen05 energy GEMM (energy_tc.cu) */
    tcJ_.ready = false;
    if constexpr (std::is_same<real, float>::value) {
        if (tcEnabled()) tcPrepareOperand(*dev_, tcJ_, dJ_.p, ldJ_, N_, N_);
    }
}

template <class real> void B200DenseGraphAnnealer<real>::setQUBO(const HostMatrix &W, sq::OptimizeMethod om) {
    sqb_throwErrorIf(W.rows != W.cols, "%s, W is not a sqare matrix.", __func__);
    sqb_throwErrorIf(!sq::isSymmetric(W), "%s, Matrix is not symmetric.", __func__);
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    clearState(solProblemSet);
    fieldsValid_ = false;
    if (nProblems_ > 1) { nProblems_ = 1; nReplicas_ = 1; cBatch_.clear(); }
    N_ = W.rows;
    m_ = N_ / 4;
    om_ = om;
    /* QUBO -> Ising on the device (formulas.cu); maximize negates W first (CUDADenseGraphAnnealer.cu:148-149) */
    ldJ_ = sq::roundUp(N_, 128);
    DevBuf<real> dW;
    dW.alloc(dev_, (size_t)N_ * ldJ_);
    dev_->h2d2D(dW.p, sizeof(real) * ldJ_, W.data, sizeof(real) * W.stride, sizeof(real) * N_, N_);
    hamiltonianFromDeviceQUBO(dW.p);
}

template <class real> void B200DenseGraphAnnealer<real>::hamiltonianFromDeviceQUBO(const real *dW) {
    dJ_.alloc(dev_, (size_t)N_ * ldJ_);
    dh_.alloc(dev_, N_);
    DevBuf<real> dc;
    dc.alloc(dev_, 1);
    devDenseHamiltonian<real>(*dev_, dh_.p, dJ_.p, ldJ_, dc.p, dW, ldJ_, N_, om_ == sq::optMaximize ? real(-1) : real(1));
    prepareTensorCoreOperand();
    dev_->d2h(&c_, dc.p, sizeof(real));
    dev_->synchronize();
    setState(solProblemSet);
}

/* extras (no reference counterpart): a synthetic random QUBO generated on the device, see randomSymmetricKernel */
template <class real> void B200DenseGraphAnnealer<real>::setQUBORandom(int N, unsigned long long seed, bool quantize, sq::OptimizeMethod om) {
    sqb_throwErrorIf(N < 1, "%s: N must be positive.", __func__);
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    clearState(solProblemSet);
    fieldsValid_ = false;
    if (nProblems_ > 1) { nProblems_ = 1; nReplicas_ = 1; cBatch_.clear(); }
    N_ = N;
    m_ = N_ / 4;
    om_ = om;
    ldJ_ = sq::roundUp(N_, 128);
    DevBuf<real> dW;
    dW.alloc(dev_, (size_t)N_ * ldJ_);
    randomSymmetricKernel<real><<<dim3((N_ + 255) / 256, N_), 256, 0, dev_->stream()>>>(dW.p, ldJ_, N_, seed, quantize ? 1 : 0);
    CUDA_CHECK(cudaGetLastError());
    ++dev_->launchCount;
    hamiltonianFromDeviceQUBO(dW.p);
}
template <class real> void B200DenseGraphAnnealer<real>::getQUBORandom(real *W, int N, int ldW, unsigned long long seed, bool quantize) const {
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    const int ld = sq::roundUp(N, 128);
    DevBuf<real> dW;
    dW.alloc(dev_, (size_t)N * ld);
    randomSymmetricKernel<real><<<dim3((N + 255) / 256, N), 256, 0, dev_->stream()>>>(dW.p, ld, N, seed, quantize ? 1 : 0);
    CUDA_CHECK(cudaGetLastError());
    ++dev_->launchCount;
    dev_->d2h2D(W, sizeof(real) * ldW, dW.p, sizeof(real) * ld, sizeof(real) * N, N);
    dev_->synchronize();
}

template <class real>
void B200DenseGraphAnnealer<real>::setHamiltonian(const HostVector &h, const HostMatrix &J, real c) {
    sqb_throwErrorIf(J.rows != J.cols || h.size != J.rows, "%s, shape mismatch between h and J.", __func__);
    sqb_throwErrorIf(!sq::isSymmetric(J), "%s, Matrix is not symmetric.", __func__);
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    clearState(solProblemSet);
    fieldsValid_ = false;
    if (nProblems_ > 1) { nProblems_ = 1; nReplicas_ = 1; cBatch_.clear(); } /* a problem batch ends here, as in setQUBO() */
    N_ = J.rows;
    m_ = N_ / 4;
    om_ = sq::optMinimize;
    c_ = c;
    uploadProblem(h.data, J.data, J.stride);
    setState(solProblemSet);
}

/* A batch of DIFFERENT problems of the same size annealed side by side: problem r is replica r of the replica
 * batch machinery, with its own J, h, c (and seed + r).  W: nProblems x N x N, row stride ldW elements. */
template <class real> void B200DenseGraphAnnealer<real>::setQUBOBatch(const real *W, int nProblems, int N, int ldW, sq::OptimizeMethod om) {
    sqb_throwErrorIf(nProblems < 1 || N < 1 || ldW < N, "%s: bad batch shape.", __func__);
    sqb_throwErrorIf(dev_ == NULL, "Device not set.");
    sqb_throwErrorIf(ringWorld_ > 1, "problem batches and ring sharding cannot be combined.");
    for (int r = 0; r < nProblems; ++r) {
        HostMatrix Wr(const_cast<real *>(W) + (size_t)r * N * ldW, N, N, ldW);
        sqb_throwErrorIf(!sq::isSymmetric(Wr), "%s, matrix %d is not symmetric.", __func__, r);
    }
    clearState(solProblemSet);
    fieldsValid_ = false;
    N_ = N;
    m_ = N_ / 4;
    om_ = om;
    nProblems_ = nProblems;
    nReplicas_ = nProblems;
    ldJ_ = sq::roundUp(N_, 128);
    dJ_.alloc(dev_, (size_t)nProblems * N_ * ldJ_);
    dh_.alloc(dev_, (size_t)nProblems * N_);
    DevBuf<real> dW, dc;
    dW.alloc(dev_, (size_t)N_ * ldJ_);
    dc.alloc(dev_, nProblems);
    for (int r = 0; r < nProblems; ++r) {
        dev_->h2d2D(dW.p, sizeof(real) * ldJ_, W + (size_t)r * N * ldW, sizeof(real) * ldW, sizeof(real) * N_, N_);
        devDenseHamiltonian<real>(*dev_, dh_.p + (size_t)r * N_, dJ_.p + (size_t)r * N_ * ldJ_, ldJ_, dc.p + r, dW.p, ldJ_, N_,
                                  om == sq::optMaximize ? real(-1) : real(1));
    }
    tcJ_.ready = false; /* energies of a batch use the CUDA-core path, one problem at a time */
    cBatch_.resize(nProblems);
    dev_->d2h(cBatch_.data(), dc.p, sizeof(real) * nProblems);
    dev_->synchronize();
    c_ = cBatch_[0];
    setState(solProblemSet);
}

template <class real> void B200DenseGraphAnnealer<real>::getHamiltonian(HostVector *h, HostMatrix *J, real *c) const {
    throwErrorIfProblemNotSet();
    h->resize(N_);
    J->resize(N_, N_);
    dev_->d2h(h->data, dh_.p, sizeof(real) * N_);
    dev_->d2h2D(J->data, sizeof(real) * J->stride, dJ_.p, sizeof(real) * ldJ_, sizeof(real) * N_, N_);
    dev_->synchronize();
    *c = c_;
}

template <class real> sq::Preferences B200DenseGraphAnnealer<real>::getPreferences() const {
    sq::Preferences prefs = Base::getPreferences();
    prefs.pushBack(sq::Preference(sq::pnDevice, "cuda"));
    return prefs;
}

template <class real> void B200DenseGraphAnnealer<real>::prepare() {
    throwErrorIfProblemNotSet();
    sqb_throwErrorIf(m_ <= 0, "# trotters must be a positive integer.");
    sqb_throwErrorIf(m_ > 32 * dev_->numSMs(), "nTrotters too large for this device.");
    if (!isRandSeedGiven()) seed((unsigned long long)time(NULL));
    setState(solRandSeedGiven);
    if (m_ == 1) selectDefaultSAAlgorithm(algo_, sq::algoSANaive);

    sqb_throwErrorIf(nReplicas_ > 1 && ringWorld_ > 1, "replica batches and ring sharding cannot be combined.");
    const int rows = m_ * nReplicas_;
    ldq_ = sq::roundUp(N_, 16);
    dq_.alloc(dev_, (size_t)rows * ldq_);
    dq2_.alloc(dev_, (size_t)rows * ldq_);
    dE_.alloc(dev_, rows);
    E_.resize(rows);
    eBack_.alloc(dev_, rows);
    hq_.assign((size_t)rows * ldq_, 0);

    /* launch geometry of the sweep */
    /* CTAs per replica: the whole device for one replica; with a batch, as few as the 32-trotters-per-CTA limit allows so
     * that many replicas run side by side in one cooperative launch */
    int G = std::min(dev_->numSMs(), (int)m_);
    if (nReplicas_ > 1) {
        const int gMin = (m_ + 31) / 32;
        G = std::max(gMin, std::min(G, dev_->numSMs() / std::min(nReplicas_, dev_->numSMs())));
        /* field mode wants at most four trotters per CTA (one accept-chain warp each): when the field rows fit in that geometry,
         * a replica gets ceil(m / 4) CTAs and fewer replicas share a launch */
        const char *feR = getenv("SQAOD_B200_SWEEP_FIELD");
        const bool fieldAllowed = ringWorld_ <= 1 && nProblems_ <= 1 && sweepModeWanted_ != 0 && !(feR && atoi(feR) == 0);
        const int Gf = std::max(gMin, std::min(dev_->numSMs(), (int)(m_ + 3) / 4));
        if (fieldAllowed && Gf > G &&
            SweepSmem<real>((m_ + Gf - 1) / Gf, packedWords64(N_), 128, 0, SW_MAX_K, SW_FIELD_WARPS, sq::roundUp(N_, 128)).total <= dev_->smemPerBlockOptin())
            G = Gf;
    }
    replicasPerLaunch_ = std::max(1, std::min(nReplicas_, dev_->numSMs() / G));
    const int maxT = (m_ + G - 1) / G;
    const int nw64 = packedWords64(N_);
    /* Shared-memory plan: the longest look-ahead window K whose cross-term table stays below 48 KiB (tables grow with
     * T K^2), and for it the deepest TMA ring that fits (12 dot warps x `stages` chunks of up to 4 KiB in flight). */
    /* warp layout: from three trotters per CTA on the accept chain is no longer what bounds the step (measured), and two
     * more warps stream rows */
    dotWarps_ = (maxT >= 3) ? SW_DOT_WARPS_WIDE : SW_DOT_WARPS;
    if (getenv("SQAOD_B200_SWEEP_WIDE")) dotWarps_ = atoi(getenv("SQAOD_B200_SWEEP_WIDE")) ? SW_DOT_WARPS_WIDE : SW_DOT_WARPS;
    int chunkElems = 0, stages = 0, K = 0;
    {
        const int forceK = getenv("SQAOD_B200_SWEEP_K") ? atoi(getenv("SQAOD_B200_SWEEP_K")) : 0;       /* tuning aids */
        const int forceS = getenv("SQAOD_B200_SWEEP_STAGES") ? atoi(getenv("SQAOD_B200_SWEEP_STAGES")) : 0;
        const int chMax = std::min((int)ldJ_, (int)(4096 / sizeof(real)));
        for (int k = SW_MAX_K; k >= 4 && K == 0; k >>= 1) {
            if (forceK ? (k != forceK) : (k > 4 && (size_t)2 * maxT * k * 2 * k * sizeof(real) > (size_t)48 * 1024)) continue;
            size_t bestRing = 0;
            for (int ci = 0; ci < 4; ++ci) { /* the whole row (or 4 KiB of it), else a smaller power of two; multiples of 128 elements */
                const int ch = (ci == 0) ? chMax : (int)(4096 / sizeof(real)) >> ci;
                if (ch < 128 || (ci > 0 && ch >= chMax)) continue;
                for (int st = 4; st >= 2; --st) {
                    if (forceS && st != forceS) continue;
                    if (SweepSmem<real>(maxT, nw64, ch, st, k, dotWarps_).total > dev_->smemPerBlockOptin()) continue;
                    const size_t ringBytes = (size_t)dotWarps_ * st * ch * sizeof(real);
                    if (ringBytes > bestRing) { bestRing = ringBytes; K = k; chunkElems = ch; stages = st; }
                }
            }
        }
        /* Field mode: no TMA ring, T rows of local fields instead (traffic: acceptance rate x one J row per attempt instead of
         * one row per attempt).  Automatic choice: whenever the rows fit next to the tables -- with this round's chain (one warp per
         * trotter, cross terms per accepted flip, table pre-pass) it wins at every size of the reference's N list where it fits
         * (m = N, ms per step classic / field: N = 128 0.099 / 0.065, 512 0.37 / 0.13, 1024 0.83 / 0.32, 2048 4.2 / 0.87; from
         * N = m = 3072 on the rows of the 21+ trotters per CTA do not fit).  set_sweep_mode() or SQAOD_B200_SWEEP_FIELD=0/1
         * override.  Not combined with ring sharding or problem batches. */
        fieldMode_ = false;
        const char *fe = getenv("SQAOD_B200_SWEEP_FIELD");
        const bool fieldAuto = true;
        const bool fieldWanted = (sweepModeWanted_ >= 0) ? sweepModeWanted_ != 0 : (fe ? atoi(fe) != 0 : fieldAuto);
        sqb_throwErrorIf(sweepModeWanted_ == 1 && (ringWorld_ > 1 || nProblems_ > 1), "field mode cannot be combined with ring sharding or problem batches.");
        if (fieldWanted && ringWorld_ <= 1 && nProblems_ <= 1) {
            for (int k = SW_MAX_K; k >= 4; k >>= 1) {
                if (forceK && k != forceK) continue;
                if (SweepSmem<real>(maxT, nw64, 128, 0, k, SW_FIELD_WARPS, ldJ_).total > dev_->smemPerBlockOptin()) continue;
                fieldMode_ = true; K = k; chunkElems = 128; stages = 0;
                /* field-mode warp layout: four accept-chain warps (one per scheduler), one helper warp, eleven field warps */
                dotWarps_ = SW_FIELD_WARPS;
                break;
            }
        }
        sqb_throwErrorIf(sweepModeWanted_ == 1 && !fieldMode_, "field mode: %d trotters per CTA x %d fields do not fit in shared memory.", maxT, (int)ldJ_);
        sqb_throwErrorIf(K == 0, "problem too large for the sweep kernel's shared memory (N=%d, m=%d).", N_, m_);
    }
    K_ = K;
    grid_ = G;
    chunkElems_ = chunkElems;
    chunksPerRow_ = (ldJ_ + chunkElems - 1) / chunkElems;
    stages_ = stages;
    nw64_ = nw64;
    smemBytes_ = SweepSmem<real>(maxT, nw64, chunkElems, stages, K, dotWarps_, fieldMode_ ? ldJ_ : 0).total;
    nWindows_ = (N_ + K - 1) / K;
    /* accept chain: the window-parallel form pays off with 1-2 trotters per CTA in the classic kernel (N = m = 128 / 256: 0.076 /
     * 0.146 vs 0.10 / 0.18 ms per step); with more trotters per CTA the per-round chain, which handles all of them in one
     * warp instruction stream, is faster (N = m = 1024: 0.89 vs 1.46 ms), and so it is in field mode (C2: 3.9 vs 4.7 ms) */
    specChain_ = getenv("SQAOD_B200_SWEEP_SPEC") ? atoi(getenv("SQAOD_B200_SWEEP_SPEC")) != 0 : (!fieldMode_ && maxT <= 2);
    fieldsValid_ = false;
    if (fieldMode_) {
        dF_.alloc(dev_, (size_t)rows * ldJ_);
        { /* max |J|: bounds the cross term of an accepted flip while its gather is in flight */
            dRowMax_.alloc(dev_, N_);
            devRowAbsMax<real>(*dev_, dRowMax_.p, dJ_.p, ldJ_, N_, N_);
            std::vector<real> rm(N_);
            dev_->d2h(rm.data(), dRowMax_.p, sizeof(real) * N_);
            dev_->synchronize();
            jAbsMax_ = *std::max_element(rm.begin(), rm.end());
            dRowMax_.release();
        }
        dev_->makeCurrent();
        CUDA_CHECK(cudaMemsetAsync(dF_.p, 0, sizeof(real) * (size_t)rows * ldJ_, dev_->stream()));
        /* the tcgen05 GEMM costs ~1 % of a step: refresh every step; the CUDA-core GEMM (fp64) only now and then */
        bool tc = false;
        if constexpr (std::is_same<real, float>::value) tc = tcJ_.ready && tcEnabled();
        /* fields are carried from step to step (written back by the sweep) and recomputed from the spins every 8 steps on the tensor
         * cores (16 on CUDA cores): after 25 steps without any recomputation they are within 5e-6 relative of a float64 evaluation
         * (tests/test_dense_annealer_gpu.py::test_field_writeback_drift_is_bounded), i.e. within the rounding of one fp32 dot product */
        fieldRefresh_ = tc ? 8 : 16;
        if (getenv("SQAOD_B200_FIELD_REFRESH")) fieldRefresh_ = std::max(1, atoi(getenv("SQAOD_B200_FIELD_REFRESH")));
        if (fieldRefreshWanted_ > 0) fieldRefresh_ = fieldRefreshWanted_;
    }
    allocHandoff();
    dStats_.alloc(dev_, 16 + 16 * (size_t)dev_->numSMs());
    launchCount_ = 0;
    CUDA_CHECK(cudaFuncSetAttribute(sweepKernelFor<real>(true, K_, fieldMode_), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes_));
    CUDA_CHECK(cudaFuncSetAttribute(sweepKernelFor<real>(false, K_, fieldMode_), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes_));
    xlist_.clear();
    qlist_.clear();
    setState(solPrepared);
}

template <class real> void B200DenseGraphAnnealer<real>::randomizeSpin() {
    throwErrorIfNotPrepared();
    launchRandomizeSpin(*dev_, dq_.p, ldq_, N_, m_ * nReplicas_, seed_, randomizeCount_++, DOM_RANDOMIZE, ringWorld_ > 1 ? yOff_ : 0,
                        nReplicas_ > 1 ? m_ : 0);
    fieldsValid_ = false;
    setState(solQSet);
}

template <class real> void B200DenseGraphAnnealer<real>::set_q(const sq::BitSet &q) {
    sqb_throwErrorIf(q.size != N_, "Dimension of q, %d, should be equal to N, %d.", q.size, N_);
    throwErrorIfNotPrepared();
    DevBuf<signed char> tmp;
    tmp.alloc(dev_, N_);
    dev_->h2d(tmp.p, q.data, N_);
    dim3 grid(m_ * nReplicas_, (N_ + 127) / 128);
    broadcastSpinRowKernel<<<grid, 128, 0, dev_->stream()>>>(dq_.p, ldq_, N_, m_ * nReplicas_, tmp.p);
    CUDA_CHECK(cudaGetLastError());
    ++dev_->launchCount;
    dev_->synchronize();
    fieldsValid_ = false;
    setState(solQSet);
}

template <class real> void B200DenseGraphAnnealer<real>::set_qset(const sq::BitSetArray &q) {
    sqb_throwErrorIf(q.size() == 0, "empty q set.");
    for (int i = 0; i < q.size(); ++i)
        sqb_throwErrorIf(q[i].size != N_, "Dimension of q, %d, should be equal to N, %d.", q[i].size, N_);
    if (nReplicas_ > 1) {
        sqb_throwErrorIf(q.size() != m_ * nReplicas_, "replica batch: expected %d x %d spin rows.", nReplicas_, m_);
        if (!isPrepared()) prepare();
    } else {
        m_ = q.size();
        prepare(); /* CUDADenseGraphAnnealer.cu:216-218: the number of trotters follows the set */
    }
    for (int y = 0; y < m_ * nReplicas_; ++y) memcpy(&hq_[(size_t)y * ldq_], q[y].data, N_);
    dev_->h2d(dq_.p, hq_.data(), hq_.size());
    dev_->synchronize();
    fieldsValid_ = false;
    setState(solQSet);
}

template <class real> void B200DenseGraphAnnealer<real>::setSpinsRaw(const signed char *q, int m) {
    /* C-ABI fast path of set_qset: q is m x N, contiguous */
    throwErrorIfProblemNotSet();
    if (nReplicas_ > 1) {
        sqb_throwErrorIf(m != m_ * nReplicas_, "replica batch: expected %d x %d spin rows, got %d.", nReplicas_, m_, m);
        if (!isPrepared()) prepare();
    } else if (m != m_ || !isPrepared()) { m_ = m; prepare(); }
    dev_->h2d2D(dq_.p, ldq_, q, N_, N_, m_ * nReplicas_);
    dev_->synchronize();
    fieldsValid_ = false;
    setState(solQSet);
}

template <class real> void B200DenseGraphAnnealer<real>::getSpinsRaw(signed char *q) const {
    throwErrorIfQNotSet();
    dev_->d2h2D(q, N_, dq_.p, ldq_, N_, m_ * nReplicas_);
    dev_->synchronize();
}

template <class real> void B200DenseGraphAnnealer<real>::syncBits() {
    xlist_.clear();
    qlist_.clear();
    dev_->d2h(hq_.data(), dq_.p, hq_.size());
    dev_->synchronize();
    for (int y = 0; y < m_ * nReplicas_; ++y) {
        sq::BitSet q(N_), x(N_);
        for (int i = 0; i < N_; ++i) {
            char v = hq_[(size_t)y * ldq_ + i];
            q(i) = v;
            x(i) = (char)((v + 1) / 2);
        }
        qlist_.pushBack(q);
        xlist_.pushBack(x);
    }
}

template <class real> const sq::BitSetArray &B200DenseGraphAnnealer<real>::get_x() const {
    if (!isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
    return xlist_;
}
template <class real> const sq::BitSetArray &B200DenseGraphAnnealer<real>::get_q() const {
    if (!isSolutionAvailable()) const_cast<This *>(this)->makeSolution();
    return qlist_;
}
template <class real> const sq::VectorType<real> &B200DenseGraphAnnealer<real>::get_E() const {
    if (!isEAvailable()) const_cast<This *>(this)->calculate_E();
    const_cast<This *>(this)->eBack_.wait(const_cast<This *>(this)->E_.data, (size_t)m_ * nReplicas_);
    return E_;
}

template <class real> void B200DenseGraphAnnealer<real>::calculate_E() {
    throwErrorIfQNotSet();
    /* E_y = -c - h.q_y - q_y^T J q_y, sign-flipped for maximize (CUDADenseGraphAnnealer.cu:260-272) */
    const real sign = (om_ == sq::optMaximize) ? real(-1) : real(1);
    bool done = false;
    if (nProblems_ > 1) { /* problem batch: rows r*m .. r*m+m-1 belong to problem r */
        for (int r = 0; r < nProblems_; ++r)
            devBatchedEnergy<real>(*dev_, dE_.p + (size_t)r * m_, dJ_.p + (size_t)r * N_ * ldJ_, ldJ_, N_, N_, dq_.p + (size_t)r * m_ * ldq_, ldq_,
                                   dq_.p + (size_t)r * m_ * ldq_, ldq_, dh_.p + (size_t)r * N_, NULL, m_, -sign, -sign * cBatch_[r]);
        done = true;
    }
    if constexpr (std::is_same<real, float>::value) {
        if (!done && tcJ_.ready && tcEnabled()) {
            tcBatchedEnergy(*dev_, dE_.p, tcJ_, dq_.p, ldq_, dq_.p, ldq_, dh_.p, NULL, m_ * nReplicas_, -sign, -sign * c_, tcWs_);
            done = true;
        }
    }
    if (!done) devBatchedEnergy<real>(*dev_, dE_.p, dJ_.p, ldJ_, N_, N_, dq_.p, ldq_, dq_.p, ldq_, dh_.p, NULL, m_ * nReplicas_, -sign, -sign * c_);
    eBack_.enqueue(dE_.p, (size_t)m_ * nReplicas_); /* asynchronous: get_E() waits for this copy's own event, not for the stream */
    setState(solEAvailable);
}

template <class real> void B200DenseGraphAnnealer<real>::makeSolution() {
    throwErrorIfQNotSet();
    syncBits();
    setState(solSolutionAvailable);
    calculate_E();
}

template <class real> real B200DenseGraphAnnealer<real>::getSystemE(real G, real beta) const {
    This *self = const_cast<This *>(this);
    sqb_throwErrorIf(nReplicas_ > 1, "getSystemE is defined per solver instance; not available on a replica batch.");
    sqb_throwErrorIf(ringWorld_ > 1, "getSystemE is not available on a shard of a trotter ring; gather the spins (multigpu.RingShardedDenseAnnealer.get_system_E).");
    self->calculate_E();
    self->eBack_.wait(self->E_.data, (size_t)m_ * nReplicas_);
    real E = E_.sum() / m_;
    if (sq::isSQAAlgorithm(algo_)) {
        real spinDotSum = (real)ringSpinDot(*dev_, dq_.p, ldq_, N_, m_);
        real coef = real(0.5) / beta * std::log(std::tanh(G * beta / m_));
        E -= spinDotSum * coef;
    }
    if (om_ == sq::optMaximize) E *= real(-1.); /* as the reference, CUDADenseGraphAnnealer.cu:372-373 */
    return E;
}

template <class real> void B200DenseGraphAnnealer<real>::annealOneStep(real G, real beta) {
    throwErrorIfQNotSet();
    clearState(solSolutionAvailable);
    SweepParams<real> P;
    P.J = dJ_.p; P.h = dh_.p; P.q = dq_.p; P.qOut = dq2_.p;
    P.ldJ = ldJ_; P.ldq = ldq_; P.N = N_; P.m = m_;
    P.seed = seed_; P.step = step_;
    const bool sqa = (algo_ == sq::algoColoring);
    if (sqa) {
        const int mAll = (ringWorld_ > 1) ? mRing_ : m_; /* the whole ring, also when this GPU holds a shard of it */
        P.twoDivM = real(2.) / real(mAll);
        P.coef = std::log(std::tanh(G * beta / mAll)) / beta;
        P.beta = beta;
        P.scaleA = beta * P.twoDivM;
        P.scaleNb = beta * P.coef;
    } else { /* annealOneStep(kT, _) for SA: CUDADenseGraphAnnealer.cu:585-602 */
        P.twoDivM = real(2.);
        P.coef = real(0.);
        P.beta = real(1.) / G;
        P.scaleA = real(2.) * P.beta;
        P.scaleNb = real(0.);
    }
    P.chunkElems = chunkElems_; P.chunksPerRow = chunksPerRow_; P.stages = stages_; P.nw64 = nw64_; P.K = K_; P.dotWarps = dotWarps_;
    HandoffLayout hl(m_, nw64_, ldq_);
    unsigned char *hb = (unsigned char *)handoff_;
    P.acceptFlags = (unsigned long long *)(hb + hl.flags); P.snapFlags = (unsigned long long *)(hb + hl.snapFlags);
    P.snapBits = (unsigned long long *)(hb + hl.snapBits);
    const bool ring = ringWorld_ > 1;
    P.mRing = ring ? mRing_ : m_;
    P.yOff = ring ? yOff_ : 0;
    P.stepFlags = (const unsigned long long *)(hb + hl.stepFlags);
    P.stepEpoch = ringEpoch_;
    for (int side = 0; side < 2; ++side) {
        P.haloQ[side] = (const signed char *)(hb + hl.haloQ) + (size_t)(side * 2 + (ringEpoch_ & 1)) * ldq_;
        unsigned char *pb = ring ? (unsigned char *)peerBase_[side] : NULL;
        P.peerFlags[side] = pb ? (unsigned long long *)(pb + hl.flags) : NULL;
        P.peerSnapFlags[side] = pb ? (unsigned long long *)(pb + hl.snapFlags) : NULL;
        P.peerSnapBits[side] = pb ? (unsigned long long *)(pb + hl.snapBits) : NULL;
    }
    if (ring) sqb_throwErrorIf(peerBase_[0] == NULL || peerBase_[1] == NULL, "ring sharding: peers not attached.");
    P.roundBase = (launchCount_ + 1ull) * (unsigned long long)(N_ + SW_FLAG_RING);
    P.snapBase = (launchCount_ + 1ull) * (unsigned long long)(nWindows_ + 2);
    P.stats = dStats_.p;
    P.F = NULL; P.ldF = 0; P.writeBackF = 0; P.uncBound = real(0); P.fieldHasH = 0;
    P.specChain = specChain_ ? 1 : 0;
    if (fieldMode_) {
        if (!fieldsValid_ || stepsSinceRefresh_ >= fieldRefresh_) refreshFields();
        ++stepsSinceRefresh_;
        P.F = dF_.p; P.ldF = ldJ_;
        P.uncBound = real(4.004) * P.scaleA * jAbsMax_;
        P.fieldHasH = fieldsHaveH_ ? 1 : 0;
        fieldsHaveH_ = (fieldRefresh_ > 1); /* this step writes h + 2 J.q back */
        P.writeBackF = (fieldRefresh_ > 1) ? 1 : 0; /* refreshed before every step otherwise */
        fieldsValid_ = (P.writeBackF != 0);
    }
    P.qReplicaStride = (size_t)m_ * ldq_;
    P.jReplicaStride = (nProblems_ > 1) ? (size_t)N_ * ldJ_ : 0;
    P.hReplicaStride = (nProblems_ > 1) ? (size_t)N_ : 0;
    P.handoffReplicaStride = hl.total;
    void *args[] = {&P};
    const void *fn = sweepKernelFor<real>(sqa, K_, fieldMode_);
    dev_->makeCurrent();
    P.tables = NULL;
    for (int base = 0; base < nReplicas_; base += replicasPerLaunch_) {
        P.replicaBase = base;
        const int nr = std::min(replicasPerLaunch_, nReplicas_ - base);
        if (fieldMode_) { /* table pre-pass for the replicas of this launch (Philox draws, -ln u, conflict masks of every window) */
            const int maxT = (m_ + grid_ - 1) / grid_;
            const SweepTabRec<real> R(maxT, K_);
            const size_t need = R.bytes * (size_t)nWindows_ * grid_ * replicasPerLaunch_;
            if (dTables_.n < need) dTables_.alloc(dev_, need);
            SweepTabParams TP;
            TP.tables = dTables_.p; TP.N = N_; TP.m = m_; TP.G = grid_; TP.K = K_; TP.nW = nWindows_; TP.replicaBase = base;
            TP.seed = seed_; TP.step = step_; TP.sqa = sqa ? 1 : 0;
            const size_t scratch = (size_t)4 * (maxT * K_ + 6 * K_) * sizeof(int);
            sweepTablesKernel<real><<<dim3((nWindows_ + 3) / 4, grid_, nr), 128, scratch, dev_->stream()>>>(TP);
            CUDA_CHECK(cudaGetLastError());
            ++dev_->launchCount;
            P.tables = dTables_.p;
        }
        CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(grid_, nr), dim3(SW_THREADS), args, smemBytes_, dev_->stream()));
        ++dev_->launchCount;
    }
    std::swap(dq_.p, dq2_.p); /* the sweep wrote the new spins into the second buffer */
    ++launchCount_;
    ++step_;
    if (ringWorld_ > 1) ringPushHalos(); /* per-sweep boundary exchange over NVLink */
}

/* F[y][j] = sum_i J[j][i] q[y][i] for every trotter (and replica): the local-field contraction of the north star, on the
 * tcgen05 split-bf16 GEMM for fp32 (energy_tc.cu), on the CUDA-core spin GEMM otherwise */
template <class real> void B200DenseGraphAnnealer<real>::setSweepMode(int mode, int fieldRefresh) {
    sqb_throwErrorIf(mode < -1 || mode > 1, "sweep mode must be -1 (automatic), 0 (classic) or 1 (field).");
    sweepModeWanted_ = mode;
    fieldRefreshWanted_ = std::max(0, fieldRefresh);
    clearState(solPrepared);
}

/* extras: the local fields the field-mode sweep carries from step to step, as H[y][j] = h[j] + 2 sum_i J[j][i] q[y][i]
 * (tests bound their drift against a fresh evaluation).  Returns false when no carried fields exist. */
template <class real> bool B200DenseGraphAnnealer<real>::getFields(real *H, int ldH) const {
    if (!fieldMode_ || !fieldsValid_ || !isPrepared()) return false;
    const int rows = m_ * nReplicas_;
    std::vector<real> F((size_t)rows * ldJ_), h(N_);
    dev_->d2h(F.data(), dF_.p, sizeof(real) * F.size());
    dev_->d2h(h.data(), dh_.p, sizeof(real) * N_);
    dev_->synchronize();
    for (int y = 0; y < rows; ++y)
        for (int j = 0; j < N_; ++j) {
            const real f = F[(size_t)y * ldJ_ + j];
            H[(size_t)y * ldH + j] = fieldsHaveH_ ? f : h[j] + real(2) * f;
        }
    return true;
}

template <class real> void B200DenseGraphAnnealer<real>::refreshFields() {
    const int rows = m_ * nReplicas_;
    bool done = false;
    if constexpr (std::is_same<real, float>::value) {
        if (tcJ_.ready && tcEnabled()) {
            tcSpinGemm(*dev_, dF_.p, ldJ_, tcJ_, dq_.p, ldq_, rows, tcWs_);
            done = true;
        }
    }
    if (!done) devSpinGemm<real>(*dev_, dF_.p, ldJ_, dJ_.p, ldJ_, dq_.p, ldq_, rows, N_, N_);
    fieldsValid_ = true;
    fieldsHaveH_ = false; /* raw J.q */
    stepsSinceRefresh_ = 0;
}

/* ---------------- ring sharding over several GPUs ---------------- */

template <class real> void B200DenseGraphAnnealer<real>::freeHandoff() {
    if (handoff_ == NULL) return;
    if (handoffIpc_) cudaFree(handoff_); else dev_->free(handoff_);
    handoff_ = NULL;
}
template <class real> void B200DenseGraphAnnealer<real>::allocHandoff() {
    freeHandoff();
    HandoffLayout hl(m_, nw64_, ldq_);
    handoffIpc_ = ringWorld_ > 1;
    if (handoffIpc_) { /* plain cudaMalloc: exportable with cudaIpcGetMemHandle (pool memory is not) */
        dev_->makeCurrent();
        CUDA_CHECK(cudaMalloc(&handoff_, hl.total * nReplicas_));
        CUDA_CHECK(cudaMemsetAsync(handoff_, 0, hl.total * nReplicas_, dev_->stream()));
        dev_->synchronize();
    } else
        handoff_ = dev_->alloc(hl.total * nReplicas_);
    for (int side = 0; side < 2; ++side) closePeer(side);
    ringEpoch_ = 0;
}
template <class real> void B200DenseGraphAnnealer<real>::closePeer(int side) {
    if (peerBase_[side] != NULL) {
        if (side == 1 && peerBase_[1] == peerBase_[0]) { peerBase_[1] = NULL; return; }
        cudaIpcCloseMemHandle(peerBase_[side]);
        if (side == 0 && peerBase_[1] == peerBase_[0]) peerBase_[1] = NULL;
        peerBase_[side] = NULL;
    }
}
template <class real> void B200DenseGraphAnnealer<real>::setNumReplicas(int n) {
    sqb_throwErrorIf(n < 1, "number of replicas must be positive.");
    sqb_throwErrorIf(nProblems_ > 1 && n != nProblems_, "a problem batch has one replica per problem (%d).", nProblems_);
    if (n != nReplicas_) clearState(solPrepared);
    nReplicas_ = n;
}

template <class real> void B200DenseGraphAnnealer<real>::ringConfigure(int rank, int world, int mGlobal) {
    sqb_throwErrorIf(world < 1 || rank < 0 || rank >= world, "ring sharding: invalid rank %d / world %d.", rank, world);
    sqb_throwErrorIf(mGlobal % world != 0 || mGlobal / world < 2, "ring sharding: n_trotters (%d) must be a multiple of the number "
                     "of GPUs (%d) with at least 2 trotters per GPU.", mGlobal, world);
    ringRank_ = rank; ringWorld_ = world; mRing_ = mGlobal;
    m_ = mGlobal / world;
    yOff_ = rank * m_;
    clearState(solPrepared);
}
template <class real> void B200DenseGraphAnnealer<real>::ringExport(unsigned char handle[64]) const {
    throwErrorIfNotPrepared();
    sqb_throwErrorIf(!handoffIpc_, "ring sharding is not configured.");
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, handoff_));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle, &h, 64);
}
template <class real> void B200DenseGraphAnnealer<real>::ringAttach(const unsigned char left[64], const unsigned char right[64]) {
    throwErrorIfNotPrepared();
    sqb_throwErrorIf(!handoffIpc_, "ring sharding is not configured.");
    dev_->makeCurrent();
    for (int side = 0; side < 2; ++side) closePeer(side);
    cudaIpcMemHandle_t h;
    memcpy(&h, left, 64);
    CUDA_CHECK(cudaIpcOpenMemHandle(&peerBase_[0], h, cudaIpcMemLazyEnablePeerAccess));
    if (memcmp(left, right, 64) == 0) peerBase_[1] = peerBase_[0]; /* two GPUs: both neighbours are the same peer */
    else {
        memcpy(&h, right, 64);
        CUDA_CHECK(cudaIpcOpenMemHandle(&peerBase_[1], h, cudaIpcMemLazyEnablePeerAccess));
    }
}

__global__ void ringPushKernel(const signed char *qFirst, const signed char *qLast, signed char *dstLeftPeer, signed char *dstRightPeer,
                               int n, unsigned long long *flagLeftPeer, unsigned long long *flagRightPeer, unsigned long long epoch) {
    /* my first trotter is the left peer's right neighbour, my last trotter the right peer's left neighbour */
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dstLeftPeer[i] = qFirst[i];
        dstRightPeer[i] = qLast[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        stReleaseSys(flagLeftPeer, epoch);
        stReleaseSys(flagRightPeer, epoch);
    }
}
template <class real> void B200DenseGraphAnnealer<real>::ringPushHalos() {
    throwErrorIfQNotSet();
    sqb_throwErrorIf(ringWorld_ <= 1, "ring sharding is not configured.");
    sqb_throwErrorIf(peerBase_[0] == NULL || peerBase_[1] == NULL, "ring sharding: peers not attached.");
    HandoffLayout hl(m_, nw64_, ldq_);
    ++ringEpoch_;
    const size_t par = ringEpoch_ & 1;
    unsigned char *L = (unsigned char *)peerBase_[0], *R = (unsigned char *)peerBase_[1];
    signed char *dstL = (signed char *)(L + hl.haloQ) + (size_t)(1 * 2 + par) * ldq_;  /* left peer's RIGHT halo */
    signed char *dstR = (signed char *)(R + hl.haloQ) + (size_t)(0 * 2 + par) * ldq_;  /* right peer's LEFT halo */
    unsigned long long *fL = (unsigned long long *)(L + hl.stepFlags) + 1, *fR = (unsigned long long *)(R + hl.stepFlags) + 0;
    ringPushKernel<<<1, 256, 0, dev_->stream()>>>(dq_.p, dq_.p + (size_t)(m_ - 1) * ldq_, dstL, dstR, N_, fL, fR, ringEpoch_);
    CUDA_CHECK(cudaGetLastError());
    ++dev_->launchCount;
}

template <class real> void B200DenseGraphAnnealer<real>::getStats(unsigned long long *accepted, unsigned long long *waits) const {
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (dStats_.p) {
        dev_->d2h(h, dStats_.p, sizeof(h));
        dev_->synchronize();
    }
    *accepted = h[0];
    *waits = h[1];
    lastBarrierWaitDot_ = h[2];
    lastBarrierWaitChain_ = h[3];
}

template <class real> void B200DenseGraphAnnealer<real>::getCounters(unsigned long long out[8]) const {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (dStats_.p) {
        dev_->d2h(out, dStats_.p, 8 * sizeof(unsigned long long));
        dev_->synchronize();
    }
}
template <class real> int B200DenseGraphAnnealer<real>::getCtaProfile(unsigned long long *out, int maxCtas) const {
    const int n = std::min(maxCtas, std::min(grid_, dev_->numSMs()));
    if (dStats_.p && n > 0) {
        dev_->d2h(out, dStats_.p + 16, (size_t)n * 16 * sizeof(unsigned long long));
        dev_->synchronize();
    }
    return n;
}
template <class real> void B200DenseGraphAnnealer<real>::getProfile(unsigned long long out[16]) const {
    for (int i = 0; i < 16; ++i) out[i] = 0;
    if (dStats_.p) {
        dev_->d2h(out, dStats_.p, 16 * sizeof(unsigned long long));
        dev_->synchronize();
    }
}

template class B200DenseGraphAnnealer<float>;
template class B200DenseGraphAnnealer<double>;

} // namespace sqb

namespace sqaod { namespace cuda {
template <> DenseGraphAnnealer<float> *newDenseGraphAnnealer<float>() { return new sqb::B200DenseGraphAnnealer<float>(); }
template <> DenseGraphAnnealer<double> *newDenseGraphAnnealer<double>() { return new sqb::B200DenseGraphAnnealer<double>(); }
}} // namespace sqaod::cuda
