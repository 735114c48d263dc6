/* cabi.cpp -- extern "C" boundary of libsqaod_b200.so (declared in include/sqaod_b200.h).
 * Thin: maps plain buffers onto sqaod::MatrixType/VectorType views (no copies), calls the C++ solver interface and
 * turns C++ exceptions into error codes, exactly where the reference's pyglue turns them into Python RuntimeErrors
 * (sqaodc/pyglue/pyglue.h:394-399). */
#include <nvtx3/nvToolsExt.h>
#include <sqaod_b200.h>
#include <sqaod_b200/sqaod_api.hpp>
#include "b200_solvers.hpp"
#include <string>
#include <stdio.h>
#include <string.h>

namespace sq = sqaod;
namespace sqc = sqaod::cuda;

static thread_local std::string g_lastError;

/* every C-ABI call is an NVTX range named after the entry point (visible in Nsight Systems / Compute timelines; header-only
 * NVTX3: a no-op unless a profiler injects itself) */
struct SqbNvtxScope {
    explicit SqbNvtxScope(const char *name) { nvtxRangePushA(name); }
    ~SqbNvtxScope() { nvtxRangePop(); }
};
#define SQB_TRY SqbNvtxScope sqbNvtxScope_(__func__); try {
#define SQB_CATCH                                                          \
    }                                                                      \
    catch (const std::exception &e) { g_lastError = e.what(); return 1; }  \
    catch (...) { g_lastError = "unknown error"; return 2; }               \
    return 0;

#define DISPATCH(dtype, ...)                                               \
    if ((dtype) == SQB_F32) { typedef float real; __VA_ARGS__; }           \
    else if ((dtype) == SQB_F64) { typedef double real; __VA_ARGS__; }     \
    else sqb_throwError("unknown dtype %d", (int)(dtype));

namespace {

template <class T> T *as(sqb_handle h) {
    sqb_throwErrorIf(h == NULL, "null handle.");
    return static_cast<T *>(h);
}
sqb::B200Device *asDev(sqb_handle h) { return as<sqb::B200Device>(h); }

template <class real> sq::MatrixType<real> mapMat(const void *p, int rows, int cols, int stride) {
    return sq::MatrixType<real>((real *)p, rows, cols, stride);
}
template <class real> sq::VectorType<real> mapVec(const void *p, int n) { return sq::VectorType<real>((real *)p, n); }

template <class S> void setPreference(S *s, const char *name, const char *str, long value) {
    sq::PreferenceName pn = sq::preferenceNameFromString(name);
    sqb_throwErrorIf(pn == sq::pnUnknown, "unknown preference name %s.", name); /* pyglue.h:310-313 */
    switch (pn) {
    case sq::pnAlgorithm: {
        sqb_throwErrorIf(str == NULL, "algorithm must be a string.");
        sq::Algorithm a = sq::algorithmFromString(str);
        sqb_throwErrorIf(a == sq::algoUnknown, "unknown algorithm %s.", str);
        s->setPreference(sq::Preference(pn, a));
        break;
    }
    case sq::pnPrecision:
    case sq::pnDevice:
        break; /* read-only */
    default:
        s->setPreference(sq::Preference(pn, (sq::SizeType)value));
    }
}
template <class S> void getPreferences(const S *s, char *buf, int buflen) {
    sq::Preferences prefs = s->getPreferences();
    std::string out;
    for (int i = 0; i < prefs.size(); ++i) {
        const sq::Preference &p = prefs[i];
        if (!out.empty()) out += ";";
        out += sq::preferenceNameToString(p.name);
        out += "=";
        char num[32];
        switch (p.name) {
        case sq::pnAlgorithm: out += sq::algorithmToString(p.algo); break;
        case sq::pnPrecision:
        case sq::pnDevice: out += p.str; break;
        default: snprintf(num, sizeof(num), "%d", (int)p.size); out += num;
        }
    }
    sqb_throwErrorIf((int)out.size() + 1 > buflen, "preference buffer too small.");
    memcpy(buf, out.c_str(), out.size() + 1);
}

void copyBitSets(signed char *dst, const sq::BitSetArray &arr, int N) {
    for (int i = 0; i < arr.size(); ++i) memcpy(dst + (size_t)i * N, arr[i].data, N);
}
sq::BitSet viewBits(const signed char *p, int n) { return sq::BitSet((char *)p, n); }

/* the C++ formulas interface takes bit / spin matrices as `real` (formulas.inc:340-345 casts on the host too) */
template <class real> sq::MatrixType<real> widen(const signed char *x, int rows, int cols) {
    sq::MatrixType<real> M(rows, cols);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) M(r, c) = (real)x[(size_t)r * cols + c];
    return M;
}
void splitPairs(signed char *d0, signed char *d1, const sq::BitSetPairArray &arr, int N0, int N1) {
    for (int i = 0; i < arr.size(); ++i) {
        memcpy(d0 + (size_t)i * N0, arr[i].bits0.data, N0);
        memcpy(d1 + (size_t)i * N1, arr[i].bits1.data, N1);
    }
}

} // namespace

extern "C" {

int sqb_version(void) { return 100; }
const char *sqb_last_error(void) { return g_lastError.c_str(); }
void sqaodc_cuda_version(int *version, int *cuda_version) {
    *version = 10003; /* interface level of sqaod 1.0.3 (common/defines.h:80) */
    int v = 0;
    cudaRuntimeGetVersion(&v);
    *cuda_version = v;
}
int sqb_device_count(int *count) {
    SQB_TRY
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    *count = n;
    SQB_CATCH
}

/* ---------------- device ---------------- */
int sqb_device_new(sqb_handle *dev) { SQB_TRY *dev = new sqb::B200Device(); SQB_CATCH }
int sqb_device_initialize(sqb_handle dev, int devNo) { SQB_TRY asDev(dev)->initialize(devNo); SQB_CATCH }
int sqb_device_finalize(sqb_handle dev) { SQB_TRY asDev(dev)->finalize(); SQB_CATCH }
int sqb_device_delete(sqb_handle dev) { SQB_TRY delete asDev(dev); SQB_CATCH }
int sqb_device_synchronize(sqb_handle dev) { SQB_TRY asDev(dev)->synchronize(); SQB_CATCH }
int sqb_device_set_stream(sqb_handle dev, void *s) { SQB_TRY asDev(dev)->setExternalStream((cudaStream_t)s); SQB_CATCH }
int sqb_device_launch_count(sqb_handle dev, unsigned long long *count, int reset) {
    SQB_TRY
    if (count) *count = asDev(dev)->launchCount;
    if (reset) asDev(dev)->launchCount = 0;
    SQB_CATCH
}
int sqb_device_num_sms(sqb_handle dev, int *n) { SQB_TRY *n = asDev(dev)->numSMs(); SQB_CATCH }

/* ---------------- dense-graph annealer ---------------- */
#define DGA(real) as<sqc::DenseGraphAnnealer<real> >(ann)
#define DGAX(real) dynamic_cast<sqb::B200DenseGraphAnnealer<real> *>(DGA(real))

int sqb_dg_annealer_new(sqb_handle *ann, int dtype) { SQB_TRY DISPATCH(dtype, *ann = sqc::newDenseGraphAnnealer<real>()) SQB_CATCH }
int sqb_dg_annealer_delete(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(DGA(real))) SQB_CATCH }
int sqb_dg_annealer_assign_device(sqb_handle ann, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_dg_annealer_seed(sqb_handle ann, unsigned long long seed, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->seed(seed)) SQB_CATCH }
int sqb_dg_annealer_set_qubo(sqb_handle ann, const void *W, int N, int stride, int optimize, int dtype) {
    SQB_TRY DISPATCH(dtype, DGA(real)->setQUBO(mapMat<real>(W, N, N, stride), (sq::OptimizeMethod)optimize)) SQB_CATCH
}
int sqb_dg_annealer_set_qubo_random(sqb_handle ann, int N, unsigned long long seed, int quantize, int optimize, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->setQUBORandom(N, seed, quantize != 0, (sq::OptimizeMethod)optimize)) SQB_CATCH
}
int sqb_dg_annealer_get_qubo_random(sqb_handle ann, void *W, int N, int ldW, unsigned long long seed, int quantize, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->getQUBORandom((real *)W, N, ldW, seed, quantize != 0)) SQB_CATCH
}
int sqb_dg_annealer_set_hamiltonian(sqb_handle ann, const void *h, const void *J, int N, int strideJ, double c, int dtype) {
    SQB_TRY DISPATCH(dtype, DGA(real)->setHamiltonian(mapVec<real>(h, N), mapMat<real>(J, N, N, strideJ), (real)c)) SQB_CATCH
}
int sqb_dg_annealer_get_hamiltonian(sqb_handle ann, void *h, void *J, int strideJ, void *c, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::SizeType N;
        DGA(real)->getProblemSize(&N);
        sq::VectorType<real> hv = mapVec<real>(h, N);
        sq::MatrixType<real> Jm = mapMat<real>(J, N, N, strideJ);
        DGA(real)->getHamiltonian(&hv, &Jm, (real *)c);
    })
    SQB_CATCH
}
int sqb_dg_annealer_get_problem_size(sqb_handle ann, int *N, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->getProblemSize(N)) SQB_CATCH }
int sqb_dg_annealer_set_preference(sqb_handle ann, const char *name, const char *str, long value, int dtype) {
    SQB_TRY DISPATCH(dtype, setPreference(DGA(real), name, str, value)) SQB_CATCH
}
int sqb_dg_annealer_get_preferences(sqb_handle ann, char *buf, int buflen, int dtype) {
    SQB_TRY DISPATCH(dtype, getPreferences(DGA(real), buf, buflen)) SQB_CATCH
}
int sqb_dg_annealer_get_num_trotters(sqb_handle ann, int *m, int dtype) { SQB_TRY DISPATCH(dtype, *m = DGAX(real)->numTrotters()) SQB_CATCH }
int sqb_dg_annealer_get_E(sqb_handle ann, void *E, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        const sq::VectorType<real> &e = DGA(real)->get_E();
        sqb_throwErrorIf(capacity < e.size, "E buffer too small (%d < %d).", capacity, e.size);
        memcpy(E, e.data, sizeof(real) * e.size);
    })
    SQB_CATCH
}
int sqb_dg_annealer_get_x(sqb_handle ann, signed char *x, int dtype) {
    SQB_TRY DISPATCH(dtype, { sq::SizeType N; DGA(real)->getProblemSize(&N); copyBitSets(x, DGA(real)->get_x(), N); }) SQB_CATCH
}
int sqb_dg_annealer_get_q(sqb_handle ann, signed char *q, int dtype) {
    SQB_TRY DISPATCH(dtype, { sq::SizeType N; DGA(real)->getProblemSize(&N); copyBitSets(q, DGA(real)->get_q(), N); }) SQB_CATCH
}
int sqb_dg_annealer_set_q(sqb_handle ann, const signed char *q, int N, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->set_q(viewBits(q, N))) SQB_CATCH }
int sqb_dg_annealer_set_qset(sqb_handle ann, const signed char *q, int m, int N, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::SizeType Np;
        DGA(real)->getProblemSize(&Np);
        sqb_throwErrorIf(N != Np, "Dimension of q, %d, should be equal to N, %d.", N, Np);
        DGAX(real)->setSpinsRaw(q, m);
    })
    SQB_CATCH
}
int sqb_dg_annealer_randomize_spin(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->randomizeSpin()) SQB_CATCH }
int sqb_dg_annealer_calculate_E(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->calculate_E()) SQB_CATCH }
int sqb_dg_annealer_prepare(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->prepare()) SQB_CATCH }
int sqb_dg_annealer_make_solution(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, DGA(real)->makeSolution()) SQB_CATCH }
int sqb_dg_annealer_get_system_E(sqb_handle ann, double G, double beta, double *E, int dtype) {
    SQB_TRY DISPATCH(dtype, *E = (double)DGA(real)->getSystemE((real)G, (real)beta)) SQB_CATCH
}
int sqb_dg_annealer_anneal_one_step(sqb_handle ann, double G, double beta, int dtype) {
    SQB_TRY DISPATCH(dtype, DGA(real)->annealOneStep((real)G, (real)beta)) SQB_CATCH
}
int sqb_dg_annealer_get_stats(sqb_handle ann, unsigned long long *accepted, unsigned long long *waits, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->getStats(accepted, waits)) SQB_CATCH
}
int sqb_dg_annealer_get_barrier_cycles(sqb_handle ann, unsigned long long *dot, unsigned long long *chain, int dtype) {
    SQB_TRY
    DISPATCH(dtype, { unsigned long long a, w; DGAX(real)->getStats(&a, &w); DGAX(real)->getBarrierStats(dot, chain); })
    SQB_CATCH
}
int sqb_dg_annealer_get_counters(sqb_handle ann, unsigned long long *out8, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->getCounters(out8)) SQB_CATCH }
int sqb_dg_annealer_get_cta_profile(sqb_handle ann, unsigned long long *out, int max_ctas, int *n, int dtype) { SQB_TRY DISPATCH(dtype, *n = DGAX(real)->getCtaProfile(out, max_ctas)) SQB_CATCH }
int sqb_dg_annealer_get_fields(sqb_handle ann, void *H, int ldH, int *valid, int dtype) { SQB_TRY DISPATCH(dtype, *valid = DGAX(real)->getFields((real *)H, ldH) ? 1 : 0) SQB_CATCH }
int sqb_dg_annealer_get_profile(sqb_handle ann, unsigned long long *out16, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->getProfile(out16)) SQB_CATCH }
int sqb_dg_annealer_set_sweep_mode(sqb_handle ann, int mode, int field_refresh, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->setSweepMode(mode, field_refresh)) SQB_CATCH }
int sqb_dg_annealer_get_sweep_mode(sqb_handle ann, int *mode, int dtype) { SQB_TRY DISPATCH(dtype, *mode = DGAX(real)->fieldMode() ? 1 : 0) SQB_CATCH }
int sqb_dg_annealer_get_spins(sqb_handle ann, signed char *q, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->getSpinsRaw(q)) SQB_CATCH }

int sqb_dg_annealer_set_qubo_batch(sqb_handle ann, const void *W, int n_problems, int N, int ldW, int optimize, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->setQUBOBatch((const real *)W, n_problems, N, ldW, (sq::OptimizeMethod)optimize)) SQB_CATCH
}
int sqb_dg_annealer_set_num_replicas(sqb_handle ann, int n, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->setNumReplicas(n)) SQB_CATCH }
int sqb_dg_annealer_get_num_replicas(sqb_handle ann, int *n, int dtype) { SQB_TRY DISPATCH(dtype, *n = DGAX(real)->numReplicas()) SQB_CATCH }
int sqb_dg_annealer_ring_configure(sqb_handle ann, int rank, int world, int m_global, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->ringConfigure(rank, world, m_global)) SQB_CATCH
}
int sqb_dg_annealer_ring_export(sqb_handle ann, unsigned char *handle64, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->ringExport(handle64)) SQB_CATCH }
int sqb_dg_annealer_ring_attach(sqb_handle ann, const unsigned char *left, const unsigned char *right, int dtype) {
    SQB_TRY DISPATCH(dtype, DGAX(real)->ringAttach(left, right)) SQB_CATCH
}
int sqb_dg_annealer_ring_push_halos(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, DGAX(real)->ringPushHalos()) SQB_CATCH }

/* ---------------- bipartite-graph annealer ---------------- */
#define BGA(real) as<sqc::BipartiteGraphAnnealer<real> >(ann)

int sqb_bg_annealer_new(sqb_handle *ann, int dtype) { SQB_TRY DISPATCH(dtype, *ann = sqc::newBipartiteGraphAnnealer<real>()) SQB_CATCH }
int sqb_bg_annealer_delete(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(BGA(real))) SQB_CATCH }
int sqb_bg_annealer_assign_device(sqb_handle ann, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_bg_annealer_seed(sqb_handle ann, unsigned long long seed, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->seed(seed)) SQB_CATCH }
int sqb_bg_annealer_set_qubo(sqb_handle ann, const void *b0, const void *b1, const void *W, int N0, int N1, int stride, int optimize, int dtype) {
    SQB_TRY
    DISPATCH(dtype, BGA(real)->setQUBO(mapVec<real>(b0, N0), mapVec<real>(b1, N1), mapMat<real>(W, N1, N0, stride), (sq::OptimizeMethod)optimize))
    SQB_CATCH
}
int sqb_bg_annealer_set_hamiltonian(sqb_handle ann, const void *h0, const void *h1, const void *J, int N0, int N1, int strideJ, double c, int dtype) {
    SQB_TRY
    DISPATCH(dtype, BGA(real)->setHamiltonian(mapVec<real>(h0, N0), mapVec<real>(h1, N1), mapMat<real>(J, N1, N0, strideJ), (real)c))
    SQB_CATCH
}
int sqb_bg_annealer_get_hamiltonian(sqb_handle ann, void *h0, void *h1, void *J, int strideJ, void *c, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::SizeType N0, N1;
        BGA(real)->getProblemSize(&N0, &N1);
        sq::VectorType<real> v0 = mapVec<real>(h0, N0), v1 = mapVec<real>(h1, N1);
        sq::MatrixType<real> Jm = mapMat<real>(J, N1, N0, strideJ);
        BGA(real)->getHamiltonian(&v0, &v1, &Jm, (real *)c);
    })
    SQB_CATCH
}
int sqb_bg_annealer_get_problem_size(sqb_handle ann, int *N0, int *N1, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->getProblemSize(N0, N1)) SQB_CATCH }
int sqb_bg_annealer_set_preference(sqb_handle ann, const char *name, const char *str, long value, int dtype) {
    SQB_TRY DISPATCH(dtype, setPreference(BGA(real), name, str, value)) SQB_CATCH
}
int sqb_bg_annealer_get_preferences(sqb_handle ann, char *buf, int buflen, int dtype) { SQB_TRY DISPATCH(dtype, getPreferences(BGA(real), buf, buflen)) SQB_CATCH }
int sqb_bg_annealer_get_num_trotters(sqb_handle ann, int *m, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::Preferences prefs = BGA(real)->getPreferences();
        for (int i = 0; i < prefs.size(); ++i) if (prefs[i].name == sq::pnNumTrotters) *m = prefs[i].nTrotters;
    })
    SQB_CATCH
}
int sqb_bg_annealer_get_E(sqb_handle ann, void *E, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        const sq::VectorType<real> &e = BGA(real)->get_E();
        sqb_throwErrorIf(capacity < e.size, "E buffer too small (%d < %d).", capacity, e.size);
        memcpy(E, e.data, sizeof(real) * e.size);
    })
    SQB_CATCH
}
int sqb_bg_annealer_get_x(sqb_handle ann, signed char *x0, signed char *x1, int dtype) {
    SQB_TRY DISPATCH(dtype, { sq::SizeType N0, N1; BGA(real)->getProblemSize(&N0, &N1); splitPairs(x0, x1, BGA(real)->get_x(), N0, N1); }) SQB_CATCH
}
int sqb_bg_annealer_get_q(sqb_handle ann, signed char *q0, signed char *q1, int dtype) {
    SQB_TRY DISPATCH(dtype, { sq::SizeType N0, N1; BGA(real)->getProblemSize(&N0, &N1); splitPairs(q0, q1, BGA(real)->get_q(), N0, N1); }) SQB_CATCH
}
int sqb_bg_annealer_set_q(sqb_handle ann, const signed char *q0, const signed char *q1, int N0, int N1, int dtype) {
    SQB_TRY DISPATCH(dtype, BGA(real)->set_q(sq::BitSetPair(viewBits(q0, N0), viewBits(q1, N1)))) SQB_CATCH
}
int sqb_bg_annealer_set_qset(sqb_handle ann, const signed char *q0, const signed char *q1, int m, int N0, int N1, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::BitSetPairArray arr;
        for (int i = 0; i < m; ++i)
            arr.pushBack(sq::BitSetPair(viewBits(q0 + (size_t)i * N0, N0), viewBits(q1 + (size_t)i * N1, N1)));
        BGA(real)->set_qset(arr);
    })
    SQB_CATCH
}
int sqb_bg_annealer_randomize_spin(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->randomizeSpin()) SQB_CATCH }
int sqb_bg_annealer_calculate_E(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->calculate_E()) SQB_CATCH }
int sqb_bg_annealer_prepare(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->prepare()) SQB_CATCH }
int sqb_bg_annealer_make_solution(sqb_handle ann, int dtype) { SQB_TRY DISPATCH(dtype, BGA(real)->makeSolution()) SQB_CATCH }
int sqb_bg_annealer_get_system_E(sqb_handle ann, double G, double beta, double *E, int dtype) {
    SQB_TRY DISPATCH(dtype, *E = (double)BGA(real)->getSystemE((real)G, (real)beta)) SQB_CATCH
}
int sqb_bg_annealer_anneal_one_step(sqb_handle ann, double G, double beta, int dtype) {
    SQB_TRY DISPATCH(dtype, BGA(real)->annealOneStep((real)G, (real)beta)) SQB_CATCH
}

/* ---------------- dense-graph brute-force searcher ---------------- */
#define DGS(real) as<sqc::DenseGraphBFSearcher<real> >(s)
#define DGSX(real) dynamic_cast<sqb::DenseBFExtras *>(DGS(real))

int sqb_dg_bf_searcher_new(sqb_handle *s, int dtype) { SQB_TRY DISPATCH(dtype, *s = sqc::newDenseGraphBFSearcher<real>()) SQB_CATCH }
int sqb_dg_bf_searcher_delete(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(DGS(real))) SQB_CATCH }
int sqb_dg_bf_searcher_assign_device(sqb_handle s, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_dg_bf_searcher_set_qubo(sqb_handle s, const void *W, int N, int stride, int optimize, int dtype) {
    SQB_TRY DISPATCH(dtype, DGS(real)->setQUBO(mapMat<real>(W, N, N, stride), (sq::OptimizeMethod)optimize)) SQB_CATCH
}
int sqb_dg_bf_searcher_get_problem_size(sqb_handle s, int *N, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->getProblemSize(N)) SQB_CATCH }
int sqb_dg_bf_searcher_set_preference(sqb_handle s, const char *name, const char *str, long value, int dtype) {
    SQB_TRY DISPATCH(dtype, setPreference(DGS(real), name, str, value)) SQB_CATCH
}
int sqb_dg_bf_searcher_get_preferences(sqb_handle s, char *buf, int buflen, int dtype) { SQB_TRY DISPATCH(dtype, getPreferences(DGS(real), buf, buflen)) SQB_CATCH }
int sqb_dg_bf_searcher_prepare(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->prepare()) SQB_CATCH }
int sqb_dg_bf_searcher_calculate_E(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->calculate_E()) SQB_CATCH }
int sqb_dg_bf_searcher_make_solution(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->makeSolution()) SQB_CATCH }
int sqb_dg_bf_searcher_search_range(sqb_handle s, int *done, unsigned long long *cur_x, int dtype) {
    SQB_TRY DISPATCH(dtype, { sq::PackedBitSet x = 0; *done = DGS(real)->searchRange(&x) ? 1 : 0; if (cur_x) *cur_x = x; }) SQB_CATCH
}
int sqb_dg_bf_searcher_search(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, DGS(real)->search()) SQB_CATCH }
int sqb_dg_bf_searcher_get_num_solutions(sqb_handle s, int *n, int dtype) { SQB_TRY DISPATCH(dtype, *n = DGS(real)->get_x().size()) SQB_CATCH }
int sqb_dg_bf_searcher_get_x(sqb_handle s, signed char *x, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::SizeType N;
        DGS(real)->getProblemSize(&N);
        const sq::BitSetArray &xs = DGS(real)->get_x();
        sqb_throwErrorIf(capacity < xs.size(), "x buffer too small (%d < %d).", capacity, xs.size());
        copyBitSets(x, xs, N);
    })
    SQB_CATCH
}
int sqb_dg_bf_searcher_get_E(sqb_handle s, void *E, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        const sq::VectorType<real> &e = DGS(real)->get_E();
        sqb_throwErrorIf(capacity < e.size, "E buffer too small (%d < %d).", capacity, e.size);
        memcpy(E, e.data, sizeof(real) * e.size);
    })
    SQB_CATCH
}
int sqb_dg_bf_searcher_set_range(sqb_handle s, unsigned long long x_begin, unsigned long long x_end, int dtype) {
    SQB_TRY DISPATCH(dtype, DGSX(real)->setRange(x_begin, x_end)) SQB_CATCH
}
int sqb_dg_bf_searcher_get_Emin(sqb_handle s, double *Emin, int dtype) { SQB_TRY DISPATCH(dtype, *Emin = DGSX(real)->getEmin()) SQB_CATCH }
int sqb_dg_bf_searcher_get_packed_x(sqb_handle s, unsigned long long *x, int capacity, int *n, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        const sq::PackedBitSetArray &p = DGSX(real)->packedSolutions();
        *n = p.size();
        for (int i = 0; i < p.size() && i < capacity; ++i) x[i] = p[i];
    })
    SQB_CATCH
}
int sqb_dg_bf_searcher_set_packed_solutions(sqb_handle s, double Emin, const unsigned long long *x, int n, int dtype) {
    SQB_TRY DISPATCH(dtype, DGSX(real)->setPackedSolutions(Emin, x, n)) SQB_CATCH
}

/* ---------------- bipartite-graph brute-force searcher ---------------- */
#define BGS(real) as<sqc::BipartiteGraphBFSearcher<real> >(s)

int sqb_bg_bf_searcher_new(sqb_handle *s, int dtype) { SQB_TRY DISPATCH(dtype, *s = sqc::newBipartiteGraphBFSearcher<real>()) SQB_CATCH }
int sqb_bg_bf_searcher_delete(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(BGS(real))) SQB_CATCH }
int sqb_bg_bf_searcher_assign_device(sqb_handle s, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_bg_bf_searcher_set_qubo(sqb_handle s, const void *b0, const void *b1, const void *W, int N0, int N1, int stride, int optimize, int dtype) {
    SQB_TRY
    DISPATCH(dtype, BGS(real)->setQUBO(mapVec<real>(b0, N0), mapVec<real>(b1, N1), mapMat<real>(W, N1, N0, stride), (sq::OptimizeMethod)optimize))
    SQB_CATCH
}
int sqb_bg_bf_searcher_get_problem_size(sqb_handle s, int *N0, int *N1, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->getProblemSize(N0, N1)) SQB_CATCH }
int sqb_bg_bf_searcher_set_preference(sqb_handle s, const char *name, const char *str, long value, int dtype) {
    SQB_TRY DISPATCH(dtype, setPreference(BGS(real), name, str, value)) SQB_CATCH
}
int sqb_bg_bf_searcher_get_preferences(sqb_handle s, char *buf, int buflen, int dtype) { SQB_TRY DISPATCH(dtype, getPreferences(BGS(real), buf, buflen)) SQB_CATCH }
int sqb_bg_bf_searcher_prepare(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->prepare()) SQB_CATCH }
int sqb_bg_bf_searcher_calculate_E(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->calculate_E()) SQB_CATCH }
int sqb_bg_bf_searcher_make_solution(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->makeSolution()) SQB_CATCH }
int sqb_bg_bf_searcher_search_range(sqb_handle s, int *done, unsigned long long *cur_x0, unsigned long long *cur_x1, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::PackedBitSet x0 = 0, x1 = 0;
        *done = BGS(real)->searchRange(&x0, &x1) ? 1 : 0;
        if (cur_x0) *cur_x0 = x0;
        if (cur_x1) *cur_x1 = x1;
    })
    SQB_CATCH
}
int sqb_bg_bf_searcher_search(sqb_handle s, int dtype) { SQB_TRY DISPATCH(dtype, BGS(real)->search()) SQB_CATCH }
int sqb_bg_bf_searcher_get_num_solutions(sqb_handle s, int *n, int dtype) { SQB_TRY DISPATCH(dtype, *n = BGS(real)->get_x().size()) SQB_CATCH }
int sqb_bg_bf_searcher_get_x(sqb_handle s, signed char *x0, signed char *x1, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::SizeType N0, N1;
        BGS(real)->getProblemSize(&N0, &N1);
        const sq::BitSetPairArray &xs = BGS(real)->get_x();
        sqb_throwErrorIf(capacity < xs.size(), "x buffer too small (%d < %d).", capacity, xs.size());
        splitPairs(x0, x1, xs, N0, N1);
    })
    SQB_CATCH
}
int sqb_bg_bf_searcher_get_E(sqb_handle s, void *E, int capacity, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        const sq::VectorType<real> &e = BGS(real)->get_E();
        sqb_throwErrorIf(capacity < e.size, "E buffer too small (%d < %d).", capacity, e.size);
        memcpy(E, e.data, sizeof(real) * e.size);
    })
    SQB_CATCH
}

/* ---------------- formulas ---------------- */
#define DGF(real) as<sqc::DenseGraphFormulas<real> >(f)
#define BGF(real) as<sqc::BipartiteGraphFormulas<real> >(f)

int sqb_dg_formulas_new(sqb_handle *f, int dtype) { SQB_TRY DISPATCH(dtype, *f = sqc::newDenseGraphFormulas<real>()) SQB_CATCH }
int sqb_dg_formulas_delete(sqb_handle f, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(DGF(real))) SQB_CATCH }
int sqb_dg_formulas_assign_device(sqb_handle f, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, DGF(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_dg_formulas_calculate_E(sqb_handle f, void *E, const void *W, int N, int strideW, const signed char *x, int nBatch, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> Ev = mapVec<real>(E, nBatch);
        DGF(real)->calculate_E(&Ev, mapMat<real>(W, N, N, strideW), widen<real>(x, nBatch, N));
    })
    SQB_CATCH
}
int sqb_dg_formulas_calculate_hamiltonian(sqb_handle f, void *h, void *J, int strideJ, void *c, const void *W, int N, int strideW, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> hv = mapVec<real>(h, N);
        sq::MatrixType<real> Jm = mapMat<real>(J, N, N, strideJ);
        DGF(real)->calculateHamiltonian(&hv, &Jm, (real *)c, mapMat<real>(W, N, N, strideW));
    })
    SQB_CATCH
}
int sqb_dg_formulas_calculate_E_from_spin(sqb_handle f, void *E, const void *h, const void *J, int N, int strideJ, double c,
                                          const signed char *q, int nBatch, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> Ev = mapVec<real>(E, nBatch);
        DGF(real)->calculate_E(&Ev, mapVec<real>(h, N), mapMat<real>(J, N, N, strideJ), (real)c, widen<real>(q, nBatch, N));
    })
    SQB_CATCH
}
int sqb_bg_formulas_new(sqb_handle *f, int dtype) { SQB_TRY DISPATCH(dtype, *f = sqc::newBipartiteGraphFormulas<real>()) SQB_CATCH }
int sqb_bg_formulas_delete(sqb_handle f, int dtype) { SQB_TRY DISPATCH(dtype, sq::deleteInstance(BGF(real))) SQB_CATCH }
int sqb_bg_formulas_assign_device(sqb_handle f, sqb_handle dev, int dtype) { SQB_TRY DISPATCH(dtype, BGF(real)->assignDevice(*asDev(dev))) SQB_CATCH }
int sqb_bg_formulas_calculate_E(sqb_handle f, void *E, const void *b0, const void *b1, const void *W, int N0, int N1, int strideW,
                                const signed char *x0, const signed char *x1, int nBatch, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> Ev = mapVec<real>(E, nBatch);
        BGF(real)->calculate_E(&Ev, mapVec<real>(b0, N0), mapVec<real>(b1, N1), mapMat<real>(W, N1, N0, strideW),
                               widen<real>(x0, nBatch, N0), widen<real>(x1, nBatch, N1));
    })
    SQB_CATCH
}
int sqb_bg_formulas_calculate_E_2d(sqb_handle f, void *E, const void *b0, const void *b1, const void *W, int N0, int N1, int strideW,
                                   const signed char *x0, int n0, const signed char *x1, int n1, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::MatrixType<real> Em = mapMat<real>(E, n1, n0, n0);
        BGF(real)->calculate_E_2d(&Em, mapVec<real>(b0, N0), mapVec<real>(b1, N1), mapMat<real>(W, N1, N0, strideW),
                                  widen<real>(x0, n0, N0), widen<real>(x1, n1, N1));
    })
    SQB_CATCH
}
int sqb_bg_formulas_calculate_hamiltonian(sqb_handle f, void *h0, void *h1, void *J, int strideJ, void *c, const void *b0,
                                          const void *b1, const void *W, int N0, int N1, int strideW, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> v0 = mapVec<real>(h0, N0), v1 = mapVec<real>(h1, N1);
        sq::MatrixType<real> Jm = mapMat<real>(J, N1, N0, strideJ);
        BGF(real)->calculateHamiltonian(&v0, &v1, &Jm, (real *)c, mapVec<real>(b0, N0), mapVec<real>(b1, N1), mapMat<real>(W, N1, N0, strideW));
    })
    SQB_CATCH
}
int sqb_bg_formulas_calculate_E_from_spin(sqb_handle f, void *E, const void *h0, const void *h1, const void *J, int N0, int N1,
                                          int strideJ, double c, const signed char *q0, const signed char *q1, int nBatch, int dtype) {
    SQB_TRY
    DISPATCH(dtype, {
        sq::VectorType<real> Ev = mapVec<real>(E, nBatch);
        BGF(real)->calculate_E(&Ev, mapVec<real>(h0, N0), mapVec<real>(h1, N1), mapMat<real>(J, N1, N0, strideJ), (real)c,
                               widen<real>(q0, nBatch, N0), widen<real>(q1, nBatch, N1));
    })
    SQB_CATCH
}

} /* extern "C" */
