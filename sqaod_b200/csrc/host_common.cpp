/* host_common.cpp -- host-side pieces of the solver API that carry no device work: error/log conventions,
 * preference names, the solver state machine, default-algorithm fallback, tile-size rounding.
 * Behaviour follows the reference (sqaodc/common/defines.cpp:32-63, Preference.cpp:10-100, Solver.cpp:23-262,
 * Common.cpp:86-104); code is new. */
#include <sqaod_b200/sqaod_api.hpp>
#include <stdarg.h>
#include <stdio.h>
#include <strings.h>

namespace sqaod {

void throwErrorAt(const char *file, unsigned long line, const char *fmt, ...) {
    char msg[512], buf[768];
    va_list va;
    va_start(va, fmt);
    vsnprintf(msg, sizeof(msg), fmt, va);
    va_end(va);
    const char *base = strrchr(file, '/');
    snprintf(buf, sizeof(buf), "%s:%d %s\n", base ? base + 1 : file, (int)line, msg);
    throw std::runtime_error(buf);
}

void throwErrorAt(const char *file, unsigned long line) { throwErrorAt(file, line, "Error"); }

/* internal invariant violated: print and abort (reference: common/defines.cpp:12-26) */
void abortAt(const char *file, unsigned long line, const char *fmt, ...) {
    char msg[512];
    va_list va;
    va_start(va, fmt);
    vsnprintf(msg, sizeof(msg), fmt, va);
    va_end(va);
    fprintf(stderr, "%s:%d %s\n", file, (int)line, msg);
    abort();
}
void abortAt(const char *file, unsigned long line) { abortAt(file, line, "aborted"); }

void log(const char *fmt, ...) {
    static int verbose = -1;
    if (verbose < 0) {
        const char *env = getenv("SQAOD_VERBOSE");
        verbose = (env != NULL && *env != '0') ? 1 : 0;
    }
    if (!verbose) return;
    va_list va;
    va_start(va, fmt);
    vfprintf(stderr, fmt, va);
    va_end(va);
    fputc('\n', stderr);
}

bool isSQAAlgorithm(Algorithm algo) { return algo == algoNaive || algo == algoColoring; }

static const struct { Algorithm a; const char *s; } kAlgoNames[] = {
    {algoNaive, "naive"}, {algoColoring, "coloring"}, {algoBruteForceSearch, "brute_force_search"},
    {algoSADefault, "sa_default"}, {algoSANaive, "sa_naive"}, {algoSAColoring, "sa_coloring"}, {algoDefault, "default"}};

const char *algorithmToString(Algorithm algo) {
    for (const auto &e : kAlgoNames) if (e.a == algo) return e.s;
    return "unknown";
}
Algorithm algorithmFromString(const char *str) {
    for (const auto &e : kAlgoNames) if (strcasecmp(e.s, str) == 0) return e.a;
    return algoUnknown;
}

static const struct { PreferenceName p; const char *s; } kPrefNames[] = {
    {pnAlgorithm, "algorithm"}, {pnNumTrotters, "n_trotters"}, {pnTileSize, "tile_size"}, {pnTileSize0, "tile_size_0"},
    {pnTileSize1, "tile_size_1"}, {pnPrecision, "precision"}, {pnDevice, "device"}, {pnExperiment, "experiment"}};

PreferenceName preferenceNameFromString(const char *name) {
    for (const auto &e : kPrefNames) {
        if (e.p == pnDevice) continue; /* read-only, not settable by name (Preference.cpp:61-78) */
        if (strcasecmp(e.s, name) == 0) return e.p;
    }
    return pnUnknown;
}
const char *preferenceNameToString(PreferenceName pn) {
    for (const auto &e : kPrefNames) if (e.p == pn) return e.s;
    return "unknown";
}

void unpackBitSet(BitSet *unpacked, PackedBitSet packed, int N) {
    unpacked->resize(N);
    for (int pos = 0; pos < N; ++pos) (*unpacked)(pos) = (char)((packed >> (N - 1 - pos)) & 1);
}

template <class real> bool isSymmetric(const MatrixType<real> &W) {
    if (W.rows != W.cols) return false;
    for (SizeType j = 0; j < W.rows; ++j)
        for (SizeType i = 0; i < j; ++i)
            if (W(i, j) != W(j, i)) return false;
    return true;
}
template bool isSymmetric<float>(const MatrixType<float> &);
template bool isSymmetric<double>(const MatrixType<double> &);

void deleteInstance(NullBase *instance) { delete instance; }

template <class real> static const char *typeString();
template <> const char *typeString<float>() { return "float"; }
template <> const char *typeString<double>() { return "double"; }

/* ---- Solver ---- */
template <class real> void Solver<real>::setPreferences(const Preferences &prefs) {
    for (Preferences::const_iterator it = prefs.begin(); it != prefs.end(); ++it) setPreference(*it);
}

/* State dependencies (Solver.cpp:40-67): setting the problem invalidates prepared/q/E/solution; (re)preparing
 * invalidates q/E/solution; setting q invalidates E/solution.  The seed flag is independent. */
template <class real> void Solver<real>::clearState(SolverState s) {
    int mask = 0;
    switch (s) {
    case solRandSeedGiven: mask = solRandSeedGiven; break;
    case solProblemSet: mask = solProblemSet | solPrepared | solQSet | solEAvailable | solSolutionAvailable; break;
    case solPrepared: mask = solPrepared | solQSet | solEAvailable | solSolutionAvailable; break;
    case solQSet: mask = solQSet | solEAvailable | solSolutionAvailable; break;
    case solEAvailable:
    case solSolutionAvailable: mask = solEAvailable | solSolutionAvailable; break;
    default: break;
    }
    solverState_ &= ~mask;
}
template <class real> void Solver<real>::setState(SolverState s) {
    clearState(s);
    solverState_ |= s;
}
template <class real> void Solver<real>::throwErrorIfProblemNotSet() const {
    sqb_throwErrorIf(!isProblemSet(), "Problem is not set.");
}
template <class real> void Solver<real>::throwErrorIfNotPrepared() const {
    throwErrorIfProblemNotSet();
    sqb_throwErrorIf(!isPrepared(), "not prepared, call prepare() in advance.");
}
template <class real> void Solver<real>::throwErrorIfQNotSet() const {
    sqb_throwErrorIf(!isQSet(), "Bits(x or q) not initialized.  Plase set or randomize in advance.");
}

/* ---- Annealer ---- */
template <class real> Preferences Annealer<real>::getPreferences() const {
    Preferences prefs;
    prefs.pushBack(Preference(pnAlgorithm, this->getAlgorithm()));
    prefs.pushBack(Preference(pnNumTrotters, m_));
    prefs.pushBack(Preference(pnPrecision, typeString<real>()));
    return prefs;
}
template <class real> void Annealer<real>::setPreference(const Preference &pref) {
    if (pref.name == pnNumTrotters) {
        sqb_throwErrorIf(pref.nTrotters <= 0, "# trotters must be a positive integer.");
        if (m_ != pref.nTrotters) Solver<real>::clearState(Solver<real>::solPrepared);
        m_ = pref.nTrotters;
    } else if (pref.name == pnAlgorithm) {
        this->selectAlgorithm(pref.algo);
    }
}
template <class real>
void Annealer<real>::selectDefaultAlgorithm(Algorithm algoOrg, Algorithm algoDef, Algorithm algoSADef) {
    if (algoOrg == algoDefault) { algo_ = algoDef; return; }
    if (algoOrg == algoSADefault) { algo_ = algoSADef; return; }
    algo_ = isSQAAlgorithm(algoOrg) ? algoDef : algoSADef;
    log("%s is not supported, selecting the default algorithm of %s.", algorithmToString(algoOrg), algorithmToString(algo_));
}
template <class real> void Annealer<real>::selectDefaultSAAlgorithm(Algorithm algoOrg, Algorithm algoSADef) {
    if (algoOrg != algoSADef) {
        algo_ = algoSADef;
        log("Selecting %s as the default SA algorithm.", algorithmToString(algo_));
    }
}

/* ---- brute-force searchers ---- */
template <class real> Preferences DenseGraphBFSearcher<real>::getPreferences() const {
    Preferences prefs;
    prefs.pushBack(Preference(pnAlgorithm, algoBruteForceSearch));
    prefs.pushBack(Preference(pnTileSize, tileSize_));
    prefs.pushBack(Preference(pnPrecision, typeString<real>()));
    return prefs;
}
static SizeType adjustedTile(SizeType requested, const char *what) {
    sqb_throwErrorIf(requested <= 0, "%s must be a positive integer.", what);
    SizeType t = roundUp(requested, 256); /* Solver.cpp:205 */
    if (t != requested) log("%s is adjusted to %d.", what, t);
    return t;
}
template <class real> void DenseGraphBFSearcher<real>::setPreference(const Preference &pref) {
    if (pref.name == pnTileSize) tileSize_ = adjustedTile(pref.tileSize, "tileSize");
}
template <class real> void DenseGraphBFSearcher<real>::search() {
    this->prepare();
    while (!searchRange(NULL)) {}
    this->makeSolution();
}
template <class real> Preferences BipartiteGraphBFSearcher<real>::getPreferences() const {
    Preferences prefs;
    prefs.pushBack(Preference(pnAlgorithm, algoBruteForceSearch));
    prefs.pushBack(Preference(pnTileSize0, tileSize0_));
    prefs.pushBack(Preference(pnTileSize1, tileSize1_));
    prefs.pushBack(Preference(pnPrecision, typeString<real>()));
    return prefs;
}
template <class real> void BipartiteGraphBFSearcher<real>::setPreference(const Preference &pref) {
    if (pref.name == pnTileSize0) tileSize0_ = adjustedTile(pref.tileSize, "tileSize0");
    if (pref.name == pnTileSize1) tileSize1_ = adjustedTile(pref.tileSize, "tileSize1");
}
template <class real> void BipartiteGraphBFSearcher<real>::search() {
    this->prepare();
    while (!searchRange(NULL, NULL)) {}
    this->makeSolution();
}

template struct Solver<float>;
template struct Solver<double>;
template struct Annealer<float>;
template struct Annealer<double>;
template struct DenseGraphBFSearcher<float>;
template struct DenseGraphBFSearcher<double>;
template struct BipartiteGraphBFSearcher<float>;
template struct BipartiteGraphBFSearcher<double>;

} // namespace sqaod
