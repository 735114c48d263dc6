/* sqaod_b200/sqaod_api.hpp -- C++ solver interface of the B200 back end.
 *
 * Source-compatible re-statement of the abstract solver API the reference exposes to its Python glue and to
 * native users (reference: sqaodc/sqaodc.h:29-61, sqaodc/common/Solver.h:17-252, common/Preference.h:8-61,
 * common/Matrix.h:12-395, common/Array.h:29-249, common/types.h:10-62, sqaodc/cuda/api.h:16-83).  Same
 * namespaces, class names, virtual method names and argument meaning, so code written against
 * sqaod::cuda::DenseGraphAnnealer<real> etc. recompiles against this header and links libsqaod_b200.so.
 * The implementation behind it is new (hand-written sm_100a CUDA, see sqaod_b200/csrc/).
 *
 * Not binary compatible with libsqaodc_cuda.so.1 (containers are laid out differently); the stable binary
 * boundary of this project is the C ABI in include/sqaod_b200.h.
 */
#pragma once
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <assert.h>
#include <vector>
#include <stdexcept>

namespace sqaod {

typedef int SizeType;
typedef int IdxType;
typedef unsigned long long PackedBitSet;

struct PackedBitSetPair {
    PackedBitSetPair() : bits0(0), bits1(0) {}
    PackedBitSetPair(PackedBitSet b0, PackedBitSet b1) : bits0(b0), bits1(b1) {}
    PackedBitSet bits0, bits1;
};

struct Dim {
    Dim() : rows(-1), cols(-1) {}
    Dim(SizeType r, SizeType c) : rows(r), cols(c) {}
    Dim transpose() const { return Dim(cols, rows); }
    SizeType rows, cols;
    friend bool operator==(const Dim &a, const Dim &b) { return a.rows == b.rows && a.cols == b.cols; }
    friend bool operator!=(const Dim &a, const Dim &b) { return !(a == b); }
};

struct NullBase {
    virtual ~NullBase() {}
};

template <class V> inline V divru(V v, int base) { return (v + base - 1) / base; }
template <class V> inline V roundUp(V v, int base) { return divru(v, base) * base; }

/* error / log conventions of the reference (common/defines.cpp:12-63): recoverable errors throw
 * std::runtime_error("file:line msg\n"); log() prints iff SQAOD_VERBOSE is set and != '0'. */
void throwErrorAt(const char *file, unsigned long line, const char *fmt, ...);
void throwErrorAt(const char *file, unsigned long line);
void abortAt(const char *file, unsigned long line, const char *fmt, ...);
void abortAt(const char *file, unsigned long line);
void log(const char *fmt, ...);
#define sqb_throwError(...) ::sqaod::throwErrorAt(__FILE__, __LINE__, __VA_ARGS__)
#define sqb_throwErrorIf(cond, ...) do { if (cond) ::sqaod::throwErrorAt(__FILE__, __LINE__, __VA_ARGS__); } while (0)

/* ---- growable array (reference: common/Array.h ArrayType) ---- */
template <class V> class ArrayType {
public:
    typedef V ValueType;
    typedef typename std::vector<V>::iterator iterator;
    typedef typename std::vector<V>::const_iterator const_iterator;
    ArrayType() {}
    explicit ArrayType(SizeType capacity) { v_.reserve(capacity); }
    void reserve(SizeType n) { v_.reserve(n); }
    void clear() { v_.clear(); }
    bool empty() const { return v_.empty(); }
    SizeType size() const { return (SizeType)v_.size(); }
    void pushBack(const V &x) { v_.push_back(x); }
    void erase(SizeType idx) { v_.erase(v_.begin() + idx); }
    template <class It> void insert(It b, It e) { v_.insert(v_.end(), b, e); }
    V &operator[](SizeType i) { return v_[i]; }
    const V &operator[](SizeType i) const { return v_[i]; }
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    const_iterator begin() const { return v_.begin(); }
    const_iterator end() const { return v_.end(); }
    V *data() { return v_.data(); }
    const V *data() const { return v_.data(); }
    friend bool operator==(const ArrayType &a, const ArrayType &b) { return a.v_ == b.v_; }
    friend bool operator!=(const ArrayType &a, const ArrayType &b) { return !(a == b); }
private:
    std::vector<V> v_;
};
typedef ArrayType<PackedBitSet> PackedBitSetArray;
typedef ArrayType<PackedBitSetPair> PackedBitSetPairArray;

/* ---- host vector / matrix views (reference: common/Matrix.h).  Row-major; `stride` in elements; owned
 * storage is 64-byte aligned with rows padded to 64 bytes; `mapped` views borrow caller memory. ---- */
template <class V> struct VectorType {
    typedef V ValueType;
    VectorType() : size(-1), data(NULL), mapped(false) {}
    explicit VectorType(SizeType n) : size(-1), data(NULL), mapped(false) { allocate(n); }
    VectorType(V *d, SizeType n) : size(n), data(d), mapped(true) {}
    VectorType(const VectorType &o) : size(-1), data(NULL), mapped(false) { copyFrom(o); }
    virtual ~VectorType() { if (!mapped) release(); }
    VectorType &operator=(const VectorType &o) { copyFrom(o); return *this; }
    VectorType &operator=(const V &v) { for (SizeType i = 0; i < size; ++i) data[i] = v; return *this; }
    void map(V *d, SizeType n) { if (!mapped) release(); data = d; size = n; mapped = true; }
    void allocate(SizeType n) {
        size = n;
        size_t bytes = roundUp((size_t)(n > 0 ? n : 1) * sizeof(V), 64);
        data = (V *)::aligned_alloc(64, bytes);
        memset(data, 0, bytes);
    }
    void release() { if (data) ::free(data); data = NULL; size = -1; }
    void resize(SizeType n) { if (mapped) { assert(n == size); return; } if (n != size) { release(); allocate(n); } }
    void copyFrom(const VectorType &o) {
        if (this == &o) return;
        if (mapped) { assert(size == o.size); } else if (size != o.size) { release(); allocate(o.size); }
        if (o.size > 0) memcpy(data, o.data, sizeof(V) * o.size);
    }
    V &operator()(IdxType i) { return data[i]; }
    const V &operator()(IdxType i) const { return data[i]; }
    V sum() const { V s = V(); for (SizeType i = 0; i < size; ++i) s += data[i]; return s; }
    V min() const { V s = data[0]; for (SizeType i = 1; i < size; ++i) if (data[i] < s) s = data[i]; return s; }
    static VectorType zeros(SizeType n) { VectorType v(n); v = V(0); return v; }
    static VectorType ones(SizeType n) { VectorType v(n); v = V(1); return v; }
    SizeType size;
    V *data;
    bool mapped;
};
template <class V> bool operator==(const VectorType<V> &a, const VectorType<V> &b) {
    if (a.size != b.size) return false;
    for (SizeType i = 0; i < a.size; ++i) if (a.data[i] != b.data[i]) return false;
    return true;
}
template <class V> bool operator!=(const VectorType<V> &a, const VectorType<V> &b) { return !(a == b); }

template <class V> struct MatrixType {
    typedef V ValueType;
    MatrixType() : rows(-1), cols(-1), stride(0), data(NULL), mapped(false) {}
    MatrixType(SizeType r, SizeType c) : rows(-1), cols(-1), stride(0), data(NULL), mapped(false) { allocate(r, c); }
    explicit MatrixType(const Dim &d) : rows(-1), cols(-1), stride(0), data(NULL), mapped(false) { allocate(d.rows, d.cols); }
    MatrixType(V *d, SizeType r, SizeType c, SizeType s) : rows(r), cols(c), stride(s), data(d), mapped(true) {}
    MatrixType(const MatrixType &o) : rows(-1), cols(-1), stride(0), data(NULL), mapped(false) { copyFrom(o); }
    virtual ~MatrixType() { if (!mapped) release(); }
    MatrixType &operator=(const MatrixType &o) { copyFrom(o); return *this; }
    Dim dim() const { return Dim(rows, cols); }
    void map(V *d, SizeType r, SizeType c, SizeType s) { if (!mapped) release(); data = d; rows = r; cols = c; stride = s; mapped = true; }
    void allocate(SizeType r, SizeType c) {
        rows = r; cols = c;
        stride = roundUp(c > 0 ? c : 1, (int)(64 / sizeof(V)));
        size_t bytes = (size_t)(r > 0 ? r : 1) * stride * sizeof(V);
        data = (V *)::aligned_alloc(64, roundUp(bytes, 64));
        memset(data, 0, bytes);
    }
    void release() { if (data) ::free(data); data = NULL; rows = cols = -1; }
    void resize(SizeType r, SizeType c) { if (mapped) { assert(r == rows && c == cols); return; } if (r != rows || c != cols) { release(); allocate(r, c); } }
    void resize(const Dim &d) { resize(d.rows, d.cols); }
    void copyFrom(const MatrixType &o) {
        if (this == &o) return;
        if (mapped) { assert(rows == o.rows && cols == o.cols); } else if (rows != o.rows || cols != o.cols) { release(); allocate(o.rows, o.cols); }
        for (SizeType r = 0; r < o.rows; ++r) memcpy(rowPtr(r), o.rowPtr(r), sizeof(V) * o.cols);
    }
    V &operator()(IdxType r, IdxType c) { return data[(size_t)r * stride + c]; }
    const V &operator()(IdxType r, IdxType c) const { return data[(size_t)r * stride + c]; }
    V *rowPtr(IdxType r) { return data + (size_t)r * stride; }
    const V *rowPtr(IdxType r) const { return data + (size_t)r * stride; }
    SizeType rows, cols, stride;
    V *data;
    bool mapped;
};
template <class V> bool operator==(const MatrixType<V> &a, const MatrixType<V> &b) {
    if (a.rows != b.rows || a.cols != b.cols) return false;
    for (SizeType r = 0; r < a.rows; ++r) for (SizeType c = 0; c < a.cols; ++c) if (a(r, c) != b(r, c)) return false;
    return true;
}

/* element-wise conversion into a newly owned container (reference: common/Matrix.h:209-210, 363-365; the pyglue formulas
 * functions widen 0/1 and +-1 int8 arrays to `real` with it) */
template <class newV, class V> MatrixType<newV> cast(const MatrixType<V> &mat) {
    MatrixType<newV> out(mat.rows, mat.cols);
    for (SizeType r = 0; r < mat.rows; ++r) for (SizeType c = 0; c < mat.cols; ++c) out(r, c) = (newV)mat(r, c);
    return out;
}
template <class newV, class V> VectorType<newV> cast(const VectorType<V> &vec) {
    VectorType<newV> out(vec.size);
    for (SizeType i = 0; i < vec.size; ++i) out(i) = (newV)vec(i);
    return out;
}

typedef VectorType<char> BitSet;
typedef ArrayType<BitSet> BitSetArray;
struct BitSetPair {
    BitSetPair() {}
    BitSetPair(const BitSet &b0, const BitSet &b1) : bits0(b0), bits1(b1) {}
    BitSet bits0, bits1;
};
inline bool operator==(const BitSetPair &a, const BitSetPair &b) { return a.bits0 == b.bits0 && a.bits1 == b.bits1; }
typedef ArrayType<BitSetPair> BitSetPairArray;
typedef MatrixType<char> BitMatrix;

/* ---- preferences (reference: common/Preference.h) ---- */
enum Algorithm { algoUnknown, algoDefault, algoNaive, algoColoring, algoBruteForceSearch, algoSADefault, algoSANaive, algoSAColoring };
bool isSQAAlgorithm(Algorithm algo);
const char *algorithmToString(Algorithm algo);
Algorithm algorithmFromString(const char *str);

enum PreferenceName { pnUnknown = 0, pnAlgorithm = 1, pnNumTrotters = 2, pnTileSize = 3, pnTileSize0 = 4, pnTileSize1 = 5,
                      pnPrecision = 6, pnDevice = 7, pnExperiment = 100 };
PreferenceName preferenceNameFromString(const char *name);
const char *preferenceNameToString(PreferenceName pn);

struct Preference {
    Preference() : name(pnUnknown) { size = 0; }
    Preference(PreferenceName n, SizeType s) : name(n) { str = NULL; size = s; }
    Preference(PreferenceName n, Algorithm a) : name(n) { str = NULL; algo = a; }
    Preference(PreferenceName n, const char *s) : name(n) { str = s; }
    PreferenceName name; /* first member, as in the reference */
    union {
        SizeType size; const char *str; Algorithm algo; SizeType tileSize; SizeType nTrotters;
        const char *precision; const char *device; int experiment;
    };
};
typedef ArrayType<Preference> Preferences;

enum OptimizeMethod { optNone = -1, optMinimize = 0, optMaximize = 1 };

/* ---- abstract solvers (reference: common/Solver.h) ---- */
template <class real> struct Solver : NullBase {
    virtual ~Solver() {}
    virtual Algorithm selectAlgorithm(Algorithm algo) = 0;
    virtual Algorithm getAlgorithm() const = 0;
    virtual Preferences getPreferences() const = 0;
    virtual void setPreference(const Preference &pref) = 0;
    template <class V> void setPreference(PreferenceName name, const V value) { setPreference(Preference(name, value)); }
    void setPreferences(const Preferences &prefs);
    virtual const VectorType<real> &get_E() const = 0;
    virtual void prepare() = 0;
    virtual void calculate_E() = 0;
    virtual void makeSolution() = 0;
protected:
    Solver() : solverState_(solNone), om_(optNone) {}
    enum SolverState { solNone = 0, solProblemSet = 1, solPrepared = 2, solEAvailable = 4, solSolutionAvailable = 8,
                       solRandSeedGiven = 16, solQSet = 32 };
    void setState(SolverState s);
    void clearState(SolverState s);
    bool isRandSeedGiven() const { return (solverState_ & solRandSeedGiven) != 0; }
    bool isProblemSet() const { return (solverState_ & solProblemSet) != 0; }
    bool isPrepared() const { return (solverState_ & solPrepared) != 0; }
    bool isQSet() const { return (solverState_ & solQSet) != 0; }
    bool isEAvailable() const { return (solverState_ & solEAvailable) != 0; }
    bool isSolutionAvailable() const { return (solverState_ & solSolutionAvailable) != 0; }
    void throwErrorIfProblemNotSet() const;
    void throwErrorIfNotPrepared() const;
    void throwErrorIfQNotSet() const;
    int solverState_;
    OptimizeMethod om_;
};

template <class real> struct BFSearcher : Solver<real> {
    virtual Algorithm selectAlgorithm(Algorithm) { return algoBruteForceSearch; }
    virtual Algorithm getAlgorithm() const { return algoBruteForceSearch; }
    virtual void search() = 0;
};

template <class real> struct Annealer : Solver<real> {
    virtual Algorithm getAlgorithm() const { return algo_; }
    virtual Preferences getPreferences() const;
    virtual void setPreference(const Preference &pref);
    using Solver<real>::setPreference;
    virtual void seed(unsigned long long seed) = 0;
    virtual void randomizeSpin() = 0;
    virtual void annealOneStep(real G, real beta) = 0;
    virtual real getSystemE(real G, real beta) const = 0;
protected:
    Annealer() : algo_(algoDefault), m_(0) {}
    void selectDefaultAlgorithm(Algorithm algoOrg, Algorithm algoDef, Algorithm algoSADef);
    void selectDefaultSAAlgorithm(Algorithm algoOrg, Algorithm algoSADef);
    Algorithm algo_;
    SizeType m_;
};

template <class real> struct DenseGraphSolver {
    virtual ~DenseGraphSolver() {}
    void getProblemSize(SizeType *N) const { *N = N_; }
    virtual void setQUBO(const MatrixType<real> &W, OptimizeMethod om = optMinimize) = 0;
    virtual const BitSetArray &get_x() const = 0;
protected:
    DenseGraphSolver() : N_(0) {}
    SizeType N_;
};

template <class real> struct BipartiteGraphSolver {
    virtual ~BipartiteGraphSolver() {}
    void getProblemSize(SizeType *N0, SizeType *N1) const { *N0 = N0_; *N1 = N1_; }
    virtual void setQUBO(const VectorType<real> &b0, const VectorType<real> &b1, const MatrixType<real> &W,
                         OptimizeMethod om = optMinimize) = 0;
    virtual const BitSetPairArray &get_x() const = 0;
protected:
    BipartiteGraphSolver() : N0_(0), N1_(0) {}
    SizeType N0_, N1_;
};

template <class real> struct DenseGraphBFSearcher : BFSearcher<real>, DenseGraphSolver<real> {
    virtual Preferences getPreferences() const;
    virtual void setPreference(const Preference &pref);
    using Solver<real>::setPreference;
    virtual bool searchRange(PackedBitSet *curXEnd) = 0;
    virtual void search();
protected:
    DenseGraphBFSearcher() : x_(0), xMax_(0), tileSize_(0) {}
    PackedBitSet x_, xMax_;
    SizeType tileSize_;
};

template <class real> struct DenseGraphAnnealer : Annealer<real>, DenseGraphSolver<real> {
    virtual void setHamiltonian(const VectorType<real> &h, const MatrixType<real> &J, real c = real(0.)) = 0;
    virtual void getHamiltonian(VectorType<real> *h, MatrixType<real> *J, real *c) const = 0;
    virtual void set_q(const BitSet &q) = 0;
    virtual void set_qset(const BitSetArray &q) = 0;
    virtual const BitSetArray &get_q() const = 0;
};

template <class real> struct BipartiteGraphBFSearcher : BFSearcher<real>, BipartiteGraphSolver<real> {
    virtual Preferences getPreferences() const;
    virtual void setPreference(const Preference &pref);
    using Solver<real>::setPreference;
    virtual bool searchRange(PackedBitSet *curX0End, PackedBitSet *curX1End) = 0;
    virtual void search();
protected:
    BipartiteGraphBFSearcher() : x0_(0), x1_(0), x0max_(0), x1max_(0), tileSize0_(0), tileSize1_(0) {}
    PackedBitSet x0_, x1_, x0max_, x1max_;
    SizeType tileSize0_, tileSize1_;
};

template <class real> struct BipartiteGraphAnnealer : Annealer<real>, BipartiteGraphSolver<real> {
    virtual void setHamiltonian(const VectorType<real> &h0, const VectorType<real> &h1, const MatrixType<real> &J,
                                real c = real(0.)) = 0;
    virtual void getHamiltonian(VectorType<real> *h0, VectorType<real> *h1, MatrixType<real> *J, real *c) const = 0;
    virtual void set_q(const BitSetPair &qPair) = 0;
    virtual void set_qset(const BitSetPairArray &qPairs) = 0;
    virtual const BitSetPairArray &get_q() const = 0;
};

/* ---- stateless formulas (reference: common/Formulas.h:7-70) ---- */
template <class real> struct DenseGraphFormulas : NullBase {
    typedef MatrixType<real> Matrix;
    typedef VectorType<real> Vector;
    virtual void calculate_E(real *E, const Matrix &W, const Vector &x) = 0;
    virtual void calculate_E(Vector *E, const Matrix &W, const Matrix &x) = 0;
    virtual void calculateHamiltonian(Vector *h, Matrix *J, real *c, const Matrix &W) = 0;
    virtual void calculate_E(real *E, const Vector &h, const Matrix &J, real c, const Vector &q) = 0;
    virtual void calculate_E(Vector *E, const Vector &h, const Matrix &J, real c, const Matrix &q) = 0;
};
template <class real> struct BipartiteGraphFormulas : NullBase {
    typedef MatrixType<real> Matrix;
    typedef VectorType<real> Vector;
    virtual void calculate_E(real *E, const Vector &b0, const Vector &b1, const Matrix &W, const Vector &x0, const Vector &x1) = 0;
    virtual void calculate_E(Vector *E, const Vector &b0, const Vector &b1, const Matrix &W, const Matrix &x0, const Matrix &x1) = 0;
    virtual void calculate_E_2d(Matrix *E, const Vector &b0, const Vector &b1, const Matrix &W, const Matrix &x0, const Matrix &x1) = 0;
    virtual void calculateHamiltonian(Vector *h0, Vector *h1, Matrix *J, real *c, const Vector &b0, const Vector &b1, const Matrix &W) = 0;
    virtual void calculate_E(real *E, const Vector &h0, const Vector &h1, const Matrix &J, real c, const Vector &q0, const Vector &q1) = 0;
    virtual void calculate_E(Vector *E, const Vector &h0, const Vector &h1, const Matrix &J, real c, const Matrix &q0, const Matrix &q1) = 0;
};

/* common helpers (reference: common/Common.cpp:78-145) */
void unpackBitSet(BitSet *unpacked, PackedBitSet packed, int N);
template <class real> bool isSymmetric(const MatrixType<real> &W);

void deleteInstance(NullBase *instance);

/* ---- CUDA back end (reference: sqaodc/cuda/api.h:16-83, sqaodc/sqaodc.h:39-61) ---- */
namespace cuda {

struct Device : NullBase {
    virtual ~Device() {}
    virtual int devNo() const = 0;
    virtual void initialize(int devNo = 0) = 0;
    virtual void finalize() = 0;
};

struct DeviceAssigner {
    virtual ~DeviceAssigner() {}
    virtual void assignDevice(Device &device) = 0;
};

template <class real> struct DenseGraphAnnealer : DeviceAssigner, sqaod::DenseGraphAnnealer<real> {};
template <class real> struct DenseGraphBFSearcher : DeviceAssigner, sqaod::DenseGraphBFSearcher<real> {};
template <class real> struct BipartiteGraphAnnealer : DeviceAssigner, sqaod::BipartiteGraphAnnealer<real> {};
template <class real> struct BipartiteGraphBFSearcher : DeviceAssigner, sqaod::BipartiteGraphBFSearcher<real> {};
template <class real> struct DenseGraphFormulas : DeviceAssigner, sqaod::DenseGraphFormulas<real> {};
template <class real> struct BipartiteGraphFormulas : DeviceAssigner, sqaod::BipartiteGraphFormulas<real> {};

Device *newDevice(int devNo = -1);
template <class real> DenseGraphBFSearcher<real> *newDenseGraphBFSearcher();
template <class real> DenseGraphAnnealer<real> *newDenseGraphAnnealer();
template <class real> BipartiteGraphBFSearcher<real> *newBipartiteGraphBFSearcher();
template <class real> BipartiteGraphAnnealer<real> *newBipartiteGraphAnnealer();
template <class real> DenseGraphFormulas<real> *newDenseGraphFormulas();
template <class real> BipartiteGraphFormulas<real> *newBipartiteGraphFormulas();

} // namespace cuda
} // namespace sqaod
