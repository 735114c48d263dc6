"""Backend-agnostic helpers: preferences, argument hygiene.  Mirrors sqaodpy/sqaod/common/{preference,common,checkers}.py
(names and behaviour), written fresh."""
import numpy as np


class Algorithm(object):  # sqaodpy/sqaod/common/preference.py:3-20
    default = 'default'
    naive = 'naive'
    coloring = 'coloring'
    brute_force_search = 'brute_force_search'
    sa_default = 'sa_default'
    sa_naive = 'sa_naive'
    sa_coloring = 'sa_coloring'

    @staticmethod
    def is_sqa(algo):
        return algo in (Algorithm.default, Algorithm.naive, Algorithm.coloring)


algorithm = Algorithm()


class Minimize(object):
    @staticmethod
    def sign(v):
        return v.copy() if hasattr(v, 'copy') else v

    @staticmethod
    def best(values):
        return min(values)

    @staticmethod
    def sort(values):
        return sorted(values)

    def __int__(self):
        return 0

    def __repr__(self):
        return 'minimize'


class Maximize(object):
    @staticmethod
    def sign(v):
        return -v

    @staticmethod
    def best(values):
        return max(values)

    @staticmethod
    def sort(values):
        return sorted(values, reverse=True)

    def __int__(self):
        return 1

    def __repr__(self):
        return 'maximize'


minimize = Minimize()
maximize = Maximize()

rtol_fp32, atol_fp32 = 1e-5, 1e-6
rtol_fp64, atol_fp64 = 1e-9, 1e-10


def is_symmetric(mat):
    if mat.dtype == np.float64:
        return np.allclose(mat, mat.T, rtol_fp64, atol_fp64)
    return np.allclose(mat, mat.T, rtol_fp32, atol_fp32)


def is_triangular(mat):
    atol = rtol_fp64 if mat.dtype == np.float64 else rtol_fp32
    if not np.any(np.abs(np.triu(mat, 1)) > atol):
        return True
    return not np.any(np.abs(np.tril(mat, -1)) > atol)


def symmetrize(mat):
    """sqaodpy/sqaod/common/common.py:55-62: symmetric matrices pass, triangular ones are mirrored, others rejected."""
    if is_symmetric(mat):
        if not np.array_equal(mat, mat.T):      # the native layer demands exact symmetry
            mat = (mat + mat.T) * mat.dtype.type(0.5)
        return mat
    if is_triangular(mat):
        return (mat + mat.T) * mat.dtype.type(0.5)
    raise RuntimeError('given matrix is not triangular nor symmetric.')


def fix_type(obj, dtype):
    """C-contiguous ndarray(s) of `dtype` (common.py:102-121)."""
    if isinstance(obj, np.ndarray):
        return np.ascontiguousarray(obj, dtype=dtype)
    try:
        return [np.ascontiguousarray(o, dtype=dtype) for o in obj]
    except TypeError:
        raise RuntimeError('Fix failed.')


def generate_random_symmetric_W(N, wmin=-0.5, wmax=0.5, dtype=np.float64):
    W = np.zeros((N, N), dtype)
    iu = np.triu_indices(N)
    W[iu] = np.random.random(len(iu[0]))
    W = W + np.tril(np.ones((N, N)), -1) * W.T
    return np.asarray(W * (wmax - wmin) + wmin, dtype)


def create_bitset_sequence(vals, nbits):
    vals = list(vals)
    x = np.empty((len(vals), nbits), np.int8)
    for i, v in enumerate(vals):
        for pos in range(nbits):
            x[i][pos] = (int(v) >> (nbits - 1 - pos)) & 1
    return x


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def check_dense_qubo(W):
    _check(isinstance(W, np.ndarray) and W.ndim == 2 and W.shape[0] == W.shape[1], 'W is not a square matrix.')


def check_dense_hJc(h, J, c):
    _check(np.ndim(h) == 1 and np.ndim(J) == 2 and J.shape[0] == J.shape[1] == h.shape[0], 'wrong shape for h, J.')
    _check(np.ndim(c) == 0 or np.size(c) == 1, 'c is not a scalar.')


def check_bipartite_qubo(b0, b1, W):
    _check(np.ndim(b0) == 1 and np.ndim(b1) == 1 and np.ndim(W) == 2, 'wrong shape for b0, b1, W.')
    _check(W.shape == (b1.shape[0], b0.shape[0]), 'dimension mismatch for b0, b1, W.')
