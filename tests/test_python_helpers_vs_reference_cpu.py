"""sqaod_b200.common (the host-side helpers of the Python package: symmetrize, fix_type, generate_random_symmetric_W,
create_bitset_sequence, the minimize / maximize tags) against the reference's own sqaod.common, whose files are staged unmodified under
oracle/_ref/refsuite by `make -C oracle glue` (build container).  Pure host logic: no GPU, no solver call."""
import importlib
import os
import sys
import types
import warnings
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, 'oracle', '_ref', 'refsuite')


@pytest.fixture(scope='module')
def ref_common():
    if not os.path.isdir(os.path.join(SUITE, 'sqaod', 'common')):
        pytest.skip('reference Python package not staged (run `make -C oracle glue` where /root/reference exists)')
    saved = {k: v for k, v in sys.modules.items() if k == 'sqaod' or k.startswith('sqaod.')}
    pkg = types.ModuleType('sqaod')
    pkg.__path__ = [os.path.join(SUITE, 'sqaod')]
    sys.modules['sqaod'] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        mod = importlib.import_module('sqaod.common')
        pref = importlib.import_module('sqaod.common.preference')
    yield mod, pref
    for k in [k for k in sys.modules if k == 'sqaod' or k.startswith('sqaod.')]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.fixture(scope='module')
def ours():
    sys.path.insert(0, ROOT)
    from sqaod_b200 import common
    return common


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_symmetrize_and_predicates(ref_common, ours, dtype):
    ref, _ = ref_common
    rng = np.random.default_rng(5)
    A = rng.random((9, 9)).astype(dtype) - dtype(0.5)
    sym = ((A + A.T) * dtype(0.5)).astype(dtype)
    for mat in (sym, np.triu(A).astype(dtype), np.tril(A).astype(dtype), np.zeros((4, 4), dtype), np.eye(3, dtype=dtype)):
        assert ours.is_symmetric(mat) == ref.is_symmetric(mat)
        a, b = ours.symmetrize(mat), ref.symmetrize(mat)
        assert a.dtype == b.dtype and np.allclose(a, b, rtol=1e-6 if dtype == np.float32 else 1e-12, atol=0)
        assert np.array_equal(a, a.T)                 # the native layer needs exact symmetry (the reference's C++ check has a tolerance)
    with pytest.raises(RuntimeError):
        ours.symmetrize(A)
    with pytest.raises(RuntimeError):
        ref.symmetrize(A)
    # is_triangular: the reference compares signed entries with the tolerance (common.py:48-52), so a lower-triangular part that is all
    # negative passes as "triangular" there; ours compares magnitudes.  They agree on matrices that really are triangular or symmetric.
    for mat in (np.triu(A), np.tril(A), np.triu(np.abs(A)) + np.tril(np.abs(A), -1)):
        mat = mat.astype(dtype)
        assert ours.is_triangular(mat) == ref.is_triangular(mat)


def test_fix_type(ref_common, ours):
    ref, _ = ref_common
    a = np.arange(12, dtype=np.float64).reshape(3, 4)[:, ::2]         # not contiguous
    for obj in (a, [a, a.T], [[1, 2, 3], [4, 5, 6]]):
        x, y = ours.fix_type(obj, np.float32), ref.fix_type(obj, np.float32)
        if isinstance(y, list):
            assert len(x) == len(y)
            for u, v in zip(x, y):
                assert u.dtype == v.dtype == np.float32 and u.flags['C_CONTIGUOUS'] and np.array_equal(u, v)
        else:
            assert x.dtype == y.dtype == np.float32 and x.flags['C_CONTIGUOUS'] and np.array_equal(x, y)
    with pytest.raises(RuntimeError):
        ours.fix_type(3, np.float32)
    with pytest.raises(RuntimeError):
        ref.fix_type(3, np.float32)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_generate_random_symmetric_W_draws_the_same_matrix(ref_common, ours, dtype):
    ref, _ = ref_common
    for N in (1, 7, 32):
        np.random.seed(1234 + N)
        a = ours.generate_random_symmetric_W(N, -0.5, 0.5, dtype)
        np.random.seed(1234 + N)
        b = ref.generate_random_symmetric_W(N, -0.5, 0.5, dtype)
        assert a.shape == b.shape and np.array_equal(a, a.T)
        assert np.allclose(a, b, rtol=0, atol=1e-7 if dtype == np.float32 else 0)


def test_create_bitset_sequence_and_optimize_tags(ref_common, ours):
    ref, pref = ref_common
    vals = [0, 1, 5, 255, 256, (1 << 40) - 3]
    assert np.array_equal(ours.create_bitset_sequence(vals, 41), ref.create_bitset_sequence(vals, 41))
    assert int(ours.minimize) == int(pref.minimize) == 0 and int(ours.maximize) == int(pref.maximize) == 1
    v = np.array([3.0, -1.0, 2.0])
    assert np.array_equal(ours.minimize.sign(v), pref.minimize.sign(v)) and np.array_equal(ours.maximize.sign(v), pref.maximize.sign(v))
    assert ours.minimize.best(list(v)) == pref.minimize.best(list(v)) and ours.maximize.best(list(v)) == pref.maximize.best(list(v))
    assert ours.minimize.sort(list(v)) == pref.minimize.sort(list(v))
    for name in ('default', 'naive', 'coloring', 'brute_force_search', 'sa_default', 'sa_naive', 'sa_coloring'):
        assert getattr(ours.algorithm, name) == getattr(pref.algorithm, name)
        assert ours.algorithm.is_sqa(name) == pref.algorithm.is_sqa(name)
