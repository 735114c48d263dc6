"""CPU-side checks of the product library: it loads, exports every symbol include/sqaod_b200.h declares, keeps the
reference's host-side semantics (preferences, state machine errors) and fails loudly without a device.  No compute."""
import ctypes as C
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from sqaod_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(lib):
    syms = lib.declared_symbols()
    assert len(syms) > 90
    missing = [s for s in syms if not hasattr(lib.lib, s)]
    assert missing == []


def test_version_symbol_of_the_reference(lib):
    # sqaodpy/sqaod/common/envcheck.py:77-99 requires ver >= 10002
    ver, cuda = C.c_int(0), C.c_int(0)
    lib.lib.sqaodc_cuda_version(C.byref(ver), C.byref(cuda))
    assert ver.value >= 10002 and cuda.value >= 12000
    assert lib.lib.sqb_version() >= 100


@pytest.mark.parametrize('dt', [0, 1])
def test_host_side_preferences_and_errors(lib, dt):
    L = lib.lib
    h = C.c_void_p()
    assert L.sqb_dg_annealer_new(C.byref(h), dt) == 0
    buf = C.create_string_buffer(256)
    assert L.sqb_dg_annealer_get_preferences(h, buf, 256, dt) == 0
    prefs = dict(kv.split('=') for kv in buf.value.decode().split(';'))
    assert prefs['algorithm'] == 'coloring' and prefs['device'] == 'cuda'
    assert prefs['precision'] == ('float' if dt == 0 else 'double')
    # algorithm fallbacks of the CUDA solver (test_dense_graph_annealer.py:454-482)
    for asked, got in (('naive', 'coloring'), ('sa_default', 'sa_naive'), ('sa_coloring', 'sa_naive'), ('default', 'coloring')):
        assert L.sqb_dg_annealer_set_preference(h, b'algorithm', asked.encode(), 0, dt) == 0
        L.sqb_dg_annealer_get_preferences(h, buf, 256, dt)
        assert ('algorithm=' + got) in buf.value.decode()
    assert L.sqb_dg_annealer_set_preference(h, b'n_trotters', None, 12, dt) == 0
    L.sqb_dg_annealer_get_preferences(h, buf, 256, dt)
    assert 'n_trotters=12' in buf.value.decode()
    assert L.sqb_dg_annealer_set_preference(h, b'n_trotters', None, 0, dt) != 0           # must be positive
    assert b'positive' in L.sqb_last_error()
    assert L.sqb_dg_annealer_set_preference(h, b'no_such_pref', None, 1, dt) != 0
    # no device assigned: problem upload must fail, not fall back
    W = np.eye(4, dtype=np.float32 if dt == 0 else np.float64)
    assert L.sqb_dg_annealer_set_qubo(h, W.ctypes.data_as(C.c_void_p), 4, 4, 0, dt) != 0
    assert b'Device not set' in L.sqb_last_error()
    assert L.sqb_dg_annealer_prepare(h, dt) != 0 and b'Problem is not set' in L.sqb_last_error()
    assert L.sqb_dg_annealer_anneal_one_step(h, C.c_double(1.), C.c_double(1.), dt) != 0
    # sweep-mode selector (host-side state only; takes effect at prepare()): -1 automatic, 0 classic, 1 field
    for mode in (-1, 0, 1):
        assert L.sqb_dg_annealer_set_sweep_mode(h, mode, 0, dt) == 0
    assert L.sqb_dg_annealer_set_sweep_mode(h, 2, 0, dt) != 0 and b'sweep mode' in L.sqb_last_error()
    mode = C.c_int(7)
    assert L.sqb_dg_annealer_get_sweep_mode(h, C.byref(mode), dt) == 0 and mode.value == 0      # nothing prepared yet
    assert L.sqb_dg_annealer_delete(h, dt) == 0
    # brute-force searcher: tile sizes are rounded up to multiples of 256 (Solver.cpp:205)
    s = C.c_void_p()
    assert L.sqb_dg_bf_searcher_new(C.byref(s), dt) == 0
    assert L.sqb_dg_bf_searcher_set_preference(s, b'tile_size', None, 1000, dt) == 0
    L.sqb_dg_bf_searcher_get_preferences(s, buf, 256, dt)
    assert 'tile_size=1024' in buf.value.decode() and 'algorithm=brute_force_search' in buf.value.decode()
    assert L.sqb_dg_bf_searcher_delete(s, dt) == 0
    assert L.sqb_dg_annealer_new(C.byref(h), 7) != 0 and b'dtype' in L.sqb_last_error()


def test_fails_loudly_without_a_gpu(lib):
    import sqaod_b200 as sq
    if sq.is_cuda_available():
        pytest.skip('a CUDA device is present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        sq.dense_graph_annealer(np.eye(4), sq.minimize, np.float32)


def test_python_helpers_match_the_reference():
    import sqaod_b200 as sq
    W = np.triu(np.arange(16, dtype=np.float64).reshape(4, 4))
    S = sq.symmetrize(W)
    assert np.array_equal(S, S.T) and np.array_equal(np.diag(S), np.diag(W))
    with pytest.raises(RuntimeError):
        sq.symmetrize(np.arange(16, dtype=np.float64).reshape(4, 4))
    x = sq.create_bitset_sequence([5, 2], 4)                 # MSB first (Common.cpp:78-93)
    assert x.tolist() == [[0, 1, 0, 1], [0, 0, 1, 0]]
    assert int(sq.minimize) == 0 and int(sq.maximize) == 1
    assert sq.algorithm.is_sqa('coloring') and not sq.algorithm.is_sqa('sa_naive')


def test_preference_and_algorithm_names_equal_the_reference_library():
    """The name tables of the solver API (sqaodc/common/Preference.cpp:8-100): `sqaod::algorithmToString / FromString`,
    `preferenceNameToString / FromString` and `isSQAAlgorithm` of libsqaod_b200.so against the same functions of the reference's own CPU
    library, compiled from its sources (oracle/_ref/libsqaodc_refcpu.so; `make -C oracle refcpu`), for every enum value and every name
    -- plus unknown inputs.  Plain functions over ints and C strings: no device, no objects cross the two libraries."""
    import ctypes as C
    ref_so = os.path.join(ROOT, 'oracle', '_ref', 'libsqaodc_refcpu.so')
    if not os.path.exists(ref_so):
        pytest.skip('reference CPU library not built (run `make -C oracle refcpu` where /root/reference exists)')
    ours, ref = C.CDLL(os.path.join(ROOT, 'sqaod_b200', 'lib', 'libsqaod_b200.so')), C.CDLL(ref_so)
    sym = {'a2s': '_ZN5sqaod17algorithmToStringENS_9AlgorithmE', 's2a': '_ZN5sqaod19algorithmFromStringEPKc',
           'p2s': '_ZN5sqaod22preferenceNameToStringENS_14PreferenceNameE', 's2p': '_ZN5sqaod24preferenceNameFromStringEPKc',
           'sqa': '_ZN5sqaod14isSQAAlgorithmENS_9AlgorithmE'}

    def fn(lib, key, restype, argtype):
        f = getattr(lib, sym[key])
        f.restype, f.argtypes = restype, [argtype]
        return f
    for lib_pair in [(ours, ref)]:
        a2s = [fn(l, 'a2s', C.c_char_p, C.c_int) for l in lib_pair]
        s2a = [fn(l, 's2a', C.c_int, C.c_char_p) for l in lib_pair]
        p2s = [fn(l, 'p2s', C.c_char_p, C.c_int) for l in lib_pair]
        s2p = [fn(l, 's2p', C.c_int, C.c_char_p) for l in lib_pair]
        sqa = [fn(l, 'sqa', C.c_bool, C.c_int) for l in lib_pair]
    names = set()
    for v in range(0, 9):                      # enum Algorithm, Preference.h:8-17
        a, b = a2s[0](v), a2s[1](v)
        assert a == b, (v, a, b)
        assert sqa[0](v) == sqa[1](v), v
        names.add(a)
    for v in list(range(0, 8)) + [100]:        # enum PreferenceName, Preference.h:26-36
        a, b = p2s[0](v), p2s[1](v)
        assert a == b, (v, a, b)
        names.add(a)
    for s in sorted(n for n in names if n) + [b'', b'nonsense', b'Coloring', b'tile_size', b'n_trotters', b'sa_naive', b'default']:
        assert s2a[0](s) == s2a[1](s), s
        assert s2p[0](s) == s2p[1](s), s


def test_product_library_carries_blackwell_sass():
    """What `cuobjdump -sass` of libsqaod_b200.so shows (B200_PROFILING.md's mnemonic table): the tcgen05 GEMM (UTCHMMA, LDTM), TMA tensor
    and bulk copies (UTMALDG, UBLKCP), mbarrier traffic (SYNCS), cp.async gathers (LDGSTS), packed fp32 adds (FADD2) -- and no legacy
    mma.sync path (every HMMA in the listing is the tail of a UTCHMMA).  Guards the build flags: a library compiled for another
    architecture would still load here."""
    import shutil
    import subprocess
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([exe, '-sass', os.path.join(ROOT, 'sqaod_b200', 'lib', 'libsqaod_b200.so')], capture_output=True, text=True,
                          timeout=600).stdout
    count = lambda k: sum(1 for l in sass.splitlines() if k in l)
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'FADD2'):
        assert count(mnemonic) > 0, mnemonic
    assert count('HMMA') == count('UTCHMMA')
    archs = [l.split('=')[1].strip() for l in sass.splitlines() if l.startswith('arch =')]
    assert archs.count('sm_100a') >= 7 and set(archs) <= {'sm_100a', 'sm_52'}     # sm_52: nvcc's empty device-link stub, no code in it


def test_product_library_links_no_vendor_math_or_collective_library():
    """north_star: "no cuBLAS, Triton or CPU fallback inside the path" -- the library's dynamic dependencies are the C/C++ runtime only
    (the CUDA runtime is linked statically); nothing of cuBLAS, cuRAND, cuDNN, cuSPARSE, NCCL or the oracle."""
    import subprocess
    out = subprocess.run(['ldd', os.path.join(ROOT, 'sqaod_b200', 'lib', 'libsqaod_b200.so')], capture_output=True, text=True).stdout.lower()
    for name in ('cublas', 'curand', 'cudnn', 'cusparse', 'cusolver', 'nccl', 'liboracle', 'sqaodc', 'torch'):
        assert name not in out, name


@pytest.mark.parametrize('dt', [0, 1])
def test_searcher_preferences_behave_like_the_reference_solver(lib, dt):
    """Backend-independent host logic of the solver base classes (sqaodc/common/Solver.cpp:190-250), ours through the C ABI against the
    reference's own library through its own Python package (oracle/_ref): how requested tile sizes are adjusted, what get_preferences
    reports for a fresh searcher, and that non-positive sizes are refused."""
    import subprocess
    import sys
    import json
    suite = os.path.join(ROOT, 'oracle', '_ref', 'refsuite')
    if not os.path.exists(os.path.join(suite, 'glue_cpu', 'cpu_dg_bf_searcher.so')):
        pytest.skip('reference CPU build absent (run `make -C oracle refcpu` where /root/reference exists)')
    sizes = [1, 255, 256, 257, 1000, 4096, 70000]
    code = r'''
import sys, json, warnings
warnings.simplefilter('ignore')
import numpy as np
sys.path.insert(0, sys.argv[1] + '/tests')
import refsuite_runner
sq = refsuite_runner.assemble('cpu')
dtype = np.float32 if sys.argv[2] == '0' else np.float64
out = {'dg': [], 'bg0': [], 'bg1': []}
s = sq.cpu.dense_graph_bf_searcher(dtype=dtype)
b = sq.cpu.bipartite_graph_bf_searcher(dtype=dtype)
out['fresh_dg'] = {k: v for k, v in s.get_preferences().items() if k in ('algorithm', 'precision')}
out['fresh_bg'] = {k: v for k, v in b.get_preferences().items() if k in ('algorithm', 'precision')}
for v in json.loads(sys.argv[3]):
    s.set_preferences(tile_size=v); out['dg'].append(s.get_preferences()['tile_size'])
    b.set_preferences(tile_size_0=v); out['bg0'].append(b.get_preferences()['tile_size_0'])
    b.set_preferences(tile_size_1=v); out['bg1'].append(b.get_preferences()['tile_size_1'])
def refused(f):
    try:
        f()
    except Exception:
        return True
    return False
out['neg'] = [refused(lambda: s.set_preferences(tile_size=0)), refused(lambda: b.set_preferences(tile_size_0=-3))]
print('REF ' + json.dumps(out))
'''
    run = subprocess.run([sys.executable, '-c', code, ROOT, str(dt), json.dumps(sizes)], capture_output=True, text=True, timeout=300)
    ref = json.loads([l for l in run.stdout.splitlines() if l.startswith('REF ')][0][4:])
    L = lib.lib
    buf = C.create_string_buffer(512)

    def prefs(getter, h):
        assert getter(h, buf, 512, dt) == 0
        return dict(kv.split('=') for kv in buf.value.decode().split(';'))
    s, b = C.c_void_p(), C.c_void_p()
    assert L.sqb_dg_bf_searcher_new(C.byref(s), dt) == 0 and L.sqb_bg_bf_searcher_new(C.byref(b), dt) == 0
    p = prefs(L.sqb_dg_bf_searcher_get_preferences, s)
    assert {k: p[k] for k in ('algorithm', 'precision')} == ref['fresh_dg']
    p = prefs(L.sqb_bg_bf_searcher_get_preferences, b)
    assert {k: p[k] for k in ('algorithm', 'precision')} == ref['fresh_bg']
    for i, v in enumerate(sizes):
        assert L.sqb_dg_bf_searcher_set_preference(s, b'tile_size', None, v, dt) == 0
        assert int(prefs(L.sqb_dg_bf_searcher_get_preferences, s)['tile_size']) == ref['dg'][i], v
        assert L.sqb_bg_bf_searcher_set_preference(b, b'tile_size_0', None, v, dt) == 0
        assert int(prefs(L.sqb_bg_bf_searcher_get_preferences, b)['tile_size_0']) == ref['bg0'][i], v
        assert L.sqb_bg_bf_searcher_set_preference(b, b'tile_size_1', None, v, dt) == 0
        assert int(prefs(L.sqb_bg_bf_searcher_get_preferences, b)['tile_size_1']) == ref['bg1'][i], v
    assert ref['neg'] == [True, True]
    assert L.sqb_dg_bf_searcher_set_preference(s, b'tile_size', None, 0, dt) != 0
    assert L.sqb_bg_bf_searcher_set_preference(b, b'tile_size_0', None, -3, dt) != 0
    assert L.sqb_dg_bf_searcher_delete(s, dt) == 0 and L.sqb_bg_bf_searcher_delete(b, dt) == 0
