"""The reference's CPython glue (sqaodpy/sqaod/cuda/src/cuda_*.cpp + sqaodc/pyglue/*.inc), compiled UNMODIFIED against
include/sqaodc/sqaodc.h and linked to libsqaod_b200.so (`make -C oracle glue`): every module loads (all its C++ symbols resolve
against the product library) and exports the reference's method table.  No device is touched."""
import importlib.util
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, 'oracle', '_ref', 'refsuite', 'glue')

ANNEALER = ('new delete assign_device seed set_qubo set_hamiltonian get_problem_size set_preferences get_preferences get_E get_x '
            'get_hamiltonian get_q set_q set_qset randomize_spin calculate_E prepare make_solution get_system_E anneal_one_step').split()
SEARCHER = ('new delete assign_device set_qubo get_problem_size set_preferences get_preferences get_x get_E prepare calculate_E '
            'make_solution search_range search').split()
FORMULAS = ('dg_formulas_new dg_formulas_delete dg_formulas_assign_device dense_graph_calculate_E dense_graph_batch_calculate_E '
            'dense_graph_calculate_hamiltonian dense_graph_calculate_E_from_spin dense_graph_batch_calculate_E_from_spin '
            'bg_formulas_new bg_formulas_delete bg_formulas_assign_device bipartite_graph_calculate_E bipartite_graph_batch_calculate_E '
            'bipartite_graph_batch_calculate_E_2d bipartite_graph_calculate_hamiltonian bipartite_graph_calculate_E_from_spin '
            'bipartite_graph_batch_calculate_E_from_spin').split()
TABLES = {'cuda_device': 'new delete initialize finalize'.split(), 'cuda_dg_annealer': ANNEALER, 'cuda_bg_annealer': ANNEALER,
          'cuda_dg_bf_searcher': SEARCHER, 'cuda_bg_bf_searcher': SEARCHER, 'cuda_formulas': FORMULAS}


@pytest.mark.parametrize('name', sorted(TABLES))
def test_reference_glue_module_loads_and_has_the_reference_method_table(name):
    if os.path.isdir('/root/reference/sqaodc'):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'glue'], stdout=subprocess.DEVNULL)
    so = os.path.join(GLUE, name + '.so')
    if not os.path.exists(so):
        pytest.skip('reference glue not built (reference tree absent)')
    spec = importlib.util.spec_from_file_location('refglue_' + name, so)
    # the init function is PyInit_<last component of the name the module was built for>
    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    missing = [f for f in TABLES[name] if not hasattr(mod, f)]
    assert not missing, missing


def test_reference_cpu_library_is_unaffected_by_the_product_library_in_the_same_process():
    """bench.py's cpu_baseline leg runs the compiled reference CPU annealer in a process that already holds libsqaod_b200.so (and the
    reference's CUDA glue would sit next to its CPU glue in a full `import sqaod`).  Both libraries define the `sqaod::` host classes
    with different layouts, so neither may bind to the other's symbols: with the product library and all six CUDA glue modules loaded
    first, the reference CPU annealer must still walk the chain recorded from it in a process of its own
    (tests/golden/refcpu_chains.npz, one worker)."""
    import sys
    code = r'''
import os, sys, importlib.util, warnings
os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]}); os.environ['OMP_NUM_THREADS'] = '1'
warnings.simplefilter('ignore')
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from sqaod_b200 import _lib                                   # the product library, RTLD_LOCAL via ctypes
glue = os.path.join(ROOT, 'oracle', '_ref', 'refsuite', 'glue')
for name in ('cuda_device', 'cuda_dg_annealer', 'cuda_bg_annealer', 'cuda_dg_bf_searcher', 'cuda_bg_bf_searcher', 'cuda_formulas'):
    spec = importlib.util.spec_from_file_location(name, os.path.join(glue, name + '.so'))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
import refsuite_runner
sq = refsuite_runner.assemble('cpu')
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'refcpu_chains.npz'))
for key in ('dense0', 'dense7'):
    N, m, seed, width = (int(v) for v in g[key + '/meta'])
    dtype = np.float32 if width == 4 else np.float64
    ann = sq.cpu.dense_graph_annealer(dtype=dtype, algorithm=str(g[key + '/algo']))
    ann.set_hamiltonian(g[key + '/h'], g[key + '/J'], dtype(g[key + '/c']))
    ann.set_preferences(n_trotters=m)
    ann.seed(seed); ann.prepare(); ann.randomize_spin()
    for k, G in enumerate(g[key + '/G']):
        ann.anneal_one_step(float(G), 1. / 0.02)
        assert np.array_equal(np.asarray(ann.get_q(), np.int8), g[key + '/q'][k + 1]), (key, k)
print('COEXIST_OK')
'''
    if not (os.path.exists(os.path.join(GLUE, 'cuda_dg_annealer.so')) and
            os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'refsuite', 'glue_cpu', 'cpu_dg_annealer.so'))):
        pytest.skip('reference glue / reference CPU build absent (reference tree absent at build time)')
    out = subprocess.run([sys.executable, '-c', code, ROOT], capture_output=True, text=True, timeout=300)
    assert 'COEXIST_OK' in out.stdout, out.stdout[-1000:] + out.stderr[-2000:]
