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
