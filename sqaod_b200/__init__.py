"""sqaod_b200 -- B200-native (sm_100a) back end for sqaod's CUDA solvers.

Drop-in for the reference's `sqaod.cuda` package: same factories and solver methods
(dense_graph_annealer, bipartite_graph_annealer, dense_graph_bf_searcher, bipartite_graph_bf_searcher, formulas),
implemented by libsqaod_b200.so through the C ABI in include/sqaod_b200.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the native library is missing)
from .common import (algorithm, minimize, maximize, symmetrize, fix_type, generate_random_symmetric_W,  # noqa: F401
                     create_bitset_sequence)
from .device import Device, active_device, set_active_device, device_count  # noqa: F401
from .solvers import (DenseGraphAnnealer, BipartiteGraphAnnealer, DenseGraphBFSearcher, BipartiteGraphBFSearcher,  # noqa: F401
                      dense_graph_annealer, bipartite_graph_annealer, dense_graph_bf_searcher, bipartite_graph_bf_searcher)
from . import formulas  # noqa: F401

__version__ = '0.1.0'


def is_cuda_available():
    try:
        return device_count() > 0
    except Exception:
        return False
