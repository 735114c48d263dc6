#!/usr/bin/env python
"""Run the REFERENCE's own Python test-suite (sqaodpy/tests, staged unmodified under oracle/_ref/refsuite by
`make -C oracle glue`) with `sqaod.cuda` bound to libsqaod_b200.so.

    python tests/refsuite_runner.py glue [pytest args]   the reference's CPython glue (sqaodc/pyglue/*.inc +
                                                         sqaodpy/sqaod/cuda/src/cuda_*.cpp), compiled unmodified against
                                                         include/sqaodc/sqaodc.h and linked to libsqaod_b200.so
    python tests/refsuite_runner.py cext [pytest args]   sqaod_b200.cext (the ctypes restatement of the same method tables)
    python tests/refsuite_runner.py full [pytest args]   the whole package as a user imports it: sqaod.py, sqaod.cpu = the reference's own CPU
                                                         back end (below) and sqaod.cuda = the reference glue over libsqaod_b200.so, all of the
                                                         reference's test classes in one process
    python tests/refsuite_runner.py cpu  [pytest args]   no GPU: the reference's CPU back end itself (`make -C oracle refcpu`: sqaodc/common +
                                                         sqaodc/cpu + the cpu_*.cpp glue, compiled unmodified against oracle/eigen_standin) under
                                                         the reference's own CPU and pure-Python test classes -- checks that build, which the
                                                         oracle is pinned against (tests/test_oracle_vs_reference_cpu.py)

The reference's top-level sqaod/__init__.py imports the CPU extension (needs Eigen: not buildable here), so the package
object is assembled here instead: sqaod.common, sqaod.py and sqaod.cuda (device.py, the four solver wrappers, formulas.py)
are the reference's files; only the six C-extension modules underneath sqaod.cuda are ours.  In the glue / cext runs sqaod.cpu is a stub
and the CPU test classes are deselected (-k cuda); in the cpu run sqaod.cpu is the reference's package over oracle/_ref/refsuite/glue_cpu
and sqaod.cuda is absent (is_cuda_available() is False, so the reference's tests define no CUDA classes)."""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, 'oracle', '_ref', 'refsuite')
CPU_CEXT = ['cpu_dg_annealer', 'cpu_bg_annealer', 'cpu_dg_bf_searcher', 'cpu_bg_bf_searcher', 'cpu_formulas']
CEXT = ['cuda_device', 'cuda_dg_annealer', 'cuda_bg_annealer', 'cuda_dg_bf_searcher', 'cuda_bg_bf_searcher', 'cuda_formulas']


_ASSEMBLED = {}


def assemble(binding):
    if binding in _ASSEMBLED:       # extension modules load once per process: a second call hands back the same package
        sys.modules['sqaod'] = _ASSEMBLED[binding]
        return _ASSEMBLED[binding]
    pkg = _assemble(binding)
    _ASSEMBLED[binding] = pkg
    return pkg


def _assemble(binding):
    sys.path.insert(0, SUITE)
    if binding == 'cext':
        sys.path.insert(1, ROOT)
    pkg = types.ModuleType('sqaod')
    pkg.__path__ = [os.path.join(SUITE, 'sqaod')]
    sys.modules['sqaod'] = pkg
    c = importlib.import_module('sqaod.common')
    pref = importlib.import_module('sqaod.common.preference')
    pkg.algorithm, pkg.minimize, pkg.maximize = pref.algorithm, pref.minimize, pref.maximize
    for obj in (pref.minimize, pref.maximize):
        # the reference's OptimizeMethod objects only define __int__; since Python 3.10 the glue's "i" format (annealer.inc:111)
        # needs __index__ -- an interpreter-version shim, nothing of the library under test
        if not hasattr(type(obj), '__index__'):
            type(obj).__index__ = lambda self: int(self)
    for k in dir(c):
        if not k.startswith('_'):
            setattr(pkg, k, getattr(c, k))
    pkg.common = c
    pkg.is_cuda_available = lambda: binding != 'cpu'
    pkg.py = importlib.import_module('sqaod.py')
    if binding == 'cpu':
        load_reference_cpu(pkg)
        return pkg
    for name in CEXT:
        full = 'sqaod.cuda.' + name
        if binding in ('glue', 'full'):
            spec = importlib.util.spec_from_file_location(full, os.path.join(SUITE, 'glue', name + '.so'))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        else:
            mod = importlib.import_module('sqaod_b200.cext.' + name)
        sys.modules[full] = mod
    pkg.cuda = importlib.import_module('sqaod.cuda')

    if binding == 'full':
        load_reference_cpu(pkg)
        return pkg

    class _NoCPU(types.ModuleType):
        def __getattr__(self, name):
            raise AttributeError('sqaod.cpu is not part of this run (%s)' % name)
    pkg.cpu = _NoCPU('sqaod.cpu')
    return pkg


def load_reference_cpu(pkg):
    """sqaod.cpu = the reference's own package (sqaod/cpu/*.py) over its own glue and CPU library built by `make -C oracle refcpu`."""
    for name in CPU_CEXT:
        full = 'sqaod.cpu.' + name
        spec = importlib.util.spec_from_file_location(full, os.path.join(SUITE, 'glue_cpu', name + '.so'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[full] = mod
    pkg.cpu = importlib.import_module('sqaod.cpu')
    return pkg.cpu


def main():
    binding = sys.argv[1] if len(sys.argv) > 1 else 'glue'
    if not os.path.isdir(os.path.join(SUITE, 'tests')):
        print('REFSUITE_ABSENT')
        return 0
    assemble(binding)
    import pytest
    select = {'cpu': 'not cuda and not version', 'full': 'not version'}.get(binding, 'cuda and not version')
    args = [os.path.join(SUITE, 'tests'), '-q', '-p', 'no:cacheprovider', '-k', select, '--rootdir', SUITE,
            '-o', 'python_files=test_*.py', '-rfE', '--tb=short'] + sys.argv[2:]
    rc = pytest.main(args)
    print('REFSUITE_RC %s %d' % (binding, int(rc)))
    return int(rc)


if __name__ == '__main__':
    sys.exit(main())
