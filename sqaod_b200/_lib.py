"""ctypes binding of libsqaod_b200.so (C ABI declared in include/sqaod_b200.h).

The product path: there is no Python/NumPy fallback.  If the shared library has not been built the import fails
loudly; if no CUDA device is present every solver call raises RuntimeError from the native layer.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libsqaod_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'sqaod_b200.h')

F32, F64 = 0, 1


def build():
    """Compile the native library for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.check_call(['make', '-C', os.path.join(_HERE, 'csrc'), '-j8'], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError('sqaod_b200: native library %s is missing; run `python -c "import __graft_entry__ as g; g.build()"` '
                          'or `make -C sqaod_b200/csrc`.  There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.sqb_last_error.restype = C.c_char_p
    return lib


lib = _load()


def declared_symbols():
    """Names of every function include/sqaod_b200.h declares (used by the symbol-export test)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sqb_[a-z0-9_A-Z]+|sqaodc_cuda_version)\s*\(', src)))


def check(rc):
    if rc != 0:
        raise RuntimeError(lib.sqb_last_error().decode('utf-8', 'replace').strip())


def dtype_code(dtype):
    import numpy as np
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise RuntimeError('dtype must be numpy.float32 or numpy.float64.')


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
