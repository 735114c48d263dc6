"""shared helpers of the cext-compatible modules"""
import ctypes as C
import numpy as np
from .. import _lib

L = _lib.lib
check = _lib.check
ptr = _lib.ptr


def dt(dtype):
    if dtype is np.float32 or dtype == np.float32:
        return _lib.F32
    if dtype is np.float64 or dtype == np.float64:
        return _lib.F64
    raise RuntimeError('dtype must be numpy.float32 or numpy.float64.')     # ASSERT_DTYPE, pyglue.h:252-254


def h(obj):
    return C.c_void_p(int(obj))


def new_handle(fn, dtype):
    p = C.c_void_p()
    check(fn(C.byref(p), dt(dtype)))
    return np.uint64(p.value)


def stride(a):
    return a.strides[0] // a.itemsize if a.ndim == 2 else a.shape[0]


def set_preferences(prefix, obj, prefs, dtype):
    fn = getattr(L, 'sqb_%s_set_preference' % prefix)
    for k, v in prefs.items():
        if isinstance(v, str):
            check(fn(h(obj), k.encode(), v.encode(), C.c_long(0), dt(dtype)))
        else:
            check(fn(h(obj), k.encode(), None, C.c_long(int(v)), dt(dtype)))


def get_preferences(prefix, obj, dtype):
    buf = C.create_string_buffer(512)
    check(getattr(L, 'sqb_%s_get_preferences' % prefix)(h(obj), buf, 512, dt(dtype)))
    out = {}
    for item in buf.value.decode().split(';'):
        if item:
            k, v = item.split('=', 1)
            out[k] = int(v) if k in ('n_trotters', 'tile_size', 'tile_size_0', 'tile_size_1', 'experiment') else v
    return out


def simple(prefix, name):
    fn = getattr(L, 'sqb_%s_%s' % (prefix, name))

    def call(obj, dtype):
        check(fn(h(obj), dt(dtype)))
    call.__name__ = name
    return call


def vec(a):
    """a vector argument as the reference glue accepts it (NpVectorType, pyglue/pyglue.h:112-131): 1-D, or 2-D with one row or
    one column"""
    a = np.asarray(a)
    if a.ndim >= 3 or (a.ndim == 2 and a.shape[0] != 1 and a.shape[1] != 1):
        raise RuntimeError('ndarray is not 1-diemsional.')
    return np.ascontiguousarray(a.reshape(-1))
