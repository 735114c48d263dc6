"""cuda_device: new / delete / initialize / finalize (sqaodpy/sqaod/cuda/src/cuda_device.cpp:5-72)"""
import ctypes as C
import numpy as np
from ._glue import L, check, h


def new():
    p = C.c_void_p()
    check(L.sqb_device_new(C.byref(p)))
    return np.uint64(p.value)


def delete(dev):
    check(L.sqb_device_delete(h(dev)))


def initialize(dev, devno):
    check(L.sqb_device_initialize(h(dev), int(devno)))


def finalize(dev):
    check(L.sqb_device_finalize(h(dev)))
