"""Device handle (mirrors sqaodpy/sqaod/cuda/device.py:6-37): one global device is created on first use."""
import ctypes as C
from . import _lib


class Device(object):
    def __init__(self, devno=0):
        self._cobj = C.c_void_p()
        _lib.check(_lib.lib.sqb_device_new(C.byref(self._cobj)))
        self._initialized = False
        if devno is not None:
            self.initialize(devno)

    def initialize(self, devno=0):
        _lib.check(_lib.lib.sqb_device_initialize(self._cobj, int(devno)))
        self._initialized = True
        self.devno = devno

    def finalize(self):
        if self._initialized:
            _lib.check(_lib.lib.sqb_device_finalize(self._cobj))
            self._initialized = False

    def synchronize(self):
        _lib.check(_lib.lib.sqb_device_synchronize(self._cobj))

    def set_stream(self, cuda_stream):
        """Run on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream); None restores the own stream."""
        _lib.check(_lib.lib.sqb_device_set_stream(self._cobj, C.c_void_p(cuda_stream or 0)))

    def launch_count(self, reset=False):
        n = C.c_ulonglong(0)
        _lib.check(_lib.lib.sqb_device_launch_count(self._cobj, C.byref(n), 1 if reset else 0))
        return n.value

    def num_sms(self):
        n = C.c_int(0)
        _lib.check(_lib.lib.sqb_device_num_sms(self._cobj, C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if self._cobj:
                _lib.lib.sqb_device_finalize(self._cobj)
                _lib.lib.sqb_device_delete(self._cobj)
                self._cobj = None
        except Exception:
            pass


_active = None


def device_count():
    n = C.c_int(0)
    _lib.check(_lib.lib.sqb_device_count(C.byref(n)))
    return n.value


def active_device():
    global _active
    if _active is None:
        import os
        _active = Device(int(os.environ.get('LOCAL_RANK', '0')) if device_count() > 1 else 0)
    return _active


def set_active_device(dev):
    global _active
    _active = dev
