"""Modules with the names and call signatures of the reference's compiled CUDA extension modules
(sqaodpy/sqaod/cuda/src/*.cpp, method tables in sqaodc/pyglue/annealer.inc:884-910, bf_searcher.inc:478-494,
formulas.inc:586-606, cuda_device.cpp:65-72), implemented over the C ABI of libsqaod_b200.so.

The reference's backend-agnostic classes (sqaodpy/sqaod/common/*_base.py) talk to a solver only through
`self._cext.<fn>(self._cobj, ..., self.dtype)`; pointing `_cext` at one of these modules makes them run on the B200 back
end unchanged.  Handles are numpy.uint64 scalars carrying the raw pointer, dtype is the numpy scalar type and always the
last argument, exactly as in the reference glue (annealer.inc:6-10, pyglue.h:252-254)."""
from . import cuda_device, cuda_dg_annealer, cuda_bg_annealer, cuda_dg_bf_searcher, cuda_bg_bf_searcher, cuda_formulas  # noqa: F401
