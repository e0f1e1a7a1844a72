"""Shared plumbing of the leaf-module drop-ins: numpy in, one batch kernel launch through the
C ABI (gelato_leaf_* in include/gelato_b200.h), fresh numpy array out -- the by-value semantics of
the reference's pybind11/Eigen modules (/root/reference/src/wrapper_air.hpp:40-43)."""
import ctypes

import numpy as np

from .. import engine as _engine

_pd = ctypes.POINTER(ctypes.c_double)
device = 0  # CUDA device the leaf kernels run on (set gelato_b200.lib._leaf.device to change it)


def arr(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def ptr(a):
    return a.ctypes.data_as(_pd) if a is not None else ctypes.cast(None, _pd)


def call(name, *args):
    L = _engine.load_library()
    rc = getattr(L, name)(device, *args)
    if rc != 0:
        raise _engine.GelatoError("%s failed (%d): %s" % (name, rc, L.gelato_last_error().decode()))
