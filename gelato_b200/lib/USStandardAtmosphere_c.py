"""USStandardAtmosphere_c (/root/reference/src/pybind_USStandardAtmosphere.cpp:28-35) on the GPU.
Scalars or arrays of altitudes; one kernel evaluates all five quantities."""
import numpy as np

from ._leaf import arr, call, ptr


def _all(z):
    zz = arr(z).ravel()
    out = np.empty((zz.size, 5))
    call("gelato_leaf_atmosphere", zz.size, ptr(zz), ptr(out))
    return out


def _col(k):
    def f(z):
        out = _all(z)[:, k]
        return float(out[0]) if np.ndim(z) == 0 else out.copy()

    return f


geopotential_altitude = _col(0)
airtemperature_at = _col(1)
airpressure_at = _col(2)
airdensity_at = _col(3)
speed_of_sound = _col(4)
