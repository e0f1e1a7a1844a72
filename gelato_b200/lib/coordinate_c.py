"""The members of coordinate_c (/root/reference/src/pybind_coordinate.cpp:28-78) that are evaluated
over whole trajectories: eci2geodetic and gravity, batched on the GPU (a single point is a batch
of one).  The scalar set-up helpers of that module (quat_from_euler, geodetic2ecef, ...) run once per
problem on the host: gelato_b200/hostmath.py."""
import numpy as np

from ._leaf import arr, call, ptr


def eci2geodetic(pos_eci, t):
    """wrapper_coordinate.hpp:193-199 -> (lat deg, lon deg, alt m); (n, 3) for n positions."""
    pos = arr(pos_eci).reshape(-1, 3)
    n = pos.shape[0]
    tt = np.broadcast_to(arr(t).ravel(), (n,)).copy()
    out = np.empty((n, 3))
    call("gelato_leaf_eci2geodetic", n, ptr(pos), ptr(tt), ptr(out))
    return out[0] if np.ndim(pos_eci) == 1 else out


def gravity(pos_eci):
    """gravity.cpp:11-57 (J2)."""
    pos = arr(pos_eci).reshape(-1, 3)
    n = pos.shape[0]
    out = np.empty((n, 3))
    call("gelato_leaf_gravity", n, ptr(pos), ptr(out))
    return out[0] if np.ndim(pos_eci) == 1 else out
