"""IIP_c (/root/reference/src/pybind_IIP.cpp:53-57) on the GPU."""
import numpy as np

from ._leaf import arr, call, ptr


def posLLH_IIP_FAA(posECEF, velECEF, fill_na=True, n_iter=5):
    """(lat deg, lon deg, 0), zeros (or NaN with fill_na=False) when there is no impact point;
    `n_iter` is accepted and ignored, like the reference (pybind_IIP.cpp:36 vs iip.cpp:37).
    Accepts one state or (n, 3) batches."""
    pos = arr(posECEF).reshape(-1, 3)
    n = pos.shape[0]
    vel = arr(velECEF, (n, 3))
    out = np.empty((n, 3))
    call("gelato_leaf_iip", n, ptr(pos), ptr(vel), 1 if fill_na else 0, ptr(out))
    return out[0] if np.ndim(posECEF) == 1 else out
