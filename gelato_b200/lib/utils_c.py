"""utils_c (/root/reference/src/pybind_utils.cpp:28-48): the aero-constraint leaves on the GPU.
Dimensional inputs, t in seconds, as in wrapper_utils.hpp:89-206."""
import numpy as np

from ._leaf import arr, call, ptr


def _aero(kind, pos, vel, quat, t, wind):
    pos = arr(pos).reshape(-1, 3)
    n = pos.shape[0]
    vel = arr(vel, (n, 3))
    q = arr(quat, (n, 4)) if quat is not None else None
    tt = arr(t).ravel()
    w = arr(wind)
    out = np.empty(n)
    call("gelato_leaf_aero", kind, n, ptr(pos), ptr(vel), ptr(q), ptr(tt), ptr(w), w.shape[0], ptr(out))
    return out


def angle_of_attack_all_array_rad(pos_eci, vel_eci, quat, t, wind):
    return _aero(0, pos_eci, vel_eci, quat, t, wind)


def dynamic_pressure_array_pa(pos_eci, vel_eci, t, wind):
    return _aero(1, pos_eci, vel_eci, None, t, wind)


def q_alpha_array_pa_rad(pos_eci, vel_eci, quat, t, wind):
    return _aero(2, pos_eci, vel_eci, quat, t, wind)


def angle_of_attack_all_rad(pos_eci, vel_eci, quat, t, wind):
    return float(_aero(0, pos_eci, vel_eci, quat, [t], wind)[0])


def dynamic_pressure_pa(pos_eci, vel_eci, t, wind):
    return float(_aero(1, pos_eci, vel_eci, None, [t], wind)[0])


def q_alpha_pa_rad(pos_eci, vel_eci, quat, t, wind):
    return float(_aero(2, pos_eci, vel_eci, quat, [t], wind)[0])
