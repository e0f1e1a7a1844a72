"""coordinate_c (/root/reference/src/pybind_coordinate.cpp:28-78) on the GPU: every function takes the reference's
arguments for one point, or arrays with a leading batch dimension (n points by one kernel launch, one thread each;
a single point is a batch of one).  eci2geodetic and gravity have kernels of their own; the rest go through
gelato_leaf_coordinate (csrc/coord_leaves.h).  Not offered: dcm_from_quat, quat_from_dcm, euler_from_dcm,
dcm_from_thrustvector, laplace_vector (no call site in the reference's live code)."""
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


# ---- the rest of the module, by function code (include/gelato_b200.h: GC_*) -------------------------------------
_CODES = {name: k for k, name in enumerate(
    "quatmult conj normalize3 normalize4 quatrot ecef2geodetic geodetic2ecef ecef2eci eci2ecef vel_ecef2eci vel_eci2ecef "
    "quat_eci2ecef quat_ecef2eci quat_ecef2nedg quat_nedg2ecef quat_eci2nedg quat_nedg2eci quat_from_euler euler_from_quat "
    "quat_nedg2body orbital_elements distance_vincenty angular_momentum_vec angular_momentum inclination_rad "
    "inclination_cosine orbit_energy angular_momentum_from_altitude orbit_energy_from_altitude laplace_vector haversine "
    "dcm_from_quat quat_from_dcm euler_from_dcm dcm_from_thrustvector".split())}
_N_OUT = {"orbital_elements": 6, "dcm_from_quat": 9, "dcm_from_thrustvector": 9, "quat_from_dcm": 4}
for _n in ("quatmult conj normalize4 quat_eci2ecef quat_ecef2eci quat_ecef2nedg quat_nedg2ecef quat_eci2nedg quat_nedg2eci "
           "quat_from_euler quat_nedg2body").split():
    _N_OUT[_n] = 4
for _n in ("distance_vincenty angular_momentum inclination_rad inclination_cosine orbit_energy "
           "angular_momentum_from_altitude orbit_energy_from_altitude haversine").split():
    _N_OUT[_n] = 1


def _leaf(name, a=None, wa=0, b=None, wb=0, t=None, fn=None):
    """One launch of function `name` over the batch the arguments imply; returns (n, width) or the single item."""
    single = True
    n = 1
    av = bv = tv = None
    if wa:
        av = arr(a).reshape(-1, wa)
        single = single and np.ndim(a) <= 1
        n = max(n, av.shape[0])
    if wb:
        bv = arr(b).reshape(-1, wb)
        single = single and np.ndim(b) <= 1
        n = max(n, bv.shape[0])
    if t is not None:
        tv = arr(t).ravel()
        single = single and np.ndim(t) == 0
        n = max(n, tv.size)
        tv = np.ascontiguousarray(np.broadcast_to(tv, (n,)))
    if wa:
        av = np.ascontiguousarray(np.broadcast_to(av, (n, wa)))
    if wb:
        bv = np.ascontiguousarray(np.broadcast_to(bv, (n, wb)))
    so = _N_OUT.get(name, 3)
    out = np.empty((n, so))
    if fn is None:
        call("gelato_leaf_coordinate", _CODES[name], n, ptr(av), wa, ptr(bv), wb, ptr(tv), ptr(out))
    else:  # test hook: the same per-item function stepped on the host (tests/emu)
        fn(_CODES[name], n, ptr(av), wa, ptr(bv), wb, ptr(tv), ptr(out))
    res = out[:, 0] if so == 1 else out
    if single:
        return float(res[0]) if so == 1 else res[0]
    return res


def quatmult(q, p, fn=None):
    return _leaf("quatmult", q, 4, p, 4, fn=fn)


def conj(q, fn=None):
    return _leaf("conj", q, 4, fn=fn)


def normalize(v, fn=None):
    """v / |v| for 3- or 4-vectors (the reference's dynamic-size Eigen vector; other sizes are not offered)."""
    w = np.shape(v)[-1]
    if w not in (3, 4):
        raise ValueError("normalize: 3- or 4-vectors")
    return _leaf("normalize%d" % w, v, w, fn=fn)


def quatrot(q, v, fn=None):
    return _leaf("quatrot", q, 4, v, 3, fn=fn)


def ecef2geodetic(x, y, z, fn=None):
    """-> (lat deg, lon deg, alt m)"""
    return _leaf("ecef2geodetic", np.stack(np.broadcast_arrays(x, y, z), axis=-1), 3, fn=fn)


def geodetic2ecef(lat, lon, alt, fn=None):
    return _leaf("geodetic2ecef", np.stack(np.broadcast_arrays(lat, lon, alt), axis=-1), 3, fn=fn)


def ecef2eci(xyz, t, fn=None):
    return _leaf("ecef2eci", xyz, 3, t=t, fn=fn)


def eci2ecef(xyz, t, fn=None):
    return _leaf("eci2ecef", xyz, 3, t=t, fn=fn)


def vel_ecef2eci(vel_ecef, pos_ecef, t, fn=None):
    return _leaf("vel_ecef2eci", vel_ecef, 3, pos_ecef, 3, t=t, fn=fn)


def vel_eci2ecef(vel_eci, pos_eci, t, fn=None):
    return _leaf("vel_eci2ecef", vel_eci, 3, pos_eci, 3, t=t, fn=fn)


def quat_eci2ecef(t, fn=None):
    return _leaf("quat_eci2ecef", t=t, fn=fn)


def quat_ecef2eci(t, fn=None):
    return _leaf("quat_ecef2eci", t=t, fn=fn)


def quat_ecef2nedg(pos_ecef, fn=None):
    return _leaf("quat_ecef2nedg", pos_ecef, 3, fn=fn)


def quat_nedg2ecef(pos_ecef, fn=None):
    return _leaf("quat_nedg2ecef", pos_ecef, 3, fn=fn)


def quat_eci2nedg(pos_eci, t, fn=None):
    return _leaf("quat_eci2nedg", pos_eci, 3, t=t, fn=fn)


def quat_nedg2eci(pos_eci, t, fn=None):
    return _leaf("quat_nedg2eci", pos_eci, 3, t=t, fn=fn)


def quat_from_euler(az, el, ro, fn=None):
    """degrees"""
    return _leaf("quat_from_euler", np.stack(np.broadcast_arrays(az, el, ro), axis=-1), 3, fn=fn)


def euler_from_quat(q, fn=None):
    """-> (azimuth, elevation, roll) degrees"""
    return _leaf("euler_from_quat", q, 4, fn=fn)


def quat_nedg2body(quat_eci2body, pos_eci, t, fn=None):
    return _leaf("quat_nedg2body", quat_eci2body, 4, pos_eci, 3, t=t, fn=fn)


def orbital_elements(pos_eci, vel_eci, fn=None):
    """-> a [m], e, inclination, ascending node, argument of perigee, true anomaly [deg]"""
    return _leaf("orbital_elements", pos_eci, 3, vel_eci, 3, fn=fn)


def distance_vincenty(lat0, lon0, lat1, lon1, fn=None):
    """degrees in, metres out"""
    return _leaf("distance_vincenty", np.stack(np.broadcast_arrays(lat0, lon0, lat1, lon1), axis=-1), 4, fn=fn)


def angular_momentum_vec(pos_eci, vel_eci, fn=None):
    return _leaf("angular_momentum_vec", pos_eci, 3, vel_eci, 3, fn=fn)


def angular_momentum(pos_eci, vel_eci, fn=None):
    return _leaf("angular_momentum", pos_eci, 3, vel_eci, 3, fn=fn)


def inclination_rad(pos_eci, vel_eci, fn=None):
    return _leaf("inclination_rad", pos_eci, 3, vel_eci, 3, fn=fn)


def inclination_cosine(pos_eci, vel_eci, fn=None):
    return _leaf("inclination_cosine", pos_eci, 3, vel_eci, 3, fn=fn)


def orbit_energy(pos_eci, vel_eci, fn=None):
    return _leaf("orbit_energy", pos_eci, 3, vel_eci, 3, fn=fn)


def angular_momentum_from_altitude(ha, hp, fn=None):
    return _leaf("angular_momentum_from_altitude", np.stack(np.broadcast_arrays(ha, hp), axis=-1), 2, fn=fn)


def orbit_energy_from_altitude(ha, hp, fn=None):
    return _leaf("orbit_energy_from_altitude", np.stack(np.broadcast_arrays(ha, hp), axis=-1), 2, fn=fn)


def laplace_vector(pos_eci, vel_eci, fn=None):
    """v x h - mu r / |r| (wrapper_coordinate.hpp:238-244)"""
    return _leaf("laplace_vector", pos_eci, 3, vel_eci, 3, fn=fn)


def _haversine(lon1, lat1, lon2, lat2, r, fn=None):
    """utils_c.haversine (wrapper_utils.hpp:37-49; exported by lib/utils_c.py): great-circle distance on a sphere of radius
    r, degrees in.  It shares the batch entry point of this module's functions."""
    a = np.stack(np.broadcast_arrays(arr(lon1), arr(lat1), arr(lon2), arr(lat2)), axis=-1)
    return _leaf("haversine", a, 4, t=r, fn=fn)


# ---- the direction-cosine-matrix helpers (pybind_coordinate.cpp:33-36, 53-56); matrices as (3, 3) arrays, C[i][j] = C(i, j)
def _mat(res):
    res = np.asarray(res)
    return res.reshape(3, 3) if res.ndim == 1 else res.reshape(-1, 3, 3)


def dcm_from_quat(q, fn=None):
    return _mat(_leaf("dcm_from_quat", q, 4, fn=fn))


def quat_from_dcm(C, fn=None):
    C = arr(C)
    return _leaf("quat_from_dcm", C.reshape(9) if C.ndim == 2 else C.reshape(-1, 9), 9, fn=fn)


def euler_from_dcm(C, fn=None):
    """degrees (azimuth, elevation, roll), wrapper_coordinate.hpp:182-186"""
    C = arr(C)
    return _leaf("euler_from_dcm", C.reshape(9) if C.ndim == 2 else C.reshape(-1, 9), 9, fn=fn)


def dcm_from_thrustvector(pos_eci, thrustvec_eci, fn=None):
    return _mat(_leaf("dcm_from_thrustvector", pos_eci, 3, thrustvec_eci, 3, fn=fn))
