"""The coordinate_c leaves that are not on the NLP path (gelato_leaf_coordinate; /root/reference/src/
pybind_coordinate.cpp:28-78): the kernel's per-item function, stepped on the host here and run through the C ABI on
the GPU, against the oracle's gmath leaves bit for bit on trajectory-like random inputs; the oracle's libm flavour is
itself bit-identical to the reference's C++ (tests/test_oracle_vs_refcpp.py)."""
import ctypes

import numpy as np
import pytest

import emu_binding
from gelato_b200.lib import coordinate_c as gc
from oracle import leaves


def _inputs(n=40, seed=11):
    rng = np.random.default_rng(seed)
    r = 6378137.0 + rng.uniform(0.0, 8.0e5, n)
    lat, lon = rng.uniform(-1.4, 1.4, n), rng.uniform(-3.1, 3.1, n)
    pos = np.column_stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)])
    vel = rng.normal(0.0, 1.0, (n, 3))
    vel = vel / np.linalg.norm(vel, axis=1)[:, None] * rng.uniform(200.0, 7900.0, n)[:, None]
    q = rng.normal(0.0, 1.0, (n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    p = rng.normal(0.0, 1.0, (n, 4))
    p /= np.linalg.norm(p, axis=1)[:, None]
    t = rng.uniform(0.0, 900.0, n)
    ang = rng.uniform(-179.0, 179.0, (n, 3)) * np.array([1.0, 0.49, 1.0])
    llh = np.column_stack([np.degrees(lat), np.degrees(lon), r - 6378137.0])
    alt = rng.uniform(1.5e5, 9.0e5, (n, 2))
    return dict(pos=pos, vel=vel, q=q, p=p, t=t, ang=ang, llh=llh, alt=alt)


# (function, arguments by name) -- every function the oracle restates
CASES = [
    ("quatmult", ("q", "p")), ("conj", ("q",)), ("normalize", ("pos",)), ("normalize", ("q",)), ("quatrot", ("q", "vel")),
    ("ecef2geodetic", ("pos0", "pos1", "pos2")), ("geodetic2ecef", ("llh0", "llh1", "llh2")), ("ecef2eci", ("pos", "t")),
    ("eci2ecef", ("pos", "t")), ("vel_ecef2eci", ("vel", "pos", "t")), ("vel_eci2ecef", ("vel", "pos", "t")),
    ("quat_eci2ecef", ("t",)), ("quat_ecef2eci", ("t",)), ("quat_ecef2nedg", ("pos",)), ("quat_nedg2ecef", ("pos",)),
    ("quat_eci2nedg", ("pos", "t")), ("quat_nedg2eci", ("pos", "t")), ("quat_from_euler", ("ang0", "ang1", "ang2")),
    ("euler_from_quat", ("q",)), ("quat_nedg2body", ("q", "pos", "t")), ("orbital_elements", ("pos", "vel")),
    ("distance_vincenty", ("llh0", "llh1", "ang1", "ang2")), ("angular_momentum_vec", ("pos", "vel")),
    ("angular_momentum", ("pos", "vel")), ("inclination_rad", ("pos", "vel")), ("inclination_cosine", ("pos", "vel")),
    ("orbit_energy", ("pos", "vel")), ("angular_momentum_from_altitude", ("alt0", "alt1")),
    ("orbit_energy_from_altitude", ("alt0", "alt1")), ("laplace_vector", ("pos", "vel")),
]


def _arg(d, name):
    if name[-1].isdigit():
        return d[name[:-1]][:, int(name[-1])]
    return d[name]


def _check(fn):
    O = leaves.get("gmath").coordinate_c
    d = _inputs()
    n = d["t"].size
    for name, argnames in CASES:
        args = [_arg(d, a) for a in argnames]
        got = getattr(gc, name)(*args, fn=fn)
        want = np.array([np.atleast_1d(getattr(O, name)(*[a[i] for a in args])) for i in range(n)])
        got2 = np.asarray(got).reshape(n, -1)
        assert got2.shape == want.shape, name
        assert np.array_equal(got2, want), (name, np.abs(got2 - want).max())
        one = getattr(gc, name)(*[a[3] for a in args], fn=fn)  # a single point, the reference's call shape
        assert np.array_equal(np.atleast_1d(one), want[3]), name
    # the four DCM helpers: (3, 3) matrices in / out
    C = gc.dcm_from_quat(d["q"], fn=fn)
    assert C.shape == (n, 3, 3)
    assert np.array_equal(C, np.array([O.dcm_from_quat(d["q"][i]) for i in range(n)]))
    assert np.array_equal(gc.dcm_from_quat(d["q"][5], fn=fn), O.dcm_from_quat(d["q"][5]))
    assert np.abs(np.einsum("nij,nkj->nik", C, C) - np.eye(3)).max() < 1e-12  # a rotation matrix (unit quaternions)
    for name in ("quat_from_dcm", "euler_from_dcm"):
        got = getattr(gc, name)(C, fn=fn)
        assert np.array_equal(got, np.array([getattr(O, name)(C[i]) for i in range(n)])), name
        assert np.array_equal(getattr(gc, name)(C[7], fn=fn), getattr(O, name)(C[7])), name
    T = gc.dcm_from_thrustvector(d["pos"], d["vel"], fn=fn)
    assert np.array_equal(T, np.array([O.dcm_from_thrustvector(d["pos"][i], d["vel"][i]) for i in range(n)]))
    par = gc.dcm_from_thrustvector(d["pos"][:4], 3.0 * d["pos"][:4], fn=fn)  # thrust along the position: the z-axis branch
    assert np.array_equal(par, np.array([O.dcm_from_thrustvector(d["pos"][i], 3.0 * d["pos"][i]) for i in range(4)]))
    assert np.all(np.isfinite(par))
    # utils_c.haversine rides on the same entry point (function code GC_HAVERSINE)
    from gelato_b200.lib import utils_c

    U = leaves.get("gmath").utils_c
    lon1, lat1, lon2, lat2 = d["ang"][:, 0], d["ang"][:, 1], d["ang"][:, 2], d["llh"][:, 0]
    got = utils_c.haversine(lon1, lat1, lon2, lat2, 6378137.0, fn=fn)
    want = np.array([U.haversine(lon1[i], lat1[i], lon2[i], lat2[i], 6378137.0) for i in range(n)])
    assert np.array_equal(got, want) and np.all(want > 0.0)
    assert utils_c.haversine(lon1[2], lat1[2], lon2[2], lat2[2], 6378137.0, fn=fn) == want[2]


def test_coordinate_leaves_match_the_oracle_on_the_host():
    emu_binding.build()
    _check(ctypes.CDLL(emu_binding.LIB).emu_coord_leaf)


@pytest.mark.gpu
def test_coordinate_leaves_match_the_oracle_on_the_gpu():
    _check(None)
