"""The hand-written oracle restatement (oracle/oracle_leaves.cpp, libm flavour) against the
REFERENCE'S OWN C++ sources compiled where they lie (oracle/_ref/libref_leaves.so, built by
`make -C oracle ref` against oracle/ref_shim where /root/reference is mounted).

Every leaf the hot path touches must agree BIT FOR BIT on random and on trajectory inputs: both
sides run the same glibc elementary functions, so any difference would be a difference in
formula or operation order between the restatement and the reference's code.
Skipped where the reference build is absent.
"""
import numpy as np
import pytest

import helpers
from oracle import leaves

pytestmark = pytest.mark.skipif(not (leaves.ref_available() or __import__("os").path.isdir(leaves.REF_SRC)),
                                reason="oracle/_ref not built and no reference tree to build it from")


@pytest.fixture(scope="module")
def pair():
    return leaves.get("libm"), leaves.get("ref")


def _states(n, seed=0):
    """n plausible flight states: position 0..400 km above the ellipsoid, velocity 0..8 km/s."""
    rng = np.random.default_rng(seed)
    lat, lon = rng.uniform(-1.4, 1.4, n), rng.uniform(-np.pi, np.pi, n)
    r = 6356752.0 + rng.uniform(21500.0, 420000.0, n)
    pos = np.stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)], axis=1)
    vel = rng.normal(size=(n, 3)) * rng.uniform(1.0, 8000.0, (n, 1))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.uniform(0.0, 900.0, n)
    return pos, vel, q, t


def _eq(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), (what, int((~same).sum()), float(np.nanmax(np.abs(a - b))))


def test_atmosphere(pair):
    A, B = pair
    hs = np.concatenate([np.linspace(-2000.0, 1.2e6, 4001), [0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0,
                         71000.0, 86000.0, 91000.0, 110000.0, 120000.0, 85999.999, 86000.001]])
    for name in ("geopotential_altitude", "airtemperature_at", "airpressure_at", "airdensity_at", "speed_of_sound"):
        fa, fb = getattr(A.USStandardAtmosphere_c, name), getattr(B.USStandardAtmosphere_c, name)
        _eq([fa(h) for h in hs], [fb(h) for h in hs], name)


def test_coordinate_functions(pair):
    A, B = pair
    pos, vel, q, t = _states(400, 1)
    ca, cb = A.coordinate_c, B.coordinate_c
    for i in range(pos.shape[0]):
        p, v, qq, tt = pos[i], vel[i], q[i], t[i]
        _eq(ca.ecef2geodetic(*p), cb.ecef2geodetic(*p), "ecef2geodetic")
        g = ca.ecef2geodetic(*p)
        _eq(ca.geodetic2ecef(*g), cb.geodetic2ecef(*g), "geodetic2ecef")
        for name in ("ecef2eci", "eci2ecef", "quat_eci2nedg", "quat_nedg2eci", "eci2geodetic"):
            _eq(getattr(ca, name)(p, tt), getattr(cb, name)(p, tt), name)
        for name in ("vel_ecef2eci", "vel_eci2ecef"):
            _eq(getattr(ca, name)(v, p, tt), getattr(cb, name)(v, p, tt), name)
        for name in ("quat_ecef2nedg", "quat_nedg2ecef", "gravity"):
            _eq(getattr(ca, name)(p), getattr(cb, name)(p), name)
        for name in ("quat_eci2ecef", "quat_ecef2eci"):
            _eq(getattr(ca, name)(tt), getattr(cb, name)(tt), name)
        _eq(ca.euler_from_quat(qq), cb.euler_from_quat(qq), "euler_from_quat")
        _eq(ca.quat_nedg2body(qq, p, tt), cb.quat_nedg2body(qq, p, tt), "quat_nedg2body")
        _eq(ca.quatmult(qq, q[i - 1]), cb.quatmult(qq, q[i - 1]), "quatmult")
        _eq(ca.conj(qq), cb.conj(qq), "conj")
        _eq(ca.quatrot(qq, v), cb.quatrot(qq, v), "quatrot")
        _eq(ca.normalize(v), cb.normalize(v), "normalize")
        C = ca.dcm_from_quat(qq)
        _eq(C, cb.dcm_from_quat(qq), "dcm_from_quat")
        _eq(ca.quat_from_dcm(C), cb.quat_from_dcm(C), "quat_from_dcm")
        _eq(ca.euler_from_dcm(C), cb.euler_from_dcm(C), "euler_from_dcm")
        _eq(ca.dcm_from_thrustvector(p, v), cb.dcm_from_thrustvector(p, v), "dcm_from_thrustvector")
        _eq(ca.dcm_from_thrustvector(p, 2.0 * p), cb.dcm_from_thrustvector(p, 2.0 * p), "dcm_from_thrustvector (parallel)")
        for name in ("angular_momentum_vec", "angular_momentum", "inclination_rad", "inclination_cosine",
                     "orbit_energy", "orbital_elements", "laplace_vector"):
            _eq(getattr(ca, name)(p, v), getattr(cb, name)(p, v), name)
    for az, el, ro in np.random.default_rng(2).uniform(-180.0, 180.0, (50, 3)):
        _eq(ca.quat_from_euler(az, el, ro), cb.quat_from_euler(az, el, ro), "quat_from_euler")
    for ha, hp in ((200e3, 200e3), (35786e3, 250e3), (500e3, 180e3)):
        _eq(ca.angular_momentum_from_altitude(ha, hp), cb.angular_momentum_from_altitude(ha, hp), "h(alt)")
        _eq(ca.orbit_energy_from_altitude(ha, hp), cb.orbit_energy_from_altitude(ha, hp), "E(alt)")
    for a in np.random.default_rng(3).uniform(-80.0, 80.0, (40, 4)):
        _eq(ca.distance_vincenty(*a), cb.distance_vincenty(*a), "distance_vincenty")
        _eq(A.utils_c.haversine(*a, 6378137.0), B.utils_c.haversine(*a, 6378137.0), "haversine")


def test_tables_aero_iip_dynamics(pair):
    A, B = pair
    inp = helpers.example_inputs()
    wind, catab = np.asarray(inp["wind_table"], dtype=np.float64), np.asarray(inp["ca_table"], dtype=np.float64)
    pos, vel, q, t = _states(600, 4)
    pos[:300] *= (6378137.0 + np.linspace(0.0, 90e3, 300))[:, None] / np.linalg.norm(pos[:300], axis=1, keepdims=True)
    vel[:300] *= 0.2
    ua, ub = A.utils_c, B.utils_c
    for x in np.concatenate([np.linspace(-10.0, 40000.0, 200), wind[1:-1, 0]]):
        _eq(ua.interp(x, wind[:, 0], wind[:, 1]), ub.interp(x, wind[:, 0], wind[:, 1]), "interp")
        _eq(ua.wind_ned(x, wind), ub.wind_ned(x, wind), "wind_ned")
    _eq(ua.angle_of_attack_all_array_rad(pos, vel, q, t, wind), ub.angle_of_attack_all_array_rad(pos, vel, q, t, wind), "aoa")
    _eq(ua.dynamic_pressure_array_pa(pos, vel, t, wind), ub.dynamic_pressure_array_pa(pos, vel, t, wind), "q")
    _eq(ua.q_alpha_array_pa_rad(pos, vel, q, t, wind), ub.q_alpha_array_pa_rad(pos, vel, q, t, wind), "q-alpha")
    for i in range(0, pos.shape[0], 3):
        _eq(ua.angle_of_attack_ab_rad(pos[i], vel[i], q[i], t[i], wind), ub.angle_of_attack_ab_rad(pos[i], vel[i], q[i], t[i], wind),
            "angle_of_attack_ab_rad")
    for i in range(pos.shape[0]):
        pe = A.coordinate_c.eci2ecef(pos[i], t[i])
        ve = A.coordinate_c.vel_eci2ecef(vel[i], pos[i], t[i])
        for fill in (True, False):
            _eq(A.IIP_c.posLLH_IIP_FAA(pe, ve, fill), B.IIP_c.posLLH_IIP_FAA(pe, ve, fill), "IIP")
    units = np.array([27442.0, 6378137.0, 1000.0])
    mass = np.random.default_rng(5).uniform(0.05, 1.0, pos.shape[0])
    param = np.array([420000.0, 140.9, 2.21, 0.0, 0.68])
    tn = t / 630.0
    da, db = A.dynamics_c, B.dynamics_c
    _eq(da.dynamics_velocity(mass, pos / units[1], vel / units[2], q, tn, param, wind, catab, units),
        db.dynamics_velocity(mass, pos / units[1], vel / units[2], q, tn, param, wind, catab, units), "dynamics_velocity")
    _eq(da.dynamics_velocity_NoAir(mass, pos / units[1], q, param, units),
        db.dynamics_velocity_NoAir(mass, pos / units[1], q, param, units), "dynamics_velocity_NoAir")
    u = np.random.default_rng(6).uniform(-2.0, 2.0, (pos.shape[0], 2))
    _eq(da.dynamics_quaternion(q, u, 1.0), db.dynamics_quaternion(q, u, 1.0), "dynamics_quaternion")
