"""GPU tier: the batch leaf kernels behind gelato_b200.lib.{dynamics_c, utils_c, coordinate_c, IIP_c,
USStandardAtmosphere_c} (C ABI gelato_leaf_*) against the oracle's gmath flavour, BIT FOR BIT, with the
reference's module / function names and argument lists (src/pybind_*.cpp)."""
import numpy as np
import pytest

import helpers
from gelato_b200.lib import IIP_c, USStandardAtmosphere_c, coordinate_c, dynamics_c, utils_c
from oracle import leaves

pytestmark = pytest.mark.gpu


def _states(n, seed):
    rng = np.random.default_rng(seed)
    lat, lon = rng.uniform(-1.4, 1.4, n), rng.uniform(-np.pi, np.pi, n)
    r = 6356752.0 + rng.uniform(21500.0, 420000.0, n)
    r[: n // 2] = 6378137.0 + rng.uniform(0.0, 90e3, n // 2)
    pos = np.stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)], axis=1)
    vel = rng.normal(size=(n, 3)) * rng.uniform(1.0, 8000.0, (n, 1))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return pos, vel, q, rng.uniform(0.0, 900.0, n)


def _eq(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), (what, int((~same).sum()), float(np.nanmax(np.abs(a - b))))


def test_dynamics_c():
    O = leaves.get("gmath").dynamics_c
    inp = helpers.example_inputs()
    wind, ca = np.asarray(inp["wind_table"], dtype=float), np.asarray(inp["ca_table"], dtype=float)
    pos, vel, q, t = _states(3000, 1)
    units = np.array([27442.0, 6378137.0, 1000.0])
    mass = np.random.default_rng(2).uniform(0.05, 1.0, pos.shape[0])
    param = np.array([420000.0, 140.9, 2.21, 0.0, 0.68])
    a = (mass, pos / units[1], vel / units[2], q, t / 630.0, param, wind, ca, units)
    _eq(dynamics_c.dynamics_velocity(*a), O.dynamics_velocity(*a), "dynamics_velocity")
    b = (mass, pos / units[1], q, param, units)
    _eq(dynamics_c.dynamics_velocity_NoAir(*b), O.dynamics_velocity_NoAir(*b), "dynamics_velocity_NoAir")
    u = np.random.default_rng(3).uniform(-2.0, 2.0, (pos.shape[0], 2))
    _eq(dynamics_c.dynamics_quaternion(q, u, 1.0), O.dynamics_quaternion(q, u, 1.0), "dynamics_quaternion")


def test_utils_c_and_coordinate_c():
    L = leaves.get("gmath")
    wind = np.asarray(helpers.example_inputs()["wind_table"], dtype=float)
    pos, vel, q, t = _states(2000, 4)
    vel[:1000] *= 0.2
    _eq(utils_c.angle_of_attack_all_array_rad(pos, vel, q, t, wind),
        L.utils_c.angle_of_attack_all_array_rad(pos, vel, q, t, wind), "alpha")
    _eq(utils_c.dynamic_pressure_array_pa(pos, vel, t, wind), L.utils_c.dynamic_pressure_array_pa(pos, vel, t, wind), "q")
    _eq(utils_c.q_alpha_array_pa_rad(pos, vel, q, t, wind), L.utils_c.q_alpha_array_pa_rad(pos, vel, q, t, wind), "qa")
    assert utils_c.q_alpha_pa_rad(pos[3], vel[3], q[3], t[3], wind) == L.utils_c.q_alpha_pa_rad(pos[3], vel[3], q[3], t[3], wind)
    _eq(coordinate_c.eci2geodetic(pos, t), [L.coordinate_c.eci2geodetic(p, tt) for p, tt in zip(pos, t)], "eci2geodetic")
    _eq(coordinate_c.eci2geodetic(pos[0], t[0]), L.coordinate_c.eci2geodetic(pos[0], t[0]), "eci2geodetic one")
    _eq(coordinate_c.gravity(pos), [L.coordinate_c.gravity(p) for p in pos], "gravity")
    k = 300
    _eq(utils_c.angle_of_attack_ab_array_rad(pos[:k], vel[:k], q[:k], t[:k], wind),
        [L.utils_c.angle_of_attack_ab_rad(pos[i], vel[i], q[i], t[i], wind) for i in range(k)], "alpha pitch / yaw")
    _eq(utils_c.angle_of_attack_ab_rad(pos[5], vel[5], q[5], t[5], wind),
        L.utils_c.angle_of_attack_ab_rad(pos[5], vel[5], q[5], t[5], wind), "alpha pitch / yaw, one point")
    alt = np.random.default_rng(8).uniform(-500.0, 60000.0, k)
    _eq(utils_c.wind_ned(alt, wind), [L.utils_c.wind_ned(a, wind) for a in alt], "wind_ned")
    _eq(utils_c.wind_ned(1234.5, wind), L.utils_c.wind_ned(1234.5, wind), "wind_ned, one altitude")


def test_iip_and_atmosphere():
    L = leaves.get("gmath")
    pos, vel, q, t = _states(1500, 5)
    pe = np.array([L.coordinate_c.eci2ecef(p, tt) for p, tt in zip(pos, t)])
    ve = np.array([L.coordinate_c.vel_eci2ecef(v, p, tt) for v, p, tt in zip(vel, pos, t)])
    for fill in (True, False):
        want = np.array([L.IIP_c.posLLH_IIP_FAA(a, b, fill) for a, b in zip(pe, ve)])
        _eq(IIP_c.posLLH_IIP_FAA(pe, ve, fill), want, "IIP fill_na=%s" % fill)
    assert np.isnan(want).any() and (np.abs(want[~np.isnan(want).any(axis=1)]) > 0).any()  # both outcomes occur
    hs = np.concatenate([np.linspace(-2000.0, 1.2e6, 3001), [0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0,
                         86000.0, 91000.0, 110000.0, 120000.0]])
    A = L.USStandardAtmosphere_c
    for name in ("geopotential_altitude", "airtemperature_at", "airpressure_at", "airdensity_at", "speed_of_sound"):
        _eq(getattr(USStandardAtmosphere_c, name)(hs), [getattr(A, name)(h) for h in hs], name)
    assert USStandardAtmosphere_c.airdensity_at(1234.5) == A.airdensity_at(1234.5)
