"""The only known answers the reference ships for its physics leaves: the derived columns of
example/example-trajectory_init.csv (written by an older output_result next to the state columns), frozen in
tests/golden/example_trajectory_kinematics.npz.  The oracle (every flavour) and the GPU leaf kernels must
reproduce them from the table's own state columns.  Tolerances are what the table's own rounding allows
(time is stored to 1e-6 s, which moves the longitude by ~1e-9 deg; the IIP columns come from an older IIP
routine and agree to 1e-7 deg); the dynamic-pressure column does not reproduce."""
import os

import numpy as np
import pytest

import helpers
from oracle import leaves

K = np.load(os.path.join(helpers.GOLDEN, "example_trajectory_kinematics.npz"))
WIND = np.asarray(helpers.example_inputs()["wind_table"], dtype=np.float64)


def _check(llh, apo, per, inc, iip, aoa_deg):
    np.testing.assert_allclose(llh[:, 0], K["lat"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(llh[:, 1], K["lon"], rtol=0, atol=5e-9)
    np.testing.assert_allclose(llh[:, 2], K["altitude"], rtol=0, atol=1e-8)
    if apo is not None:
        np.testing.assert_allclose(apo, K["apogee"], rtol=5e-10, atol=1e-6)
        np.testing.assert_allclose(per, K["perigee"], rtol=1e-12, atol=1e-6)
        np.testing.assert_allclose(inc, K["inclination"], rtol=0, atol=1e-12)
    has = ~np.isnan(K["lat_iip"])  # blank cells: no impact point (orbital) -> the leaf returns zeros (iip.cpp:49-128)
    assert 0 < (~has).sum() < has.sum()
    np.testing.assert_allclose(iip[has, 0], K["lat_iip"][has], rtol=0, atol=2e-7)
    np.testing.assert_allclose(iip[has, 1], K["lon_iip"][has], rtol=0, atol=2e-7)
    assert not iip[~has].any()
    np.testing.assert_allclose(aoa_deg, K["aoa_deg"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("flavour", ["libm", "gmath", "ref"])
def test_oracle_reproduces_the_shipped_trajectory_table(flavour):
    if flavour == "ref" and not leaves.ref_available():
        pytest.skip("oracle/_ref not built")
    L = leaves.get(flavour)
    C = L.coordinate_c
    pos, vel, quat, t = K["pos"], K["vel"], K["quat"], K["t"]
    llh = np.array([C.eci2geodetic(p, tt) for p, tt in zip(pos, t)])
    oe = np.array([C.orbital_elements(p, v) for p, v in zip(pos, vel)])
    a, e = oe[:, 0], oe[:, 1]
    pe = np.array([C.eci2ecef(p, tt) for p, tt in zip(pos, t)])
    ve = np.array([C.vel_eci2ecef(v, p, tt) for v, p, tt in zip(vel, pos, t)])
    iip = np.array([L.IIP_c.posLLH_IIP_FAA(x, y) for x, y in zip(pe, ve)])
    aoa = np.degrees(L.utils_c.angle_of_attack_all_array_rad(pos, vel, quat, t, WIND))
    _check(llh, a * (1 + e) - 6378137.0, a * (1 - e) - 6378137.0, oe[:, 2], iip, aoa)


@pytest.mark.gpu
def test_gpu_leaf_kernels_reproduce_the_shipped_trajectory_table():
    from gelato_b200.lib import IIP_c, coordinate_c, utils_c

    L = leaves.get("gmath")
    pos, vel, quat, t = K["pos"], K["vel"], K["quat"], K["t"]
    llh = coordinate_c.eci2geodetic(pos, t)
    pe = np.array([L.coordinate_c.eci2ecef(p, tt) for p, tt in zip(pos, t)])
    ve = np.array([L.coordinate_c.vel_eci2ecef(v, p, tt) for v, p, tt in zip(vel, pos, t)])
    iip = IIP_c.posLLH_IIP_FAA(pe, ve)
    aoa = np.degrees(utils_c.angle_of_attack_all_array_rad(pos, vel, quat, t, WIND))
    _check(llh, None, None, None, iip, aoa)
