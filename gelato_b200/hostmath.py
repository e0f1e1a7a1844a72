"""Small host-side geodesy helpers used ONLY at problem set-up (launch state).

They play the role the reference's `coordinate_c` module plays in
/root/reference/Trajectory_Optimization.py:69-76,140-149 (geodetic2ecef,
ecef2eci, vel_ecef2eci, quat_eci2nedg, quat_from_euler, quatmult).  The values
they produce are *inputs* of the hot path (`condition["init"]`), not part of it,
so plain numpy is fine here; the hot path itself lives in csrc/.
"""
import math

import numpy as np

MU = 3.986004418e14
OMEGA = 7.2921151467e-5
RA = 6378137.0
F = 1.0 / 298.257223563
RB = RA * (1.0 - F)
E2 = (RA * RA - RB * RB) / RA / RA
EP2 = (RA * RA - RB * RB) / RB / RB


def geodetic2ecef(lat_deg, lon_deg, alt):
    lat = lat_deg * math.pi / 180.0
    lon = lon_deg * math.pi / 180.0
    n = RA / math.sqrt(1.0 - E2 * math.sin(lat) * math.sin(lat))
    return np.array(
        [
            (n + alt) * math.cos(lat) * math.cos(lon),
            (n + alt) * math.cos(lat) * math.sin(lon),
            (n * (1.0 - E2) + alt) * math.sin(lat),
        ]
    )


def ecef2geodetic_rad(p):
    r = math.sqrt(p[0] * p[0] + p[1] * p[1])
    theta = math.atan2(p[2] * RA, r * RB)
    lat = math.atan2(p[2] + EP2 * RB * math.sin(theta) ** 3, r - E2 * RA * math.cos(theta) ** 3)
    lon = math.atan2(p[1], p[0])
    n = RA / math.sqrt(1.0 - E2 * math.sin(lat) ** 2)
    return lat, lon, r / math.cos(lat) - n


def ecef2eci(a, t):
    c, s = math.cos(OMEGA * t), math.sin(OMEGA * t)
    return np.array([a[0] * c - a[1] * s, a[0] * s + a[1] * c, a[2]])


def eci2ecef(a, t):
    c, s = math.cos(OMEGA * t), math.sin(OMEGA * t)
    return np.array([a[0] * c + a[1] * s, -a[0] * s + a[1] * c, a[2]])


def vel_ecef2eci(vel_ecef, pos_ecef, t):
    pos_eci = ecef2eci(pos_ecef, t)
    return ecef2eci(vel_ecef, t) + np.cross([0.0, 0.0, OMEGA], pos_eci)


def quatmult(q, p):
    return np.array(
        [
            q[0] * p[0] - q[1] * p[1] - q[2] * p[2] - q[3] * p[3],
            q[0] * p[1] + q[1] * p[0] + q[2] * p[3] - q[3] * p[2],
            q[0] * p[2] - q[1] * p[3] + q[2] * p[0] + q[3] * p[1],
            q[0] * p[3] + q[1] * p[2] - q[2] * p[1] + q[3] * p[0],
        ]
    )


def quat_eci2nedg(pos_eci, t):
    q_eci2ecef = np.array([math.cos(OMEGA * t / 2.0), 0.0, 0.0, math.sin(OMEGA * t / 2.0)])
    lat, lon, _ = ecef2geodetic_rad(eci2ecef(pos_eci, t))
    c_hl, s_hl = math.cos(lon / 2.0), math.sin(lon / 2.0)
    c_hp, s_hp = math.cos(lat / 2.0), math.sin(lat / 2.0)
    r2 = math.sqrt(2.0)
    q_ecef2ned = np.array(
        [c_hl * (c_hp - s_hp) / r2, s_hl * (c_hp + s_hp) / r2, -c_hl * (c_hp + s_hp) / r2, s_hl * (c_hp - s_hp) / r2]
    )
    return quatmult(q_eci2ecef, q_ecef2ned)


def quat_from_euler(az_deg, el_deg, ro_deg):
    az, el, ro = (a * math.pi / 180.0 for a in (az_deg, el_deg, ro_deg))
    qz = np.array([math.cos(az / 2), 0.0, 0.0, math.sin(az / 2)])
    qy = np.array([math.cos(el / 2), 0.0, math.sin(el / 2), 0.0])
    qx = np.array([math.cos(ro / 2), math.sin(ro / 2), 0.0, 0.0])
    return quatmult(quatmult(qz, qy), qx)


def angular_momentum_from_altitude(ha, hp):
    """Target angular momentum of the terminal orbit (reference:
    src/wrapper_coordinate.hpp:252-258); only + - * / sqrt."""
    ra = RA + ha
    rp = RA + hp
    a = (ra + rp) / 2.0
    vp = math.sqrt(MU * (2.0 / rp - 1.0 / a))
    return rp * vp


def orbit_energy_from_altitude(ha, hp):
    """Target specific orbital energy (reference: src/wrapper_coordinate.hpp:260-265)."""
    ra = RA + ha
    rp = RA + hp
    a = (ra + rp) / 2.0
    return -MU / 2.0 / a
