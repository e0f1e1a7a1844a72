"""Oracle restatement (TEST INFRASTRUCTURE ONLY) of the reference's forward-simulation initial guess:
/root/reference/initialize.py:37-111 `dynamics_init`, :114-179 `rocket_simulation`, :182-221
`zerolift_turn_correct`, :229-235 `integrate_runge_kutta_4d`, :238-319 `initialize_xdict_6DoF_2`.

Same operations in the same order, on the oracle's leaves (oracle/leaves.py: libm flavour = the reference's C++
bit for bit, gmath flavour = the bit-twin of the CUDA kernels).  tests/test_initguess.py checks it live against the
reference's own functions where /root/reference exists.

Two things the reference leaves to its environment are fixed here, and said so:
  * `norm` / `sqrt`: initialize.py takes them from `from lib.utils_c import *`, but the pybind module does not export
    them (/root/reference/src/pybind_utils.cpp:28-48; the legacy lib/utils.py did, through `from numpy.linalg import
    norm` and `from math import sqrt`) -- as shipped the forward simulation stops with a NameError.  The legacy
    meaning is used: norm(v) = sqrt(v . v) with numpy's dot (BLAS ddot: fused multiply-add, ascending index -- the
    `dot` argument selects numpy itself or the sequential-FMA statement of it, as in oracle/nlp.py).
  * `condition["rf_m"]`: read at :283 only to be printed; it plays no part in the numbers.
"""
import math

import numpy as np

from . import leaves as _leaves


def _np_interp(x, xp, fp):
    """numpy.interp for a scalar x (numpy/_core/src/multiarray/compiled_base.c: binary search for the interval,
    slope * (x - xp[j]) + fp[j]; left / right values outside)."""
    n = len(xp)
    if x <= xp[0]:
        return fp[0]
    if x >= xp[n - 1]:
        return fp[n - 1]
    j = int(np.searchsorted(xp, x, side="right")) - 1
    slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])
    res = slope * (x - xp[j]) + fp[j]
    if res != res:
        res = slope * (x - xp[j + 1]) + fp[j + 1]
        if res != res and fp[j] == fp[j + 1]:
            res = fp[j]
    return res


class ForwardSimulation:
    def __init__(self, leaves, dot="numpy"):
        self.crd, self.utl, self.atm = leaves.coordinate_c, leaves.utils_c, leaves.USStandardAtmosphere_c
        self.dot, self.lib = dot, leaves.lib

    def norm(self, v):
        if self.dot == "numpy":
            return float(np.sqrt(np.dot(v, v)))
        # the sequential-FMA statement of the dot product (oracle_leaves.cpp o_seqfma_matmul: product of the first
        # pair, then fused multiply-adds in ascending index -- what the BLAS ddot kernel does for three elements)
        vv = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(1)
        self.lib.o_seqfma_matmul(_leaves._p(vv), _leaves._p(vv), 1, vv.size, 1, _leaves._p(out))
        return math.sqrt(out[0])

    def air_velocity(self, pos_eci, vel_eci, t, wind):
        c, u, a = self.crd, self.utl, self.atm
        pos_llh = c.ecef2geodetic(pos_eci[0], pos_eci[1], pos_eci[2])
        altitude_m = a.geopotential_altitude(pos_llh[2])
        vel_ecef = c.vel_eci2ecef(vel_eci, pos_eci, t)
        vel_wind_ned = u.wind_ned(altitude_m, wind)
        vel_wind_eci = c.quatrot(c.quat_nedg2eci(pos_eci, t), vel_wind_ned)
        return altitude_m, c.ecef2eci(vel_ecef, t) - vel_wind_eci

    def dynamics_init(self, x, u, t, param, zlt, wind, ca):  # initialize.py:37-111
        c, a = self.crd, self.atm
        mass, pos_eci, vel_eci, quat = x[0], x[1:4], x[4:7], x[7:11]
        altitude_m, vel_air_eci = self.air_velocity(pos_eci, vel_eci, t, wind)
        rho = a.airdensity_at(altitude_m)
        p = a.airpressure_at(altitude_m)
        mach_number = self.norm(vel_air_eci) / a.speed_of_sound(altitude_m)
        coeff = _np_interp(mach_number, ca[:, 0], ca[:, 1])
        ret = np.zeros(11)
        aero_n_eci = 0.5 * rho * self.norm(vel_air_eci) * -vel_air_eci * param[2] * coeff
        thrust_n = param[0] - param[4] * p
        if zlt:
            thrustdir_eci = c.normalize(vel_air_eci)
        else:
            thrustdir_eci = c.quatrot(c.conj(quat), np.array([1.0, 0.0, 0.0]))
        thrust_n_eci = thrustdir_eci * thrust_n
        acc_eci = c.gravity(pos_eci) + (thrust_n_eci + aero_n_eci) / mass
        omega = np.deg2rad(np.array([0.0, u[0], u[1], u[2]]))
        ret[0] = -param[1]
        ret[1:4] = vel_eci
        ret[4:7] = acc_eci
        ret[7:11] = 0.5 * c.quatmult(quat, omega)
        return ret

    def zerolift_turn_correct(self, x, t, wind):  # :182-221
        c = self.crd
        pos_eci, vel_eci = x[1:4], x[4:7]
        _, vel_air_eci = self.air_velocity(pos_eci, vel_eci, t, wind)
        xb = c.normalize(vel_air_eci)
        yb = c.normalize(np.cross(vel_air_eci, pos_eci))
        zb = np.cross(xb, yb)
        q0 = 0.5 * math.sqrt(1.0 + xb[0] + yb[1] + zb[2])
        q1 = 0.25 / q0 * (yb[2] - zb[1])
        q2 = 0.25 / q0 * (zb[0] - xb[2])
        q3 = 0.25 / q0 * (xb[1] - yb[0])
        return c.normalize(np.array((q0, q1, q2, q3)))

    @staticmethod
    def rk4(function, x, t, dt):  # :229-235
        k1 = function(x, t)
        k2 = function(x + dt / 2.0 * k1, t + dt / 2.0)
        k3 = function(x + dt / 2.0 * k2, t + dt / 2.0)
        k4 = function(x + dt * k3, t + dt)
        return x + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0 * dt

    def rocket_simulation(self, x_init, u_table, pdict, t_init, t_out, dt=0.1):  # :114-179
        c = self.crd
        x = np.array(x_init, dtype=np.float64)  # the reference works on the caller's array: the jettison below mutates it
        x_map, t_map, u_map = [x], [t_init], [[0.0, 0.0, 0.0]]
        t = t_init
        t_final = t_out[-1] if hasattr(t_out, "__iter__") else t_out
        event_index = -1
        param = np.zeros(5)
        wind, ca = pdict["wind_table"], pdict["ca_table"]
        prm = pdict["params"]
        while t < t_final:
            tn = t + dt
            if event_index < len(prm) - 1:
                if tn > prm[event_index + 1]["time"]:
                    event_index += 1
                    param[0] = prm[event_index]["thrust"]
                    param[1] = prm[event_index]["massflow"]
                    param[2] = prm[event_index]["reference_area"]
                    param[4] = prm[event_index]["nozzle_area"]
                    x[0] -= prm[event_index]["mass_jettison"]  # in place: also the last recorded state
            u = np.array([_np_interp(t, u_table[:, 0], u_table[:, i + 1]) for i in range(3)])
            x = self.rk4(lambda xa, ta: self.dynamics_init(xa, u, ta, param, False, wind, ca), x, t, dt)
            t = t + dt
            if prm[event_index]["attitude"] == "zero-lift-turn":
                x[7:11] = self.zerolift_turn_correct(x, t, wind)
            x[7:11] = c.normalize(x[7:11])
            t_map.append(t)
            x_map.append(x)
            u_map.append(u)
        x_map, u_map, t_map = np.array(x_map), np.array(u_map), np.array(t_map)
        t_out = np.atleast_1d(np.asarray(t_out, dtype=np.float64))
        x_out = np.array([[_np_interp(to, t_map, x_map[:, i]) for i in range(11)] for to in t_out])
        u_out = np.array([[_np_interp(to, t_map, u_map[:, i]) for i in range(3)] for to in t_out])
        return x_out, u_out

    def initialize_xdict(self, x_init, pdict, unitdict, dt=0.005):  # :238-319 (mode "LGR", no display)
        ps = pdict["ps_params"]
        S = pdict["num_sections"]
        time_nodes, time_x_nodes = np.array([]), np.array([])
        for i in range(S):
            to, tf = pdict["params"][i]["time"], pdict["params"][i]["timeFinishAt"]
            tau = ps.tau(i)
            tau_x = np.hstack((-1.0, tau))
            time_nodes = np.hstack((time_nodes, tau * (tf - to) / 2.0 + (tf + to) / 2.0))
            time_x_nodes = np.hstack((time_x_nodes, tau_x * (tf - to) / 2.0 + (tf + to) / 2.0))
        xdict = {"t": (np.array([e["time"] for e in pdict["params"]]) / unitdict["t"]).ravel()}
        u_nodes = np.vstack([[[pdict["params"][i]["pitchrate_init"], pdict["params"][i]["yawrate_init"]]] * ps.nodes(i)
                             for i in range(S)])
        xdict["u"] = (u_nodes / unitdict["u"]).ravel()
        u_table = np.hstack((time_nodes.reshape(-1, 1), np.column_stack((np.zeros(len(u_nodes)), u_nodes))))
        x_nodes, _ = self.rocket_simulation(x_init, u_table, pdict, time_nodes[0], time_x_nodes, dt)
        xdict["mass"] = x_nodes[:, 0] / unitdict["mass"]
        xdict["position"] = (x_nodes[:, 1:4] / unitdict["position"]).ravel()
        xdict["velocity"] = (x_nodes[:, 4:7] / unitdict["velocity"]).ravel()
        xdict["quaternion"] = (x_nodes[:, 7:11]).ravel()
        return xdict, u_table, time_x_nodes
