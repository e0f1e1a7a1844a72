"""CPU oracle of the reference's result table (TEST INFRASTRUCTURE ONLY): a restatement of
/root/reference/output_result.py:37-263.  The per-node quantities come from
oracle_leaves.cpp:output_row (C++, either flavour); the bookkeeping around them -- which section's
parameters a node uses (:130-148, including the node-after-the-boundary rule), time rounding (:69),
control-rate interpolation (:110-111), numpy norm of the velocity (:103) -- is restated here.
"""
import ctypes

import numpy as np

from . import leaves as _leaves

COLUMNS = [
    "event", "time", "stage", "section", "thrust", "mass", "lat", "lon", "lat_IIP", "lon_IIP", "downrange", "altitude",
    "altitude_apogee", "altitude_perigee", "inclination", "argument_perigee", "lon_ascending_node", "true_anomaly",
    "pos_ECI_X", "pos_ECI_Y", "pos_ECI_Z", "vel_ECI_X", "vel_ECI_Y", "vel_ECI_Z", "vel_ground_NED_X", "vel_ground_NED_Y",
    "vel_ground_NED_Z", "quat_ECI2BODY_0", "quat_ECI2BODY_1", "quat_ECI2BODY_2", "quat_ECI2BODY_3", "accel_BODY_X",
    "aero_BODY_X", "heading_NED2BODY", "pitch_NED2BODY", "roll_NED2BODY", "vel_inertial",
    "flightpath_vel_inertial_geocentric", "azimuth_vel_inertial_geocentric", "thrust_direction_ECI_X",
    "thrust_direction_ECI_Y", "thrust_direction_ECI_Z", "rate_BODY_X", "rate_BODY_Y", "rate_BODY_Z", "vel_ground", "vel_air",
    "AOA_total", "AOA_pitch", "AOA_yaw", "dynamic_pressure", "Q_alpha", "M",
]
# column of the per-node kernel output (output_row) -> table column
KERNEL_COLUMNS = [
    "thrust", "lat", "lon", "lat_IIP", "lon_IIP", "downrange", "altitude", "altitude_apogee", "altitude_perigee",
    "inclination", "argument_perigee", "lon_ascending_node", "true_anomaly", "vel_ground_NED_X", "vel_ground_NED_Y",
    "vel_ground_NED_Z", "accel_BODY_X", "aero_BODY_X", "heading_NED2BODY", "pitch_NED2BODY", "roll_NED2BODY",
    "flightpath_vel_inertial_geocentric", "azimuth_vel_inertial_geocentric", "thrust_direction_ECI_X",
    "thrust_direction_ECI_Y", "thrust_direction_ECI_Z", "vel_ground", "vel_air", "AOA_total", "AOA_pitch", "AOA_yaw",
    "dynamic_pressure", "Q_alpha", "M",
]


def node_sections(pdict, n_rows):
    """Section whose parameters row i uses, and the event names / stages of the table
    (output_result.py:126-148: the section counter advances AFTER the row that reaches the boundary)."""
    ps = pdict["ps_params"]
    sec = np.zeros(n_rows, dtype="i4")
    event = [""] * n_rows
    stage = [None] * n_rows
    section = 0
    event[0] = pdict["params"][0]["name"]
    for i in range(n_rows):
        sec[i] = section
        stage[i] = pdict["params"][section]["rocketStage"]
        if i >= ps.index_start_u(section) + ps.nodes(section) + section:
            event[i] = pdict["params"][section + 1]["name"]
            section += 1
    return sec, event, stage


def assemble(xdict, unitdict, tx_res, tu_res, pdict, kernel_rows, sec, event, stage):
    """The table from the per-node kernel output [n][34] and the pass-through / host columns."""
    n = len(tx_res)
    pos = xdict["position"].reshape(-1, 3) * unitdict["position"]
    vel = xdict["velocity"].reshape(-1, 3) * unitdict["velocity"]
    quat = xdict["quaternion"].reshape(-1, 4)
    u = xdict["u"].reshape(-1, 2) * unitdict["u"]
    out = {"event": event, "time": np.asarray(tx_res).round(6), "stage": stage, "section": sec,
           "mass": xdict["mass"] * unitdict["mass"]}
    for k in range(3):
        out["pos_ECI_" + "XYZ"[k]] = pos[:, k]
        out["vel_ECI_" + "XYZ"[k]] = vel[:, k]
    for k in range(4):
        out["quat_ECI2BODY_%d" % k] = quat[:, k]
    out["vel_inertial"] = np.linalg.norm(vel, axis=1)
    out["rate_BODY_X"] = np.zeros(n)
    out["rate_BODY_Y"] = np.interp(tx_res, tu_res, u[:, 0])
    out["rate_BODY_Z"] = np.interp(tx_res, tu_res, u[:, 1])
    for j, name in enumerate(KERNEL_COLUMNS):
        out[name] = kernel_rows[:, j]
    return {c: out[c] for c in COLUMNS}


def node_inputs(xdict, unitdict, tx_res, pdict, sec):
    mass = np.ascontiguousarray(xdict["mass"] * unitdict["mass"])
    pos = np.ascontiguousarray(xdict["position"].reshape(-1, 3) * unitdict["position"])
    vel = np.ascontiguousarray(xdict["velocity"].reshape(-1, 3) * unitdict["velocity"])
    quat = np.ascontiguousarray(xdict["quaternion"].reshape(-1, 4), dtype=np.float64)
    t = np.ascontiguousarray(tx_res, dtype=np.float64)
    prm = pdict["params"]
    thrust = np.array([prm[s]["thrust"] for s in sec], dtype=np.float64)
    area = np.array([prm[s]["reference_area"] for s in sec], dtype=np.float64)
    nozzle = np.array([prm[s]["nozzle_area"] for s in sec], dtype=np.float64)
    return mass, pos, vel, quat, t, thrust, area, nozzle


def output_result(xdict, unitdict, tx_res, tu_res, pdict, flavour="libm"):
    """dict of columns in the reference's order (pandas.DataFrame(result) gives the reference's frame)."""
    L = _leaves.get(flavour).lib
    n = len(tx_res)
    sec, event, stage = node_sections(pdict, n)
    mass, pos, vel, quat, t, thrust, area, nozzle = node_inputs(xdict, unitdict, tx_res, pdict, sec)
    wind = np.ascontiguousarray(pdict["wind_table"], dtype=np.float64)
    ca = np.ascontiguousarray(pdict["ca_table"], dtype=np.float64)
    rows = np.empty((n, len(KERNEL_COLUMNS)))
    P = ctypes.POINTER(ctypes.c_double)
    p = lambda a: a.ctypes.data_as(P)  # noqa: E731
    L.o_output_rows(ctypes.c_int(n), p(mass), p(pos), p(vel), p(quat), p(t), p(thrust), p(area), p(nozzle), p(wind),
                    ctypes.c_int(wind.shape[0]), p(ca), ctypes.c_int(ca.shape[0]),
                    ctypes.c_double(pdict["LaunchCondition"]["lat"]), ctypes.c_double(pdict["LaunchCondition"]["lon"]), p(rows))
    return assemble(xdict, unitdict, tx_res, tu_res, pdict, rows, sec, event, stage)
