"""The reference's result table (`output_result(xdict, unitdict, tx_res, tu_res, pdict)`,
/root/reference/output_result.py:37-263) with its per-node body on the GPU.

The reference walks the state nodes in a Python loop and makes ~30 pybind11 leaf calls per node; here
ONE kernel launch (`gelato_leaf_output_table`, one thread per node) computes the 34 derived quantities
of every node -- of one trajectory, or of many solved scenarios at once (`output_tables`).  The host keeps
what is bookkeeping in the reference too: which section's parameters a row uses, event names, time
rounding, control-rate interpolation and the pass-through state columns.
"""
import ctypes

import numpy as np

from . import engine as _engine

COLUMNS = (
    "event time stage section thrust mass lat lon lat_IIP lon_IIP downrange altitude altitude_apogee altitude_perigee "
    "inclination argument_perigee lon_ascending_node true_anomaly pos_ECI_X pos_ECI_Y pos_ECI_Z vel_ECI_X vel_ECI_Y "
    "vel_ECI_Z vel_ground_NED_X vel_ground_NED_Y vel_ground_NED_Z quat_ECI2BODY_0 quat_ECI2BODY_1 quat_ECI2BODY_2 "
    "quat_ECI2BODY_3 accel_BODY_X aero_BODY_X heading_NED2BODY pitch_NED2BODY roll_NED2BODY vel_inertial "
    "flightpath_vel_inertial_geocentric azimuth_vel_inertial_geocentric thrust_direction_ECI_X thrust_direction_ECI_Y "
    "thrust_direction_ECI_Z rate_BODY_X rate_BODY_Y rate_BODY_Z vel_ground vel_air AOA_total AOA_pitch AOA_yaw "
    "dynamic_pressure Q_alpha M"
).split()
# order of the kernel's 34 outputs (include/gelato_b200.h, gelato_leaf_output_table)
KERNEL_COLUMNS = (
    "thrust lat lon lat_IIP lon_IIP downrange altitude altitude_apogee altitude_perigee inclination argument_perigee "
    "lon_ascending_node true_anomaly vel_ground_NED_X vel_ground_NED_Y vel_ground_NED_Z accel_BODY_X aero_BODY_X "
    "heading_NED2BODY pitch_NED2BODY roll_NED2BODY flightpath_vel_inertial_geocentric azimuth_vel_inertial_geocentric "
    "thrust_direction_ECI_X thrust_direction_ECI_Y thrust_direction_ECI_Z vel_ground vel_air AOA_total AOA_pitch AOA_yaw "
    "dynamic_pressure Q_alpha M"
).split()
_pd = ctypes.POINTER(ctypes.c_double)


def row_sections(pdict, n_rows):
    """(section index, event name, stage) of every table row.  A row belongs to the section that is
    current when the loop reaches it; the counter moves on after the row that closes a section, which is
    also the row that carries the next event's name (output_result.py:126-148)."""
    ps, prm = pdict["ps_params"], pdict["params"]
    sec = np.zeros(n_rows, dtype="i4")
    event, stage = [""] * n_rows, [None] * n_rows
    event[0] = prm[0]["name"]
    cur = 0
    for i in range(n_rows):
        sec[i], stage[i] = cur, prm[cur]["rocketStage"]
        if i >= ps.index_start_u(cur) + ps.nodes(cur) + cur:
            event[i] = prm[cur + 1]["name"]
            cur += 1
    return sec, event, stage


def _kernel_inputs(xdict, unitdict, tx_res, pdict, sec):
    prm = pdict["params"]
    pick = lambda key: np.array([prm[s][key] for s in sec], dtype=np.float64)  # noqa: E731
    return (np.ascontiguousarray(xdict["mass"] * unitdict["mass"], dtype=np.float64),
            np.ascontiguousarray(xdict["position"].reshape(-1, 3) * unitdict["position"], dtype=np.float64),
            np.ascontiguousarray(xdict["velocity"].reshape(-1, 3) * unitdict["velocity"], dtype=np.float64),
            np.ascontiguousarray(xdict["quaternion"].reshape(-1, 4), dtype=np.float64),
            np.ascontiguousarray(tx_res, dtype=np.float64), pick("thrust"), pick("reference_area"), pick("nozzle_area"))


def kernel_rows(inputs, wind, ca, launch_lat, launch_lon, device=0, fn=None):
    """[n][34] derived quantities for concatenated node inputs (one launch)."""
    mass = inputs[0]
    n = mass.size
    out = np.empty((n, len(KERNEL_COLUMNS)))
    wind = np.ascontiguousarray(wind, dtype=np.float64)
    ca = np.ascontiguousarray(ca, dtype=np.float64)
    ptrs = [a.ctypes.data_as(_pd) for a in inputs]
    if fn is None:
        L = _engine.load_library()
        rc = L.gelato_leaf_output_table(device, n, *ptrs, wind.ctypes.data_as(_pd), wind.shape[0], ca.ctypes.data_as(_pd),
                                        ca.shape[0], float(launch_lat), float(launch_lon), out.ctypes.data_as(_pd))
        if rc != 0:
            raise _engine.GelatoError("gelato_leaf_output_table failed (%d): %s" % (rc, L.gelato_last_error().decode()))
    else:  # test hook: the same per-node function stepped on the host (tests/emu)
        fn(n, *ptrs, wind.ctypes.data_as(_pd), wind.shape[0], ca.ctypes.data_as(_pd), ca.shape[0],
           ctypes.c_double(launch_lat), ctypes.c_double(launch_lon), out.ctypes.data_as(_pd))
    return out


def _assemble(xdict, unitdict, tx_res, tu_res, rows, sec, event, stage):
    pos = xdict["position"].reshape(-1, 3) * unitdict["position"]
    vel = xdict["velocity"].reshape(-1, 3) * unitdict["velocity"]
    quat = xdict["quaternion"].reshape(-1, 4)
    u = xdict["u"].reshape(-1, 2) * unitdict["u"]
    tab = {"event": event, "time": np.asarray(tx_res).round(6), "stage": stage, "section": sec,
           "mass": xdict["mass"] * unitdict["mass"], "vel_inertial": np.linalg.norm(vel, axis=1),
           "rate_BODY_X": np.zeros(len(tx_res)), "rate_BODY_Y": np.interp(tx_res, tu_res, u[:, 0]),
           "rate_BODY_Z": np.interp(tx_res, tu_res, u[:, 1])}
    for k, axis in enumerate("XYZ"):
        tab["pos_ECI_" + axis], tab["vel_ECI_" + axis] = pos[:, k], vel[:, k]
    for k in range(4):
        tab["quat_ECI2BODY_%d" % k] = quat[:, k]
    for j, name in enumerate(KERNEL_COLUMNS):
        tab[name] = rows[:, j]
    return {c: tab[c] for c in COLUMNS}


def output_result(xdict, unitdict, tx_res, tu_res, pdict, device=0, as_frame=True, _fn=None):
    """Drop-in for the reference's output_result: a pandas DataFrame with the reference's columns
    (as_frame=False: the dict of columns)."""
    sec, event, stage = row_sections(pdict, len(tx_res))
    rows = kernel_rows(_kernel_inputs(xdict, unitdict, tx_res, pdict, sec), pdict["wind_table"], pdict["ca_table"],
                       pdict["LaunchCondition"]["lat"], pdict["LaunchCondition"]["lon"], device=device, fn=_fn)
    tab = _assemble(xdict, unitdict, tx_res, tu_res, rows, sec, event, stage)
    if not as_frame:
        return tab
    import pandas as pd

    return pd.DataFrame(tab)


def output_tables(solutions, device=0):
    """Result tables of many solved scenarios with ONE kernel launch.  solutions: list of
    (xdict, unitdict, tx_res, tu_res, pdict) that share the wind / CA tables and launch site
    (dispersed scenarios with their own wind table go one call per distinct table).  Returns a list of
    column dicts."""
    metas, parts = [], []
    for xdict, unitdict, tx_res, tu_res, pdict in solutions:
        sec, event, stage = row_sections(pdict, len(tx_res))
        metas.append((sec, event, stage))
        parts.append(_kernel_inputs(xdict, unitdict, tx_res, pdict, sec))
    cat = [np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in range(8)]
    p0 = solutions[0][4]
    rows = kernel_rows(cat, p0["wind_table"], p0["ca_table"], p0["LaunchCondition"]["lat"], p0["LaunchCondition"]["lon"],
                       device=device)
    out, at = [], 0
    for (xdict, unitdict, tx_res, tu_res, pdict), (sec, event, stage) in zip(solutions, metas):
        n = len(tx_res)
        out.append(_assemble(xdict, unitdict, tx_res, tu_res, rows[at: at + n], sec, event, stage))
        at += n
    return out
