"""The reference's forward-simulation initial guess (/root/reference/initialize.py:114-179 `rocket_simulation`,
:238-319 `initialize_xdict_6DoF_2`) with the integration on the GPU, ONE THREAD PER SCENARIO.

The reference integrates the 3-DoF equations of motion in a Python loop (its dt = 0.005 s: ~126 000 Runge-Kutta
steps, ~half a million right-hand sides of ~10 pybind leaf calls each) once per settings file.  A dispersed study
needs one initial guess per scenario; `gelato_init_rocket_simulation` (include/gelato_b200.h) runs them all in one
launch.  The host keeps what is bookkeeping in the reference too: the mesh times, the rate table, the scaling into
xdict.

Same argument meaning as the reference's functions; `flag_display` is accepted and ignored (plotting is not part of
this path).  `condition` is accepted for signature compatibility: the reference reads `condition["rf_m"]` only to
print it (:279-285).
"""
import ctypes

import numpy as np

from . import engine as _engine

_pd = ctypes.POINTER(ctypes.c_double)
EVENT_COLUMNS = ("time", "thrust", "massflow", "reference_area", "nozzle_area", "mass_jettison")


def event_rows(pdict):
    """[n_ev][6] rows of the kernel's event table and the zero-lift-turn flags from pdict["params"]."""
    prm = pdict["params"]
    rows = np.array([[float(e[c]) for c in EVENT_COLUMNS] for e in prm], dtype=np.float64)
    zlt = np.array([1 if e["attitude"] == "zero-lift-turn" else 0 for e in prm], dtype=np.int32)
    return rows, zlt


def _stack(arrays, width):
    """(contiguous array, scenario stride in doubles): one copy shared (stride 0) or one per scenario."""
    if isinstance(arrays, np.ndarray) and arrays.ndim == width:
        a = np.ascontiguousarray(arrays, dtype=np.float64)
        return a, 0
    a = np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in arrays]))
    return a, int(np.prod(a.shape[1:]))


def rocket_simulation_batch(x_inits, u_table, pdicts, t_init, t_out, dt=0.1, device=0, fn=None):
    """`rocket_simulation` for len(pdicts) scenarios sharing the event schedule's shape, the rate table and t_out.
    x_inits: one 11-vector or one per scenario.  Returns (x_out [n][n_out][11], u_out [n][n_out][3])."""
    n = len(pdicts)
    ev_z = [event_rows(p) for p in pdicts]
    zlt = ev_z[0][1]
    for _, z in ev_z[1:]:
        if not np.array_equal(z, zlt):
            raise ValueError("the scenarios of one batch must share the attitude modes of their events")
    x_inits = np.asarray(x_inits, dtype=np.float64)
    x0, sx = (np.ascontiguousarray(x_inits), 0) if x_inits.ndim == 1 else (np.ascontiguousarray(x_inits), 11)
    if x0.shape[-1] != 11 or (sx and x0.shape[0] != n):
        raise ValueError("x_init must hold 11 values (one vector, or one per scenario)")
    same = lambda key: all(p[key] is pdicts[0][key] for p in pdicts)  # noqa: E731
    ev, se = _stack([e for e, _ in ev_z], 2)
    wind, sw = _stack(pdicts[0]["wind_table"] if same("wind_table") else [p["wind_table"] for p in pdicts], 2)
    ca, sc = _stack(pdicts[0]["ca_table"] if same("ca_table") else [p["ca_table"] for p in pdicts], 2)
    n_ev, n_wind, n_ca = zlt.size, wind.shape[-2], ca.shape[-2]
    u_table = np.ascontiguousarray(u_table, dtype=np.float64)
    t_out = np.ascontiguousarray(np.atleast_1d(np.asarray(t_out, dtype=np.float64)))
    strides = np.array([sx, se, sw, sc], dtype=np.int64)
    x_out = np.empty((n, t_out.size, 11))
    u_out = np.empty((n, t_out.size, 3))
    args = (n, x0.ctypes.data_as(_pd), ev.ctypes.data_as(_pd), zlt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n_ev,
            u_table.ctypes.data_as(_pd), u_table.shape[0], wind.ctypes.data_as(_pd), n_wind, ca.ctypes.data_as(_pd), n_ca,
            strides.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_double(float(t_init)), t_out.ctypes.data_as(_pd),
            t_out.size, ctypes.c_double(float(dt)), x_out.ctypes.data_as(_pd), u_out.ctypes.data_as(_pd))
    if fn is None:
        L = _engine.load_library()
        rc = L.gelato_init_rocket_simulation(device, *args)
        if rc != 0:
            raise _engine.GelatoError("gelato_init_rocket_simulation failed (%d): %s" % (rc, L.gelato_last_error().decode()))
    else:  # test hook: the same per-thread function stepped on the host (tests/emu)
        fn(*args)
    return x_out, u_out


def rocket_simulation(x_init, u_table, pdict, t_init, t_out, dt=0.1, device=0, fn=None):
    """initialize.py:114-179: returns (x_out [n_out][11], u_out [n_out][3]).  Unlike the reference it leaves the
    caller's x_init untouched (the reference applies the first jettison to it in place)."""
    x_out, u_out = rocket_simulation_batch(np.asarray(x_init, dtype=np.float64), u_table, [pdict], t_init, t_out, dt, device, fn)
    return x_out[0], u_out[0]


def mesh_times(pdict, mode="LGR"):
    """(time_nodes, time_x_nodes) of the mesh (initialize.py:259-275)."""
    ps = pdict["ps_params"]
    time_nodes, time_x_nodes = np.array([]), np.array([])
    for i in range(pdict["num_sections"]):
        to, tf = pdict["params"][i]["time"], pdict["params"][i]["timeFinishAt"]
        tau = ps.tau(i)
        tau_x = np.hstack((-1.0, tau)) if mode in ("LG", "LGR") else tau
        time_nodes = np.hstack((time_nodes, tau * (tf - to) / 2.0 + (tf + to) / 2.0))
        time_x_nodes = np.hstack((time_x_nodes, tau_x * (tf - to) / 2.0 + (tf + to) / 2.0))
    return time_nodes, time_x_nodes


def rate_table(pdict, time_nodes):
    """(u_nodes [N][2], u_table [N][4]) from the events' initial pitch / yaw rates (initialize.py:287-302)."""
    ps, prm = pdict["ps_params"], pdict["params"]
    u_nodes = np.vstack([[[prm[i]["pitchrate_init"], prm[i]["yawrate_init"]]] * ps.nodes(i)
                         for i in range(pdict["num_sections"])])
    u_table = np.hstack((time_nodes.reshape(-1, 1), np.column_stack((np.zeros(len(u_nodes)), u_nodes))))
    return u_nodes, u_table


def _interp1d_extrapolate(x, y, x_new):
    """scipy.interpolate.interp1d(x, y, axis=0, fill_value="extrapolate")(x_new) for sorted x, linear kind: the
    two-weight form of scipy >= 1.10 (`_call_linear`: searchsorted, clip(1, n-1),
    (x_new - x_lo)/(x_hi - x_lo) * y_hi + (x_hi - x_new)/(x_hi - x_lo) * y_lo)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    y2 = y.reshape(y.shape[0], -1)
    hi = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(int)
    lo = hi - 1
    x_lo, x_hi = x[lo], x[hi]
    out = ((x_new - x_lo) / (x_hi - x_lo))[:, None] * y2[hi] + ((x_hi - x_new) / (x_hi - x_lo))[:, None] * y2[lo]
    return out.reshape((len(x_new),) + y.shape[1:])


def initialize_xdict_6DoF_from_file(x_ref, pdict, condition, unitdict, mode="LGL", flag_display=False):
    """initialize.py:322-413: the initial xdict by interpolating a reference trajectory table (a DataFrame or any
    mapping of the columns time, mass, pos_ECI_X/Y/Z, vel_ECI_X/Y/Z, quat_ECI2BODY_0..3, rate_BODY_Y/Z) onto the
    mesh.  Host code, as in the reference (it runs once per solve); same values as the reference's function with
    the SciPy of this image, bit for bit (tests/test_initguess.py)."""
    time_nodes, time_x_nodes = mesh_times(pdict, mode)
    t_ref = np.asarray(x_ref["time"], dtype=np.float64)

    def cols(names):
        return np.column_stack([np.asarray(x_ref[n], dtype=np.float64) for n in names])

    xdict = {"t": (np.array([e["time"] for e in pdict["params"]]) / unitdict["t"]).ravel()}
    xdict["mass"] = (_interp1d_extrapolate(t_ref, np.asarray(x_ref["mass"], dtype=np.float64), time_x_nodes) / unitdict["mass"]).ravel()
    xdict["position"] = (_interp1d_extrapolate(t_ref, cols(["pos_ECI_X", "pos_ECI_Y", "pos_ECI_Z"]), time_x_nodes)
                         / unitdict["position"]).ravel()
    xdict["velocity"] = (_interp1d_extrapolate(t_ref, cols(["vel_ECI_X", "vel_ECI_Y", "vel_ECI_Z"]), time_x_nodes)
                         / unitdict["velocity"]).ravel()
    xdict["quaternion"] = _interp1d_extrapolate(
        t_ref, cols(["quat_ECI2BODY_0", "quat_ECI2BODY_1", "quat_ECI2BODY_2", "quat_ECI2BODY_3"]), time_x_nodes).ravel()
    xdict["u"] = (_interp1d_extrapolate(t_ref, cols(["rate_BODY_Y", "rate_BODY_Z"]), time_nodes) / unitdict["u"]).ravel()
    return xdict


def _xdict_from_nodes(pdict, unitdict, u_nodes, x_nodes):
    xdict = {"t": (np.array([e["time"] for e in pdict["params"]]) / unitdict["t"]).ravel(),
             "u": (u_nodes / unitdict["u"]).ravel(),
             "mass": x_nodes[:, 0] / unitdict["mass"],
             "position": (x_nodes[:, 1:4] / unitdict["position"]).ravel(),
             "velocity": (x_nodes[:, 4:7] / unitdict["velocity"]).ravel(),
             "quaternion": (x_nodes[:, 7:11]).ravel()}
    return xdict


def initialize_xdict_6DoF_2(x_init, pdict, condition, unitdict, mode="LGR", dt=0.005, flag_display=False, device=0, fn=None):
    """initialize.py:238-319: the NLP's initial xdict from one forward simulation."""
    return initialize_xdict_batch(x_init, [pdict], unitdict, mode, dt, device, fn)[0]


def initialize_xdict_batch(x_inits, pdicts, unitdict, mode="LGR", dt=0.005, device=0, fn=None):
    """One initial xdict per scenario from ONE launch.  The scenarios share the mesh (sections, nodes, event times and
    initial rates -- what `gelato_b200.scenarios` keeps fixed) and differ in x_init, event parameters and tables."""
    time_nodes, time_x_nodes = mesh_times(pdicts[0], mode)
    u_nodes, u_table = rate_table(pdicts[0], time_nodes)
    x_nodes, _ = rocket_simulation_batch(x_inits, u_table, pdicts, time_nodes[0], time_x_nodes, dt, device, fn)
    return [_xdict_from_nodes(p, unitdict, u_nodes, x_nodes[k]) for k, p in enumerate(pdicts)]
