"""Shared helpers of the test-suite: problem fixtures, oracle construction and
dictionary comparison."""
import copy
import json
import os

import numpy as np

from gelato_b200 import plan as gplan
from gelato_b200 import problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INPUTS = os.path.join(GOLDEN, "example_inputs.json")
USER_EVENT = "IIP_END"


def example_inputs():
    """The shipped example's inputs (settings, events, wind / CA tables, initial
    trajectory) as frozen by tests/golden/make_golden.py."""
    return problem.load_inputs_json(INPUTS)


def example_problem(coord=None, factor=1, max_nodes=20, mutate=None):
    """(pdict, unitdict, condition, xdict0); factor > 1 refines the mesh (C2)."""
    inp = example_inputs()
    if mutate is not None:
        mutate(inp)
    return problem.problem_from_inputs(inp, coord=coord, factor=factor, max_nodes=max_nodes)


def oracle_nlp(p, u, c, flavour, dot, user=True):
    from oracle import leaves, nlp, user_builtin

    L = leaves.get(flavour)
    ue = user_builtin.perigee_ratio_at(L, USER_EVENT) if user else None
    return nlp.OracleNLP(p, u, c, flavour, dot, user_eq=ue)


def compiled_plan(p, u, c, coord=None, user=True):
    return gplan.CompiledPlan(p, u, c, user_eq=gplan.PerigeeAtEvent(USER_EVENT) if user else None, coord=coord)


def perturbed(xdict, seed=7, scale=1e-3):
    """x0 + scale*U(-1,1): every defect becomes O(scale)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in xdict.items():
        out[k] = v + scale * rng.uniform(-1.0, 1.0, v.shape) * (0.0 if k == "t" else 1.0)
    return out


def copy_x(xdict):
    return {k: np.array(v, dtype=np.float64, copy=True) for k, v in xdict.items()}


def flatten_funcs(f):
    """funcs dict -> {key: 1-D float array} (absent groups omitted)."""
    return {k: np.atleast_1d(np.asarray(v, dtype=np.float64)) for k, v in f.items() if v is not None}


def flatten_sens(s):
    """funcsSens dict -> {"group/var": (rows, cols, data, shape)} ; dense blocks get rows=cols=None."""
    out = {}
    for k, blk in s.items():
        if blk is None:
            continue
        for var, b in blk.items():
            if isinstance(b, dict):
                out["%s/%s" % (k, var)] = (b["coo"][0], b["coo"][1], np.asarray(b["coo"][2]), tuple(b["shape"]))
            else:
                out["%s/%s" % (k, var)] = (None, None, np.asarray(b, dtype=np.float64), np.shape(b))
    return out


def assert_funcs_equal(fa, fb):
    """Bit-exact comparison of two funcs dicts (None-ness included)."""
    assert list(fa.keys()) == list(fb.keys())
    for k in fa:
        a, b = fa[k], fb[k]
        assert (a is None) == (b is None), k
        if a is None:
            continue
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.array_equal(a, b), (k, float(np.max(np.abs(a - b))))


def assert_sens_equal(sa, sb):
    """Bit-exact comparison of two funcsSens dicts: same groups, same variable
    order, same (row, col) arrays and dtype, same shapes, same values."""
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        a, b = sa[k], sb[k]
        assert (a is None) == (b is None), k
        if a is None:
            continue
        assert list(a.keys()) == list(b.keys()), (k, list(a.keys()), list(b.keys()))
        for var in a:
            av, bv = a[var], b[var]
            if isinstance(av, dict):
                assert tuple(av["shape"]) == tuple(bv["shape"]), (k, var)
                for i in range(2):
                    assert av["coo"][i].dtype == np.int32 and bv["coo"][i].dtype == np.int32, (k, var)
                    assert np.array_equal(av["coo"][i], bv["coo"][i]), (k, var, "index %d" % i)
                da, db = np.asarray(av["coo"][2]), np.asarray(bv["coo"][2])
                assert da.shape == db.shape, (k, var)
                assert np.array_equal(da, db), (k, var, int(np.count_nonzero(da - db)), float(np.max(np.abs(da - db))))
            else:
                assert np.shape(av) == np.shape(bv), (k, var)
                assert np.array_equal(av, bv), (k, var, float(np.max(np.abs(np.asarray(av) - np.asarray(bv)))))


def assert_sens_within_noise(s, npz, name, floor_factor=2.0):
    """funcsSens `s` against the reference fixture: identical sparsity; values within
    1e-10 relative + floor_factor x the block's finite-difference noise floor
    (`jn` arrays of tests/golden/example_reference.npz)."""
    for k, (r, c, d, shape) in flatten_sens(s).items():
        ref = npz["%s/j/%s/data" % (name, k)]
        if r is not None:
            assert r.dtype == np.int32 and c.dtype == np.int32, k
            assert np.array_equal(r, npz["%s/j/%s/rows" % (name, k)]), k
            assert np.array_equal(c, npz["%s/j/%s/cols" % (name, k)]), k
        assert tuple(shape) == tuple(npz["%s/j/%s/shape" % (name, k)].tolist()), k
        jn = npz["%s/jn/%s" % (name, k)]
        floor = floor_factor * float(jn.max()) if jn.size else 0.0
        np.testing.assert_allclose(d, ref, rtol=1e-10, atol=floor, err_msg=k)


def variant_inputs(name):
    """Problem variants exercising the constraint branches the shipped example leaves cold."""
    inp = copy.deepcopy(example_inputs())
    s = inp["settings"]
    fc = s["FlightConstraint"]
    if name == "example":
        pass
    elif name == "fuel_inclination":  # non-Payload objective, inclination row, radius-style target
        s["OptimizationMode"] = "Fuel"
        s["TerminalCondition"]["inclination"] = 42.2
        s["TerminalCondition"]["altitude_perigee"] = None
    elif name == "all_aero":  # max-q rows, alpha over a whole section, several Q-alpha sections
        fc["dynamic_pressure_max"] = {"KICKTURN": {"value": 60000.0, "range": "all"},
                                      "ZEROLIFT_END": {"value": 50000.0, "range": "initial"}}
        fc["AOA_max"]["ZEROLIFT_START"] = {"value": 5.0, "range": "all"}
        fc["Q_alpha_max"]["KICKTURN"] = {"value": 40000.0, "range": "initial"}
    elif name == "waypoints":  # lat/lon/altitude and IIP rows, equality and both inequality bounds, 2 antennas
        fc["waypoint"] = {
            "FAIRING": {"altitude": {"exact": 100000.0, "min": 90000.0, "max": 120000.0},
                        "lat": {"min": 40.0, "max": 45.0}, "lon": {"exact": 146.0},
                        "lon_IIP": {"min": 145.0, "max": 170.0}, "lat_IIP": {"exact": 40.0}},
            "SEIG": {"lat_IIP": {"min": 30.0}, "lon": {"max": 150.0}},
            "IIP_END": {"altitude": {"min": 150000.0}, "lon_IIP": {"max": 200.0}},
        }
        fc["antenna"]["ANT2"] = {"lon": 145.0, "lat": 40.0, "altitude": 10.0,
                                 "elevation_min": {"SEIG": 2.0, "FAIRING": 1.0, "SECO": 0.5}}
    elif name == "neg_area":  # stage 2 with a NEGATIVE reference area: the reference takes the air formula
        # (reference_area != 0.0, con_dynamics.py:257) but the vacuum Jacobian branches (> 0.0, :403,454)
        s["RocketStage"]["2"]["reference_area"] = -0.5
    elif name == "three_stage":  # BASELINE.json configs[2]: a third stage (coast, burn, coast) after SEP2
        s["RocketStage"]["3"] = {"mass_dry": 150.0, "mass_propellant": 600.0, "dropMass": {}, "Isp_vac": 320.0,
                                 "reference_area": 0.0, "ignition_at": "TEIG", "cutoff_at": "TECO",
                                 "separation_at": "SEP3"}
        ev = inp["events"]
        base = dict(ev[-1])
        ev[-2]["rocketStage"] = 3  # SEP2 starts the third stage's coast
        ev.pop()  # SIMEND is re-appended last
        def event(name, time, ref, on, thrust, att, nodes, pr=0.0):
            e = dict(base)
            e.update(name=name, time=time, time_ref=ref, rocketStage=3, engineOn=on, thrust=thrust, attitude=att,
                     pitchrate_init=pr, yawrate_init=0.0, num_nodes=nodes)
            return e
        ev += [event("TEIG", 640.0, "SEP2", True, 5000.0, "pitch-yaw", 7, -0.02),
               event("TECO", 700.0, float("nan"), False, 0.0, "hold", 3),
               event("SEP3", 720.0, "TECO", False, 0.0, "hold", 2),
               event("SIMEND", 725.0, "SEP3", False, 0.0, "hold", 2)]
    elif name == "iip_orbital":  # IIP rows where there is no impact point (orbital speed): the leaf returns
        # zeros (iip.cpp:49-128), the rows become constants and their finite differences exact zeros
        fc["waypoint"] = {"SECO": {"lat_IIP": {"min": -10.0}, "lon_IIP": {"exact": 150.0}},
                          "SEP2": {"lon_IIP": {"max": 190.0}},
                          "KICKTURN": {"lat_IIP": {"max": 60.0}, "lon_IIP": {"min": 100.0}}}
    elif name == "bare":  # no waypoint / antenna blocks, no aero rows, no user constraint rows
        fc.pop("waypoint")
        fc.pop("antenna")
        fc["AOA_max"] = {}
        fc["Q_alpha_max"] = {}
    else:
        raise KeyError(name)
    return inp
