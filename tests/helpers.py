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


def row_noise_floor(rows, jn):
    """Largest `jn` per constraint row of one block, broadcast back to the block's slots."""
    jn = np.asarray(jn, dtype=np.float64)
    if rows is None or jn.size == 0:
        return np.full(jn.shape, float(jn.max()) if jn.size else 0.0)
    rows = np.asarray(rows)
    top = np.zeros(int(rows.max()) + 1)
    np.maximum.at(top, rows, jn)
    return top[rows]


def group_noise_floors(npz, name):
    """Per-slot finite-difference noise floor of every Jacobian block of a reference fixture: the largest
    `jn` over the slots of the SAME constraint row, over all variable blocks of the group.  One row = one
    leaf value f_c and one scale, so all its slots share the quantum ulp(f)/dx x scale by which a last-bit
    difference in f moves any of them; a single slot's own `jn` (the reference's value re-evaluated with
    inputs one ulp away, 64 draws) can miss that quantum when the nudges happen not to flip its last bit,
    its row neighbours do not.  Returns {"group/var": floor per slot}."""
    pre = "%s/jn/" % name
    keys = [k[len(pre):] for k in npz.files if k.startswith(pre)]
    top = {}
    for k in keys:
        g = k.split("/")[0]
        jn = npz[pre + k].ravel()
        rk = "%s/j/%s/rows" % (name, k)
        rows = npz[rk].astype(np.int64) if rk in npz.files else np.zeros(jn.size, dtype=np.int64)
        if jn.size == 0:
            continue
        t = top.setdefault(g, np.zeros(0))
        if t.size <= rows.max():
            t = np.concatenate((t, np.zeros(int(rows.max()) + 1 - t.size)))
        np.maximum.at(t, rows, jn)
        top[g] = t
    out = {}
    for k in keys:
        g = k.split("/")[0]
        jn = npz[pre + k]
        rk = "%s/j/%s/rows" % (name, k)
        if jn.size == 0:
            out[k] = np.zeros(jn.shape)
        elif rk in npz.files:
            out[k] = top[g][npz[rk].astype(np.int64)]
        else:
            out[k] = np.full(jn.shape, top[g][0])
    return out


def assert_sens_within_noise(s, npz, name, floor_factor=2.0):
    """funcsSens `s` against the reference fixture: identical sparsity; every value within 1e-10 relative
    + floor_factor x the finite-difference noise floor of ITS OWN ROW (`group_noise_floors`).  Slots of rows
    without finite differences have a zero floor and must meet 1e-10 outright."""
    floors = group_noise_floors(npz, name)
    for k, (r, c, d, shape) in flatten_sens(s).items():
        ref = npz["%s/j/%s/data" % (name, k)]
        if r is not None:
            assert r.dtype == np.int32 and c.dtype == np.int32, k
            assert np.array_equal(r, npz["%s/j/%s/rows" % (name, k)]), k
            assert np.array_equal(c, npz["%s/j/%s/cols" % (name, k)]), k
        assert tuple(shape) == tuple(npz["%s/j/%s/shape" % (name, k)].tolist()), k
        floor = floor_factor * floors[k].reshape(np.shape(ref))
        err = np.abs(np.asarray(d, dtype=np.float64) - ref)
        bad = err > 1e-10 * np.abs(ref) + floor
        assert not bad.any(), (k, int(bad.sum()), float(err[bad].max()))


def true_jacobian(objfunc, xdict, h=1e-6):
    """Dense d(funcs)/d(x) of a callback by 4th-order central differences (Richardson): truncation ~h^4,
    rounding ~1e-16/h (h = 1e-6: 6 m / 1 mm/s, well inside one interval of the wind, drag and atmosphere
    tables) -- two orders closer to the derivative than the reference's forward difference with
    dx = 1e-8, which makes it the yardstick for "how far from the truth" both the
    reference's and the kernels' Jacobian values are.  Returns ({group: row offset}, J[rows, n_vars],
    {var: column offset})."""
    keys = list(xdict.keys())
    col0, o = {}, 0
    for k in keys:
        col0[k] = o
        o += xdict[k].size
    nvar = o

    def fvec(x):
        f, _ = objfunc({k: v.copy() for k, v in x.items()})
        return np.concatenate([np.atleast_1d(np.asarray(f[g], dtype=np.float64)) for g in f if g != "obj" and f[g] is not None])

    f0, _ = objfunc({k: v.copy() for k, v in xdict.items()})
    row0, r = {}, 0
    for g, v in f0.items():
        if g != "obj" and v is not None:
            row0[g] = r
            r += np.atleast_1d(np.asarray(v)).size
    J = np.zeros((r, nvar))
    for k in keys:
        for i in range(xdict[k].size):
            vals = []
            for step in (h, -h, 2 * h, -2 * h):
                x = {kk: vv.copy() for kk, vv in xdict.items()}
                x[k][i] += step
                vals.append(fvec(x))
            J[:, col0[k] + i] = (8.0 * (vals[0] - vals[1]) - (vals[2] - vals[3])) / (12.0 * h)
    return row0, J, col0


def derivative_errors(s, row0, J, col0):
    """{"group/var": |value - true derivative| per COO slot} for the sparse blocks of a funcsSens dict."""
    out = {}
    for key, (r, c, d, shape) in flatten_sens(s).items():
        if r is None:
            continue
        g, var = key.split("/")
        out[key] = np.abs(d - J[row0[g] + r.astype(np.int64), col0[var] + c.astype(np.int64)])
    return out


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
        problem.three_stage_inputs(inp)
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


def assert_as_close_to_truth_as_reference(s, npz, name, row0, J, col0, factor=1.5):
    """The statement that matters to the NLP solver: block by block, the Jacobian values in `s` are as close
    to the TRUE derivative (`true_jacobian`) as the reference's own values (fixture `npz`) are -- maximum and
    root-mean-square error at most `factor` x the reference's, plus twice the block's finite-difference noise
    floor (what separates two equally valid last-bit roundings of one quotient)."""
    eg = derivative_errors(s, row0, J, col0)
    fs = flatten_sens(s)
    floors = group_noise_floors(npz, name)
    report = {}
    for key, e in eg.items():
        r, c, _, _ = fs[key]
        g, var = key.split("/")
        ref = npz["%s/j/%s/data" % (name, key)]
        er = np.abs(ref - J[row0[g] + r.astype(np.int64), col0[var] + c.astype(np.int64)])
        floor = floors[key]
        if e.size == 0:
            continue
        rms = lambda a: float(np.sqrt(np.mean(np.square(a))))  # noqa: E731
        report[key] = (float(e.max()), float(er.max()), rms(e), rms(er))
        assert e.max() <= factor * er.max() + 2.0 * floor.max(), (key, report[key])
        assert rms(e) <= factor * rms(er) + 2.0 * rms(floor), (key, report[key])
    return report
