"""Drive the REAL reference Python layer (/root/reference/lib/con_*.py and the
objfunc/sens bodies of /root/reference/Trajectory_Optimization.py) on top of the
oracle's physics leaves.  Only usable where /root/reference exists (this
container); the GPU box uses the golden fixtures this produces instead.

The reference's five pybind11 modules cannot be built here (Eigen3 is absent,
/root/reference/CMakeLists.txt:13), so `lib.dynamics_c` & co. are satisfied by
oracle/leaves.py namespaces registered in sys.modules before `lib` is imported.
Nothing is copied: the reference files are imported / exec'd where they lie.
"""
import importlib
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "lib"))


def _inject(leaves):
    for name, ns in leaves.modules().items():
        mod = types.ModuleType("lib." + name)
        mod.__dict__.update(vars(ns))
        mod.__all__ = list(vars(ns).keys())
        sys.modules["lib." + name] = mod


def load(leaves):
    """Returns a namespace dict holding the reference's con_* modules."""
    assert available()
    for p in (REF, os.path.join(REF, "example")):
        if p not in sys.path:
            sys.path.insert(0, p)
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.") or k == "user_constraints"]:
        del sys.modules[k]
    _inject(leaves)
    ns = {}
    ns["con_a"] = importlib.import_module("lib.con_init_terminal_knot")
    ns["con_traj"] = importlib.import_module("lib.con_trajectory")
    ns["con_aero"] = importlib.import_module("lib.con_aero")
    ns["con_dynamics"] = importlib.import_module("lib.con_dynamics")
    ns["con_wp"] = importlib.import_module("lib.con_waypoint")
    ns["con_user"] = importlib.import_module("lib.con_user")
    cg = importlib.import_module("lib.cost_gradient")
    ns["cost_6DoF"], ns["cost_jac"] = cg.cost_6DoF, cg.cost_jac
    uc = importlib.import_module("user_constraints")
    ns["equality_user"], ns["inequality_user"] = uc.equality_user, uc.inequality_user
    ns["PSparams"] = importlib.import_module("lib.SectionParameters").PSparams
    return ns


def _slice_source(start_marker, end_marker, include_end=True):
    src = open(os.path.join(REF, "Trajectory_Optimization.py")).read().split("\n")
    a = next(i for i, l in enumerate(src) if l.startswith(start_marker))
    b = next(i for i, l in enumerate(src) if i > a and l.startswith(end_marker))
    return "\n".join(src[a : b + (1 if include_end else 0)])


def reference_setup(leaves, settings_path=None):
    """exec the reference's own set-up block (Trajectory_Optimization.py:55-177)
    -> (pdict, unitdict, condition).  Needs pandas (present here)."""
    import json

    import numpy as np
    import pandas as pd

    settings_path = settings_path or os.path.join(REF, "example", "example-settings.json")
    ns = load(leaves)
    g = {"np": np, "pd": pd, "PSparams": ns["PSparams"], "sys": sys}
    g.update(vars(sys.modules["lib.coordinate_c"]))
    with open(settings_path) as f:
        g["settings"] = json.load(f)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(settings_path))
    try:
        exec(_slice_source("wind = pd.read_csv", 'condition["OptimizationMode"]'), g)
    finally:
        os.chdir(cwd)
    return g["pdict"], g["unitdict"], g["condition"]


def reference_callbacks(leaves, pdict, unitdict, condition):
    """exec the reference's `objfunc` and `sens` definitions
    (Trajectory_Optimization.py:194-312) against the given problem."""
    ns = load(leaves)
    g = dict(ns)
    g.update(pdict=pdict, unitdict=unitdict, condition=condition)
    exec(_slice_source("def objfunc(xdict):", "optProb = Optimization", include_end=False), g)
    return g["objfunc"], g["sens"]


def reference_output_result(leaves):
    """The reference's own output_result function (/root/reference/output_result.py:37-263), imported
    where it lies on top of the given leaves."""
    load(leaves)
    for k in [k for k in sys.modules if k == "output_result"]:
        del sys.modules[k]
    return importlib.import_module("output_result").output_result


def result_times(xdict, pdict, unitdict):
    """tx_res / tu_res as the reference's driver builds them (Trajectory_Optimization.py:477-492)."""
    import numpy as np

    tu, tx = np.array([]), np.array([])
    ps = pdict["ps_params"]
    for i in range(pdict["num_sections"]):
        to, tf = xdict["t"][i], xdict["t"][i + 1]
        tau_x = np.hstack((-1.0, ps.tau(i)))
        tu = np.hstack((tu, (ps.tau(i) * (tf - to) / 2 + (tf + to) / 2) * unitdict["t"]))
        tx = np.hstack((tx, (tau_x * (tf - to) / 2 + (tf + to) / 2) * unitdict["t"]))
    return tx, tu


def reference_initialize(leaves):
    """The reference's own initialize.py (forward-simulation initial guess), imported where it lies.  Its plotting
    imports (tools.plot_output, output_result -> matplotlib / pandas plotting) are replaced by empty stand-ins, and
    `norm` / `sqrt` -- which initialize.py expects from `from lib.utils_c import *` but the pybind module does not
    export (the legacy lib/utils.py did) -- are supplied with their legacy meaning (numpy.linalg.norm, math.sqrt)."""
    import math

    import numpy as np

    load(leaves)
    sys.modules["lib.utils_c"].__dict__.update(norm=np.linalg.norm, sqrt=math.sqrt)
    sys.modules["lib.utils_c"].__all__ = list(sys.modules["lib.utils_c"].__dict__.keys())
    for name in ("tools", "tools.plot_output", "output_result"):
        mod = types.ModuleType(name)
        mod.display_6DoF = mod.output_result = lambda *a, **k: None
        sys.modules[name] = mod
    for k in [k for k in sys.modules if k == "initialize"]:
        del sys.modules[k]
    return importlib.import_module("initialize")
