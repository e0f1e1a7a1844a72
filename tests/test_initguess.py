"""The forward-simulation initial guess (/root/reference/initialize.py:37-319): the oracle's
restatement (oracle/initguess.py) live against the reference's own functions, and the batched CUDA kernel against the
oracle's gmath flavour bit for bit."""
import numpy as np
import pytest

import helpers
import refharness
from gelato_b200 import problem
from oracle import initguess, leaves


def _case(coord=None, dt=0.5):
    p, u, c, _ = helpers.example_problem(coord=coord)
    x_init = np.concatenate(([c["init"]["mass"]], c["init"]["position"], c["init"]["velocity"], c["init"]["quaternion"]))
    return p, u, c, x_init, dt


@pytest.mark.skipif(not refharness.available(), reason="/root/reference not present on this machine")
def test_oracle_forward_simulation_matches_the_reference():
    """rocket_simulation / zerolift_turn_correct / dynamics_init of the reference (imported where they lie, on the
    reference-faithful leaves) against the oracle's restatement: the node states of the shipped example's mesh from
    a 0.5 s Runge-Kutta integration over all 13 events (stage separations and the zero-lift turn included)."""
    L = leaves.get("libm")
    ini = refharness.reference_initialize(L)
    p, u, c, x_init, dt = _case()
    F = initguess.ForwardSimulation(L, "numpy")
    xd, u_table, t_x = F.initialize_xdict(x_init.copy(), p, u, dt)
    x_ref, u_ref = ini.rocket_simulation(x_init.copy(), u_table, p, u_table[0, 0], t_x, dt)
    assert np.array_equal(xd["mass"] * u["mass"], x_ref[:, 0] / u["mass"] * u["mass"])
    assert np.array_equal(xd["position"], (x_ref[:, 1:4] / u["position"]).ravel())
    assert np.array_equal(xd["velocity"], (x_ref[:, 4:7] / u["velocity"]).ravel())
    assert np.array_equal(xd["quaternion"], x_ref[:, 7:11].ravel())
    # single pieces on a mid-flight state
    x = x_ref[30].copy()
    prm = np.array([p["params"][3]["thrust"], p["params"][3]["massflow"], p["params"][3]["reference_area"], 0.0,
                    p["params"][3]["nozzle_area"]])
    uu = np.array([0.0, -0.3, 0.05])
    assert np.array_equal(ini.dynamics_init(x, uu, 55.0, prm, False, p["wind_table"], p["ca_table"]),
                          F.dynamics_init(x, uu, 55.0, prm, False, p["wind_table"], p["ca_table"]))
    assert np.array_equal(ini.zerolift_turn_correct(x, 55.0, p["wind_table"]), F.zerolift_turn_correct(x, 55.0, p["wind_table"]))


def _emu_fn():
    import ctypes

    import emu_binding
    emu_binding.build()
    return ctypes.CDLL(emu_binding.LIB).emu_rocket_simulation


def _dispersed(p, k):
    """scenario k of a small dispersed batch: thrust, mass flow, initial mass and winds perturbed"""
    import copy
    q = dict(p)
    q["params"] = copy.deepcopy(p["params"])
    for e in q["params"]:
        e["thrust"] *= 1.0 + 0.01 * k
        e["massflow"] *= 1.0 - 0.005 * k
    q["wind_table"] = p["wind_table"] * np.array([1.0, 1.0 + 0.1 * k, 1.0 - 0.2 * k])
    return q


def test_kernel_thread_function_matches_the_oracle_bit_for_bit():
    """The kernel's per-thread function (initguess.h, stepped on the host) against the oracle on the gmath leaves with
    numpy's norm stated as sequential FMAs: identical node states and rate history for a 3-scenario dispersed batch,
    and the drop-in initialize_xdict_6DoF_2 gives the oracle's xdict."""
    from gelato_b200 import initialize
    L = leaves.get("gmath")
    F = initguess.ForwardSimulation(L, "seqfma")
    p, u, c, x_init, dt = _case()
    fn = _emu_fn()
    pd = [_dispersed(p, k) for k in range(3)]
    x0 = np.stack([x_init * np.r_[1.0 + 0.002 * k, np.ones(10)] for k in range(3)])
    t_nodes, t_x = initialize.mesh_times(p)
    _, u_table = initialize.rate_table(p, t_nodes)
    u_table[:, 2] = -0.2 + 0.001 * np.arange(len(u_table))  # a rate history that exercises the interpolation
    x_out, u_out = initialize.rocket_simulation_batch(x0, u_table, pd, t_nodes[0], t_x, dt, fn=fn)
    for k in range(3):
        xo, uo = F.rocket_simulation(x0[k].copy(), u_table, pd[k], t_nodes[0], t_x, dt)
        assert np.array_equal(x_out[k], xo), k
        assert np.array_equal(u_out[k], uo), k
    assert not np.array_equal(x_out[0], x_out[2])
    xd = initialize.initialize_xdict_6DoF_2(x_init, p, c, u, dt=dt, fn=fn)
    xo, _, _ = F.initialize_xdict(x_init.copy(), p, u, dt)
    assert set(xd) == set(xo)
    for key in xo:
        assert np.array_equal(xd[key], xo[key]), key


def test_forward_simulation_edge_cases():
    """Output times before the start, on a recorded time, past the end; a single output time; dt that does not divide
    the span."""
    from gelato_b200 import initialize
    L = leaves.get("gmath")
    F = initguess.ForwardSimulation(L, "seqfma")
    p, u, c, x_init, _ = _case()
    fn = _emu_fn()
    t_nodes, _ = initialize.mesh_times(p)
    _, u_table = initialize.rate_table(p, t_nodes)
    for t_out, dt in ((np.array([-1.0, 0.0, 0.7, 1.4, 2.1, 35.0, 90.0]), 0.7), (np.array([12.5]), 0.3),
                      (np.array([5.0, 5.0, 6.0]), 1.0)):
        xo, uo = F.rocket_simulation(x_init.copy(), u_table, p, 0.0, t_out, dt)
        x_out, u_out = initialize.rocket_simulation(x_init, u_table, p, 0.0, t_out, dt, fn=fn)
        assert np.array_equal(x_out, xo) and np.array_equal(u_out, uo), (t_out, dt)


@pytest.mark.gpu
def test_forward_simulation_kernel_on_the_gpu():
    """gelato_init_rocket_simulation through the C-ABI: a 64-scenario dispersed batch in one launch equals the oracle
    (three scenarios checked in full), and equals itself scenario by scenario."""
    from gelato_b200 import initialize
    L = leaves.get("gmath")
    F = initguess.ForwardSimulation(L, "seqfma")
    p, u, c, x_init, dt = _case()
    pd = [_dispersed(p, k % 7) for k in range(64)]
    x0 = np.stack([x_init * np.r_[1.0 + 0.002 * (k % 5), np.ones(10)] for k in range(64)])
    t_nodes, t_x = initialize.mesh_times(p)
    _, u_table = initialize.rate_table(p, t_nodes)
    x_out, u_out = initialize.rocket_simulation_batch(x0, u_table, pd, t_nodes[0], t_x, dt)
    for k in (0, 9, 63):
        xo, uo = F.rocket_simulation(x0[k].copy(), u_table, pd[k], t_nodes[0], t_x, dt)
        assert np.array_equal(x_out[k], xo), k
        assert np.array_equal(u_out[k], uo), k
    one, _ = initialize.rocket_simulation(x0[9], u_table, pd[9], t_nodes[0], t_x, dt)
    assert np.array_equal(one, x_out[9])
    xd = initialize.initialize_xdict_batch(x0, pd, u, dt=dt)
    assert len(xd) == 64 and np.array_equal(xd[9]["mass"], x_out[9][:, 0] / u["mass"])


@pytest.mark.skipif(not refharness.available(), reason="/root/reference not present on this machine")
def test_file_based_initial_guess_matches_the_reference():
    """initialize_xdict_6DoF_from_file (interpolation of the shipped trajectory table onto the mesh) against the
    reference's own function, with the SciPy of this image: every array bit for bit, LGR and LGL meshes."""
    import os

    import pandas as pd

    from gelato_b200 import initialize
    ini = refharness.reference_initialize(leaves.get("libm"))
    p, u, c, _ = helpers.example_problem()
    x_ref = pd.read_csv(os.path.join(refharness.REF, "example", "example-trajectory_init.csv"))
    for mode in ("LGR", "LGL"):
        want = ini.initialize_xdict_6DoF_from_file(x_ref, p, c, u, mode, False)
        got = initialize.initialize_xdict_6DoF_from_file(x_ref, p, c, u, mode, False)
        assert list(got) == list(want)
        for k in want:
            assert np.array_equal(np.asarray(want[k]).ravel(), got[k]), (mode, k)
