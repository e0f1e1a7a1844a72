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
