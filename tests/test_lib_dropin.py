"""gelato_b200.lib: the reference's per-group functions (lib/con_*.py, cost_gradient.py, jac_fd.py;
SURVEY.md 8(b) "mid boundary") as views of the fused kernels' results.  Runs on the host emulator of
the kernels in the CPU tier and on the CUDA engine in the GPU tier."""
import numpy as np
import pytest

import helpers
from gelato_b200 import lib as glib
from gelato_b200 import problem
from gelato_b200.callbacks import PerigeeAtEvent
from gelato_b200.lib import (con_aero, con_dynamics, con_init_terminal_knot, con_trajectory, con_user, con_waypoint,
                             cost_gradient, jac_fd)
from oracle import leaves

# reference function -> (module here, funcs / funcsSens key); names as in /root/reference/lib
VALUE_FUNCS = {
    "eqcon_init": (con_init_terminal_knot, "equality_init"), "eqcon_time": (con_init_terminal_knot, "equality_time"),
    "eqcon_dyn_mass": (con_dynamics, "equality_dynamics_mass"), "eqcon_dyn_pos": (con_dynamics, "equality_dynamics_position"),
    "eqcon_dyn_vel": (con_dynamics, "equality_dynamics_velocity"),
    "eqcon_dyn_quat": (con_dynamics, "equality_dynamics_quaternion"),
    "eqcon_knot": (con_init_terminal_knot, "equality_knot_LGR"),
    "eqcon_terminal": (con_init_terminal_knot, "equality_6DoF_LGR_terminal"),
    "eqcon_rate": (con_trajectory, "equality_6DoF_rate"), "eqcon_pos": (con_waypoint, "equality_posLLH"),
    "eqcon_iip": (con_waypoint, "equality_IIP"), "eqcon_user": (con_user, "equality_user"),
    "ineqcon_alpha": (con_aero, "inequality_max_alpha"), "ineqcon_q": (con_aero, "inequality_max_q"),
    "ineqcon_qalpha": (con_aero, "inequality_max_qalpha"), "ineqcon_mass": (con_trajectory, "inequality_mass"),
    "ineqcon_kick": (con_trajectory, "inequality_kickturn"), "ineqcon_time": (con_init_terminal_knot, "inequality_time"),
    "ineqcon_pos": (con_waypoint, "inequality_posLLH"), "ineqcon_iip": (con_waypoint, "inequality_IIP"),
    "ineqcon_antenna": (con_waypoint, "inequality_antenna"), "ineqcon_user": (con_user, "inequality_user"),
}
JAC_NAMES = {
    "eqcon_init": "equality_jac_init", "eqcon_time": "equality_jac_time", "eqcon_dyn_mass": "equality_jac_dynamics_mass",
    "eqcon_dyn_pos": "equality_jac_dynamics_position", "eqcon_dyn_vel": "equality_jac_dynamics_velocity",
    "eqcon_dyn_quat": "equality_jac_dynamics_quaternion", "eqcon_knot": "equality_jac_knot_LGR",
    "eqcon_terminal": "equality_jac_6DoF_LGR_terminal", "eqcon_rate": "equality_jac_6DoF_rate",
    "eqcon_pos": "equality_jac_posLLH", "eqcon_iip": "equality_jac_IIP", "eqcon_user": "equality_jac_user",
    "ineqcon_alpha": "inequality_jac_max_alpha", "ineqcon_q": "inequality_jac_max_q",
    "ineqcon_qalpha": "inequality_jac_max_qalpha", "ineqcon_mass": "inequality_jac_mass",
    "ineqcon_kick": "inequality_jac_kickturn", "ineqcon_time": "inequality_jac_time",
    "ineqcon_pos": "inequality_jac_posLLH", "ineqcon_iip": "inequality_jac_IIP",
    "ineqcon_antenna": "inequality_jac_antenna", "ineqcon_user": "inequality_jac_user",
}


def _emu_factory(plan):
    import emu_binding

    return emu_binding.EmuEngine(plan)


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def setup(request):
    Lg = leaves.get("gmath")
    inp = helpers.variant_inputs("waypoints")
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c)
    glib.configure(user_eq=PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                   engine_factory=_emu_factory if request.param == "emu" else None)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    yield p, u, c, helpers.perturbed(x0), O
    glib.reset()


def test_every_group_function_matches_the_oracle_with_two_evaluations(setup):
    p, u, c, x, O = setup
    xa = helpers.copy_x(x)
    fo, _ = O.objfunc(xa)
    so, _ = O.sens(xa)
    funcs = {"obj": cost_gradient.cost_6DoF(x, c)}
    sens = {"obj": cost_gradient.cost_jac(x, c)}
    for key, (mod, name) in VALUE_FUNCS.items():
        funcs[key] = getattr(mod, name)(x, p, u, c)
    for key, (mod, _) in VALUE_FUNCS.items():
        sens[key] = getattr(mod, JAC_NAMES[key])(x, p, u, c)
    helpers.assert_funcs_equal(fo, {k: funcs[k] for k in fo})
    helpers.assert_sens_equal(so, {k: sens[k] for k in so})
    eng = glib.problem_for(p, u, c).prob.engine
    assert eng.calls == 2  # 23 value functions = ONE residual evaluation, 23 Jacobian functions = ONE Jacobian evaluation
    # a new decision vector invalidates both caches; results are copies the caller owns
    before = funcs["eqcon_dyn_vel"].copy()
    x2 = helpers.perturbed(x, seed=11)
    r2 = con_dynamics.equality_dynamics_velocity(x2, p, u, c)
    assert eng.calls == 3 and not np.array_equal(r2, before)
    assert np.array_equal(funcs["eqcon_dyn_vel"], before)
    assert con_aero.inequality_length_max_qalpha(x, p, u, c) == len(funcs["ineqcon_qalpha"])
    assert con_aero.inequality_length_max_q(x, p, u, c) == 0 and funcs["ineqcon_q"] is None
    assert con_trajectory.equality_length_6DoF_rate(x, p, u, c) == len(funcs["eqcon_rate"])
    x3 = helpers.copy_x(x2)  # the drop-in never writes the caller's xdict
    con_dynamics.equality_jac_dynamics_velocity(x2, p, u, c)
    for k in x2:
        assert np.array_equal(x2[k], x3[k]), k


def test_jac_fd_routes_builtins_to_the_kernel_and_differences_python_callables_on_the_host(setup):
    p, u, c, x, O = setup
    xa = helpers.copy_x(x)
    so, _ = O.sens(xa)
    got = jac_fd.jac_fd(con_user.equality_user, x, p, u, c)
    assert list(got.keys()) == list(so["eqcon_user"].keys())
    for k in got:
        assert np.array_equal(got[k], so["eqcon_user"][k]), k

    def host_con(xdict, pdict, unitdict, condition):  # an arbitrary Python constraint: stays on the host
        return np.array([xdict["mass"][0] * xdict["t"][-1], xdict["u"][3] ** 2])

    xb = helpers.copy_x(x)
    J = jac_fd.jac_fd(host_con, xb, p, u, c)
    assert set(J) == set(x) and J["mass"].shape == (2, x["mass"].size)
    np.testing.assert_allclose(J["mass"][0, 0], x["t"][-1], rtol=1e-6)
    np.testing.assert_allclose(J["u"][1, 3], 2 * x["u"][3], rtol=1e-5, atol=1e-7)
    assert np.count_nonzero(J["position"]) == 0
