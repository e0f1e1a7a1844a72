"""The callbacks inside a solver loop: the experimental interior-point solver (gelato_b200/ipsolve.py -- NOT IPOPT, and
not converging to IPOPT's tolerance on this problem, see its header) run for a fixed number of iterations on the CPU
oracle's callbacks and on the kernels' code (host emulator here, the GPU in test_gpu_parity.py) must produce the SAME
iterates bit for bit -- the drop-in property a real solve rests on -- and must make progress on the constraints."""
import numpy as np

import emu_binding
import helpers
from gelato_b200 import callbacks, ipsolve, nlpshim, problem
from oracle import leaves


def _solve(objfunc, sens, x0, p, c, iters):
    # a copy: the reference's (and the oracle's) `sens` perturbs and restores the dictionary it is given IN PLACE, which
    # leaves fl(fl(x+dx)-dx) in it -- the registration call would hand the next solve a different start
    x0 = helpers.copy_x(x0)
    opt = nlpshim.attach_structure(nlpshim.register(objfunc, sens, x0, c), p)
    return ipsolve.IPSolver({"max_iter": iters})(opt, sens=sens)


def test_oracle_and_kernel_callbacks_give_identical_iterates():
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    a = _solve(lambda x: O.objfunc(x), lambda x, f=None: O.sens(x), x0, p, c, 12)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                                   engine_factory=emu_binding.EmuEngine)
    b = _solve(prob.objfunc, prob.sens, x0, p, c, 12)
    assert np.array_equal(problem.xdict_to_vector(a.xStar), problem.xdict_to_vector(b.xStar))
    assert a.userSensCalls == b.userSensCalls and a.userObjCalls == b.userObjCalls and a.nit == b.nit == 12
    # from a violation of 6.7 at the initial guess
    assert a.constr_violation < 0.5 and a.history[0][1] > 5.0
    assert a.optInform["value"] == 1  # iteration limit: the stand-in does not claim convergence


def test_state_elimination_solver_gives_identical_iterates_on_both_callback_sets():
    """gelato_b200/redsqp.py (inner Newton on the state equations, reduced derivatives by the implicit function
    theorem, Levenberg-Marquardt + SLSQP outside) driven by the oracle's callbacks and by the emulated kernels: a few
    evaluations of each phase end at the same point bit for bit, and the inner solve leaves the state equations at
    round-off."""
    from gelato_b200 import redsqp
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                                   engine_factory=emu_binding.EmuEngine)
    ends = []
    for objfunc, sens in ((lambda x: O.objfunc(x), lambda x, f=None: O.sens(x)), (prob.objfunc, prob.sens)):
        opt = nlpshim.attach_structure(nlpshim.register(objfunc, sens, helpers.copy_x(x0), c), p)
        sol = redsqp.ReducedSQP({"phase1_evals": 6, "max_iter": 4})(opt, sens=sens)
        ends.append(problem.xdict_to_vector(sol.xStar))
        assert sol.userSensCalls > 0 and sol.userObjCalls > sol.userSensCalls
    assert np.array_equal(ends[0], ends[1])


def _callback_sets(p, u, c, Lg):
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                                   engine_factory=emu_binding.EmuEngine)
    return ((lambda x: O.objfunc(x), lambda x, f=None: O.sens(x)), (prob.objfunc, prob.sens))


def test_dependent_terminal_row_is_found_and_penalised_with_the_right_sign():
    """The shipped example asks for a circular orbit through two rows (orbit energy, angular momentum) that are
    dependent at every feasible point; redsqp.py finds the pair from the singular values of the reduced equality
    Jacobian and carries the angular-momentum row (negative wherever the other rows hold) by a penalty f - lam c with
    lam > 0.  Same decision and same first penalty level on both callback sets."""
    from gelato_b200 import redsqp
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    ends = []
    for objfunc, sens in _callback_sets(p, u, c, Lg):
        opt = nlpshim.register(objfunc, sens, helpers.copy_x(x0), c)
        S = redsqp.ReducedSQP({"phase1_evals": 60, "max_iter": 12, "level_iter": 8})
        sol = S(opt, sens=sens)
        R = S.reduced
        assert sol.dependent_rows == [("eqcon_terminal", 1)] and sol.penalty_sign == 1.0
        assert [R.row_names[i] for i in R.e_rows].count("eqcon_terminal") == 2
        ends.append(problem.xdict_to_vector(sol.xStar))
        assert sol.status != 0  # 12 iterations are not a solve
    assert np.array_equal(ends[0], ends[1])


def test_recorded_solution_of_the_example_passes_the_termination_test():
    """tests/golden/example_solution.npz is the end point of `tests/scripts/solve_example.py --arm cpu --save ...`
    (1 056 major iterations from the reference's initial guess, 7 penalty levels; payload 27 817.29 kg).  At that
    point every row of the ORIGINAL problem is within 1e-8 (one objfunc of the oracle), and the solver restarted there
    one penalty level below the recorded multiplier runs two short levels, finds the objective settled and the dual
    residual at the noise floor, and reports convergence (status 3) -- on both callback sets, ending at the same point
    bit for bit, with payload and event times within 1e-6 relative of the record."""
    import os

    from gelato_b200 import redsqp
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_solution.npz"))
    xs = problem.vector_to_xdict(rec["x"], p["M"], p["N"], p["num_sections"])
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    f, fail = O.objfunc(helpers.copy_x(xs))
    assert not fail
    for key, val in f.items():
        if val is None:
            continue
        if key.startswith("eqcon"):
            assert np.abs(val).max() <= 1e-8, key
        elif key.startswith("ineqcon"):
            assert np.min(val) >= -1e-8, key
    assert abs(xs["mass"][0] * u["mass"] - float(rec["payload_kg"])) <= 1e-9 * float(rec["payload_kg"])
    ends = []
    for objfunc, sens in _callback_sets(p, u, c, Lg):
        opt = nlpshim.register(objfunc, sens, helpers.copy_x(xs), c)
        sol = redsqp.ReducedSQP({"phase1_evals": 0, "penalty0": float(rec["multiplier"]) / 10.0 ** 0.5, "penalty_levels": 2,
                                 "level_iter": 50, "max_iter": 300})(opt, sens=sens)
        assert sol.status == 3 and "converged in objective" in sol.message, sol.message
        assert sol.constr_violation <= 1e-8 and len(sol.penalty_levels) == 2
        assert abs(sol.xStar["mass"][0] * u["mass"] / float(rec["payload_kg"]) - 1.0) <= 1e-6
        assert np.allclose(sol.xStar["t"] * u["t"], rec["event_times_s"], rtol=1e-6, atol=1e-6)
        ends.append(problem.xdict_to_vector(sol.xStar))
    assert np.array_equal(ends[0], ends[1])
