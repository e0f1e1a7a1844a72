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
