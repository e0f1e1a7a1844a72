#!/usr/bin/env python
"""End-to-end solver loop on the shipped example with the pyoptsparse stand-in (gelato_b200/nlpshim.py) and the
experimental interior-point solver (gelato_b200/ipsolve.py -- NOT IPOPT; it does not reach IPOPT's tolerance on this
problem, see its header and profiles/r02_solver_attempts.txt), once on the CPU oracle's callbacks and once on the CUDA
callbacks: same solver, same problem, same start.  Prints the state it reaches (objective, constraint violation,
payload, event times), iteration counts and the time spent inside the callbacks under pyoptsparse's names
(userObjTime / userSensTime / calls, Trajectory_Optimization.py:511-517).

    python tests/scripts/solve_example.py --arm cpu|gpu|both [--maxiter 300] [--factor 1] [--solver ip|trust-constr]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from gelato_b200 import nlpshim, problem  # noqa: E402


def run(arm, factor, maxiter, solver="redsqp", verbose=0, scenario=0, of=8):
    from gelato_b200 import scenarios
    from oracle import leaves

    Lg = leaves.get("gmath")
    if scenario == 0:
        p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c, factor=factor)
    else:  # a dispersed copy (masses, thrust, wind: scenarios.disperse), as the batched solves use them
        inp = scenarios.disperse(helpers.example_inputs(), of, seed=20260117)[scenario]
        p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c, factor=factor, max_nodes=20)
    if arm == "cpu":
        O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")

        def objfunc(x):
            return O.objfunc(x)

        def sens(x, f=None):
            return O.sens(x)
    else:
        from gelato_b200 import callbacks

        prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c)
        objfunc, sens = prob.objfunc, prob.sens
    from gelato_b200 import ipsolve

    opt = nlpshim.attach_structure(nlpshim.register(objfunc, sens, helpers.copy_x(x0), c), p)
    if solver == "redsqp":
        from gelato_b200 import redsqp
        extra = json.loads(os.environ.get("REDSQP_OPTIONS", "{}"))  # experiments: {"start_radius": 0.1} ...
        sol = redsqp.ReducedSQP(dict({"max_iter": maxiter, "verbose": verbose}, **extra))(opt, sens=sens)
    elif solver == "ip":
        sol = ipsolve.IPSolver({"max_iter": maxiter})(opt, sens=sens)
    else:
        sol = nlpshim.TrustConstr({"maxiter": maxiter})(opt, sens=sens)
    t_events = sol.xStar["t"] * u["t"]
    out = {
        "arm": arm, "scenario": scenario, "solver": solver + " (NOT IPOPT)", "penalty_levels": getattr(sol, "penalty_levels", None), "optimality": getattr(sol, "optimality", None),
        "reduced_evaluations": getattr(sol, "reduced_evaluations", None), "nodes": int(p["N"]), "nit": sol.nit, "status": sol.status,
        "message": getattr(sol, "message", ""), "obj": float(sol.fStar),
        "constr_violation": sol.constr_violation,
        "payload_kg": float(sol.xStar["mass"][0] * u["mass"]) if c["OptimizationMode"] == "Payload" else None,
        "event_times_s": [float(v) for v in t_events],
        "optTime": sol.optTime, "userObjTime": sol.userObjTime, "userObjCalls": sol.userObjCalls,
        "userSensTime": sol.userSensTime, "userSensCalls": sol.userSensCalls,
    }
    return out, problem.xdict_to_vector(sol.xStar)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["cpu", "gpu", "both"])
    ap.add_argument("--maxiter", type=int, default=1800)
    ap.add_argument("--solver", default="redsqp", choices=["redsqp", "ip", "trust-constr"])
    ap.add_argument("--verbose", type=int, default=0)
    ap.add_argument("--factor", type=int, default=1)
    ap.add_argument("--save", default="", help="write the end point of the (first) arm to this .npz")
    ap.add_argument("--scenario", type=int, default=0, help="0: the shipped example; k > 0: the k-th dispersed copy of --of")
    ap.add_argument("--of", type=int, default=8)
    a = ap.parse_args()
    res = {}
    for arm in (["cpu", "gpu"] if a.arm == "both" else [a.arm]):
        res[arm] = run(arm, a.factor, a.maxiter, a.solver, a.verbose, a.scenario, a.of)
        print(json.dumps(res[arm][0]), flush=True)
        if a.save and len(res) == 1:
            o = res[arm][0]
            np.savez(a.save, x=res[arm][1], payload_kg=o["payload_kg"], event_times_s=np.array(o["event_times_s"]), obj=o["obj"],
                     multiplier=o["penalty_levels"][-1]["lam"] if o["penalty_levels"] else np.nan,
                     optimality=o["optimality"], constr_violation=o["constr_violation"], nit=o["nit"])
    if len(res) == 2:
        xa, xb = res["cpu"][1], res["gpu"][1]
        print(json.dumps({"max_abs_diff_x": float(np.max(np.abs(xa - xb))), "identical_iterates": bool(np.array_equal(xa, xb))}))


if __name__ == "__main__":
    main()
