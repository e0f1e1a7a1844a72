#!/usr/bin/env python
"""Concurrent solves of dispersed scenarios on one GPU through the coalescing server
(gelato_b200/server.py), with the pyoptsparse stand-in (scipy trust-constr -- NOT IPOPT -- capped at a
few iterations: it is there to generate a realistic callback pattern, not to converge).

    python tests/scripts/batch_solve.py [n_scenarios] [maxiter]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from gelato_b200 import nlpshim, problem, scenarios, server  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    scen = scenarios.disperse(helpers.example_inputs(), n, seed=20260117)
    plans, x0s, conds = [], [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si)
        plans.append(helpers.compiled_plan(p, u, c))
        x0s.append(problem.xdict_to_vector(x0))
        conds.append(c)
    t0 = time.perf_counter()
    sols, stats = server.solve_batch(plans, x0s, conds, lambda: nlpshim.TrustConstr({"maxiter": maxiter}))
    wall = time.perf_counter() - t0
    print(json.dumps({
        "scenarios": n, "maxiter": maxiter, "wall_s": wall, "callback_calls": stats["calls"], "launches": stats["launches"],
        "mean_batch": stats["calls"] / max(1, stats["launches"]), "largest_batch": stats["largest_batch"],
        "callback_s_per_solver_mean": sum(s.userObjTime + s.userSensTime for s in sols) / n,
        "objectives": [float(s.fStar) for s in sols[:4]],
    }))


if __name__ == "__main__":
    main()
