"""Batched solver runs of dispersed scenarios on one GPU (BASELINE.json metric iii: batched NLP solves per hour).

Each scenario is an independent NLP: one host thread per scenario runs a solver on callbacks with the
reference's signatures, and the per-GPU coalescing server (server.py) turns whatever `objfunc` / `sens` requests are
pending into ONE batched launch each.  Ranks own disjoint scenario blocks (scenarios.partition); nothing is exchanged
while solving.

The solver is the experimental interior-point stand-in of ipsolve.py -- NOT IPOPT -- and it does not reach IPOPT's
tolerance on this problem (ipsolve.py header), so this module reports what it can measure honestly: fixed-budget solver
runs per hour (every run = `iters` interior-point iterations, the callback pattern of a real solve) and how many of them
converged.  `solves_per_hour` is only filled in when every run converged."""
import time

import numpy as np

from . import ipsolve, plan as gplan, problem, scenarios, server

USER_EVENT = "IIP_END"


def solve_dispersed(inputs, n_total, world, rank, device=0, iters=100, engine_factory=None, max_workers=None):
    own = scenarios.partition(n_total, world, rank)
    scen = scenarios.disperse(inputs, n_total, seed=20260117)
    plans, x0s, conds = [], [], []
    for k in own:
        p, u, c, x0 = problem.problem_from_inputs(scen[k])
        plans.append(gplan.CompiledPlan(p, u, c, user_eq=gplan.PerigeeAtEvent(USER_EVENT)))
        x0s.append(problem.xdict_to_vector(x0))
        conds.append(c)
    t0 = time.perf_counter()
    sols, stats = server.solve_batch(plans, x0s, conds, lambda: ipsolve.IPSolver({"max_iter": iters}), device=device,
                                     engine_factory=engine_factory, max_workers=max_workers)
    wall = time.perf_counter() - t0
    n = len(plans)
    converged = sum(1 for s in sols if s.status == 0)
    return {
        "solver": "gelato_b200/ipsolve.py interior-point stand-in -- NOT IPOPT; fixed budget of %d iterations per run" % iters,
        "scenarios": n, "converged": converged, "wall_s": wall,
        "runs_per_hour": n / wall * 3600.0, "solves_per_hour": (n / wall * 3600.0) if converged == n else None,
        "callback_calls": stats["calls"], "launches": stats["launches"],
        "mean_batch": stats["calls"] / max(1, stats["launches"]), "largest_batch": stats["largest_batch"],
        "userObjTime_mean_s": float(np.mean([s.userObjTime for s in sols])),
        "userSensTime_mean_s": float(np.mean([s.userSensTime for s in sols])),
        "userObjCalls_mean": float(np.mean([s.userObjCalls for s in sols])),
        "userSensCalls_mean": float(np.mean([s.userSensCalls for s in sols])),
        "constr_violation_max": float(max(s.constr_violation for s in sols)),
        "objective_mean": float(np.mean([s.fStar for s in sols])),
    }
