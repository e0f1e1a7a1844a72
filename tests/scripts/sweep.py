#!/usr/bin/env python
"""BASELINE.json configs[4]: standalone callback sweep over mesh sizes (one NLP, no batching) plus the
three-stage 250-section problem (configs[2]).  For each size: device time of the two kernels (CUDA events,
100 back-to-back launches), host-buffer call times, and the CPU oracle's time for the same pair where it
finishes in seconds.  Writes JSON lines to gpurun_out/sweep.jsonl.  Run on the GPU box."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from gelato_b200 import callbacks, problem  # noqa: E402


def main():
    import torch

    out = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "w")
    cases = [("example", 1), ("example", 2), ("example", 15), ("three_stage", 62), ("example", 150), ("example", 1500)]
    for variant, factor in cases:
        inp = helpers.variant_inputs(variant)
        t0 = time.perf_counter()
        p, u, c, x0 = problem.problem_from_inputs(inp, factor=factor, max_nodes=20)
        prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), reuse_output=False)
        t_plan = time.perf_counter() - t0
        P, E = prob.plan, prob.engine
        x = helpers.perturbed(x0)
        xv = problem.xdict_to_vector(x)
        ec = P.eval_counts()
        xd = torch.from_numpy(xv).cuda()
        gd = torch.empty(P.n_rows, dtype=torch.float64, device="cuda")
        vd = torch.empty(P.n_vals, dtype=torch.float64, device="cuda")
        E.fill_template(vd.data_ptr(), 1)
        torch.cuda.synchronize()
        E.time_kernel(0, xd.data_ptr(), gd.data_ptr(), 1, 5)
        res_ms = E.time_kernel(0, xd.data_ptr(), gd.data_ptr(), 1, 100)
        E.time_kernel(1, xd.data_ptr(), vd.data_ptr(), 1, 5)
        jac_ms = E.time_kernel(1, xd.data_ptr(), vd.data_ptr(), 1, 100)

        def timed(fn, reps):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) / reps * 1e3

        reps = 20 if P.N < 20000 else 5
        obj_ms = timed(lambda: prob.objfunc(x), reps)
        sens_ms = timed(lambda: prob.sens(x), reps)
        prob2 = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), reuse_output=True)
        sens_reuse_ms = timed(lambda: prob2.sens(x), reps)
        prob2.close()
        prob3 = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), fuse_pair=True)
        fused_ms = timed(lambda: (prob3.objfunc(x), prob3.sens(x)), reps)
        prob3.close()
        row = {"variant": variant, "nodes": P.N, "sections": P.S, "n_vars": P.n_vars, "n_rows": P.n_rows,
               "n_vals": int(P.n_vals), "n_xdep": P.n_xdep, "evals_objfunc": ec["objfunc"], "evals_sens": ec["sens"],
               "plan_compile_s": t_plan, "k_residuals_ms": res_ms, "k_jacobian_ms": jac_ms,
               "device_evals_per_s": (ec["objfunc"] + ec["sens"]) / ((res_ms + jac_ms) * 1e-3),
               "objfunc_call_ms": obj_ms, "sens_call_fresh_arrays_ms": sens_ms, "sens_call_ms": sens_reuse_ms,
               "callback_pairs_per_s": 1e3 / (obj_ms + sens_reuse_ms),
               "fused_pair_call_ms": fused_ms, "fused_callback_pairs_per_s": 1e3 / fused_ms,
               "note": "sens_call_ms: the drop-in default (reuse_output: one page-locked buffer kept across calls, only the "
                       "x-dependent values cross PCIe); sens_call_fresh_arrays_ms: reuse_output=False (full copy into a new array); "
                       "fused_pair_call_ms: objfunc then sens at the same x with fuse_pair=True (one pair evaluation)"}
        if P.N <= 1000:
            from oracle import leaves

            flav = "ref" if leaves.ref_available() else "libm"
            O = helpers.oracle_nlp(p, u, c, flav, "numpy")
            xa = helpers.copy_x(x)
            row["cpu_objfunc_ms"] = timed(lambda: O.objfunc(xa), 3)
            row["cpu_sens_ms"] = timed(lambda: O.sens(xa), 3)
            row["cpu_leaves"] = flav
        print(json.dumps(row))
        out.write(json.dumps(row) + "\n")
        out.flush()
        prob.close()
        del xd, gd, vd


if __name__ == "__main__":
    main()
