import json
lines=open('/root/repo/profiles/r02k_solve_example.jsonl').read().strip().split('\n')
cpu=json.loads(lines[0]); gpu=json.loads(lines[1]); same=json.loads(lines[2])
bs=json.loads(open('/root/repo/profiles/r02j_bench_solves.json').read())["solves"]
out=[]
out.append("""Round 2, last session: a converged host-side solve of the shipped example through the callbacks
(gelato_b200/redsqp.py -- NOT IPOPT, which cannot be installed in the image).

Why the 20 earlier configurations failed (profiles/r02_solver_attempts.txt) and what was changed
 * The example asks for a CIRCULAR 200 km orbit through two equality rows, c1 = E / E_t - 1 (orbit energy) and
   c2 = h / h_t - 1 (angular momentum) (con_init_terminal_knot.py:365-368).  On the surface c1 = 0 the angular
   momentum is at its maximum exactly where c2 = 0, so c2 <= 0 on the rest of the feasible set and its gradient there
   lies in the span of c1's: no finite multipliers exist, the optimum is not a KKT point.
 * redsqp.py finds such a row (singular values of the row-normalised reduced equality Jacobian: one of them 1e-2 of
   the largest and 5e-2 of the next or less; 4.9e-4 of the largest at the phase-1 point of the example, 6e-9 at the
   solution), takes it out of the constraint list and puts - lam c2 into the objective: an exact penalty that is
   SMOOTH on the feasible set of the other rows, so SLSQP sees a regular problem.  lam grows by sqrt(10) per level,
   150 major iterations per level, warm start; the first level starts with four runs of 25 iterations inside a box of
   +-0.15 around the current point (SLSQP's first steps, identity Hessian, otherwise leave the region where the
   linearisations hold on some scenarios: without it scenario 2 of 8 ended 8 % below in payload).  Of the pair, the row with the smaller gradient (angular momentum) is
   the one to penalise: with the energy row SLSQP wandered (objective -1.003 .. -0.98 after 120 iterations).
 * Theory for a row that is a negative-definite quadratic on the tangent space of the others: violation ~ 1 / lam^2,
   objective error ~ 1 / lam.  Measured (nominal problem; the GPU box's host and the build container give the same
   digits, BLAS limited to one thread inside the solver):

      lam        objective          c2 (row value)   lam^2 c2   scaled optimality error   major iterations (cumulative)""")
for l in gpu['penalty_levels']:
    out.append("   %9.0f   %.10f   %+.3e      %+.3f     %.2e                  %d" % (l['lam'], l['obj'], l['row_value'], l['lam']**2*l['row_value'], l['kkt_scaled'], l['major_iterations']))
out.append("""
   The row value falls by ~10 per level (1 / lam^2; the column shows the level's best point, and below ~1e-10 the
   row -- computed as h / h_t - 1 -- is at its own resolution) and the objective steps shrink by ~sqrt(10) per level
   ({steps}): the continuation converges as the theory says.

CPU oracle callbacks against CUDA callbacks, same solver, same start (tests/scripts/solve_example.py --arm both, GPU box,
profiles/r02k_solve_example.jsonl): identical iterates, max |x_cpu - x_gpu| = {diff}.
                          CPU oracle callbacks      CUDA callbacks
   status                 {c[status]}                         {g[status]}
   major iterations       {c[nit]}                      {g[nit]}
   payload [kg]           {c[payload_kg]:.6f}              {g[payload_kg]:.6f}
   constraint violation   {c[constr_violation]:.2e}                  {g[constr_violation]:.2e}
   scaled optimality err  {c[optimality]:.2e}                  {g[optimality]:.2e}
   optTime [s]            {c[optTime]:.1f}                     {g[optTime]:.1f}
   userObjTime [s]/calls  {c[userObjTime]:.2f} / {c[userObjCalls]}            {g[userObjTime]:.2f} / {g[userObjCalls]}
   userSensTime [s]/calls {c[userSensTime]:.2f} / {c[userSensCalls]}             {g[userSensTime]:.2f} / {g[userSensCalls]}
   event times [s]        {ev}
   (13 event times identical in both arms.)
""".format(steps=", ".join("%.1e" % abs(b["obj"]-a["obj"]) for a,b in zip(gpu["penalty_levels"][:-1], gpu["penalty_levels"][1:])), diff=same["max_abs_diff_x"], c=cpu, g=gpu, ev=", ".join("%.4f"%t for t in gpu['event_times_s'])))
assert cpu['event_times_s']==gpu['event_times_s']
out.append(""" * Termination.  The scaled optimality error (IPOPT's definition, ORIGINAL problem, multiplier lam on c2, adjoint
   multipliers for the 857 state equations) sits at 1e-4 .. 5e-3 on every level, over scenarios and hosts: the dual
   residual carries lam x (error of the reference's forward-difference Jacobian, up to eps / dx = 2.2e-8 per entry with
   dx = 1e-8, Trajectory_Optimization.py:167) plus the noise of the other finite-difference rows (aerodynamic rows:
   up to 3e-4 per entry) times their multipliers.  No solver that sees only these callbacks can certify more.  Some
   runs dip under IPOPT's acceptable_tol = 1e-4 on some level by chance (the table above: no level; scenario 5: 6e-5
   and 7e-5 at lam = 1e4 / 3.2e4; profiles/r02j_solve_example_cpu_verbose.log, an earlier variant that stopped at the
   first such point: status 0 "solved to acceptable level" at lam = 3.2e4, 463 major iterations, payload
   27 818.63 kg, 3e-5 from the end point above -- what IPOPT's own acceptable-level exit would leave on this
   problem).  The default therefore runs the continuation until two successive levels agree in the objective to 1e-6
   relative with every row of the original problem within 1e-8 and the scaled optimality error within 5e-3 (the
   measured floor, stated in the status message), and reports status 3 ("converged in objective and constraints;
   optimality error at the noise floor"); status 0 is kept for IPOPT's own two tests.
 * Reproducibility.  Same machine, either callback set: bit-identical.  ACROSS paths (another host's library rounding,
   another flavour of the physics leaves, the solver before / after the boxed start) the end points of one scenario
   differ by 2e-5 .. 1e-3 in payload: nominal 27 817.29 / 27 817.82 / 27 817.87 kg (and 27 795.96 kg once, 8e-4 lower),
   k = 1: 27 578.46 / 27 579.58 / 27 580.53 / 27 581.04 / 27 581.13, k = 2: 27 632.64 / 27 656.36 / 27 688.10,
   k = 7: 27 802.83 / 27 836.53 / 27 857.65.  With an optimality error of ~3e-4 that no run gets under and a valley
   that flat (payload changes of 3e-4 over parameter changes of order one), that spread is the resolution of THIS
   problem on THESE callbacks (forward differences, dx = 1e-8): "settled to 1e-6 between levels" bounds the continuation
   error of one path, not the distance between two paths.
""")
def batch(bs, title):
    return """%s
   wall %.1f s for %d runs (%.0f runs / hour; solves_per_hour: %s); statuses %s; major iterations %s
   payload [kg] %s
   mean optTime %.1f s of which %.1f s inside objfunc (%d calls) and %.2f s inside sens (%d calls): the host side of
   the solver (SciPy SLSQP, sparse LU of the state equations, Python) is ~80 %% of a solve with the CUDA callbacks.
""" % (title, bs['wall_s'], bs['scenarios'], bs['runs_per_hour'], bs['solves_per_hour'], bs['statuses'], bs['major_iterations'], [round(v,2) for v in bs['payload_kg']], bs['optTime_mean_s'], bs['userObjTime_mean_s'], bs['userObjCalls_mean'], bs['userSensTime_mean_s'], bs['userSensCalls_mean'])
b1=json.loads(open('/root/repo/profiles/r02k_bench.json').read())["solves"]
br=json.loads(open('/root/repo/profiles/r02k_bench_reference.json').read())["solves"]
out.append(batch(bs, "Batched solves, first run (bench.py --solve-scenarios 8, one B200 + the box's host cores; profiles/r02j_bench_solves.json):\n8 dispersed scenarios (1 % masses and thrust, 20 % wind), one worker process each, all sharing GPU 0."))
out.append("""   Two of the eight did not converge in THAT run: k = 2 (the dependency was not visible at the phase-1 point with the
   first detection threshold) and k = 7 (phase 1 accepted parameters for which the state equations have no solution);
   both causes were fixed after it (detection by the gap in the singular values, re-checked after short runs; phase 1
   rejects unconverged inner solves).
""")
out.append(batch(b1, "Final session, default bench.py on one B200 (profiles/r02k_bench.json): 4 dispersed scenarios, 4 worker processes, CUDA callbacks: %.0f solves / hour" % b1["solves_per_hour"]))
out.append("""Same four scenarios in the reference arm (bench.py --impl reference, profiles/r02k_bench_reference.json): the solver on the
CPU oracle's callbacks (the reference's own C++ leaves), 4 worker processes: %d of %d converged, wall %.1f s, %.0f solves / hour;
mean optTime %.1f s of which %.1f s inside objfunc and %.1f s inside sens.
""" % (br["converged_total"], br["scenarios_total"], br["wall_s"], br["solves_per_hour"], br["optTime_mean_s"], br["userObjTime_mean_s"], br["userSensTime_mean_s"]))
import os
for n in (2,4,8):
    f='/root/repo/profiles/r02k_bench_solves_%dgpu.json'%n if n==2 else '/root/repo/profiles/r02k_bench_%dgpu.json'%n
    if os.path.exists(f):
        b=json.loads(open(f).read())["solves"]
        out.append(batch(b, "Batched solves, %d GPUs (%s): %d scenarios in total; wall = max over ranks %.1f s; solves_per_hour (whole job) %s; the lists below are rank 0's" % (n, os.path.basename(f), b.get('scenarios_total', b['scenarios']), b.get('wall_s_max_over_ranks', b['wall_s']), b['solves_per_hour'])))
out.append(open('/tmp/exp/local_table.txt').read())
open('/root/repo/profiles/r02j_solver_convergence.txt','w').write("\n".join(out))
