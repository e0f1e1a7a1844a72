"""State-elimination SQP: a host-side NLP solver for the trajectory problem behind the pyoptsparse call shape
(`solver(optProb, sens=sens)`), used where the reference calls IPOPT / SNOPT through pyoptsparse
(/root/reference/Trajectory_Optimization.py:419-462).  NOT IPOPT -- neither library can be installed in this image --
and not a general NLP solver: it uses the structure of THIS problem.

The collocation NLP has 11 state values per state node and, for them, exactly as many "state equations": the initial
conditions, the collocation defects and the knot conditions (eqcon_init, eqcon_dyn_*, eqcon_knot).  Given the
parameters p -- the rate controls u, the event times t and the state entries the initial conditions leave free (the
lift-off mass in payload mode) -- those equations determine every state: they ARE the integration of the equations
of motion, a well-conditioned square system that Newton's method solves in 2-3 iterations with the sparse Jacobian
`sens` returns.  What is left is a small dense problem in p (146 variables, 39 degrees of freedom for the shipped
example): the objective, the remaining equality rows (fixed times, rate continuity, terminal orbit, user rows) and the
inequality rows as functions of p, with derivatives by the implicit function theorem,
    dc/dp = J[c, p] - J[c, s] (J[F, s])^-1 J[F, p],
handled in two phases: a bounded Levenberg-Marquardt iteration (SciPy least_squares, trust-region reflective) on the
violated rows, then SciPy's SLSQP with a scaled objective.

STATUS (profiles/r02_solver_attempts.txt): on the shipped example this reaches a FEASIBLE trajectory from the
reference's initial guess -- every row of the original problem within 1e-8 after ~200 major iterations, where the
full-space interior-point and SQP variants of ipsolve.py never got below 1e-4 -- but it does NOT reach IPOPT's
optimality tolerance: the scaled first-order error stays at 1e-3 .. 1e-2 and the objective keeps creeping along a
nearly flat feasible valley (different runs stop 0.4 % apart in payload).  `status` says which: 0 optimal, 2 feasible
but not optimal, 1 neither.  Converged-solution parity and solves per hour therefore remain unmeasured.

Every function value the solver sees comes from the two callbacks `objfunc` / `sens` -- the oracle's on the CPU or
the CUDA kernels' -- so bit-identical callbacks give bit-identical iterates; the time spent inside them is recorded
under pyoptsparse's names (userObjTime / userSensTime / calls, Trajectory_Optimization.py:511-517).

Termination: first-order optimality of the reduced problem, checked after every major iteration with least-squares
multipliers on the active set -- IPOPT's scaled test (`tol`, default 1e-6) -- plus `constr_tol` on every row of the
original problem.
"""
import time
import types

import numpy as np
import scipy.optimize as so
import scipy.sparse.linalg as spla

from .ipsolve import IPSolver, Solution

STATE_VARS = ("mass", "position", "velocity", "quaternion")
STATE_ROWS = ("eqcon_init", "eqcon_dyn_mass", "eqcon_dyn_pos", "eqcon_dyn_vel", "eqcon_dyn_quat", "eqcon_knot")


class InnerFailure(Exception):
    pass


class ReducedSQP:
    """solver = ReducedSQP({"tol": 1e-6, "max_iter": 600}); sol = solver(optProb, sens=sens)."""

    DEFAULTS = {"tol": 1e-6, "constr_tol": 1e-8, "max_iter": 260, "inner_tol": 1e-12, "inner_iter": 25, "restarts": 4,
                "active_tol": 1e-7, "u_scale": 1.0, "phase1_evals": 60, "obj_scale": 0.01, "predict_radius": 0.05, "slsqp_ftol": 1e-15, "verbose": 0}

    def __init__(self, options=None):
        self.opt = dict(self.DEFAULTS)
        for k, v in (options or {}).items():
            if k in self.opt:
                self.opt[k] = v  # IPOPT-only options (linear_solver, output_file ...) are ignored

    def __call__(self, prob, sens=None, **_):
        o = self.opt
        t_start = time.perf_counter()
        names = [v[0] for v in prob.vars]
        sizes = [v[1] for v in prob.vars]
        offs = np.concatenate(([0], np.cumsum(sizes))).astype(int)
        n = int(offs[-1])
        col0 = dict(zip(names, offs[:-1]))
        x0 = np.concatenate([v[2] for v in prob.vars]).astype(float)
        xl = np.concatenate([np.full(v[1], -np.inf if v[3] is None else v[3]) for v in prob.vars])
        xu = np.concatenate([np.full(v[1], np.inf if v[4] is None else v[4]) for v in prob.vars])
        cons = list(prob.cons)
        roffs = np.concatenate(([0], np.cumsum([g[1] for g in cons]))).astype(int)
        m = int(roffs[-1])
        rows_of = {g[0]: np.arange(roffs[i], roffs[i + 1]) for i, g in enumerate(cons)}
        for g in cons:
            is_eq = g[3] is not None and g[2] == g[3]
            if (is_eq and g[2] != 0.0) or (not is_eq and (g[3] is not None or g[2] != 0.0)):
                raise NotImplementedError("constraint groups other than c(x) = 0 and c(x) >= 0")
        eq_names = [g[0] for g in cons if g[3] is not None and g[2] == g[3]]
        missing = [r for r in STATE_ROWS if r not in rows_of] + [v for v in STATE_VARS if v not in col0]
        if missing:
            raise ValueError("not a GELATO trajectory problem (missing %s)" % ", ".join(missing))
        f_rows = np.concatenate([rows_of[r] for r in STATE_ROWS])
        e_rows = np.concatenate([rows_of[r] for r in eq_names if r not in STATE_ROWS] or [np.zeros(0, int)]).astype(int)
        i_rows = np.concatenate([rows_of[g[0]] for g in cons if g[0] not in eq_names] or [np.zeros(0, int)]).astype(int)
        state_cols = np.concatenate([np.arange(col0[v], col0[v] + sizes[names.index(v)]) for v in STATE_VARS])
        stat = {"obj_t": 0.0, "obj_n": 0, "sens_t": 0.0, "sens_n": 0}

        def xdict(xv):
            return {nm: xv[offs[i]: offs[i + 1]].copy() for i, nm in enumerate(names)}

        def values(xv):
            t0 = time.perf_counter()
            f, fail = prob.objfunc(xdict(xv))
            stat["obj_t"] += time.perf_counter() - t0
            stat["obj_n"] += 1
            if fail:
                raise InnerFailure("objfunc reported failure")
            c = np.concatenate([np.atleast_1d(np.asarray(f[g[0]], dtype=float)) for g in cons])
            return float(np.asarray(f[prob.obj]).ravel()[0]), c, f

        def jacobian(xv, f):
            t0 = time.perf_counter()
            s, fail = sens(xdict(xv), f)
            stat["sens_t"] += time.perf_counter() - t0
            stat["sens_n"] += 1
            if fail:
                raise InnerFailure("sens reported failure")
            grad = np.zeros(n)
            for var, blk in s[prob.obj].items():
                grad[col0[var]: col0[var] + np.size(blk)] = np.ravel(blk)
            return grad, IPSolver._jac(s, cons, m, n, col0).tocsc()

        # ---- which state entries the state equations determine: all but the ones the initial conditions leave free ----
        _, c0, f0 = values(x0)
        g0, J0 = jacobian(x0, f0)
        k_free = state_cols.size - f_rows.size
        if k_free < 0:
            raise ValueError("more state equations than state values")
        pinned = np.asarray(np.abs(J0[rows_of["eqcon_init"]]).sum(axis=0)).ravel() > 0
        node0 = np.concatenate([np.arange(col0[v], col0[v] + sizes[names.index(v)] // (sizes[names.index("mass")]))
                                for v in STATE_VARS])
        free = [int(cidx) for cidx in node0 if not pinned[cidx]]
        if len(free) != k_free:
            raise ValueError("cannot tell which %d initial state values are free (found %d)" % (k_free, len(free)))
        s_cols = np.setdiff1d(state_cols, free)
        p_cols = np.setdiff1d(np.arange(n), s_cols)
        # the outer iteration works on q = w p: SLSQP starts from the identity as its Hessian, and an identity in the
        # raw rate controls (bounds +-9 around values of ~0.3) makes its first steps run into those bounds
        w = np.ones(p_cols.size)
        if "u" in col0:
            w[np.searchsorted(p_cols, np.arange(col0["u"], col0["u"] + sizes[names.index("u")]))] = o["u_scale"]
        bounds = list(zip(xl[p_cols] * w, xu[p_cols] * w))

        # ---- the inner solve: states from parameters ----
        S = {"x": x0.copy(), "lu": None, "dsdp": None, "p_lin": None, "x_lin": None, "cache": {}, "evals": 0}

        def factor(J):
            S["lu"] = spla.splu(J[f_rows][:, s_cols].tocsc())

        factor(J0)

        def newton(xv):
            """The state equations solved from xv (parameters fixed): chord steps with the kept factors while they
            contract well, otherwise Newton steps with a fresh Jacobian, damped by backtracking on the residual.
            Returns (x, obj, c, f, converged); an unconverged result is the last iterate -- finite values the outer
            line search can still compare."""
            obj, c, f = values(xv)
            r = np.abs(c[f_rows]).max()
            if not np.isfinite(r):
                raise InnerFailure("non-finite state residual at the start")
            for _ in range(o["inner_iter"]):
                if r < o["inner_tol"]:
                    return xv, obj, c, f, True
                xt = xv.copy()
                xt[s_cols] -= S["lu"].solve(c[f_rows])
                obj_t, c_t, f_t = values(xt)
                r_t = np.abs(c_t[f_rows]).max()
                if np.isfinite(r_t) and r_t <= 0.5 * r:
                    xv, obj, c, f, r = xt, obj_t, c_t, f_t, r_t
                    continue
                _, J = jacobian(xv, f)
                factor(J)
                d = -S["lu"].solve(c[f_rows])
                alpha, best = 1.0, None
                while alpha > 1e-4:
                    xt = xv.copy()
                    xt[s_cols] += alpha * d
                    obj_t, c_t, f_t = values(xt)
                    r_t = np.abs(c_t[f_rows]).max()
                    if np.isfinite(r_t) and (best is None or r_t < best[4]):
                        best = (xt, obj_t, c_t, f_t, r_t)
                    if np.isfinite(r_t) and r_t < (1.0 - 0.3 * alpha) * r:
                        break
                    alpha *= 0.5
                if best is None:
                    raise InnerFailure("no finite trial point")
                xv, obj, c, f, r = best
            return xv, obj, c, f, r < 1e4 * o["inner_tol"]

        def solve_states(pv):
            """x(p): from the last states (plus, for a short step, the first-order prediction of the last linearisation)"""
            xv = S["x"].copy()
            if S["dsdp"] is not None and np.abs(pv - S["p_lin"]).max() < o["predict_radius"]:
                xv[s_cols] = S["x_lin"][s_cols] + S["dsdp"] @ (pv - S["p_lin"])
            xv[p_cols] = pv
            return newton(xv)

        def at(qv, want_jac):
            """the reduced problem at q = w p: values, and with want_jac the derivatives with respect to q"""
            qv = np.asarray(qv, dtype=float)
            pv = qv / w
            key = qv.tobytes()
            e = S["cache"].get(key)
            if e is None:
                try:
                    xv, obj, c, f, ok = solve_states(pv)
                    S["x"] = xv
                    e = {"x": xv, "obj": obj, "c": c, "f": f, "ok": True, "converged": ok}
                    if not ok:
                        S["failures"] = S.get("failures", 0) + 1
                        if o["verbose"] > 1:
                            print("      state equations left at residual %.1e" % np.abs(c[f_rows]).max(), flush=True)
                except InnerFailure as exc:
                    S["failures"] = S.get("failures", 0) + 1
                    if o["verbose"] > 1:
                        print("      inner failure:", exc, " |dp| =", np.abs(pv - S["x"][p_cols]).max(), flush=True)
                    # no finite trajectory at all: a finite, very bad point, so that the line search backs off
                    e = {"x": None, "obj": 1e3, "c": None, "f": None, "ok": False, "converged": False}
                S["cache"] = {key: e} if want_jac else dict(list(S["cache"].items())[-3:] + [(key, e)])
                S["evals"] += 1
            if want_jac and not e["ok"]:
                # SLSQP accepted a point where the trajectory cannot be integrated (its line search gave up): hand it
                # the last linearisation; the next line search starts from values that send it back
                if S.get("last_good") is None:
                    raise InnerFailure("derivatives requested where the state equations have no solution")
                return dict(S["last_good"], ok=False, obj=e["obj"], c=None)
            if want_jac and "Je" not in e:
                grad, J = jacobian(e["x"], e["f"])
                factor(J)
                dsdp = -S["lu"].solve(J[f_rows][:, p_cols].toarray())
                S["dsdp"], S["p_lin"], S["x_lin"] = dsdp, pv.copy(), e["x"].copy()
                e["g"] = (grad[p_cols] + dsdp.T @ grad[s_cols]) / w
                e["Je"] = (J[e_rows][:, p_cols].toarray() + J[e_rows][:, s_cols] @ dsdp) / w
                e["Ji"] = (J[i_rows][:, p_cols].toarray() + J[i_rows][:, s_cols] @ dsdp) / w
                S["last_good"] = e
            return e

        big_e, big_i = np.full(e_rows.size, 1e3), np.full(i_rows.size, -1e3)

        def fun(pv):
            return at(pv, False)["obj"]

        def ceq(pv):
            e = at(pv, False)
            return e["c"][e_rows] if e["ok"] else big_e

        def cin(pv):
            e = at(pv, False)
            return e["c"][i_rows] if e["ok"] else big_i

        # ---- first-order optimality of the reduced problem ----
        def kkt(pv, e):
            ci = e["c"][i_rows]
            act = np.where(ci <= o["active_tol"])[0]
            at_l = np.where(pv - xl[p_cols] * w <= 1e-9)[0]
            at_u = np.where(xu[p_cols] * w - pv <= 1e-9)[0]
            # g = Je' lam + Ji[act]' mu + zl - zu,  mu, zl, zu >= 0
            A = np.hstack([e["Je"].T, e["Ji"][act].T, np.eye(pv.size)[:, at_l], -np.eye(pv.size)[:, at_u]])
            lo = np.concatenate([np.full(e_rows.size, -np.inf), np.zeros(act.size + at_l.size + at_u.size)])
            colsc = np.maximum(np.abs(A).max(axis=0), 1e-300)  # column scaling: the rows' gradients differ by 1e6
            r = so.lsq_linear(A / colsc, e["g"], bounds=(lo, np.full(lo.size, np.inf)), method="bvls", lsmr_tol=None)
            mult = r.x / colsc
            resid = np.abs(A @ mult - e["g"]).max()
            s_d = max(100.0, np.abs(mult).sum() / max(mult.size, 1)) / 100.0  # IPOPT's multiplier scaling
            return resid / s_d, mult

        hist = {"it": 0, "best": None, "kkt": np.inf}

        class Done(Exception):
            pass

        def feas(c):
            return max(np.abs(c[np.concatenate([f_rows, e_rows])]).max(), -min(0.0, c[i_rows].min()) if i_rows.size else 0.0)

        def callback(pv):
            hist["it"] += 1
            e = at(np.asarray(pv), True)
            if not e["ok"]:
                return
            th = feas(e["c"])
            err, _ = kkt(np.asarray(pv), e)
            if o["verbose"]:
                print("%4d  obj %.10f  infeas %.2e  kkt %.2e  evals %d" % (hist["it"], e["obj"], th, err, S["evals"]), flush=True)
            if th <= o["constr_tol"] and (hist["best"] is None or err < hist["kkt"]):
                hist["best"], hist["kkt"] = (np.array(pv), e), err
            if th <= o["constr_tol"] and err <= o["tol"]:
                raise Done()
            if hist["it"] >= o["max_iter"]:
                raise Done()

        # the reduced problem of this call, for diagnostics (tests/scripts): q = w p
        self.reduced = types.SimpleNamespace(at=at, kkt=kkt, feas=feas, p_cols=p_cols, s_cols=s_cols, w=w, lower=xl[p_cols] * w,
                                             upper=xu[p_cols] * w, e_rows=e_rows, i_rows=i_rows, state=S, xdict=xdict,
                                             row_names=[g[0] for g in cons for _ in range(g[1])])
        cons_sq = [{"type": "eq", "fun": ceq, "jac": lambda pv: at(pv, True)["Je"]},
                   {"type": "ineq", "fun": cin, "jac": lambda pv: at(pv, True)["Ji"]}]
        pv = x0[p_cols] * w
        status, message = 1, "iteration limit"
        # ---- phase 1: towards feasibility with a trust region.  From the reference's initial guess SLSQP's first steps
        # (identity Hessian, linear objective) leave the region where the linearised path constraints mean anything
        # and the iteration turns chaotic; a bounded Levenberg-Marquardt iteration on the violated rows
        # [c_E; min(c_I, 0)] gets close to the feasible set in short, safe steps. ----
        hist["phase"] = 1
        if o["phase1_evals"] > 0:
            def resid(q):
                e = at(q, False)
                if not e["ok"]:
                    return np.full(e_rows.size + i_rows.size, 1e3)
                return np.concatenate([e["c"][e_rows], np.minimum(e["c"][i_rows], 0.0)])

            def resjac(q):
                e = at(q, True)
                viol = (e["c"][i_rows] < 0.0) if e.get("c") is not None else np.ones(i_rows.size, bool)
                return np.vstack([e["Je"], e["Ji"] * viol[:, None]])

            ls = so.least_squares(resid, pv, jac=resjac, bounds=(xl[p_cols] * w, xu[p_cols] * w), method="trf", xtol=1e-14,
                                  ftol=1e-14, gtol=1e-14, max_nfev=o["phase1_evals"])
            pv = np.asarray(ls.x)
            if o["verbose"]:
                e = at(pv, False)
                print("phase 1: %d evaluations, |p - p0| %.3e, infeasibility %.2e" % (ls.nfev, np.abs(pv / w - x0[p_cols]).max(),
                                                                                    feas(e["c"]) if e["ok"] else np.inf), flush=True)
        hist["phase"] = 2
        sf = o["obj_scale"]  # SLSQP's initial Hessian is the identity: the scale sets the length of its first steps
        for attempt in range(o["restarts"] + 1):
            try:
                res = so.minimize(lambda q: sf * fun(q), pv, jac=lambda q: sf * at(q, True)["g"], method="SLSQP", bounds=bounds, constraints=cons_sq,
                                  callback=callback, options={"maxiter": max(1, o["max_iter"] - hist["it"]), "ftol": o["slsqp_ftol"]})
                pv = np.asarray(res.x)
                message = "SLSQP: " + str(res.message)
            except Done:
                pass
            if hist["best"] is not None and hist["kkt"] <= o["tol"]:
                status, message = 0, "optimal: constraint violation <= %.0e, scaled optimality error <= %.0e" % (o["constr_tol"], o["tol"])
                break
            if hist["it"] >= o["max_iter"]:
                break
            if hist["best"] is not None:  # restart the quasi-Newton matrix from the best feasible point so far
                pv = hist["best"][0]
        if hist["best"] is not None:
            pv, e = hist["best"]
            if status != 0:
                status, message = 2, "feasible (violation <= %.0e) but not optimal: scaled optimality error %.1e" % (o["constr_tol"], hist["kkt"])
        else:
            e = at(pv, True)
            if not e["ok"]:
                e = {"x": S["x"], "obj": np.nan, "c": values(S["x"])[1]}
        sol = Solution()
        sol.xStar = xdict(e["x"])
        sol.fStar = e["obj"]
        sol.nit = hist["it"]
        sol.status = status
        sol.message = message
        sol.constr_violation = float(feas(e["c"]))
        sol.optimality = float(hist["kkt"])
        sol.reduced_evaluations = S["evals"]
        sol.q = np.array(pv)
        sol.optTime = time.perf_counter() - t_start
        sol.userObjTime, sol.userObjCalls = stat["obj_t"], stat["obj_n"]
        sol.userSensTime, sol.userSensCalls = stat["sens_t"], stat["sens_n"]
        return sol
