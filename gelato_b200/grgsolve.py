"""A host-side reduced-space NLP solver for the callbacks of this package -- NOT IPOPT.

The reference hands `objfunc` / `sens` to pyoptsparse's IPOPT wrapper
(/root/reference/Trajectory_Optimization.py:454-458, options `example-settings.json:92-97`).  Neither
pyoptsparse nor IPOPT can be installed in this image, so converged solutions -- payload mass, event times,
callback counts, time spent in the callbacks, solves per hour -- are produced with this stand-in instead.
It is labelled NOT-IPOPT wherever its numbers appear.

Method (generalised reduced gradient; Abadie & Carpentier, Lasdon et al.): a transcribed trajectory problem
has almost as many equality rows as variables (the shipped example: 964 collocation / knot / boundary rows,
1 003 variables, 39 degrees of freedom).  The variables are split into m_E DEPENDENT ones (chosen once by a
column-pivoted QR of the equality Jacobian, so that their square block c_y is well conditioned) and the few
INDEPENDENT ones u.  For given u the dependent variables solve c_E(y, u) = 0 by Newton's method on the sparse
block c_y (scipy SuperLU: the sparse factorisation stays on the host, as the north star prescribes), which
makes every iterate feasible to ~1e-12; the outer problem

        min f(y(u), u)   s.t.   c_I(y(u), u) >= 0,   bounds

has a few dozen variables, reduced gradients from one adjoint solve per row (c_y^-T), and is solved by a dense
SQP method (scipy SLSQP).  Every residual and every Jacobian value comes from the callbacks -- one `objfunc` +
one `sens` per Newton iteration -- so the CPU oracle and the CUDA kernels drive it identically.

Interface: `GRGSolver(options)(optProb, sens=sens) -> Solution`, for `nlpshim.Optimization` problems (the
call the reference makes on pyoptsparse's classes).
"""
import time

import numpy as np
import scipy.optimize as so
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .ipsolve import IPSolver, Solution


class _Infeasible(Exception):
    pass


class GRGSolver:
    # outer_eq: equality groups kept as constraints of the OUTER problem instead of being solved by the inner Newton
    # iteration -- the event-point rows (terminal orbit, waypoints, user constraint), which for given controls are
    # shooting conditions: far from the solution they have no root nearby.  state_vars: variable groups that are
    # always dependent (the collocation rows are an implicit integration of them for given controls).
    DEFAULTS = {"tol": 1e-6, "max_iter": 400, "newton_tol": 1e-11, "newton_iter": 25, "verbose": 0, "ftol": 1e-13,
                "outer_eq": ("eqcon_terminal", "eqcon_pos", "eqcon_iip", "eqcon_user"),
                "state_vars": ("mass", "position", "velocity", "quaternion")}

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
        eq_all = [g for g in prob.cons if g[3] is not None and g[2] == g[3]]
        ineq = [g for g in prob.cons if g not in eq_all]
        eq = [g for g in eq_all if g[0] not in o["outer_eq"]]     # solved by the inner Newton iteration
        oeq = [g for g in eq_all if g[0] in o["outer_eq"]]        # equality constraints of the outer problem
        for g in ineq:
            if g[3] is not None or g[2] != 0.0:
                raise NotImplementedError("inequality groups other than c(x) >= 0")
        for g in eq:
            if g[2] != 0.0:
                raise NotImplementedError("equality groups other than c(x) = 0")
        mE, mI, mO = sum(g[1] for g in eq), sum(g[1] for g in ineq), sum(g[1] for g in oeq)
        stat = {"obj_t": 0.0, "obj_n": 0, "sens_t": 0.0, "sens_n": 0, "newton": 0}

        def xdict(xv):
            return {nm: xv[offs[i]: offs[i + 1]].copy() for i, nm in enumerate(names)}

        def evaluate(xv, want_jac):
            xd = xdict(xv)
            t0 = time.perf_counter()
            f, fail = prob.objfunc(xd)
            stat["obj_t"] += time.perf_counter() - t0
            stat["obj_n"] += 1
            assert not fail
            cE = np.concatenate([np.atleast_1d(np.asarray(f[g[0]], dtype=float)) for g in eq]) if eq else np.zeros(0)
            # outer rows: the shooting equalities first, then the inequalities
            cI = np.concatenate([np.atleast_1d(np.asarray(f[g[0]], dtype=float)) for g in oeq + ineq]) if (oeq or ineq) else np.zeros(0)
            obj = float(np.asarray(f[prob.obj]).ravel()[0])
            if not want_jac:
                return obj, cE, cI, None, None, None
            t0 = time.perf_counter()
            s, fail = sens(xd, f)
            stat["sens_t"] += time.perf_counter() - t0
            stat["sens_n"] += 1
            assert not fail
            grad = np.zeros(n)
            for var, blk in s[prob.obj].items():
                grad[col0[var]: col0[var] + np.size(blk)] = np.ravel(blk)
            return obj, cE, cI, grad, IPSolver._jac(s, eq, mE, n, col0).tocsc(), IPSolver._jac(s, oeq + ineq, mO + mI, n, col0).tocsc()

        # ---- dependent / independent split: column-pivoted QR of the equality Jacobian at the start ----
        import scipy.linalg as sla

        if n > 6000:
            raise NotImplementedError("the dense basis selection is meant for transcriptions of a few thousand variables")
        _, _, _, _, JE0, _ = evaluate(x0, True)
        weight = np.ones(n)
        for nm in o["state_vars"]:
            if nm in col0:
                weight[col0[nm]: col0[nm] + sizes[names.index(nm)]] = 1e4  # pivots first: always dependent
        R, piv = sla.qr(JE0.toarray() * weight, mode="r", pivoting=True)
        d = np.abs(np.diag(R))
        rank = int((d > 1e-11 * d[0]).sum())
        if rank < mE:
            raise ValueError("the equality Jacobian is rank deficient at the starting point (%d of %d rows)" % (rank, mE))
        dep, ind = np.sort(piv[:mE]), np.sort(piv[mE:])
        k = ind.size

        cache = {}

        def restore(u, y_start):
            """Newton on the dependent variables: c_E(y, u) = 0.  Returns everything evaluated at the solution."""
            x = np.empty(n)
            x[ind] = u
            x[dep] = y_start
            best = None
            for itn in range(o["newton_iter"]):
                obj, cE, cI, g, JE, JI = evaluate(x, True)
                stat["newton"] += 1
                res = float(np.abs(cE).max(initial=0.0))
                if not np.isfinite(res):
                    raise _Infeasible()
                if res <= o["newton_tol"]:
                    lu = spla.splu(JE[:, dep].tocsc())
                    return x, obj, cI, g, JE, JI, lu
                lu = spla.splu(JE[:, dep].tocsc())
                dy = lu.solve(-cE)
                if not np.all(np.isfinite(dy)):
                    raise _Infeasible()
                # damped: the step is halved while it does not reduce the residual
                a = 1.0
                for _ in range(12):
                    xt = x.copy()
                    xt[dep] = x[dep] + a * dy
                    _, cEt, _, _, _, _ = evaluate(xt, False)
                    rt = float(np.abs(cEt).max(initial=0.0))
                    if np.isfinite(rt) and rt < (1.0 - 1e-4 * a) * res:
                        break
                    a *= 0.5
                else:
                    raise _Infeasible()
                x = xt
                if best is not None and res > 0.9 * best and itn > 15:
                    raise _Infeasible()
                best = res if best is None else min(best, res)
            raise _Infeasible()

        state = {"y": x0[dep].copy(), "last_ok": None}

        def at(u):
            key = u.tobytes()
            if key not in cache:
                try:
                    x, obj, cI, g, JE, JI, lu = restore(u, state["y"])
                except _Infeasible:
                    cache.clear()
                    cache[key] = None
                    return None
                # reduced gradients: one adjoint solve per row (objective + inequality rows)
                rows = sp.vstack((sp.csr_matrix(g[dep]), JI[:, dep].tocsr())).toarray()  # (1 + mI) x mE
                lam = lu.solve(rows.T, trans="T")                                         # c_y^-T rows^T
                Nmat = JE[:, ind]
                red = np.vstack((g[ind][None, :], JI[:, ind].toarray())) - (Nmat.T @ lam).T
                cache.clear()
                cache[key] = (x, obj, cI, red)
                state["y"] = x[dep].copy()
                state["last_ok"] = cache[key]
            return cache[key]

        big = {"f": None}

        def fun(u):
            r = at(u)
            if r is None:
                return big["f"] + 1.0 if big["f"] is not None else 1e6
            big["f"] = r[1] if big["f"] is None else max(big["f"], r[1])
            return r[1]

        def jac(u):
            r = at(u) or state["last_ok"]
            return r[3][0]

        def con(u):
            r = at(u)
            if r is None:
                return -np.ones(mI)
            return r[2][mO:]

        def con_jac(u):
            r = at(u) or state["last_ok"]
            return r[3][1 + mO:]

        def ceq(u):
            r = at(u)
            if r is None:
                return np.ones(mO)
            return r[2][:mO]

        def ceq_jac(u):
            r = at(u) or state["last_ok"]
            return r[3][1: 1 + mO]

        history = []

        def callback(u):
            r = state["last_ok"]
            if r is not None:
                vio = float(max(np.abs(r[2][:mO]).max(initial=0.0), np.maximum(-r[2][mO:], 0.0).max(initial=0.0)))
                history.append((r[1], vio))
                if o["verbose"] and len(history) % o["verbose"] == 0:
                    print("it %4d  obj %.10f  outer violation %.2e  newton %d" % (len(history), r[1], vio, stat["newton"]))

        u0 = x0[ind].copy()
        if at(u0) is None:
            raise RuntimeError("Newton's method could not make the starting point feasible")
        res = so.minimize(fun, u0, jac=jac, method="SLSQP", bounds=list(zip(xl[ind], xu[ind])),
                          constraints=([{"type": "eq", "fun": ceq, "jac": ceq_jac}] if mO else [])
                          + ([{"type": "ineq", "fun": con, "jac": con_jac}] if mI else []),
                          options={"maxiter": o["max_iter"], "ftol": o["ftol"]}, callback=callback)
        r = at(res.x) or state["last_ok"]
        x, obj, cI, red = r
        # IPOPT-style optimality error of the result: least-squares multipliers of the active rows
        _, cE, cO, g, JE, JI = evaluate(x, True)[:6]
        cI = cO[mO:]
        act = np.concatenate((np.ones(mO, dtype=bool), cI <= 1e-7))
        A = sp.vstack((JE, JI.tocsr()[np.flatnonzero(act)])).tocsc() if act.any() else JE
        K = sp.bmat([[sp.identity(n), A.T], [A, -1e-12 * sp.identity(A.shape[0])]], format="csc")
        v = spla.splu(K).solve(np.concatenate((-g, np.zeros(A.shape[0]))))
        dual = float(np.abs(v[:n]).max())

        sol = Solution()
        sol.xStar = xdict(x)
        sol.fStar = obj
        sol.optTime = time.perf_counter() - t_start
        sol.userObjTime, sol.userObjCalls = stat["obj_t"], stat["obj_n"]
        sol.userSensTime, sol.userSensCalls = stat["sens_t"], stat["sens_n"]
        sol.constr_violation = float(max(np.abs(cE).max(initial=0.0), np.abs(cO[:mO]).max(initial=0.0),
                                         np.maximum(-cI, 0.0).max(initial=0.0)))
        sol.dual_infeasibility = dual
        sol.bound_violation = float(max(np.maximum(xl - x, 0.0).max(), np.maximum(x - xu, 0.0).max()))
        sol.nit, sol.status, sol.message, sol.history = int(res.nit), int(res.status), str(res.message), history
        sol.newton_iterations = stat["newton"]
        sol.independent_variables = int(k)
        sol.optInform = {"value": int(res.status), "text": str(res.message)}
        return sol
