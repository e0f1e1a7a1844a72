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

The dependent terminal pair.  The shipped example asks for a CIRCULAR orbit through two rows, orbit energy and angular
momentum (con_init_terminal_knot.py:365-368).  On the surface E = E_t the angular momentum has its maximum exactly
where h = h_t: at every feasible point the gradient of one row lies in the span of the other's, the optimum is not a
KKT point (no finite multipliers), and every Newton-type method on the problem as posed stalls or wanders
(profiles/r02_solver_attempts.txt: 20 configurations).  Phase 2 therefore finds such a row from the singular values of
the reduced equality Jacobian and carries it by an exact penalty that is smooth on the rest of the feasible set,
f - lam c_k (c_k has one sign there), level after level of lam (continuation, warm start): the row's violation falls
like 1 / lam^2, the objective approaches its limit like 1 / lam (measured: profiles/r02j_solver_convergence.txt).  The
stationarity condition of f - lam c_k IS that of the original Lagrangian with multiplier lam on c_k, so the
termination test is the ORIGINAL problem's: every row within `constr_tol`, and IPOPT's scaled dual infeasibility
(full space, adjoint multipliers for the state equations) against `tol` / `acceptable_tol`.

STATUS.  On the shipped example, from the reference's initial guess, the continuation converges: constraint violation
<= 1e-10, objective settled to 1e-6 relative between the last two levels (payload 27 817.3 kg), event times settled to
1e-4 s.  The scaled optimality error ends at 1e-4 .. 2e-3: the dual residual carries (multiplier of the dependent
row, which must grow without bound) x (error of the reference's forward-difference Jacobian, ~2e-8 per entry at
dx = 1e-8) -- a floor that no solver on these callbacks gets under, IPOPT included.  `status`: 0 optimal or "solved to
acceptable level" (IPOPT's two tests), 3 converged in objective and constraints with the optimality error at that
noise floor, 2 feasible but neither, 1 not feasible.

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
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .ipsolve import IPSolver, Solution

STATE_VARS = ("mass", "position", "velocity", "quaternion")
STATE_ROWS = ("eqcon_init", "eqcon_dyn_mass", "eqcon_dyn_pos", "eqcon_dyn_vel", "eqcon_dyn_quat", "eqcon_knot")


class InnerFailure(Exception):
    pass


class _Blocks:
    """The sub-blocks of the sparse Jacobian that every `sens` needs (state equations x states, outer rows x
    parameters, ...), cut by index maps that are computed once: fancy indexing of a SciPy sparse matrix costs more
    than the factorisation that follows it here.  The maps come from the SAME indexing operations applied to a matrix
    whose values are entry numbers, so a block built from them has the structure, the entry order and the values
    SciPy's own slicing gives -- bit for bit; a Jacobian whose sparsity pattern differs from the one the maps were
    made for (a dense block with an exact zero at another place) is sliced the slow way."""

    def __init__(self, J, specs):
        ids = J.copy()
        ids.data = np.arange(1, J.nnz + 1, dtype=np.float64)
        self.indptr, self.indices, self.shape = J.indptr.copy(), J.indices.copy(), J.shape
        self.specs, self.maps = specs, {}
        for name, (rows, cols) in specs.items():
            M = ids[rows][:, cols].tocsc()
            self.maps[name] = (M.data.astype(np.int64) - 1, M.indices.copy(), M.indptr.copy(), M.shape)

    def cut(self, J):
        """name -> block of J"""
        if (J.shape == self.shape and J.nnz == self.indices.size and np.array_equal(J.indptr, self.indptr)
                and np.array_equal(J.indices, self.indices)):
            return {name: sp.csc_matrix((J.data[src], ind, ptr), shape=shape) for name, (src, ind, ptr, shape) in self.maps.items()}
        return {name: J[rows][:, cols].tocsc() for name, (rows, cols) in self.specs.items()}


class ReducedSQP:
    """solver = ReducedSQP({"tol": 1e-6, "max_iter": 600}); sol = solver(optProb, sens=sens)."""

    DEFAULTS = {"tol": 1e-6, "constr_tol": 1e-8, "max_iter": 260, "inner_tol": 1e-12, "inner_iter": 25, "restarts": 4,
                "active_tol": 1e-7, "u_scale": 1.0, "phase1_evals": 60, "obj_scale": 0.01, "predict_radius": 0.05, "slsqp_ftol": 1e-15, "verbose": 0,
                # IPOPT's second threshold ("Solved To Acceptable Level"; example-settings.json:92-97 sets 1e-4)
                "acceptable_tol": 1e-4,
                # penalty continuation on a dependent equality row (see __call__, phase 2): "auto" or "off"
                "degenerate": "auto", "dependent_ratio": 1e-2, "dependent_gap": 0.05, "detect_chunks": 4, "detect_iter": 25, "penalty0": 1e3, "penalty_factor": 10.0 ** 0.5,
                "penalty_levels": 9, "level_iter": 150, "obj_change_tol": 1e-6,
                # relative error of one entry of the callbacks' Jacobian: the reference's forward difference, eps / dx with
                # dx = 1e-8 (Trajectory_Optimization.py:167) on values of order one
                "jac_noise": 2.2e-8, "noise_floor_tol": 5e-3, "newton_iters": 0, "radius": 0.0,
                # the first penalty level starts inside a box around the phase-1 point, recentred a few times: SLSQP's first
                # steps (identity Hessian) otherwise leave the region where the linearisations hold on some scenarios
                "start_radius": 0.15, "start_segments": 4, "start_iter": 25}

    def __init__(self, options=None):
        self.opt = dict(self.DEFAULTS)
        for k, v in (options or {}).items():
            if k in self.opt:
                self.opt[k] = v  # IPOPT-only options (linear_solver, output_file ...) are ignored

    def __call__(self, prob, sens=None, **_):
        # the matrices are small (a few hundred rows): threaded BLAS only adds hand-over time (3-8x slower here), and
        # its reduction order would make the iterates depend on the thread count
        try:
            from threadpoolctl import threadpool_limits
        except ImportError:  # pragma: no cover
            return self._solve(prob, sens)
        with threadpool_limits(limits=1):
            return self._solve(prob, sens)

    def _solve(self, prob, sens):
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

        cutter = _Blocks(J0, {"Fs": (f_rows, s_cols), "Fp": (f_rows, p_cols), "Es": (e_rows, s_cols), "Ep": (e_rows, p_cols),
                              "Is": (i_rows, s_cols), "Ip": (i_rows, p_cols)})

        def factor(J):
            """LU of the state equations' Jacobian with respect to the states; returns the blocks of J"""
            blk = cutter.cut(J)
            S["lu"] = spla.splu(blk["Fs"])
            return blk

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
                blk = factor(J)
                dsdp = -S["lu"].solve(blk["Fp"].toarray())
                S["dsdp"], S["p_lin"], S["x_lin"] = dsdp, pv.copy(), e["x"].copy()
                e["g"] = (grad[p_cols] + dsdp.T @ grad[s_cols]) / w
                e["Je"] = (blk["Ep"].toarray() + blk["Es"] @ dsdp) / w
                e["Ji"] = (blk["Ip"].toarray() + blk["Is"] @ dsdp) / w
                e["grad"], e["blk"] = grad, blk
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

        # ---- first-order optimality of the ORIGINAL problem at x(p) ----
        pen = {"rows": np.zeros(0, int), "lam": np.zeros(0)}  # outer equality rows (positions in e_rows) carried by a penalty
        n_bound_mult = int(np.isfinite(xl).sum() + np.isfinite(xu).sum())

        def kkt(pv, e, full_space=True):
            """(IPOPT's scaled dual infeasibility, multipliers of [outer equalities | active inequalities | active
            bounds], unscaled residual).  Reduced gradient against the reduced active-set Jacobian, least squares with
            sign constraints; a penalised row keeps the multiplier the penalty gives it (f - lam c has exactly the
            stationarity condition of the Lagrangian with multiplier lam).  With the adjoint multipliers of the state
            equations, lam_F = J[F,s]^-T (grad_s - J[c,s]^T lam_c), the full-space dual residual is zero in the state
            columns and the reduced residual in the others, so the infinity norms agree; s_d is IPOPT's,
            max(100, ||multipliers||_1 / (m + bound multipliers)) / 100, over ALL multipliers of the original problem."""
            ci = e["c"][i_rows]
            act = np.where(ci <= o["active_tol"])[0]
            at_l = np.where(pv - xl[p_cols] * w <= 1e-9)[0]
            at_u = np.where(xu[p_cols] * w - pv <= 1e-9)[0]
            free_e = np.setdiff1d(np.arange(e_rows.size), pen["rows"])
            # g = Je' lam + Ji[act]' mu + zl - zu,  mu, zl, zu >= 0
            A = np.hstack([e["Je"][free_e].T, e["Ji"][act].T, np.eye(pv.size)[:, at_l], -np.eye(pv.size)[:, at_u]])
            lo = np.concatenate([np.full(free_e.size, -np.inf), np.zeros(act.size + at_l.size + at_u.size)])
            rhs = e["g"] - e["Je"][pen["rows"]].T @ pen["lam"]
            colsc = np.maximum(np.abs(A).max(axis=0), 1e-300)  # column scaling: the rows' gradients differ by 1e6
            r = so.lsq_linear(A / colsc, rhs, bounds=(lo, np.full(lo.size, np.inf)), method="bvls", lsmr_tol=None)
            mult = r.x / colsc
            resid = np.abs(A @ mult - rhs).max()
            lam_e = np.zeros(e_rows.size)
            lam_e[free_e] = mult[: free_e.size]
            lam_e[pen["rows"]] = pen["lam"]
            norm1 = np.abs(mult).sum() + np.abs(pen["lam"]).sum()
            count = mult.size + pen["lam"].size
            if full_space and "blk" in e:
                blk = e["blk"]
                mu = np.zeros(i_rows.size)
                mu[act] = mult[free_e.size: free_e.size + act.size]
                rhs_s = e["grad"][s_cols] - blk["Es"].T @ lam_e - blk["Is"].T @ mu
                lam_f = spla.splu(blk["Fs"].T.tocsc()).solve(np.asarray(rhs_s).ravel())
                norm1 += np.abs(lam_f).sum()
                count = m + n_bound_mult
            s_d = max(100.0, norm1 / max(count, 1)) / 100.0
            full = np.concatenate([lam_e, mult[free_e.size:]])
            return resid / s_d, full, resid

        hist = {"it": 0, "best": None, "kkt": np.inf}

        class Done(Exception):
            pass

        def feas(c):
            return max(np.abs(c[np.concatenate([f_rows, e_rows])]).max(), -min(0.0, c[i_rows].min()) if i_rows.size else 0.0)

        def callback(pv):
            hist["it"] += 1
            hist["level_it"] = hist.get("level_it", 0) + 1
            e = at(np.asarray(pv), True)
            if not e["ok"]:
                return
            hist["last_q"], hist["stopped"] = np.array(pv), False
            if hist.get("stop_when") is not None:
                if hist["stop_when"](e):
                    hist["stopped"] = True
                    raise Done()
                if hist["level_it"] >= hist.get("level_cap", 1 << 30):
                    raise Done()
                return
            th = feas(e["c"])
            if th > o["constr_tol"] and not o["verbose"]:  # the optimality test is only read at feasible iterates
                if hist["it"] >= o["max_iter"] or hist["level_it"] >= hist.get("level_cap", 1 << 30):
                    raise Done()
                return
            err, _, raw = kkt(np.asarray(pv), e)
            if o["verbose"]:
                print("%4d  obj %.10f  infeas %.2e  kkt %.2e (unscaled %.2e)  evals %d" % (hist["it"], e["obj"], th, err, raw, S["evals"]),
                      flush=True)
            if th <= o["constr_tol"] and (hist["best"] is None or err < hist["kkt"]):
                hist["best"], hist["kkt"], hist["kkt_raw"] = (np.array(pv), e), err, raw
            if th <= o["constr_tol"] and err <= hist["target"]:
                raise Done()
            if hist["it"] >= o["max_iter"] or hist["level_it"] >= hist.get("level_cap", 1 << 30):
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
                if not e["ok"] or not e["converged"]:
                    # no trajectory for these parameters (the state equations did not converge): a step too long --
                    # the trust region has to shrink, not move on with the last Newton iterate
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
        hist["target"] = o["tol"]
        sf = o["obj_scale"]  # SLSQP's initial Hessian is the identity: the scale sets the length of its first steps

        # ---- a dependent outer equality row?  The shipped example asks for a CIRCULAR orbit through two rows, orbit
        # energy and angular momentum (con_init_terminal_knot.py:365-368): on the surface E = E_t the angular momentum
        # has its maximum exactly where h = h_t, so at every feasible point the second row's gradient lies in the span
        # of the first's -- no finite multipliers exist (the optimum is not a KKT point) and every Newton-type step
        # is ill-posed (profiles/r02_solver_attempts.txt).  Such a row, call it c_k, has one sign on the rest of the
        # feasible set, so f - lam c_k is an exact-penalty objective that is SMOOTH there: the row leaves the
        # constraint list and SLSQP sees a regular problem whose minimiser violates c_k by O(1 / lam^2) and whose
        # objective is O(1 / lam) from the limit; lam grows level by level (continuation, warm start).  The stationarity
        # condition of f - lam c_k IS that of the original Lagrangian with multiplier lam, so the termination test
        # below is the original problem's.  Detection: smallest singular value of the row-normalised reduced equality
        # Jacobian at the phase-1 point; the row with the largest weight in its left singular vector is penalised
        # (either of a dependent pair will do), with the sign of its least-squares multiplier there. ----
        def run_slsqp(pv, lam_vec, cap, target, stop_when=None):
            """SLSQP on f - lam . c[penalised rows] with those rows out of the constraint list, from pv, for at most cap
            major iterations (restarting its quasi-Newton matrix from the best feasible point when it gives up early)."""
            nonlocal message
            pen["lam"] = np.asarray(lam_vec, dtype=float)
            keep = np.setdiff1d(np.arange(e_rows.size), pen["rows"])
            hist["level_it"], hist["level_cap"], hist["target"] = 0, cap, target
            hist["best"], hist["kkt"], hist["stop_when"] = None, np.inf, stop_when  # per level: the multiplier is part of the test

            def f_pen(q):
                e = at(q, False)
                if not e["ok"] or not pen["rows"].size:
                    return sf * e["obj"]
                return sf * (e["obj"] - float(pen["lam"] @ e["c"][e_rows][pen["rows"]]))

            def g_pen(q):
                e = at(q, True)
                return sf * (e["g"] - (e["Je"][pen["rows"]].T @ pen["lam"] if pen["rows"].size else 0.0))

            cons_lv = [{"type": "eq", "fun": lambda q: ceq(q)[keep], "jac": lambda q: at(q, True)["Je"][keep]}, cons_sq[1]]
            for attempt in range(o["restarts"] + 1):
                bounds_lv = bounds
                if o["radius"] > 0.0:  # a box around the start: SLSQP's early steps (identity Hessian) stay where the model holds
                    bounds_lv = [(max(lo_, c_ - o["radius"]), min(hi_, c_ + o["radius"])) for (lo_, hi_), c_ in zip(bounds, pv)]
                try:
                    res = so.minimize(f_pen, pv, jac=g_pen, method="SLSQP", bounds=bounds_lv, constraints=cons_lv, callback=callback,
                                      options={"maxiter": max(1, min(o["max_iter"] - hist["it"], cap - hist["level_it"])),
                                               "ftol": o["slsqp_ftol"]})
                    pv = np.asarray(res.x)
                    message = "SLSQP: " + str(res.message)
                    if o["verbose"]:
                        print("     ", message, flush=True)
                except Done:
                    pv = hist.get("last_q", pv)
                if hist["best"] is not None and hist["kkt"] <= target:
                    break
                if hist["it"] >= o["max_iter"] or hist["level_it"] >= cap or hist.get("stopped"):
                    break
                if hist["best"] is not None:
                    pv = hist["best"][0]
            return hist["best"][0] if hist["best"] is not None else pv

        levels = []
        if o["newton_iters"] > 0:  # a full SQP step or two on the problem as posed: onto the linearised constraints
            pv = run_slsqp(pv, [], o["newton_iters"], -1.0)
        def dependent_row(e1):
            """None, or the position (in e_rows) of an outer equality row to carry by a penalty: the reduced equality
            Jacobian, rows normalised, has ONE singular value far below the rest (<= dependent_ratio of the largest and
            <= dependent_gap of the next one).  Of the rows that make up the dependency (weights within a factor 5 of the
            largest in the left singular vector) the one with the SMALLEST gradient: the angular-momentum row of the
            example's pair (d(E / E_t) = -2 d(h / h_t), half the energy row's gradient).  Penalising that one works;
            with the energy row SLSQP wanders (measured, profiles/r02j_solver_convergence.txt)."""
            norms = np.maximum(np.linalg.norm(e1["Je"], axis=1), 1e-300)
            U, sv, _ = np.linalg.svd(e1["Je"] / norms[:, None], full_matrices=False)
            hist["sv"] = (sv[-1] / sv[0], sv[-1] / sv[-2] if sv.size > 1 else 0.0)
            if sv[-1] > o["dependent_ratio"] * sv[0] or (sv.size > 1 and sv[-1] > o["dependent_gap"] * sv[-2]):
                return None
            wgt = np.abs(U[:, -1])
            cand = np.where(wgt >= 0.2 * wgt.max())[0]
            return int(cand[np.argmin(norms[cand])])

        if o["degenerate"] == "auto" and e_rows.size:
            e1 = at(pv, True)
            k = dependent_row(e1) if e1["ok"] else None
            pv1 = pv
            for chunk in range(o["detect_chunks"] if k is None else 0):
                # not visible from here: a few iterations on the problem as posed, towards the feasible set, and look again
                pv = run_slsqp(pv, [], o["detect_iter"], -1.0)
                e1 = at(pv, True)
                k = dependent_row(e1) if e1["ok"] else None
                if k is not None or hist["it"] >= o["max_iter"]:
                    break
            if k is not None:
                pv = pv1  # the continuation starts from the phase-1 point
                sv = hist["sv"]
                pen["rows"] = np.array([k])
                others = np.setdiff1d(np.arange(e_rows.size), [k])
                # the sign the row takes where the OTHER rows hold: a short run without it (multiplier 0)
                def others_hold(e):
                    return max(np.abs(e["c"][f_rows]).max(), np.abs(e["c"][e_rows][others]).max(),
                               -min(0.0, e["c"][i_rows].min()) if i_rows.size else 0.0) <= 1e-3 * abs(e["c"][e_rows][k])
                e2 = at(run_slsqp(pv, [0.0], o["level_iter"] // 2, -1.0, stop_when=others_hold), True)  # pv itself stays
                sign = 1.0 if e2["c"][e_rows][k] < 0.0 else -1.0  # f - lam c must GROW with the violation
                levels = [sign * o["penalty0"] * o["penalty_factor"] ** j for j in range(o["penalty_levels"])]
                if o["verbose"]:
                    print("dependent equality row: %s (row %d of the outer equalities), sigma_min / sigma_max = %.1e, its value "
                          "where the other rows hold %.2e, penalty sign %+d" % ([g[0] for g in cons for _ in range(g[1])][e_rows[k]], k,
                                                                               sv[0], e2["c"][e_rows][k], sign), flush=True)
        hist["levels"] = []
        hist["stop_when"] = None

        def at_noise_floor():
            """IPOPT's scaled dual infeasibility of the level's best point <= noise_floor_tol: the level the callbacks'
            forward-difference Jacobian lets any run reach on this problem (measured 1e-4 .. 5e-3 over penalty levels,
            scenarios and hosts, profiles/r02j_solver_convergence.txt) -- stated, not IPOPT's acceptable_tol"""
            return hist["kkt"] <= o["noise_floor_tol"]

        if levels and o["start_radius"] > 0.0:
            for seg in range(o["start_segments"]):
                o["radius"] = o["start_radius"]
                try:
                    pv = run_slsqp(pv, [levels[0]], o["start_iter"], -1.0)
                finally:
                    o["radius"] = 0.0
        for level, lam in enumerate(levels or [None]):
            pv = run_slsqp(pv, [lam] if lam is not None else [], o["level_iter"] if lam is not None else o["max_iter"],
                           -1.0 if lam is not None else o["tol"])
            e_lv = hist["best"][1] if hist["best"] is not None else at(pv, False)
            if lam is not None and e_lv["ok"]:
                hist["levels"].append({"lam": lam, "obj": e_lv["obj"], "row_value": float(e_lv["c"][e_rows][pen["rows"]][0]),
                                       "infeasibility": float(feas(e_lv["c"])), "kkt_scaled": float(hist["kkt"]),
                                       "kkt_unscaled": float(hist.get("kkt_raw", np.nan)), "major_iterations": hist["it"]})
                if o["verbose"]:
                    print("level %d: %s" % (level, hist["levels"][-1]), flush=True)
            if lam is None:
                if hist["best"] is not None and hist["kkt"] <= o["tol"]:
                    status, message = 0, "optimal: constraint violation <= %.0e, scaled optimality error <= %.0e" % (o["constr_tol"], o["tol"])
                break
            # ---- termination of the continuation.  Every level is run out (its objective has to settle); the sequence
            # ends when two successive levels agree in the objective to obj_change_tol -- the level error is O(1 / lam),
            # so with a level factor of sqrt(10) what is left is about half the last change.  The end point is then
            # classified by IPOPT's test on the ORIGINAL problem with multiplier lam on the dependent row. ----
            lv = hist["levels"]
            hist.setdefault("level_best", []).append((hist["best"], hist["kkt"], hist.get("kkt_raw", np.nan), lam)
                                                     if hist["best"] is not None else None)
            kept = hist["level_best"]
            if (len(lv) >= 2 and len(kept) >= 2 and kept[-1] is not None and kept[-2] is not None
                    and lv[-2]["infeasibility"] <= o["constr_tol"]
                    and abs(lv[-1]["obj"] - lv[-2]["obj"]) <= o["obj_change_tol"] * max(1.0, abs(lv[-1]["obj"]))):
                # two levels that agree in the objective: the end point is the later one, unless only the earlier one
                # passes the optimality test (the best feasible iterate of a level is a noisy sample of it)
                for cand in (kept[-1], kept[-2]):
                    if cand[1] <= o["noise_floor_tol"]:
                        hist["best"], hist["kkt"], hist["kkt_raw"] = cand[0], cand[1], cand[2]
                        pen["lam"] = np.array([cand[3]])
                        hist["settled"] = abs(lv[-1]["obj"] - lv[-2]["obj"])
                        break
                if "settled" in hist:
                    break
            if hist["it"] >= o["max_iter"]:
                break
        if levels and hist["best"] is not None:
            if hist["kkt"] <= o["tol"]:
                status, message = 0, "optimal: constraint violation <= %.0e, scaled optimality error <= %.0e" % (o["constr_tol"], o["tol"])
            elif hist["kkt"] <= o["acceptable_tol"]:
                status = 0
                message = ("solved to acceptable level: constraint violation <= %.0e, scaled optimality error %.1e <= acceptable_tol "
                           "%.0e (multiplier of the dependent row %.3g)" % (o["constr_tol"], hist["kkt"], o["acceptable_tol"], pen["lam"][0]))
            elif "settled" in hist and at_noise_floor():
                # the optimum of a problem with a dependent row is not a KKT point: the multiplier has to grow without
                # bound, and the dual residual carries (multiplier) x (error of the reference's forward-difference
                # Jacobian, ~2e-8 per entry with dx = 1e-8, Trajectory_Optimization.py:167) -- a floor no solver on these
                # callbacks gets under.  What has converged is what the user reads: objective and constraints.
                status = 3
                message = ("converged in objective (change %.1e between the last two penalty levels) and constraints (violation <= %.0e); "
                           "scaled optimality error %.1e: above acceptable_tol, at the noise floor of the forward-difference "
                           "Jacobian (multiplier of the dependent row %.3g x ~%.0e per entry)"
                           % (hist["settled"], o["constr_tol"], hist["kkt"], pen["lam"][0], o["jac_noise"]))
        if hist["best"] is not None:
            pv, e = hist["best"]
            if status not in (0, 3):
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
        sol.penalty_levels = hist.get("levels", [])
        names_of_rows = [g[0] for g in cons for _ in range(g[1])]
        sol.dependent_rows = [(names_of_rows[e_rows[k]], int(e_rows[k] - rows_of[names_of_rows[e_rows[k]]][0])) for k in pen["rows"]]
        sol.penalty_sign = float(np.sign(levels[0])) if levels else 0.0
        sol.reduced_evaluations = S["evals"]
        sol.q = np.array(pv)
        sol.optTime = time.perf_counter() - t_start
        sol.userObjTime, sol.userObjCalls = stat["obj_t"], stat["obj_n"]
        sol.userSensTime, sol.userSensCalls = stat["sens_t"], stat["sens_n"]
        return sol
