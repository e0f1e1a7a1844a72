"""An EXPERIMENTAL host-side interior-point NLP solver for the callbacks of this package -- NOT IPOPT, and on the
shipped example it does NOT reach IPOPT's tolerance (status below).

The reference hands `objfunc` / `sens` to pyoptsparse's IPOPT wrapper
(/root/reference/Trajectory_Optimization.py:454-458; options `example-settings.json:92-97`: tol 1e-6,
limited-memory Hessian -- pyoptsparse's default, the callbacks give first derivatives only --, MUMPS for the
sparse KKT systems).  Neither pyoptsparse nor IPOPT can be installed in this image (no wheel, no network), so this
module restates the METHOD (Waechter & Biegler 2006) in numpy / scipy to drive the very same callbacks, CPU oracle or
CUDA, through a realistic solve loop: same iterates from both (the callbacks are bit-identical), callback counts
and times under pyoptsparse's names.

    min f(x)  s.t.  c_E(x) = 0,  c_I(x) >= 0,  x_L <= x <= x_U

* primal-dual log-barrier with slacks for the inequality rows, monotone (Fiacco-McCormick) barrier update,
  fraction-to-the-boundary rule, gradient-based row scaling (IPOPT's nlp_scaling_max_gradient = 100);
* Hessian of the Lagrangian, three models: `lbfgs` (default; damped limited-memory BFGS in compact form, the KKT matrix
  [[sigma I + Sigma, J^T], [J, -D]] stays SPARSE -- scipy SuperLU -- and the rank-2m part goes through the Woodbury
  identity: the sparse KKT solve stays on the host, as the north star prescribes); `sparse-fd` (the exact sparse
  Hessian from 13 + S + 1 structurally coloured finite differences of the Lagrangian's gradient: cheap with batched
  CUDA callbacks); `reduced-fd` (the exact Hessian on the null space of the equality Jacobian);
* globalisation: the filter line search of the paper (or an l1 merit function) with a second-order correction;
  when it fails the quasi-Newton memory is dropped, then least-norm steps towards feasibility stand in for the
  restoration phase.

STATUS (round 2, profiles/r02_solver_attempts.txt): on the shipped example (1 003 variables, 964 equality rows, 39
degrees of freedom) no configuration converges to tol = 1e-6.  The default reaches objective -1.0209 with a constraint
violation of ~3e-3 in 300 iterations and then creeps; scipy's trust-constr creeps likewise (violation 3.8e-4 after
3 000 iterations).  What was learnt: (1) the two terminal rows -- orbit energy and angular momentum,
con_init_terminal_knot.py:362-370 -- have nearly parallel gradients at a near-circular target orbit, so the equality
Jacobian's smallest singular value is 3e-5 of its largest and their multipliers are ~1e4..1e5: every Newton-type
step is dominated by them; (2) the loose variable bounds of the reference's registration (position +-10 Earth radii
...) dominate the barrier function at mu = 0.1 and pull the iterates to the centre of the box, so by default bounds get
no barrier term here (`ignore_bounds`; the result reports `bound_violation`); (3) with exact curvature the
tangent-space Hessian has eigenvalues down to 1e-6 of the largest far from the solution, so plain line-search Newton
steps leave the region where the linearisation means anything.  IPOPT deals with all three (inertia and delta_c
regularisation, a true restoration phase, adaptive scaling); a converged-solution comparison needs IPOPT itself:
`nlpshim.register` is the reference's registration block for any pyoptsparse-compatible class.

Interface: `IPSolver(options)(optProb, sens=sens) -> Solution`, for `nlpshim.Optimization` problems
(the same call the reference makes on pyoptsparse's classes).
"""
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Solution:
    pass


class _LBFGS:
    """Damped limited-memory BFGS approximation B = sigma I - U M^-1 U^T of the Hessian of the Lagrangian
    (compact form of Byrd, Nocedal & Schnabel 1994), U = [sigma S, Y]."""

    def __init__(self, n, memory, sigma_min=1e-8, sigma_max=1e8):
        self.n, self.m = n, memory
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.reset()

    def reset(self, sigma=1.0):
        self.S, self.Y = [], []
        self.sigma = sigma
        self._U = self._M = None

    def times(self, v):
        if not self.S:
            return self.sigma * v
        U, M = self.factors()
        return self.sigma * v - U @ np.linalg.solve(M, U.T @ v)

    def update(self, s, y):
        ss = float(s @ s)
        if ss < 1e-30:
            return
        Bs = self.times(s)
        sBs = float(s @ Bs)
        sy = float(s @ y)
        if sy < 0.2 * sBs:  # Powell damping keeps the approximation positive definite
            theta = 0.8 * sBs / (sBs - sy)
            y = theta * y + (1.0 - theta) * Bs
            sy = float(s @ y)
        if sy <= 1e-12 * np.sqrt(ss * float(y @ y)):
            return
        self.S.append(s.copy())
        self.Y.append(y.copy())
        if len(self.S) > self.m:
            self.S.pop(0)
            self.Y.pop(0)
        self.sigma = min(max(sy / ss, self.sigma_min), self.sigma_max)  # IPOPT limited_memory_initialization = scalar1
        self._U = None

    def factors(self):
        """(U [n, 2k], M [2k, 2k]) with B = sigma I - U M^-1 U^T."""
        if self._U is None:
            S, Y = np.array(self.S).T, np.array(self.Y).T
            SY = S.T @ Y
            Lm = np.tril(SY, -1)
            D = np.diag(np.diag(SY))
            sg = self.sigma
            self._U = np.hstack((sg * S, Y))
            self._M = np.block([[sg * (S.T @ S), Lm], [Lm.T, -D]])
        return self._U, self._M


class _ReducedHessian:
    """W ~ sigma I + Z C Z^T: the EXACT Hessian of the Lagrangian on the null space of the equality Jacobian (Z: an
    orthonormal basis, n x k with k = n - m_E, a few dozen for a transcribed trajectory problem), sigma I on its
    complement.  Z^T W Z comes from k directional finite differences of the Lagrangian's gradient (k extra `sens`
    calls per iteration -- one batched launch for the CUDA callbacks), is symmetrised and its eigenvalues are lifted to
    stay positive, so the KKT matrix keeps the right inertia without a correction loop.  The cross term Z^T W Y is
    dropped, as in reduced-Hessian SQP."""

    def __init__(self, n, floor_rel=1e-6):
        self.n = n
        self.floor_rel = floor_rel
        self.sigma = 1.0
        self.Z = None
        self.C = None
        self.S = []  # (interface of _LBFGS, for the log line)
        self.idx = None

    def reset(self, sigma=None):
        pass

    def update(self, s, y):
        pass

    def times(self, v):
        if self.Z is None:
            return self.sigma * v
        return self.sigma * v + self.Z @ (self.C @ (self.Z.T @ v))

    def additive(self):
        """(U, C^-1) with W = sigma I + U C U^T, or None."""
        if self.Z is None:
            return None
        return self.Z, np.linalg.inv(self.C)

    def choose_basis(self, JE):
        """Independent variables: the columns a pivoted QR of the equality Jacobian leaves for last."""
        import scipy.linalg as sla

        if JE.shape[1] > 6000:
            raise NotImplementedError("dense basis selection is meant for transcriptions of a few thousand variables")
        R, piv = sla.qr(JE.toarray(), mode="r", pivoting=True)
        d = np.abs(np.diag(R))
        rank = int((d > 1e-11 * d[0]).sum())
        self.idx = np.sort(piv[rank:])

    def build(self, JE, lagr_grad, x, eps):
        """lagr_grad(x) -> J_E(x)^T lam_E + J_I(x)^T lam_I with the CURRENT multipliers (the objective is linear)."""
        n, mE = self.n, JE.shape[0]
        if self.idx is None:
            self.choose_basis(JE)
        k = self.idx.size
        if k == 0:
            self.Z = None
            return 0
        K = sp.bmat([[sp.identity(n), JE.T], [JE, -1e-12 * sp.identity(mE)]], format="csc")
        E = np.zeros((n + mE, k))
        E[self.idx, np.arange(k)] = 1.0
        Zr = spla.splu(K).solve(E)[:n]
        Z, R = np.linalg.qr(Zr)
        if np.abs(np.diag(R)).min() < 1e-8 * np.abs(np.diag(R)).max():  # the independent set went stale: choose again
            self.choose_basis(JE)
            return self.build(JE, lagr_grad, x, eps)
        g0 = lagr_grad(x)
        H = np.empty((k, k))
        for i in range(k):
            H[:, i] = Z.T @ ((lagr_grad(x + eps * Z[:, i]) - g0) / eps)
        H = 0.5 * (H + H.T)
        w, V = np.linalg.eigh(H)
        floor = max(1e-8, self.floor_rel * float(np.abs(w).max()))
        w = np.maximum(np.abs(w), floor)  # |lambda|: a saddle direction is treated as a convex one of the same curvature
        self.sigma = 0.5 * float(w.min())
        self.Z = Z
        self.C = (V * (w - self.sigma)) @ V.T
        return k


class _SparseFDHessian:
    """The Hessian of the Lagrangian as a SPARSE matrix, from finite differences of the Lagrangian's gradient with a
    structural colouring.  In a transcribed trajectory problem the only non-linear couplings are inside one
    collocation node -- its state row (mass, position, velocity, quaternion: 11 variables) and control row (2) --
    and between a node and the event times (everything else: D.X, knots, rates, are linear).  So perturbing variable
    c of EVERY node at once gives column c of every node block from one `sens` call (13 calls), and one call per
    event time gives the time columns: 13 + S + 1 extra Jacobian evaluations per iteration -- one batched launch for
    the CUDA callbacks -- buy a Newton step instead of a quasi-Newton one."""

    def __init__(self, n, structure, col0, eps):
        self.n, self.eps = n, eps
        M, N, S = structure["M"], structure["N"], structure["S"]
        o = {k: col0[k] for k in ("mass", "position", "velocity", "quaternion", "u", "t")}
        self.t_idx = o["t"] + np.arange(S + 1)
        # members[r] = variables of the group of state row r (control row of its node included)
        node_of_row = np.full(M, -1)
        for ua, xa, nn in structure["sections"]:
            node_of_row[xa + 1: xa + 1 + nn] = ua + np.arange(nn)
        self.colours = []  # (perturbed variables, rows of H, columns of H)
        members = []
        for r in range(M):
            m = [o["mass"] + r] + [o["position"] + 3 * r + k for k in range(3)] + [o["velocity"] + 3 * r + k for k in range(3)] \
                + [o["quaternion"] + 4 * r + k for k in range(4)]
            if node_of_row[r] >= 0:
                m += [o["u"] + 2 * node_of_row[r], o["u"] + 2 * node_of_row[r] + 1]
            members.append(np.array(m))
        for c in range(13):
            pert, rows, cols = [], [], []
            for r in range(M):
                if c >= members[r].size:
                    continue
                v = members[r][c]
                pert.append(v)
                rows.append(members[r])
                cols.append(np.full(members[r].size, v))
            self.colours.append((np.array(pert), np.concatenate(rows), np.concatenate(cols)))
        self.S = []
        self.sigma = 0.0
        self.W = sp.csc_matrix((n, n))

    def reset(self, sigma=None):
        pass

    def update(self, s, y):
        pass

    def times(self, v):
        return self.W @ v

    def build(self, lagr_grad, x):
        n, eps = self.n, self.eps
        g0 = lagr_grad(x)
        R, C, V = [], [], []
        for pert, rows, cols in self.colours:
            xp = x.copy()
            xp[pert] += eps
            v = (lagr_grad(xp) - g0) / eps
            R.append(rows)
            C.append(cols)
            V.append(v[rows])
        for tk in self.t_idx:
            xp = x.copy()
            xp[tk] += eps
            v = (lagr_grad(xp) - g0) / eps
            nz = np.flatnonzero(v)
            R += [nz, np.full(nz.size, tk)]
            C += [np.full(nz.size, tk), nz]
            V += [v[nz], v[nz]]
        H = sp.coo_matrix((np.concatenate(V), (np.concatenate(R), np.concatenate(C))), shape=(n, n)).tocsc()
        # the time x time block was entered twice (as column and as row)
        T = sp.csc_matrix((np.ones(self.t_idx.size), (self.t_idx, self.t_idx)), shape=(n, n))
        H = H - 0.5 * (T @ H @ T)
        self.W = (0.5 * (H + H.T)).tocsc()
        return len(self.colours) + self.t_idx.size


class IPSolver:
    """solver = IPSolver({"tol": 1e-6, "max_iter": 2000}); sol = solver(optProb, sens=sens)."""

    DEFAULTS = {"tol": 1e-6, "max_iter": 2000, "mu_init": 0.1, "memory": 12, "bound_push": 1e-2,
                "scaling_max_gradient": 100.0, "acceptable_tol": 1e-4, "acceptable_iter": 15, "verbose": 0,
                "sigma_max": 1e8, "recalc_y": False, "hessian": "lbfgs", "fd_eps": 1e-5, "floor_rel": 1e-2, "ignore_bounds": True, "globalization": "filter", "feasible": False, "radius": 0.05, "delta_c": 1e-9, "feasible_iter": 4,
                "feasible_tol": 1e-9}

    def __init__(self, options=None):
        self.opt = dict(self.DEFAULTS)
        for k, v in (options or {}).items():
            if k in self.opt:
                self.opt[k] = v  # IPOPT-only options (linear_solver, output_file ...) are ignored

    # ------------------------------------------------------------------
    def __call__(self, prob, sens=None, **_):
        o = self.opt
        t_start = time.perf_counter()
        names = [v[0] for v in prob.vars]
        sizes = [v[1] for v in prob.vars]
        offs = np.concatenate(([0], np.cumsum(sizes))).astype(int)
        n = int(offs[-1])
        col0 = dict(zip(names, offs[:-1]))
        x = np.concatenate([v[2] for v in prob.vars]).astype(float)
        xl = np.concatenate([np.full(v[1], -np.inf if v[3] is None else v[3]) for v in prob.vars])
        xu = np.concatenate([np.full(v[1], np.inf if v[4] is None else v[4]) for v in prob.vars])
        if o["ignore_bounds"]:
            xl[:], xu[:] = -np.inf, np.inf
        eq = [g for g in prob.cons if g[3] is not None and g[2] == g[3]]
        ineq = [g for g in prob.cons if g not in eq]
        for g in ineq:
            if g[3] is not None or g[2] != 0.0:
                raise NotImplementedError("inequality groups other than c(x) >= 0")
        for g in eq:
            if g[2] != 0.0:
                raise NotImplementedError("equality groups other than c(x) = 0")
        mE, mI = sum(g[1] for g in eq), sum(g[1] for g in ineq)
        stat = {"obj_t": 0.0, "obj_n": 0, "sens_t": 0.0, "sens_n": 0}

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
            cI = np.concatenate([np.atleast_1d(np.asarray(f[g[0]], dtype=float)) for g in ineq]) if ineq else np.zeros(0)
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
            return obj, cE, cI, grad, self._jac(s, eq, mE, n, col0), self._jac(s, ineq, mI, n, col0)

        # ---- starting point: inside the bounds (IPOPT bound_push / bound_frac) ----
        span = np.where(np.isfinite(xu - xl), xu - xl, np.inf)
        pl = np.minimum(o["bound_push"] * np.maximum(1.0, np.abs(np.where(np.isfinite(xl), xl, 0.0))), 0.01 * span)
        pu = np.minimum(o["bound_push"] * np.maximum(1.0, np.abs(np.where(np.isfinite(xu), xu, 0.0))), 0.01 * span)
        x = np.minimum(np.maximum(x, xl + pl), xu - pu)
        hasL, hasU = np.isfinite(xl), np.isfinite(xu)

        f0, cE, cI, g, JE, JI = evaluate(x, True)
        # gradient-based scaling of the objective and of every constraint row
        gmax = o["scaling_max_gradient"]

        def row_scale(J):
            if J.shape[0] == 0:
                return np.zeros(0)
            mx = np.abs(J).max(axis=1).toarray().ravel()
            return np.minimum(1.0, gmax / np.maximum(mx, 1e-300))

        dE, dI = row_scale(JE), row_scale(JI)
        df = min(1.0, gmax / max(np.abs(g).max(), 1e-300))
        DE, DI = sp.diags(dE), sp.diags(dI)

        def scaled(ev):
            ob, ce, ci, gr, je, ji = ev
            return (df * ob, dE * ce, dI * ci, None if gr is None else df * gr, None if je is None else (DE @ je).tocsr(),
                    None if ji is None else (DI @ ji).tocsr())

        f0, cE, cI, g, JE, JI = scaled((f0, cE, cI, g, JE, JI))
        s = np.maximum(cI, o["bound_push"])
        mu = o["mu_init"]
        zL = np.where(hasL, mu / np.maximum(x - xl, 1e-300), 0.0)
        zU = np.where(hasU, mu / np.maximum(xu - x, 1e-300), 0.0)
        zs = mu / s
        lamI = -zs.copy()
        lamE = self._ls_multipliers(g, JE, JI, zL, zU, lamI)
        exact = o["hessian"] == "reduced-fd"
        newton = o["hessian"] == "sparse-fd"
        if newton and getattr(prob, "structure", None) is None:
            raise ValueError("hessian = sparse-fd needs the problem's structure (nlpshim.register attaches it)")
        if newton:
            B = _SparseFDHessian(n, prob.structure, col0, o["fd_eps"])
        else:
            B = _ReducedHessian(n, o["floor_rel"]) if exact else _LBFGS(n, o["memory"], sigma_max=o["sigma_max"])
        delta_last = 0.0
        radius = o["radius"]

        def lagr_grad_at(lE, lI):
            def fn(xv):
                t0 = time.perf_counter()
                sj, fail = sens(xdict(xv), None)
                stat["sens_t"] += time.perf_counter() - t0
                stat["sens_n"] += 1
                je = DE @ self._jac(sj, eq, mE, n, col0)
                ji = DI @ self._jac(sj, ineq, mI, n, col0)
                return je.T @ lE + ji.T @ lI
            return fn

        last = (0.0, 0.0, "")
        nu = 1.0
        filt = None  # the filter of the current barrier problem
        theta_max = theta_min = 0.0
        fails_in_a_row = 0
        tol = o["tol"]
        history = []
        status, message = 1, "maximum number of iterations exceeded"
        acceptable = 0
        fails = 0
        it = 0
        for it in range(o["max_iter"] + 1):
            # ---- optimality error of the NLP (mu = 0) and of the barrier problem ----
            rx = g + JE.T @ lamE + JI.T @ lamI - zL + zU
            s_d = max(100.0, (np.abs(lamE).sum() + np.abs(lamI).sum() + np.abs(zL).sum() + np.abs(zU).sum() + np.abs(zs).sum())
                      / max(1, mE + mI + 2 * n + mI)) / 100.0
            viol = max(np.abs(cE).max(initial=0.0), np.abs(cI - s).max(initial=0.0))

            def compl(m):
                dL, dU = np.where(hasL, x - xl, 1.0), np.where(hasU, xu - x, 1.0)
                return max(np.abs(np.where(hasL, dL * zL - m, 0.0)).max(initial=0.0),
                           np.abs(np.where(hasU, dU * zU - m, 0.0)).max(initial=0.0), np.abs(s * zs - m).max(initial=0.0))

            du_inf = max(np.abs(rx).max(), np.abs(lamI + zs).max(initial=0.0)) / s_d
            E0 = max(du_inf, viol, compl(0.0) / s_d)
            history.append((f0 / df, viol, du_inf, mu))
            if o["verbose"] and (it % o["verbose"] == 0):
                print("it %4d  obj %.8f  viol %.2e  dual %.2e  mu %.1e  filter %d  mem %d  fails %d  |lamE| %.2e  alpha %.2e  |dx| %.2e  sigma %.2e %s"
                      % (it, f0 / df, viol, du_inf, mu, 0 if filt is None else len(filt), len(B.S), fails,
                         np.abs(lamE).max(initial=0.0), last[0], last[1], B.sigma, last[2]))
            if E0 <= tol:
                status, message = 0, "converged to tol %g" % tol
                break
            acceptable = acceptable + 1 if E0 <= o["acceptable_tol"] else 0
            if acceptable >= o["acceptable_iter"]:
                status, message = 0, "converged to acceptable_tol %g" % o["acceptable_tol"]
                break
            if it == o["max_iter"]:
                break
            while mu > tol / 10.0 and max(du_inf, viol, compl(mu) / s_d) <= 10.0 * mu:
                mu = max(tol / 10.0, min(0.2 * mu, mu ** 1.5))
                filt = None
            tau = max(0.99, 1.0 - mu)

            if exact:
                B.build(JE, lagr_grad_at(lamE, lamI), x, o["fd_eps"])
            if newton:
                B.build(lagr_grad_at(lamE, lamI), x)
            # ---- Newton step of the barrier problem ----
            SigL = np.where(hasL, zL / np.maximum(x - xl, 1e-300), 0.0)
            SigU = np.where(hasU, zU / np.maximum(xu - x, 1e-300), 0.0)
            Sigs = zs / s
            r_x = g + JE.T @ lamE + JI.T @ lamI - np.where(hasL, mu / np.maximum(x - xl, 1e-300), 0.0) \
                + np.where(hasU, mu / np.maximum(xu - x, 1e-300), 0.0)
            r_I = (cI - s) - (lamI + mu / s) / Sigs
            rhs = -np.concatenate((r_x, cE, r_I))
            # inertia control without an inertia: the step must see positive curvature of the regularised Hessian
            # (Chiang & Zavala 2016); otherwise delta_w grows, as in IPOPT's correction loop
            delta_w = 0.0 if not newton else (0.0 if delta_last == 0.0 else max(delta_last / 3.0, 1e-8))
            sol_vec = None
            for attempt in range(40):
                sol_vec = self._kkt_solve(B, SigL + SigU + delta_w, JE, JI, 1.0 / Sigs, rhs, n, mE, mI)
                ok_step = sol_vec is not None and np.all(np.isfinite(sol_vec))
                if ok_step and newton:
                    dxt = sol_vec[:n]
                    curv = float(dxt @ B.times(dxt)) + float(dxt @ ((SigL + SigU + delta_w) * dxt))
                    ok_step = curv >= 1e-9 * float(dxt @ dxt)
                    # trust radius (Levenberg-Marquardt style): a nearly flat tangent-space Hessian must not send the
                    # iterate where the linearisation means nothing; more regularisation shortens the tangential step
                    if ok_step and np.abs(dxt).max() > radius and delta_w < 1e8:
                        ok_step = False
                if ok_step:
                    break
                delta_w = 1e-4 if delta_w == 0.0 else (8.0 * delta_w if delta_last > 0.0 or attempt > 0 else 100.0 * delta_w)
                if delta_w > 1e20:
                    sol_vec = None
                    break
            if sol_vec is None or not np.all(np.isfinite(sol_vec)):
                status, message = 2, "KKT system could not be solved"
                break
            delta_last = delta_w
            dx, dlE, dlI = sol_vec[:n], sol_vec[n: n + mE], sol_vec[n + mE:]
            ds = (dlI + lamI + mu / s) / Sigs
            dzL = np.where(hasL, mu / np.maximum(x - xl, 1e-300) - zL - SigL * dx, 0.0)
            dzU = np.where(hasU, mu / np.maximum(xu - x, 1e-300) - zU + SigU * dx, 0.0)
            dzs = mu / s - zs - Sigs * ds

            def max_step(v, dv, t):
                neg = dv < 0
                return min(1.0, float(np.min(-t * v[neg] / dv[neg]))) if neg.any() else 1.0

            a_max = min(max_step(np.where(hasL, x - xl, 1.0), np.where(hasL, dx, 0.0), tau),
                        max_step(np.where(hasU, xu - x, 1.0), np.where(hasU, -dx, 0.0), tau), max_step(s, ds, tau))
            a_z = min(max_step(zL, dzL, tau), max_step(zU, dzU, tau), max_step(zs, dzs, tau))

            # ---- filter line search (Waechter & Biegler, Alg. A) with a second-order correction ----
            def barrier(xv, sv, fv):
                return fv - mu * (np.log(np.where(hasL, xv - xl, 1.0)).sum() + np.log(np.where(hasU, xu - xv, 1.0)).sum()
                                  + np.log(sv).sum())

            gphi_d = float(g @ dx) - mu * (np.where(hasL, dx / np.maximum(x - xl, 1e-300), 0.0).sum()
                                           - np.where(hasU, dx / np.maximum(xu - x, 1e-300), 0.0).sum() + (ds / s).sum())
            theta0 = np.abs(cE).sum() + np.abs(cI - s).sum()
            phi0 = barrier(x, s, f0)
            if filt is None:  # (re)started with every new barrier parameter
                filt = []
                theta_max = 1e4 * max(1.0, theta0)
                theta_min = 1e-4 * max(1.0, theta0)
            g_th, g_ph, eta = 1e-5, 1e-8, 1e-8

            def in_filter(th, ph):
                return th >= theta_max or any(th >= (1.0 - g_th) * t_ and ph >= p_ - g_ph * t_ for t_, p_ in filt)

            use_merit = o["globalization"] == "merit"
            if use_merit:
                lam_inf = max(np.abs(lamE + dlE).max(initial=0.0), np.abs(lamI + dlI).max(initial=0.0))
                nu = max(nu, 1.5 * lam_inf) if nu >= lam_inf else 1.5 * lam_inf + 1.0
                Dmerit = gphi_d - nu * theta0
            alpha = a_max
            accepted = False
            armijo_step = False
            soc_done = False
            step_dx, step_ds = dx, ds
            a_min = 1e-10
            for ls in range(50):
                xt = x + alpha * step_dx
                st = s + alpha * step_ds
                ft, cEt, cIt, _, _, _ = scaled(evaluate(xt, False))
                thetat = np.abs(cEt).sum() + np.abs(cIt - st).sum()
                phit = barrier(xt, st, ft)
                ok = np.isfinite(phit) and np.isfinite(thetat) and not in_filter(thetat, phit)
                if use_merit:  # l1 exact-penalty merit function instead of the filter
                    ok = np.isfinite(phit) and np.isfinite(thetat) and \
                        phit + nu * thetat <= phi0 + nu * theta0 + 1e-4 * alpha * min(Dmerit, 0.0) + 1e-13 * abs(phi0)
                    armijo_step = True
                elif ok:
                    switching = gphi_d < 0 and theta0 <= theta_min and alpha * (-gphi_d) ** 2.3 > theta0 ** 1.1
                    if switching:
                        ok = phit <= phi0 + eta * alpha * gphi_d + 10.0 * np.finfo(float).eps * abs(phi0)
                        armijo_step = ok
                    else:
                        ok = thetat <= (1.0 - g_th) * theta0 or phit <= phi0 - g_ph * theta0
                if ok:
                    accepted = True
                    break
                if ls == 0 and not soc_done and np.isfinite(thetat) and thetat >= theta0:
                    # second-order correction: re-solve with the constraint values of the trial point added
                    soc_done = True
                    rhs2 = -np.concatenate((r_x, alpha * cE + cEt, alpha * r_I + (cIt - st)))
                    sv = self._kkt_solve(B, SigL + SigU + delta_w, JE, JI, 1.0 / Sigs, rhs2, n, mE, mI, reuse=True)
                    if sv is not None and np.all(np.isfinite(sv)):
                        dx2 = sv[:n]
                        ds2 = (sv[n + mE:] + lamI + mu / s) / Sigs
                        a2 = min(max_step(np.where(hasL, x - xl, 1.0), np.where(hasL, dx2, 0.0), tau),
                                 max_step(np.where(hasU, xu - x, 1.0), np.where(hasU, -dx2, 0.0), tau), max_step(s, ds2, tau))
                        x2, s2 = x + a2 * dx2, s + a2 * ds2
                        f2, cE2, cI2, _, _, _ = scaled(evaluate(x2, False))
                        th2 = np.abs(cE2).sum() + np.abs(cI2 - s2).sum()
                        phi2 = barrier(x2, s2, f2)
                        soc_ok = np.isfinite(phi2) and not in_filter(th2, phi2) and (
                            th2 <= (1.0 - g_th) * theta0 or phi2 <= phi0 - g_ph * theta0)
                        if use_merit:
                            soc_ok = np.isfinite(phi2) and phi2 + nu * th2 <= phi0 + nu * theta0 + 1e-4 * a2 * min(Dmerit, 0.0)
                        if soc_ok:
                            step_dx, step_ds, alpha = dx2, ds2, a2
                            accepted = True
                            break
                alpha *= 0.5
                if alpha < a_min:
                    break
            if accepted and not armijo_step:
                filt.append(((1.0 - g_th) * theta0, phi0 - g_ph * theta0))
            if not accepted:
                fails += 1
                if len(B.S) > 0 and fails_in_a_row == 0:  # a poor quasi-Newton model: drop the memory and try again from sigma I
                    B.reset(B.sigma)
                    fails_in_a_row += 1
                    continue
                # restoration: least-norm steps towards feasibility until the filter accepts the point
                filt.append(((1.0 - g_th) * theta0, phi0 - g_ph * theta0))
                xr, sr, thr = x.copy(), s.copy(), theta0
                cEr, cIr, JEr, JIr = cE, cI, JE, JI
                ok = False
                for _r in range(20):
                    dxr = self._feasibility_step(JEr, JIr, cEr, cIr - sr, n, mE, mI)
                    if dxr is None:
                        break
                    a = min(max_step(np.where(hasL, xr - xl, 1.0), np.where(hasL, dxr, 0.0), tau),
                            max_step(np.where(hasU, xu - xr, 1.0), np.where(hasU, -dxr, 0.0), tau))
                    moved = False
                    for _ in range(30):
                        xt = xr + a * dxr
                        ft, cEt, cIt, _, _, _ = scaled(evaluate(xt, False))
                        st = np.maximum(cIt, np.minimum(sr, mu))
                        tht = np.abs(cEt).sum() + np.abs(cIt - st).sum()
                        if np.isfinite(tht) and tht < (1.0 - 1e-4 * a) * thr:
                            moved = True
                            break
                        a *= 0.5
                    if not moved:
                        break
                    xr, sr, thr = xt, st, tht
                    if thr <= 0.9 * theta0 and not in_filter(thr, barrier(xr, sr, ft)):
                        ok = True
                        break
                    _, cEr, cIr, _, JEr, JIr = scaled(evaluate(xr, True))
                if not ok:
                    status, message = 3, "restoration failed"
                    break
                alpha, step_dx, step_ds = 1.0, xr - x, sr - s
                dlE, dlI = np.zeros(mE), np.zeros(mI)
                B.reset(B.sigma)
            fails_in_a_row = 0
            if newton:  # the radius follows the line search: full steps widen it, short ones narrow it
                if alpha >= 0.99:
                    radius = min(2.0 * radius, 1.0)
                elif alpha < 0.25:
                    radius = max(0.5 * radius, 1e-4)
            last = (alpha, float(np.abs(alpha * step_dx).max()), "armijo" if armijo_step else ("soc" if step_dx is not dx else "theta"))
            # ---- accept: primal and equality multipliers with alpha, bound multipliers with their own step ----
            x_new, s_new = x + alpha * step_dx, s + alpha * step_ds
            lamE_new = lamE + alpha * dlE
            lamI_new = lamI + alpha * dlI
            zL = zL + a_z * dzL
            zU = zU + a_z * dzU
            zs = zs + a_z * dzs
            # keep the bound multipliers within a factor of the primal estimate mu / slack (IPOPT kappa_Sigma = 1e10)
            ks = 1e10
            zL = np.where(hasL, np.clip(zL, mu / (ks * np.maximum(x_new - xl, 1e-300)), ks * mu / np.maximum(x_new - xl, 1e-300)), 0.0)
            zU = np.where(hasU, np.clip(zU, mu / (ks * np.maximum(xu - x_new, 1e-300)), ks * mu / np.maximum(xu - x_new, 1e-300)), 0.0)
            zs = np.clip(zs, mu / (ks * s_new), ks * mu / s_new)
            f_new, cE_new, cI_new, g_new, JE_new, JI_new = scaled(evaluate(x_new, True))
            if o["feasible"]:
                # back onto the constraint manifold: least-norm Newton steps on the equality rows (quadratic convergence;
                # each costs one callback pair).  The quasi-Newton pairs are then differences between FEASIBLE points, so the
                # approximation learns the curvature along the manifold -- the only curvature the tangential step needs.
                for _r in range(o["feasible_iter"]):
                    if np.abs(cE_new).max(initial=0.0) <= o["feasible_tol"]:
                        break
                    dxr = self._feasibility_step(JE_new, JI_new[:0], cE_new, np.zeros(0), n, mE, 0)
                    if dxr is None:
                        break
                    a = min(1.0, max_step(np.where(hasL, x_new - xl, 1.0), np.where(hasL, dxr, 0.0), tau),
                            max_step(np.where(hasU, xu - x_new, 1.0), np.where(hasU, -dxr, 0.0), tau))
                    x_try = x_new + a * dxr
                    ev_try = scaled(evaluate(x_try, True))
                    if not np.isfinite(ev_try[1]).all() or np.abs(ev_try[1]).max(initial=0.0) >= np.abs(cE_new).max(initial=0.0):
                        break
                    x_new = x_try
                    f_new, cE_new, cI_new, g_new, JE_new, JI_new = ev_try
                s_new = np.maximum(cI_new, np.minimum(s_new, mu))
            if o["recalc_y"]:  # least-squares equality multipliers at the new point (IPOPT recalc_y)
                lamE_new = self._ls_multipliers(g_new, JE_new, JI_new, zL, zU, lamI_new, fallback=lamE_new)
            yk = (g_new + JE_new.T @ lamE_new + JI_new.T @ lamI_new) - (g + JE.T @ lamE_new + JI.T @ lamI_new)
            B.update(x_new - x, yk)
            x, s, lamE, lamI = x_new, s_new, lamE_new, lamI_new
            f0, cE, cI, g, JE, JI = f_new, cE_new, cI_new, g_new, JE_new, JI_new

        sol = Solution()
        sol.xStar = xdict(x)
        sol.fStar = f0 / df
        sol.optTime = time.perf_counter() - t_start
        sol.userObjTime, sol.userObjCalls = stat["obj_t"], stat["obj_n"]
        sol.userSensTime, sol.userSensCalls = stat["sens_t"], stat["sens_n"]
        sol.constr_violation = float(max(np.abs(cE / dE).max(initial=0.0), np.maximum(-(cI / dI), 0.0).max(initial=0.0)))
        sol.dual_infeasibility = float(history[-1][2]) if history else float("nan")
        sol.nit, sol.status, sol.message, sol.history = it, status, message, history
        sol.line_search_failures = fails
        sol.optInform = {"value": status, "text": message}
        return sol

    # ------------------------------------------------------------------
    @staticmethod
    def _jac(s, groups, nrow, nvar, col0):
        rows, cols, data = [], [], []
        r0 = 0
        for name, n, _, _, wrt, _ in groups:
            for var, blk in s[name].items():
                if wrt is not None and var not in wrt:
                    continue
                if isinstance(blk, dict):
                    r, c, d = blk["coo"]
                    r, c, d = np.asarray(r), np.asarray(c), np.asarray(d, dtype=float)
                else:
                    dense = np.atleast_2d(np.asarray(blk, dtype=float))
                    r, c = np.nonzero(dense)
                    d = dense[r, c]
                rows.append(r + r0)
                cols.append(c + col0[var])
                data.append(d)
            r0 += n
        if not rows:
            return sp.csr_matrix((nrow, nvar))
        return sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(nrow, nvar)).tocsr()

    @staticmethod
    def _ls_multipliers(g, JE, JI, zL, zU, lamI, fallback=None):
        """Least-squares estimate of the equality multipliers: min || g + JE^T lam + JI^T lamI - zL + zU ||."""
        mE = JE.shape[0]
        if mE == 0:
            return np.zeros(0)
        n = JE.shape[1]
        K = sp.bmat([[sp.identity(n), JE.T], [JE, -1e-10 * sp.identity(mE)]], format="csc")
        rhs = np.concatenate((-(g + JI.T @ lamI - zL + zU), np.zeros(mE)))
        zero = np.zeros(mE) if fallback is None else fallback
        try:
            lam = spla.splu(K).solve(rhs)[n:]
        except RuntimeError:
            return zero
        if not np.all(np.isfinite(lam)) or (fallback is None and np.abs(lam).max() > 1e3):
            return zero
        return lam

    def _kkt_solve(self, B, diag_x, JE, JI, inv_sigs, rhs, n, mE, mI, reuse=False):
        """Solve [[B + diag_x, JE^T, JI^T], [JE, -dc, 0], [JI, 0, -1/Sigma_s - dc]] v = rhs with B = sigma I - U M^-1 U^T:
        sparse LU of the sigma I part, Woodbury for the rank-2k part."""
        if not reuse:
            dc = self.opt["delta_c"]
            Wx = sp.diags(B.sigma + diag_x)
            if getattr(B, "W", None) is not None:
                Wx = Wx + B.W
            K0 = sp.bmat([[Wx, JE.T, JI.T],
                          [JE, -dc * sp.identity(mE), None],
                          [JI, None, sp.diags(-(inv_sigs + dc))]], format="csc")
            try:
                self._lu = spla.splu(K0)
            except RuntimeError:
                return None
            self._W = None
            add = B.additive() if hasattr(B, "additive") else None
            if add is not None:  # W = sigma I + U C U^T
                U, Cinv = add
                Ubar = np.vstack((U, np.zeros((mE + mI, U.shape[1]))))
                KU = self._lu.solve(Ubar)
                self._W = (Ubar, -KU, Cinv + Ubar.T @ KU)
            elif B.S:  # W = sigma I - U M^-1 U^T
                U, M = B.factors()
                Ubar = np.vstack((U, np.zeros((mE + mI, U.shape[1]))))
                KU = self._lu.solve(Ubar)
                self._W = (Ubar, KU, M - Ubar.T @ KU)
        v = self._lu.solve(rhs)
        if self._W is not None:
            Ubar, KU, C = self._W
            try:
                v = v + KU @ np.linalg.solve(C, Ubar.T @ v)
            except np.linalg.LinAlgError:
                return None
        return v

    @staticmethod
    def _feasibility_step(JE, JI, cE, cIs, n, mE, mI):
        J = sp.vstack((JE, JI)).tocsr()
        c = np.concatenate((cE, cIs))
        K = sp.bmat([[sp.identity(n), J.T], [J, -1e-8 * sp.identity(mE + mI)]], format="csc")
        try:
            v = spla.splu(K).solve(np.concatenate((np.zeros(n), -c)))
        except RuntimeError:
            return None
        return v[:n] if np.all(np.isfinite(v)) else None
