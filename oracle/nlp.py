"""CPU oracle of GELATO's NLP callbacks (TEST INFRASTRUCTURE ONLY).

A numpy restatement of what the reference's `objfunc` / `sens`
(/root/reference/Trajectory_Optimization.py:194-312) compute through
lib/con_dynamics.py, con_aero.py, con_waypoint.py, con_init_terminal_knot.py,
con_trajectory.py, con_user.py, jac_fd.py and cost_gradient.py -- same group
keys, same row order, same COO (row, col, data) order, same forward-difference
protocol INCLUDING the in-place perturb/restore on the caller's `xdict`
(so the `fl(fl(x+dx)-dx)` residue later columns and later groups see is
reproduced by construction, not modelled).  Each method cites the reference
lines it follows.  The physics leaves come from oracle/leaves.py.

It is pinned against the real reference Python layer in this container by
tests/golden/make_golden.py (golden fixtures are committed) and by
tests/test_oracle_vs_reference.py (skipped where /root/reference is absent).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; gelato_b200/ never does.

`dot` selects how D.X is formed: "numpy" (BLAS, as the reference) or "seqfma"
(left-to-right fused multiply-add accumulation, the CUDA kernels' order).
"""
import math

import numpy as np

from . import leaves as _leaves

EQ_GROUPS = (
    "eqcon_init eqcon_time eqcon_dyn_mass eqcon_dyn_pos eqcon_dyn_vel eqcon_dyn_quat eqcon_knot "
    "eqcon_terminal eqcon_rate eqcon_pos eqcon_iip eqcon_user"
).split()
INEQ_GROUPS = (
    "ineqcon_alpha ineqcon_q ineqcon_qalpha ineqcon_mass ineqcon_kick ineqcon_time ineqcon_pos "
    "ineqcon_iip ineqcon_antenna ineqcon_user"
).split()
GROUPS = EQ_GROUPS + INEQ_GROUPS


def _coo(rows, cols, data, shape):
    return {
        "coo": [np.asarray(rows, dtype="i4"), np.asarray(cols, dtype="i4"), np.asarray(data, dtype="f8")],
        "shape": shape,
    }


def _cat(parts, dtype=None):
    if len(parts) == 0:
        return np.zeros(0, dtype=dtype or "f8")
    return np.concatenate([np.asarray(p).ravel() for p in parts])


class OracleNLP:
    def __init__(self, pdict, unitdict, condition, flavour="libm", dot="numpy", user_eq=None, user_ineq=None):
        self.p = pdict
        self.u = unitdict
        self.c = condition
        self.leaves = _leaves.get(flavour)
        self.dot_mode = dot
        self.user_eq = user_eq
        self.user_ineq = user_ineq
        self.ps = pdict["ps_params"]
        self.S = pdict["num_sections"]
        self.N = pdict["N"]
        self.M = pdict["M"]
        self.dx = pdict["dx"]
        L = self.leaves
        self.dyn, self.utl, self.crd, self.iip = L.dynamics_c, L.utils_c, L.coordinate_c, L.IIP_c

    # ------------------------------------------------------------------
    def _dot(self, D, X):
        if self.dot_mode == "numpy":
            return D.dot(X)
        X2 = X.reshape(X.shape[0], -1)
        out = np.empty((D.shape[0], X2.shape[1]))
        self.leaves.lib.o_seqfma_matmul(
            _leaves._p(np.ascontiguousarray(D)), _leaves._p(np.ascontiguousarray(X2)), D.shape[0], D.shape[1],
            X2.shape[1], _leaves._p(out))
        return out.reshape((D.shape[0],) + X.shape[1:])

    def _param(self, i):
        prm = self.p["params"][i]
        return np.array([prm["thrust"], prm["massflow"], prm["reference_area"], 0.0, prm["nozzle_area"]])

    def _rhs_vel(self, i, mass, pos, vel, quat, t_nodes, units):
        """con_dynamics.py:257-286 / :345-351 -- air branch iff reference_area != 0."""
        param = self._param(i)
        if param[2] == 0.0:
            return self.dyn.dynamics_velocity_NoAir(mass, pos, quat, param, units)
        return self.dyn.dynamics_velocity(
            mass, pos, vel, quat, t_nodes, param, self.p["wind_table"], self.p["ca_table"], units)

    # ================= objective (cost_gradient.py:29-47) ==============
    def cost(self, x):
        return -x["mass"][0] if self.c["OptimizationMode"] == "Payload" else x["t"][-1]

    def cost_jac(self, x):
        if self.c["OptimizationMode"] == "Payload":
            g = np.zeros(x["mass"].size)
            g[0] = -1.0
            return {"mass": g}
        g = np.zeros(x["t"].size)
        g[-1] = 1.0
        return {"t": g}

    # ================= init / time (con_init_terminal_knot.py:41-171) ==
    def eq_init(self, x):
        c, u = self.c, self.u
        parts = []
        if c["OptimizationMode"] != "Payload":
            parts.append(x["mass"][0] - c["init"]["mass"] / u["mass"])
        parts.append(x["position"][0:3] - c["init"]["position"] / u["position"])
        parts.append(x["velocity"][0:3] - c["init"]["velocity"] / u["velocity"])
        parts.append(x["quaternion"][0:4] - c["init"]["quaternion"])
        return np.concatenate(parts, axis=None)

    def eq_init_jac(self, x):
        M = self.M
        off = 0 if self.c["OptimizationMode"] == "Payload" else 1
        nrow = 10 + off
        jac = {}
        if off:
            jac["mass"] = _coo([0], [0], [1.0], (nrow, M))
        jac["position"] = _coo(np.arange(off, off + 3), np.arange(3), np.ones(3), (nrow, M * 3))
        jac["velocity"] = _coo(np.arange(off + 3, off + 6), np.arange(3), np.ones(3), (nrow, M * 3))
        jac["quaternion"] = _coo(np.arange(off + 6, off + 10), np.arange(4), np.ones(4), (nrow, M * 4))
        return jac

    def _timed_events(self):
        """[(i, i_ref)] for events whose time is tied to another event (:137-143)."""
        prm, idx = self.p["params"], self.p["event_index"]
        return [(i, idx[prm[i]["time_ref"]]) for i in range(1, self.S + 1) if prm[i]["time_ref"] in idx]

    def eq_time(self, x):
        prm, ut, t = self.p["params"], self.u["t"], x["t"]
        rows = [t[0] - prm[0]["time"] / ut]
        for i, ir in self._timed_events():
            rows.append(t[i] - t[ir] - (prm[i]["time"] - prm[ir]["time"]) / ut)
        return np.concatenate(rows, axis=None)

    def eq_time_jac(self, x):
        r, c, d = [0], [0], [1.0]
        for k, (i, ir) in enumerate(self._timed_events()):
            r += [k + 1, k + 1]
            c += [i, ir]
            d += [1.0, -1.0]
        return {"t": _coo(r, c, d, (len(self._timed_events()) + 1, len(x["t"])))}

    # ================= dynamics: mass (con_dynamics.py:34-113) =========
    def eq_dyn_mass(self, x):
        um, ut, t = self.u["mass"], self.u["t"], x["t"]
        out = []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            m = x["mass"][xa:xb]
            to, tf = t[i], t[i + 1]
            if self.p["params"][i]["engineOn"]:
                lh = self._dot(self.ps.D(i), m)
                rh = np.full(n, -self.p["params"][i]["massflow"] / um * (tf - to) * ut / 2.0)
                out.append(lh - rh)
            else:
                out.append(m[1:] - m[0])
        return np.concatenate(out, axis=None)

    def eq_dyn_mass_jac(self, x):
        um, ut = self.u["mass"], self.u["t"]
        mr, mc, md, tr, tc, td = [], [], [], [], [], []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            if self.p["params"][i]["engineOn"]:
                mr.append(np.repeat(np.arange(ua, ub), n + 1))
                mc.append(np.tile(np.arange(xa, xb), n))
                md.append(self.ps.D(i).ravel(order="C"))
                mf = self.p["params"][i]["massflow"]
                tr += [np.arange(ua, ub), np.arange(ua, ub)]
                tc += [np.full(n, i), np.full(n, i + 1)]
                td += [np.full(n, -mf / um * ut / 2.0), np.full(n, mf / um * ut / 2.0)]
            else:
                mr += [np.arange(ua, ub), np.arange(ua, ub)]
                mc += [np.full(n, xa), np.arange(xa + 1, xb)]
                md += [np.full(n, -1.0), np.full(n, 1.0)]
        return {
            "mass": _coo(_cat(mr, "i4"), _cat(mc, "i4"), _cat(md), (self.N, self.M)),
            "t": _coo(_cat(tr, "i4"), _cat(tc, "i4"), _cat(td), (self.N, self.S + 1)),
        }

    # ================= dynamics: position (con_dynamics.py:116-213) ====
    def eq_dyn_pos(self, x):
        up, uv, ut, t = self.u["position"], self.u["velocity"], self.u["t"], x["t"]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        out = []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            to, tf = t[i], t[i + 1]
            lh = self._dot(self.ps.D(i), pos[xa:xb])
            rh = vel[xa:xb][1:] * uv * (tf - to) * ut / 2.0 / up
            out.append((lh - rh).ravel())
        return np.concatenate(out, axis=None)

    def eq_dyn_pos_jac(self, x):
        up, uv, ut, t = self.u["position"], self.u["velocity"], self.u["t"], x["t"]
        vel = x["velocity"].reshape(-1, 3)
        pr, pc, pd, vr, vc, vd, tr, tc, td = ([] for _ in range(9))
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            to, tf = t[i], t[i + 1]
            Di = self.ps.D(i).ravel()
            rows3 = np.arange(ua * 3, ub * 3)
            vr.append(rows3)
            vc.append(np.arange((xa + 1) * 3, xb * 3))
            vd.append(np.full(n * 3, -uv * (tf - to) * ut / 2.0 / up))
            rh_to = vel[xa:xb][1:].ravel() * uv * ut / 2.0 / up
            tr += [rows3, rows3]
            tc += [np.full(n * 3, i), np.full(n * 3, i + 1)]
            td += [rh_to, -rh_to]
            for k in range(3):
                pr.append(np.repeat(np.arange(ua * 3 + k, ub * 3 + k, 3), n + 1))
                pc.append(np.tile(np.arange(xa * 3 + k, xb * 3 + k, 3), n))
                pd.append(Di)
        N3, M3 = self.N * 3, self.M * 3
        return {
            "position": _coo(_cat(pr, "i4"), _cat(pc, "i4"), _cat(pd), (N3, M3)),
            "velocity": _coo(_cat(vr, "i4"), _cat(vc, "i4"), _cat(vd), (N3, M3)),
            "t": _coo(_cat(tr, "i4"), _cat(tc, "i4"), _cat(td), (N3, self.S + 1)),
        }

    # ================= dynamics: velocity (con_dynamics.py:216-496) ====
    def _units3(self):
        return np.array([self.u["mass"], self.u["position"], self.u["velocity"]])

    def eq_dyn_vel(self, x):
        ut, t = self.u["t"], x["t"]
        mass = x["mass"]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        quat = x["quaternion"].reshape(-1, 4)
        units = self._units3()
        out = []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            to, tf = t[i], t[i + 1]
            tn = self.ps.time_nodes(i, to, tf)
            lh = self._dot(self.ps.D(i), vel[xa:xb])
            f = self._rhs_vel(i, mass[xa + 1 : xb], pos[xa + 1 : xb], vel[xa + 1 : xb], quat[xa + 1 : xb], tn[1:], units)
            rh = f * (tf - to) * ut / 2.0
            out.append((lh - rh).ravel())
        return np.concatenate(out, axis=None)

    def eq_dyn_vel_jac(self, x):
        dx, ut, t = self.dx, self.u["t"], x["t"]
        mass = x["mass"]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        quat = x["quaternion"].reshape(-1, 4)
        units = self._units3()
        acc = {k: ([], [], []) for k in ("mass", "position", "velocity", "quaternion", "t")}

        def put(key, r, c, d):
            acc[key][0].append(np.asarray(r).ravel())
            acc[key][1].append(np.asarray(c).ravel())
            acc[key][2].append(np.asarray(d).ravel())

        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            to, tf = t[i], t[i + 1]
            tn = self.ps.time_nodes(i, to, tf)
            air = self._param(i)[2] > 0.0
            m_, p_, v_, q_ = mass[xa + 1 : xb], pos[xa + 1 : xb], vel[xa + 1 : xb], quat[xa + 1 : xb]  # views

            def f(tnodes=tn):
                return self._rhs_vel(i, m_, p_, v_, q_, tnodes[1:], units)

            f_c = f()
            rows_nodes = np.arange(ua * 3, ub * 3)  # (node j, comp c) -> 3j+c
            xs = np.arange(xa + 1, xb)

            m_ += dx
            f_p = f()
            m_ -= dx
            put("mass", rows_nodes, np.repeat(xs, 3), -(f_p - f_c) / dx * (tf - to) * ut / 2.0)

            for k in range(3):
                p_[:, k] += dx
                f_p = f()
                p_[:, k] -= dx
                put("position", rows_nodes, np.repeat(xs * 3 + k, 3), -(f_p - f_c) / dx * (tf - to) * ut / 2.0)

            D = self.ps.D(i)
            sub = np.zeros((n * 3, (n + 1) * 3))
            for k in range(3):
                sub[k::3, k::3] = D
            if air:
                for k in range(3):
                    v_[:, k] += dx
                    f_p = f()
                    v_[:, k] -= dx
                    rh = -(f_p - f_c) / dx * (tf - to) * ut / 2.0
                    for j in range(n):
                        sub[j * 3 : j * 3 + 3, (j + 1) * 3 + k] += rh[j]
            for ki in range(3):
                for kj in range(3):
                    put("velocity", np.repeat(np.arange(ua * 3 + ki, ub * 3 + ki, 3), n + 1),
                        np.tile(np.arange(xa * 3 + kj, xb * 3 + kj, 3), n), sub[ki::3, kj::3])

            for k in range(4):
                q_[:, k] += dx
                f_p = f()
                q_[:, k] -= dx
                put("quaternion", rows_nodes, np.repeat(xs * 4 + k, 3), -(f_p - f_c) / dx * (tf - to) * ut / 2.0)

            to_p = to + dx
            if air:
                f_p = f(self.ps.time_nodes(i, to_p, tf))
                rh_to = -(f_p * (tf - to_p) - f_c * (tf - to)).ravel() / dx * ut / 2.0
                tf_p = tf + dx
                f_p = f(self.ps.time_nodes(i, to, tf_p))
                rh_tf = -(f_p * (tf_p - to) - f_c * (tf - to)).ravel() / dx * ut / 2.0
            else:
                rh_to = f_c.ravel() * ut / 2.0
                rh_tf = -rh_to
            put("t", rows_nodes, np.full(n * 3, i), rh_to)
            put("t", rows_nodes, np.full(n * 3, i + 1), rh_tf)

        N3, M = self.N * 3, self.M
        shapes = {"mass": (N3, M), "position": (N3, M * 3), "velocity": (N3, M * 3), "quaternion": (N3, M * 4),
                  "t": (N3, self.S + 1)}
        return {k: _coo(_cat(a[0], "i4"), _cat(a[1], "i4"), _cat(a[2]), shapes[k]) for k, a in acc.items()}

    # ================= dynamics: quaternion (con_dynamics.py:499-632) ==
    def _holds(self, i):
        return self.p["params"][i]["attitude"] in ["hold", "vertical"]

    def eq_dyn_quat(self, x):
        uu, ut, t = self.u["u"], self.u["t"], x["t"]
        quat = x["quaternion"].reshape(-1, 4)
        u = x["u"].reshape(-1, 2)
        out = []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            q = quat[xa:xb]
            to, tf = t[i], t[i + 1]
            if self._holds(i):
                out.append((q[1:] - q[0]).ravel())
            else:
                lh = self._dot(self.ps.D(i), q)
                rh = self.dyn.dynamics_quaternion(q[1:], u[ua:ub], uu) * (tf - to) * ut / 2.0
                out.append((lh - rh).ravel())
        return np.concatenate(out, axis=None)

    def eq_dyn_quat_jac(self, x):
        dx, uu, ut, t = self.dx, self.u["u"], self.u["t"], x["t"]
        quat = x["quaternion"].reshape(-1, 4)
        u = x["u"].reshape(-1, 2)
        acc = {k: ([], [], []) for k in ("quaternion", "u", "t")}

        def put(key, r, c, d):
            acc[key][0].append(np.asarray(r).ravel())
            acc[key][1].append(np.asarray(c).ravel())
            acc[key][2].append(np.asarray(d).ravel())

        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            to, tf = t[i], t[i + 1]
            rows = np.arange(ua * 4, ub * 4)
            if self._holds(i):
                put("quaternion", rows, np.tile(np.arange(xa * 4, (xa + 1) * 4), n), np.full(4 * n, -1.0))
                put("quaternion", rows, np.arange((xa + 1) * 4, xb * 4), np.full(4 * n, 1.0))
                continue
            q_ = quat[xa + 1 : xb]
            u_ = u[ua:ub]
            D = self.ps.D(i)
            sub = np.zeros((n * 4, (n + 1) * 4))
            for k in range(4):
                sub[k::4, k::4] = D
            f_c = self.dyn.dynamics_quaternion(q_, u_, uu)
            for k in range(4):
                q_[:, k] += dx
                f_p = self.dyn.dynamics_quaternion(q_, u_, uu)
                q_[:, k] -= dx
                rh = -(f_p - f_c) / dx * (tf - to) * ut / 2.0
                for j in range(n):
                    sub[j * 4 : j * 4 + 4, (j + 1) * 4 + k] += rh[j]
            put("quaternion", np.repeat(rows, (n + 1) * 4), np.tile(np.arange(xa * 4, xb * 4), n * 4), sub)
            for k in range(2):
                u_[:, k] += dx
                f_p = self.dyn.dynamics_quaternion(q_, u_, uu)
                u_[:, k] -= dx
                put("u", rows, np.repeat(np.arange(ua, ub) * 2 + k, 4), -(f_p - f_c) / dx * (tf - to) * ut / 2.0)
            rh_to = f_c.ravel() * ut / 2.0
            put("t", rows, np.full(n * 4, i), rh_to)
            put("t", rows, np.full(n * 4, i + 1), -rh_to)
        N4 = self.N * 4
        shapes = {"quaternion": (N4, self.M * 4), "u": (N4, self.N * 2), "t": (N4, self.S + 1)}
        return {k: _coo(_cat(a[0], "i4"), _cat(a[1], "i4"), _cat(a[2]), shapes[k]) for k, a in acc.items()}

    # ================= knot (con_init_terminal_knot.py:174-326) ========
    def _stage_sections(self):
        """[(section_ig, section_sep, stage)] for stages that separate (:188-201)."""
        names = [p["name"] for p in self.p["params"]]
        out = []
        for stage in self.p["RocketStage"].values():
            if stage["separation_at"] is not None:
                out.append((names.index(stage["ignition_at"]), names.index(stage["separation_at"]), stage))
        return out

    def eq_knot(self, x):
        um = self.u["mass"]
        mass = x["mass"]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        quat = x["quaternion"].reshape(-1, 4)
        rows = []
        sep = []
        for ig, sp, stage in self._stage_sections():
            sep.append(sp)
            mass_stage = stage["mass_dry"] + stage["mass_propellant"] + sum(
                [item["mass"] for item in stage["dropMass"].values()])
            rows.append(mass[self.ps.index_start_x(ig)] - mass[self.ps.index_start_x(sp)] - mass_stage / um)
        for i in range(1, self.S):
            xa = self.ps.index_start_x(i)
            if i not in sep:
                rows.append(mass[xa] - mass[xa - 1] + self.p["params"][i]["mass_jettison"] / um)
            rows.append(pos[xa] - pos[xa - 1])
            rows.append(vel[xa] - vel[xa - 1])
            rows.append(quat[xa] - quat[xa - 1])
        return np.concatenate(rows, axis=None)

    def eq_knot_jac(self, x):
        acc = {k: ([], [], []) for k in ("mass", "position", "velocity", "quaternion")}
        r = 0
        sep = []
        for ig, sp, _ in self._stage_sections():
            sep.append(sp)
            acc["mass"][0].extend([r, r])
            acc["mass"][1].extend([self.ps.index_start_x(ig), self.ps.index_start_x(sp)])
            acc["mass"][2].extend([1.0, -1.0])
            r += 1
        for i in range(1, self.S):
            xa = self.ps.index_start_x(i)
            if i not in sep:
                acc["mass"][0].extend([r, r])
                acc["mass"][1].extend([xa - 1, xa])
                acc["mass"][2].extend([-1.0, 1.0])
                r += 1
            for key, w in (("position", 3), ("velocity", 3), ("quaternion", 4)):
                acc[key][0].extend(list(range(r, r + w)) * 2)
                acc[key][1].extend(list(range((xa - 1) * w, xa * w)) + list(range(xa * w, (xa + 1) * w)))
                acc[key][2].extend([-1.0] * w + [1.0] * w)
                r += w
        M = self.M
        shapes = {"mass": (r, M), "position": (r, M * 3), "velocity": (r, M * 3), "quaternion": (r, M * 4)}
        return {k: _coo(a[0], a[1], a[2], shapes[k]) for k, a in acc.items()}

    # ================= terminal (con_init_terminal_knot.py:329-405) ====
    def eq_terminal(self, x):
        c = self.c
        pos_f = x["position"][-3:] * self.u["position"]
        vel_f = x["velocity"][-3:] * self.u["velocity"]
        GMe = 3.986004418e14
        if c["altitude_perigee"] is not None and c["altitude_apogee"] is not None:
            c_t = self.crd.angular_momentum_from_altitude(c["altitude_perigee"], c["altitude_apogee"])
            e_t = self.crd.orbit_energy_from_altitude(c["altitude_perigee"], c["altitude_apogee"])
        else:
            c_t = c["radius"] * c["vel_tangential_geocentric"]
            vf = c["vel_tangential_geocentric"] / math.cos(math.radians(c["flightpath_vel_inertial_geocentric"]))
            e_t = vf**2 / 2.0 - GMe / c["radius"]
        rows = [(self.crd.orbit_energy(pos_f, vel_f) / e_t) - 1.0, (self.crd.angular_momentum(pos_f, vel_f) / c_t) - 1.0]
        if c["inclination"] is not None:
            rows.append(self.crd.inclination_rad(pos_f, vel_f) - math.radians(c["inclination"]))
        return np.concatenate(rows, axis=None)

    def eq_terminal_jac(self, x):
        dx = self.dx
        f_c = self.eq_terminal(x)
        nrow = 3 if self.c["inclination"] is not None else 2
        ncol = self.M * 3
        jac = {}
        for key in ("position", "velocity"):
            r, c, d = [], [], []
            for j in range(ncol - 3, ncol):
                x[key][j] += dx
                f_p = self.eq_terminal(x)
                x[key][j] -= dx
                r += list(range(nrow))
                c += [j] * nrow
                d += ((f_p - f_c) / dx).tolist()
            jac[key] = _coo(r, c, d, (nrow, ncol))
        return jac

    # ================= rate / mass / kick (con_trajectory.py) ==========
    def _rate_rows(self):
        """[(coef_col_pairs...)] per row: list of (col, coef) -- :160-347."""
        rows = []
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            att = self.p["params"][i]["attitude"]
            if att in ["hold", "vertical"]:
                rows += [[(c, 1.0)] for c in range(ua * 2, ub * 2)]
            elif att in ("kick-turn", "pitch"):
                rows += [[(ua * 2, -1.0), ((ua + j) * 2, 1.0)] for j in range(1, n)]
                rows += [[((ua + j) * 2 + 1, 1.0)] for j in range(n)]
            elif att == "pitch-yaw":
                rows += [[(ua * 2, -1.0), ((ua + j) * 2, 1.0)] for j in range(1, n)]
                rows += [[(ua * 2 + 1, -1.0), ((ua + j) * 2 + 1, 1.0)] for j in range(1, n)]
            elif att == "same-rate":
                rows += [[(ua * 2 - 2, -1.0), ((ua + j) * 2, 1.0)] for j in range(n)]
                rows += [[(ua * 2 - 1, -1.0), ((ua + j) * 2 + 1, 1.0)] for j in range(n)]
            elif att in ("zero-lift-turn", "free"):
                pass
            else:
                raise SystemExit("ERROR: UNKNOWN ATTITUDE OPTION! ({})".format(att))
        return rows

    def eq_rate(self, x):
        u = x["u"]
        vals = []
        for row in self._rate_rows():
            if len(row) == 1:
                vals.append(u[row[0][0]])
            else:  # u[later] - u[first]
                vals.append(u[row[1][0]] - u[row[0][0]])
        if len(vals) == 0:
            return np.array([])
        return np.array(vals)

    def eq_rate_jac(self, x):
        """COO order of con_trajectory.py:249-347: per block, the -1 column entries
        first, then the +1 entries."""
        r, c, d = [], [], []
        row0 = 0
        for i in range(self.S):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            att = self.p["params"][i]["attitude"]

            def block(first_col, cols):
                nonlocal row0
                k = len(cols)
                if first_col is not None:
                    r.extend(range(row0, row0 + k)); c.extend([first_col] * k); d.extend([-1.0] * k)
                r.extend(range(row0, row0 + k)); c.extend(cols); d.extend([1.0] * k)
                row0 += k

            if att in ["hold", "vertical"]:
                block(None, list(range(ua * 2, (ua + n) * 2)))
            elif att in ("kick-turn", "pitch"):
                block(ua * 2, list(range((ua + 1) * 2, (ua + n) * 2, 2)))
                block(None, list(range(ua * 2 + 1, (ua + n) * 2 + 1, 2)))
            elif att == "pitch-yaw":
                block(ua * 2, list(range((ua + 1) * 2, (ua + n) * 2, 2)))
                block(ua * 2 + 1, list(range((ua + 1) * 2 + 1, (ua + n) * 2 + 1, 2)))
            elif att == "same-rate":
                block(ua * 2 - 2, list(range(ua * 2, (ua + n) * 2, 2)))
                block(ua * 2 - 1, list(range(ua * 2 + 1, (ua + n) * 2 + 1, 2)))
        return {"u": _coo(r, c, d, (row0, self.N * 2))}

    def _stage_ig_co(self):
        names = [p["name"] for p in self.p["params"]]
        return [(self.ps.index_start_x(names.index(st["ignition_at"])),
                 self.ps.index_start_x(names.index(st["cutoff_at"])), st) for st in self.p["RocketStage"].values()]

    def ineq_mass(self, x):
        out = []
        for ig, co, st in self._stage_ig_co():
            d_mass = st["mass_propellant"]
            if st["dropMass"] is not None:
                d_mass += sum([item["mass"] for item in st["dropMass"].values()])
            out.append(-x["mass"][ig] + x["mass"][co] + d_mass / self.u["mass"])
        return out  # a Python list, as the reference returns (con_trajectory.py:61)

    def ineq_mass_jac(self, x):
        r, c, d = [], [], []
        for k, (ig, co, _) in enumerate(self._stage_ig_co()):
            r += [k, k]; c += [ig, co]; d += [-1.0, 1.0]
        return {"mass": _coo(r, c, d, (len(self._stage_ig_co()), len(x["mass"])))}

    def _kick_sections(self):
        return [i for i in range(self.S - 1) if "kick" in self.p["params"][i]["attitude"]]

    def ineq_kick(self, x):
        u = x["u"].reshape(-1, 2) * self.u["u"]
        out = [-u[self.ps.get_index(i)[0] : self.ps.get_index(i)[1], 0] for i in self._kick_sections()]
        if not out:
            return np.zeros(0, dtype="f8")
        return np.concatenate(out, axis=None)

    def ineq_kick_jac(self, x):
        r, c, d = [], [], []
        nrow = 0
        for i in self._kick_sections():
            ua, ub, _, _, n = self.ps.get_index(i)
            r.extend(range(nrow, nrow + n)); c.extend(range(ua * 2, ub * 2, 2)); d.extend([-self.u["u"]] * n)
            nrow += n
        return {"u": _coo(r, c, d, (nrow, len(x["u"])))}

    def _free_time_sections(self):
        prm, idx = self.p["params"], self.p["event_index"]
        return [i for i in range(self.S) if not (prm[i]["time_ref"] in idx and prm[i + 1]["time_ref"] in idx)]

    def ineq_time(self, x):
        t = x["t"]
        return np.array([t[i + 1] - t[i] for i in self._free_time_sections()])

    def ineq_time_jac(self, x):
        r, c, d = [], [], []
        for k, i in enumerate(self._free_time_sections()):
            r += [k, k]; c += [i, i + 1]; d += [-1.0, 1.0]
        return {"t": _coo(r, c, d, (len(self._free_time_sections()), len(x["t"])))}

    # ================= aero (con_aero.py) ==============================
    _AERO = {
        "alpha": ("AOA_max", True, True),
        "q": ("dynamic_pressure_max", False, False),
        "qalpha": ("Q_alpha_max", True, True),
    }

    def _aero_eval(self, kind, pos_e, vel_e, quat, t_e, units):
        """*_array_dimless of con_aero.py:48-86."""
        pos = pos_e * units[0]
        vel = vel_e * units[1]
        t = t_e * units[2]
        w = self.p["wind_table"]
        if kind == "alpha":
            return self.utl.angle_of_attack_all_array_rad(pos, vel, quat, t, w) / units[3]
        if kind == "q":
            return self.utl.dynamic_pressure_array_pa(pos, vel, t, w) / units[3]
        return self.utl.q_alpha_array_pa_rad(pos, vel, quat, t, w) / units[3]

    def _aero_sections(self, kind):
        key, in_rad, _ = self._AERO[kind]
        out = []
        for i in range(self.S - 1):
            name = self.p["params"][i]["name"]
            if name in self.c[key]:
                lim = self.c[key][name]["value"]
                if in_rad:
                    lim = lim * np.pi / 180.0
                rng = self.c[key][name]["range"]
                n = self.ps.nodes(i)
                nk = n + 1 if rng == "all" else 1
                out.append((i, lim, rng, nk))
        return out

    def ineq_aero(self, x, kind):
        units = [self.u["position"], self.u["velocity"], self.u["t"], 1.0]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        quat = x["quaternion"].reshape(-1, 4)
        t = x["t"]
        out = []
        for i, lim, rng, nk in self._aero_sections(kind):
            ua, ub, xa, xb, n = self.ps.get_index(i)
            units[3] = lim
            to, tf = t[i], t[i + 1]
            if rng == "all":
                tn = self.ps.time_nodes(i, to, tf)
                out.append(1.0 - self._aero_eval(kind, pos[xa:xb], vel[xa:xb], quat[xa:xb], tn, units))
            elif rng == "initial":
                out.append(1.0 - self._aero_eval(kind, pos[xa : xa + 1], vel[xa : xa + 1], quat[xa : xa + 1],
                                                 np.array([to]), units)[0])
        if len(out) == 0:
            return None
        return np.concatenate(out, axis=None)

    def ineq_aero_jac(self, x, kind):
        dx = self.dx
        has_quat = self._AERO[kind][2]
        secs = self._aero_sections(kind)
        nrow = sum(s[3] for s in secs)
        if nrow == 0:
            return None
        units = [self.u["position"], self.u["velocity"], self.u["t"], 1.0]
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        quat = x["quaternion"].reshape(-1, 4)
        t = x["t"]
        acc = {k: ([], [], []) for k in ("position", "velocity", "quaternion", "t")}
        r0 = 0
        for i, lim, rng, nk in secs:
            ua, ub, xa, xb, n = self.ps.get_index(i)
            units[3] = lim
            to, tf = t[i], t[i + 1]
            ki = list(range(nk))
            p_, v_, q_ = pos[xa:xb][ki], vel[xa:xb][ki], quat[xa:xb][ki]  # copies (fancy index)
            tk = self.ps.time_nodes(i, to, tf)[ki]

            def f(tt=tk):
                return self._aero_eval(kind, p_, v_, q_, tt, units)

            f_c = f()
            g = {"position": np.zeros((nk, 3)), "velocity": np.zeros((nk, 3)), "quaternion": np.zeros((nk, 4))}
            for key, arr, w in (("position", p_, 3), ("velocity", v_, 3), ("quaternion", q_, 4)):
                if key == "quaternion" and not has_quat:
                    continue
                for j in range(w):
                    arr[:, j] += dx
                    f_p = f()
                    arr[:, j] -= dx
                    g[key][:, j] = (f_p - f_c) / dx
            to_p = to + dx
            g_to = (f(self.ps.time_nodes(i, to_p, tf)[ki]) - f_c) / dx
            tf_p = tf + dx
            g_tf = (f(self.ps.time_nodes(i, to, tf_p)[ki]) - f_c) / dx
            rows = np.arange(r0, r0 + nk)
            for key, w in (("position", 3), ("velocity", 3), ("quaternion", 4)):
                if key == "quaternion" and not has_quat:
                    continue
                for j in range(w):
                    acc[key][0].append(rows)
                    acc[key][1].append(np.arange(xa * w + j, (xa + nk) * w + j, w))
                acc[key][2].append(-g[key].ravel(order="F"))
            acc["t"][0].extend([rows, rows])
            acc["t"][1].extend([np.full(nk, i), np.full(nk, i + 1)])
            acc["t"][2].extend([-g_to, -g_tf])
            r0 += nk
        M = self.M
        shapes = {"position": (nrow, M * 3), "velocity": (nrow, M * 3), "quaternion": (nrow, M * 4),
                  "t": (nrow, self.S + 1)}
        return {k: _coo(_cat(a[0], "i4"), _cat(a[1], "i4"), _cat(a[2]), shapes[k]) for k, a in acc.items()}

    # ================= waypoint / IIP / antenna (con_waypoint.py) ======
    def _wp_sections(self):
        if "waypoint" not in self.c:
            return None
        return [(i, self.c["waypoint"][self.p["params"][i]["name"]]) for i in range(self.S - 1)
                if self.p["params"][i]["name"] in self.c["waypoint"]]

    @staticmethod
    def _no_downrange(wp):
        if "downrange" in wp:
            raise NotImplementedError(
                "downrange waypoints: the reference emits a malformed COO block for them "
                "(con_waypoint.py:704,917,934); unsupported here")

    def _llh_rows(self, wp, eq):
        """[(component, scale_kind, sign, ref)] for position-LLH rows in the
        reference's order (:533-556 eq, :743-778 ineq)."""
        rows = []
        for comp, key in ((0, "lat"), (1, "lon"), (2, "altitude")):
            if key in wp:
                if eq:
                    if "exact" in wp[key]:
                        rows.append((comp, key, +1.0, wp[key]["exact"]))
                else:
                    if "min" in wp[key]:
                        rows.append((comp, key, +1.0, wp[key]["min"]))
                    if "max" in wp[key]:
                        rows.append((comp, key, -1.0, wp[key]["max"]))
        return rows

    def _posllh(self, x, eq):
        secs = self._wp_sections()
        if secs is None:
            return None
        pos = x["position"].reshape(-1, 3)
        out = []
        for i, wp in secs:
            self._no_downrange(wp)
            xa = self.ps.index_start_x(i)
            llh = self.crd.eci2geodetic(pos[xa] * self.u["position"], x["t"][i] * self.u["t"])
            for comp, key, sgn, ref in self._llh_rows(wp, eq):
                if key == "altitude":
                    v = (llh[2] / ref) - 1.0 if sgn > 0 else -(llh[2] / ref) + 1.0
                else:
                    den = 90.0 if key == "lat" else 180.0
                    v = (llh[comp] - ref) / den if sgn > 0 else -(llh[comp] - ref) / den
                out.append(v)
        if len(out) == 0:
            return None
        return np.concatenate(out, axis=None)

    def _posllh_grad(self, pos_row, t_i):
        """posLLH_gradient (:562-580): in place on the xdict row."""
        dx, up, ut = self.dx, self.u["position"], self.u["t"]
        f = lambda p, tt: self.crd.eci2geodetic(p * up, tt * ut)  # noqa: E731
        f_c = f(pos_row, t_i)
        gp = np.zeros((3, 3))
        for j in range(3):
            pos_row[j] += dx
            f_p = f(pos_row, t_i)
            pos_row[j] -= dx
            gp[:, j] = (f_p - f_c) / dx
        t_p = t_i + dx
        gt = (f(pos_row, t_p) - f_c) / dx
        return gp, gt

    def _posllh_jac(self, x, eq):
        f_c = self._posllh(x, eq)
        if f_c is None:
            return None
        nrow = len(f_c)
        pos = x["position"].reshape(-1, 3)
        pr, pc, pd, tr, tc, td = [], [], [], [], [], []
        r = 0
        for i, wp in self._wp_sections():
            xa = self.ps.index_start_x(i)
            for comp, key, sgn, ref in self._llh_rows(wp, eq):
                gp, gt = self._posllh_grad(pos[xa], x["t"][i])
                den = 90.0 if key == "lat" else (180.0 if key == "lon" else ref)
                pr += [r] * 3; pc += list(range(xa * 3, (xa + 1) * 3))
                if sgn > 0:
                    pd += list(gp[comp, :] / den); td.append(gt[comp] / den)
                else:
                    pd += list(-gp[comp, :] / den); td.append(-gt[comp] / den)
                tr.append(r); tc.append(i)
                r += 1
        return {"position": _coo(pr, pc, pd, (nrow, self.M * 3)), "t": _coo(tr, tc, td, (nrow, self.S + 1))}

    def eq_pos(self, x): return self._posllh(x, True)
    def eq_pos_jac(self, x): return self._posllh_jac(x, True)
    def ineq_pos(self, x): return self._posllh(x, False)
    def ineq_pos_jac(self, x): return self._posllh_jac(x, False)

    def _iip_rows(self, wp, eq):
        rows = []
        for comp, key, den in ((0, "lat_IIP", 90.0), (1, "lon_IIP", 180.0)):
            if key in wp:
                if eq:
                    if "exact" in wp[key]:
                        rows.append((comp, den, +1.0, wp[key]["exact"]))
                else:
                    if "min" in wp[key]:
                        rows.append((comp, den, +1.0, wp[key]["min"]))
                    if "max" in wp[key]:
                        rows.append((comp, den, -1.0, wp[key]["max"]))
        return rows

    def _iip_from_eci(self, p_, v_, t_):
        up, uv, ut = self.u["position"], self.u["velocity"], self.u["t"]
        return self.iip.posLLH_IIP_FAA(self.crd.eci2ecef(p_ * up, t_ * ut),
                                       self.crd.vel_eci2ecef(v_ * uv, p_ * up, t_ * ut))

    def _iip(self, x, eq):
        secs = self._wp_sections()
        if secs is None:
            return None
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        out = []
        for i, wp in secs:
            xa = self.ps.index_start_x(i)
            llh = self._iip_from_eci(pos[xa], vel[xa], x["t"][i])
            for comp, den, sgn, ref in self._iip_rows(wp, eq):
                out.append((llh[comp] - ref) / den if sgn > 0 else (ref - llh[comp]) / den)
        if len(out) == 0:
            return None
        return np.concatenate(out, axis=None)

    def _iip_grad(self, pos_row, vel_row, t_i):
        """posLLH_IIP_gradient (:210-240): pos_j and vel_j interleaved, in place."""
        dx = self.dx
        f_c = self._iip_from_eci(pos_row, vel_row, t_i)
        gp, gv = np.zeros((3, 3)), np.zeros((3, 3))
        for j in range(3):
            pos_row[j] += dx
            f_p = self._iip_from_eci(pos_row, vel_row, t_i)
            pos_row[j] -= dx
            gp[:, j] = (f_p - f_c) / dx
            vel_row[j] += dx
            f_p = self._iip_from_eci(pos_row, vel_row, t_i)
            vel_row[j] -= dx
            gv[:, j] = (f_p - f_c) / dx
        t_p = t_i + dx
        gt = (self._iip_from_eci(pos_row, vel_row, t_p) - f_c) / dx
        return gp, gv, gt

    def _iip_jac(self, x, eq):
        f_c = self._iip(x, eq)
        if f_c is None:
            return None
        nrow = len(f_c)
        pos = x["position"].reshape(-1, 3)
        vel = x["velocity"].reshape(-1, 3)
        acc = {k: ([], [], []) for k in ("position", "velocity", "t")}
        r = 0
        for i, wp in self._wp_sections():
            xa = self.ps.index_start_x(i)
            for comp, den, sgn, ref in self._iip_rows(wp, eq):
                gp, gv, gt = self._iip_grad(pos[xa], vel[xa], x["t"][i])
                for key, g in (("position", gp), ("velocity", gv)):
                    acc[key][0].extend([r] * 3)
                    acc[key][1].extend(range(xa * 3, (xa + 1) * 3))
                    acc[key][2].extend(g[comp, :] / den if sgn > 0 else -g[comp, :] / den)
                acc["t"][0].append(r); acc["t"][1].append(i)
                acc["t"][2].append(gt[comp] / den if sgn > 0 else -gt[comp] / den)
                r += 1
        shapes = {"position": (nrow, self.M * 3), "velocity": (nrow, self.M * 3), "t": (nrow, self.S + 1)}
        return {k: _coo(a[0], a[1], a[2], shapes[k]) for k, a in acc.items()}

    def eq_iip(self, x): return self._iip(x, True)
    def eq_iip_jac(self, x): return self._iip_jac(x, True)
    def ineq_iip(self, x): return self._iip(x, False)
    def ineq_iip_jac(self, x): return self._iip_jac(x, False)

    def _antenna_rows(self):
        if "antenna" not in self.c:
            return None
        rows = []
        for ant in self.c["antenna"].values():
            p_ant = self.crd.geodetic2ecef(ant["lat"], ant["lon"], ant["altitude"])
            for i in range(self.S - 1):
                name = self.p["params"][i]["name"]
                if name in ant["elevation_min"]:
                    rows.append((i, p_ant, ant["elevation_min"][name]))
        return rows

    def _sin_elev(self, pos_row, t_, p_ant):
        """sin_elevation (:45-51)."""
        pos = pos_row * self.u["position"]
        to = t_ * self.u["t"]
        p_ecef = self.crd.eci2ecef(pos, to)
        direction = self.crd.normalize(p_ecef - p_ant)
        vertical = self.crd.quatrot(self.crd.quat_nedg2ecef(p_ant), np.array([0, 0, -1.0]))
        return float(self._dot(np.asarray(direction).reshape(1, 3), np.asarray(vertical).reshape(3, 1)).ravel()[0])

    def ineq_antenna(self, x):
        rows = self._antenna_rows()
        if rows is None:
            return None
        pos = x["position"].reshape(-1, 3)
        out = [self._sin_elev(pos[self.ps.index_start_x(i)], x["t"][i], p_ant) - np.sin(el * np.pi / 180.0)
               for i, p_ant, el in rows]
        if len(out) == 0:
            return None
        return np.concatenate(out, axis=None)

    def ineq_antenna_jac(self, x):
        f_c = self.ineq_antenna(x)
        if f_c is None:
            return None
        dx = self.dx
        pos = x["position"].reshape(-1, 3)
        pr, pc, pd, tr, tc, td = [], [], [], [], [], []
        for r, (i, p_ant, _) in enumerate(self._antenna_rows()):
            xa = self.ps.index_start_x(i)
            row = pos[xa]
            t_i = x["t"][i]
            fc = self._sin_elev(row, t_i, p_ant)
            for j in range(3):
                row[j] += dx
                fp = self._sin_elev(row, t_i, p_ant)
                row[j] -= dx
                pd.append((fp - fc) / dx)
            t_p = t_i + dx
            td.append((self._sin_elev(row, t_p, p_ant) - fc) / dx)
            pr += [r] * 3; pc += list(range(xa * 3, (xa + 1) * 3)); tr.append(r); tc.append(i)
        nrow = len(f_c)
        return {"position": _coo(pr, pc, pd, (nrow, self.M * 3)), "t": _coo(tr, tc, td, (nrow, self.S + 1))}

    # ================= user constraints (con_user.py, jac_fd.py) =======
    def jac_fd(self, con, x):
        """lib/jac_fd.py:29-62 -- dense forward differences over every variable,
        in place, key order = xdict insertion order."""
        dx = self.dx
        g0 = con(x, self.p, self.u, self.c)
        nrows = len(g0) if hasattr(g0, "__len__") else 1
        jac = {}
        for key, val in x.items():
            jac[key] = np.zeros((nrows, val.size))
            for i in range(val.size):
                x[key][i] += dx
                g_p = con(x, self.p, self.u, self.c)
                jac[key][:, i] = (g_p - g0) / dx
                x[key][i] -= dx
        return jac

    def _user(self, x, fn):
        return None if fn is None else fn(x, self.p, self.u, self.c)

    def _user_jac(self, x, fn):
        if fn is not None and fn(x, self.p, self.u, self.c) is not None:
            return self.jac_fd(fn, x)
        return None

    # ================= the two callbacks ===============================
    def objfunc(self, x):
        """Trajectory_Optimization.py:194-242."""
        f = {"obj": self.cost(x)}
        f["eqcon_init"] = self.eq_init(x)
        f["eqcon_time"] = self.eq_time(x)
        f["eqcon_dyn_mass"] = self.eq_dyn_mass(x)
        f["eqcon_dyn_pos"] = self.eq_dyn_pos(x)
        f["eqcon_dyn_vel"] = self.eq_dyn_vel(x)
        f["eqcon_dyn_quat"] = self.eq_dyn_quat(x)
        f["eqcon_knot"] = self.eq_knot(x)
        f["eqcon_terminal"] = self.eq_terminal(x)
        f["eqcon_rate"] = self.eq_rate(x)
        f["eqcon_pos"] = self.eq_pos(x)
        f["eqcon_iip"] = self.eq_iip(x)
        f["eqcon_user"] = self._user(x, self.user_eq)
        f["ineqcon_alpha"] = self.ineq_aero(x, "alpha")
        f["ineqcon_q"] = self.ineq_aero(x, "q")
        f["ineqcon_qalpha"] = self.ineq_aero(x, "qalpha")
        f["ineqcon_mass"] = self.ineq_mass(x)
        f["ineqcon_kick"] = self.ineq_kick(x)
        f["ineqcon_time"] = self.ineq_time(x)
        f["ineqcon_pos"] = self.ineq_pos(x)
        f["ineqcon_iip"] = self.ineq_iip(x)
        f["ineqcon_antenna"] = self.ineq_antenna(x)
        f["ineqcon_user"] = self._user(x, self.user_ineq)
        return f, False

    def sens(self, x, funcs=None):
        """Trajectory_Optimization.py:245-312 (evaluation ORDER matters: groups
        perturb `x` in place and leave rounding residue for later groups)."""
        s = {"obj": self.cost_jac(x)}
        s["eqcon_init"] = self.eq_init_jac(x)
        s["eqcon_time"] = self.eq_time_jac(x)
        s["eqcon_dyn_mass"] = self.eq_dyn_mass_jac(x)
        s["eqcon_dyn_pos"] = self.eq_dyn_pos_jac(x)
        s["eqcon_dyn_vel"] = self.eq_dyn_vel_jac(x)
        s["eqcon_dyn_quat"] = self.eq_dyn_quat_jac(x)
        s["eqcon_knot"] = self.eq_knot_jac(x)
        s["eqcon_terminal"] = self.eq_terminal_jac(x)
        s["eqcon_rate"] = self.eq_rate_jac(x)
        s["eqcon_pos"] = self.eq_pos_jac(x)
        s["eqcon_iip"] = self.eq_iip_jac(x)
        s["eqcon_user"] = self._user_jac(x, self.user_eq)
        s["ineqcon_alpha"] = self.ineq_aero_jac(x, "alpha")
        s["ineqcon_q"] = self.ineq_aero_jac(x, "q")
        s["ineqcon_qalpha"] = self.ineq_aero_jac(x, "qalpha")
        s["ineqcon_mass"] = self.ineq_mass_jac(x)
        s["ineqcon_kick"] = self.ineq_kick_jac(x)
        s["ineqcon_time"] = self.ineq_time_jac(x)
        s["ineqcon_pos"] = self.ineq_pos_jac(x)
        s["ineqcon_iip"] = self.ineq_iip_jac(x)
        s["ineqcon_antenna"] = self.ineq_antenna_jac(x)
        s["ineqcon_user"] = self._user_jac(x, self.user_ineq)
        return s, False
