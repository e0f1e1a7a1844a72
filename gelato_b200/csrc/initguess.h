/* initguess.h -- the forward-simulation initial guess as a per-thread function: ONE THREAD PER SCENARIO integrates
 * the 3-DoF equations of motion through the whole event schedule with the classical Runge-Kutta scheme and
 * interpolates the states at the requested times (the LGR state nodes of the mesh).
 *
 * Reference: /root/reference/initialize.py:37-111 dynamics_init, :114-179 rocket_simulation, :182-221
 * zerolift_turn_correct, :229-235 integrate_runge_kutta_4d -- same operations in the same order (the oracle's
 * restatement, oracle/initguess.py, is live-checked against those functions; this is its bit-twin on gmath).  The
 * reference runs it once per settings file in Python (126 000 steps of 4 right-hand sides at its dt = 0.005 s);
 * a dispersed-scenario study needs one per scenario, which is what the batch is for.
 *
 * Details that matter for the bits:
 *  - norm(v) is numpy.linalg.norm: sqrt of a BLAS dot, i.e. fused multiply-adds in ascending index (np_norm3);
 *  - the drag-table and rate-table look-ups are numpy.interp (np_interp), not the C++ `interp` of the NLP path;
 *  - a stage's jettison is applied IN PLACE to the state array that is also the last recorded state, so the
 *    recorded history sees it one step early: the interpolation below runs one step behind the integration;
 *  - t is advanced as t = t + dt (the rounding of the accumulated time decides when an event switches).
 * host+device: tests/emu steps the same function on the CPU.
 */
#ifndef GELATO_B200_INITGUESS_H_
#define GELATO_B200_INITGUESS_H_

#include "physics.h"

enum { GI_TIME = 0, GI_THRUST, GI_MASSFLOW, GI_REF_AREA, GI_NOZZLE_AREA, GI_JETTISON, GI_COLS };

/* air-relative velocity in ECI axes and the geopotential altitude (initialize.py:61-72, :203-210) */
P_HD Vec3 init_air_velocity(Vec3 pos, Vec3 vel, double t, const Tables& tb, double* altitude_m) {
  const Geodetic g = ecef2geodetic<1>(pos);
  const double alt = geopotential_altitude(g.alt);
  double s, c;
  gm_sincos(P_OMEGA * t, &s, &c);
  const Vec3 vel_ecef = vel_eci2ecef_cs(vel, pos, c, s);
  const Vec3 wind = v3(interp_table(alt, tb.wind, tb.wind + 1, tb.n_wind, 3), interp_table(alt, tb.wind, tb.wind + 2, tb.n_wind, 3), 0.0);
  const Vec3 wind_eci = quatrot(quat_ned2eci_cs(pos, P_OMEGA * t, c, s), wind);
  *altitude_m = alt;
  return sub3(rot_ecef2eci(vel_ecef, c, s), wind_eci);
}

/* dynamics_init (zlt = False: the reference never switches it on, the zero-lift turn is the correction below).
 * x = mass, position[3], velocity[3], quaternion[4]; u = roll, pitch, yaw rates [deg/s]; ev = one row of the event table */
P_HD void init_dynamics(const double* x, const double* u, double t, const double* ev, const Tables& tb, double* ret) {
  const double mass = x[0];
  const Vec3 pos = v3(x[1], x[2], x[3]), vel = v3(x[4], x[5], x[6]);
  const Quat q = q4(x[7], x[8], x[9], x[10]);
  double alt;
  const Vec3 va = init_air_velocity(pos, vel, t, tb, &alt);
  const AirState as = us76(alt, 3);
  const double vn = np_norm3(va);
  const double mach = vn / as.a;
  const double ca = np_interp(mach, tb.ca, tb.ca + 1, tb.n_ca, 2);
  const double k = 0.5 * as.rho * vn; /* 0.5 * rho * norm(v) * -v * area * ca, left to right */
  const Vec3 aero = v3(k * -va.x * ev[GI_REF_AREA] * ca, k * -va.y * ev[GI_REF_AREA] * ca, k * -va.z * ev[GI_REF_AREA] * ca);
  const double thrust = ev[GI_THRUST] - ev[GI_NOZZLE_AREA] * as.P;
  const Vec3 tdir = quatrot(quatconj(q), v3(1.0, 0.0, 0.0));
  const Vec3 thr = v3(tdir.x * thrust, tdir.y * thrust, tdir.z * thrust);
  const Vec3 grav = gravity_eci(pos);
  const double d2r = 0.017453292519943295; /* numpy.deg2rad: x * (pi / 180) */
  const Quat dq = quatmult(q, q4(0.0 * d2r, u[0] * d2r, u[1] * d2r, u[2] * d2r));
  ret[0] = -ev[GI_MASSFLOW];
  ret[1] = vel.x; ret[2] = vel.y; ret[3] = vel.z;
  ret[4] = grav.x + (thr.x + aero.x) / mass;
  ret[5] = grav.y + (thr.y + aero.y) / mass;
  ret[6] = grav.z + (thr.z + aero.z) / mass;
  ret[7] = 0.5 * dq.w; ret[8] = 0.5 * dq.x; ret[9] = 0.5 * dq.y; ret[10] = 0.5 * dq.z;
}

/* dynamic-vector normalize of the C++ wrapper (wrapper_coordinate.hpp:64-68): v / v.norm(), sequential sum */
P_HD void init_normalize(double* v, int n) {
  double s = v[0] * v[0];
  for (int i = 1; i < n; i++) s = s + v[i] * v[i];
  const double nrm = gm_sqrt(s);
  for (int i = 0; i < n; i++) v[i] = v[i] / nrm;
}

/* zerolift_turn_correct: body x axis along the air velocity, zero roll (initialize.py:182-221) */
P_HD void init_zerolift_quat(const double* x, double t, const Tables& tb, double* qout) {
  const Vec3 pos = v3(x[1], x[2], x[3]), vel = v3(x[4], x[5], x[6]);
  double alt;
  const Vec3 va = init_air_velocity(pos, vel, t, tb, &alt);
  double xb[3] = {va.x, va.y, va.z};
  init_normalize(xb, 3);
  const Vec3 cr = cross3(va, pos);
  double yb[3] = {cr.x, cr.y, cr.z};
  init_normalize(yb, 3);
  const Vec3 zb = cross3(v3(xb[0], xb[1], xb[2]), v3(yb[0], yb[1], yb[2]));
  const double q0 = 0.5 * gm_sqrt(1.0 + xb[0] + yb[1] + zb.z);
  qout[0] = q0;
  qout[1] = 0.25 / q0 * (yb[2] - zb.y);
  qout[2] = 0.25 / q0 * (zb.x - xb[2]);
  qout[3] = 0.25 / q0 * (xb[1] - yb[0]);
  init_normalize(qout, 4);
}

/* rocket_simulation for one scenario.  ev[n_ev][GI_COLS] (times ascending), zlt[n_ev] (attitude == zero-lift-turn),
 * u_table[n_u][4] (time, roll, pitch, yaw), t_out[n_out] ascending; x_out[n_out][11]; u_out[n_out][3] or null (the
 * reference's second return value: the rate history, whose entry k is the rate used by the step that ENDED at time k). */
struct InitRecord {
  double x[11], u[3];
};
P_HD void init_emit(const InitRecord& a, double ta, const InitRecord& b, double tb_, double to, double* xo, double* uo) {
  for (int i = 0; i < 11; i++) {
    const double slope = (b.x[i] - a.x[i]) / (tb_ - ta);
    xo[i] = slope * (to - ta) + a.x[i];
  }
  if (uo)
    for (int i = 0; i < 3; i++) {
      const double slope = (b.u[i] - a.u[i]) / (tb_ - ta);
      uo[i] = slope * (to - ta) + a.u[i];
    }
}
P_HD void init_emit_flat(const InitRecord& a, double* xo, double* uo) {
  for (int i = 0; i < 11; i++) xo[i] = a.x[i];
  if (uo)
    for (int i = 0; i < 3; i++) uo[i] = a.u[i];
}

P_HD void rocket_simulation_thread(const double* x_init, const double* ev, const int32_t* zlt, int n_ev, const double* u_table,
                                   int n_u, const Tables& tb, double t_init, const double* t_out, int n_out, double dt,
                                   double* x_out, double* u_out) {
  InitRecord prev, cur;
  double k1[11], k2[11], k3[11], k4[11], xa[11], u[3];
  for (int i = 0; i < 11; i++) cur.x[i] = prev.x[i] = x_init[i];
  for (int i = 0; i < 3; i++) cur.u[i] = prev.u[i] = 0.0;
  double t = t_init, t_prev = t_init;
  const double t_final = t_out[n_out - 1];
  int event_index = -1, io = 0;
  double evp[GI_COLS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; /* param starts as zeros(5) */
  bool have_prev = false;
  while (t < t_final) {
    const double tn = t + dt;
    if (event_index < n_ev - 1 && tn > ev[(event_index + 1) * GI_COLS + GI_TIME]) {
      event_index += 1;
      for (int c = 0; c < GI_COLS; c++) evp[c] = ev[event_index * GI_COLS + c];
      cur.x[0] -= evp[GI_JETTISON]; /* in place: the recorded state of this time sees the jettison too */
    }
    /* `cur` (recorded at time t) is final now: emit the outputs that fall before it */
    if (!have_prev) {
      while (io < n_out && t_out[io] <= t) { /* numpy.interp: left value up to and including the first abscissa */
        init_emit_flat(cur, x_out + io * 11, u_out ? u_out + io * 3 : nullptr);
        io++;
      }
    } else {
      while (io < n_out && t_out[io] < t) {
        init_emit(prev, t_prev, cur, t, t_out[io], x_out + io * 11, u_out ? u_out + io * 3 : nullptr);
        io++;
      }
    }
    for (int i = 0; i < 3; i++) u[i] = np_interp(t, u_table, u_table + 1 + i, n_u, 4);
    /* integrate_runge_kutta_4d */
    init_dynamics(cur.x, u, t, evp, tb, k1);
    for (int i = 0; i < 11; i++) xa[i] = cur.x[i] + dt / 2.0 * k1[i];
    init_dynamics(xa, u, t + dt / 2.0, evp, tb, k2);
    for (int i = 0; i < 11; i++) xa[i] = cur.x[i] + dt / 2.0 * k2[i];
    init_dynamics(xa, u, t + dt / 2.0, evp, tb, k3);
    for (int i = 0; i < 11; i++) xa[i] = cur.x[i] + dt * k3[i];
    init_dynamics(xa, u, t + dt, evp, tb, k4);
    prev = cur;
    for (int i = 0; i < 11; i++) cur.x[i] = cur.x[i] + (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0 * dt;
    for (int i = 0; i < 3; i++) cur.u[i] = u[i];
    t_prev = t;
    t = t + dt;
    have_prev = true;
    /* prm[event_index] with event_index == -1 is Python's LAST row (no event has started yet) */
    if (zlt[event_index >= 0 ? event_index : n_ev - 1]) init_zerolift_quat(cur.x, t, tb, cur.x + 7);
    init_normalize(cur.x + 7, 4);
  }
  /* the last interval and everything at or beyond the last recorded time */
  while (io < n_out) {
    if (have_prev && t_out[io] < t) init_emit(prev, t_prev, cur, t, t_out[io], x_out + io * 11, u_out ? u_out + io * 3 : nullptr);
    else init_emit_flat(cur, x_out + io * 11, u_out ? u_out + io * 3 : nullptr);
    io++;
  }
}

#endif /* GELATO_B200_INITGUESS_H_ */
