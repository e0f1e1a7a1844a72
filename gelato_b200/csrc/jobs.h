/* jobs.h -- the work items of the two fused kernels, as per-thread functions.
 *
 * A kernel block has a role (a row of the block table) and runs its job in two
 * phases separated by one block barrier:
 *   phase 1  every thread evaluates ONE leaf (a right-hand side, an aero
 *            quantity, an event-point function) on its own perturbed copy of
 *            the inputs and parks the result in shared memory;
 *   phase 2  every thread forms its finite-difference quotients against the
 *            centre evaluation of its group and scatters them to their COO
 *            slots (Jacobian kernel), or combines D.X with the right-hand side
 *            into residual rows (residual kernel).
 * The functions are host+device so tests/emu can step through the same code on
 * the CPU; the shipped library only runs them inside CUDA kernels.
 *
 * Reference formulas: see the citations on each function; the perturbation
 * protocol (which inputs carry fl(fl(x+dx)-dx) residue when a given column is
 * evaluated) is DESIGN.md "H3" / SURVEY.md A.4.
 */
#ifndef GELATO_B200_JOBS_H_
#define GELATO_B200_JOBS_H_

#include "../../include/gelato_b200.h"
#include "physics.h"

#define GB_THREADS 128
#define GB_DYN_JAC_NODES 8   /* 16 lanes per node */
#define GB_DYN_RES_NODES 32  /* warp 0 = one thread per node */
#define GB_ROWS16 8          /* 16-lane jobs per block (aero / event Jacobian) */

/* block roles */
enum { BR_DYN = 0, BR_AERO = 1, BR_EVT = 2, BR_LIN = 3 };
/* block table columns */
enum { BT_ROLE = 0, BT_JOB, BT_START, BT_COUNT, BT_COLS };

struct PlanView {
  int S, N, M, n_vars, n_rows;
  long long n_vals;
  int payload_mode;
  int off_pos, off_vel, off_quat, off_u, off_t; /* offsets into x (mass at 0) */
  const int32_t* sec_i32;
  const int64_t* sec_i64;
  const double* sec_f64;
  long long sec_f64_sstride; /* per-scenario stride (0 = shared) */
  const double* d_pool;
  const double* tau_pool;
  const double* wind;
  long long wind_sstride;
  int n_wind;
  const double* ca;
  int n_ca;
  Units un;
  const double* unit_mass_scen; /* [n_scen] or NULL */
  int n_lin;
  const int32_t* lin_i32;
  const double* lin_f64;
  const double* lin_const_scen; /* [n_scen][n_lin] or NULL */
  int n_aero;
  const int32_t* aero_i32;
  const int64_t* aero_i64;
  const double* aero_f64;
  const uint8_t* rc_aero;
  int n_evt;
  const int32_t* evt_i32;
  const int64_t* evt_i64;
  const double* evt_f64;
};

/* scratch shared by the threads of one block */
struct BlockScratch {
  double f[GB_THREADS][4];
  double q[GB_THREADS][4];
};

/* fl(fl(x + dx) - dx): what a perturb/restore cycle leaves behind */
P_HD double residue(double x, double dx) {
  double y = x + dx;
  return y - dx;
}
P_HD double residue_n(double x, double dx, int n) {
  for (int i = 0; i < n; i++) x = residue(x, dx);
  return x;
}

P_HD Units scen_units(const PlanView& P, int scen) {
  Units u = P.un;
  if (P.unit_mass_scen) u.mass = P.unit_mass_scen[scen];
  return u;
}
P_HD SecParam sec_param(const PlanView& P, int scen, int sec) {
  const double* f = P.sec_f64 + (long long)scen * P.sec_f64_sstride + (long long)sec * GS_F64_COLS;
  SecParam sp;
  sp.thrust = f[GS_THRUST];
  sp.massflow = f[GS_MASSFLOW];
  sp.ref_area = f[GS_REF_AREA];
  sp.nozzle_area = f[GS_NOZZLE_AREA];
  return sp;
}
P_HD Tables scen_tables(const PlanView& P, int scen) {
  Tables tb;
  tb.wind = P.wind + (long long)scen * P.wind_sstride;
  tb.n_wind = P.n_wind;
  tb.ca = P.ca;
  tb.n_ca = P.n_ca;
  return tb;
}

/* PSparams.time_nodes (SectionParameters.py:77-81): row 0 = to, row r>=1 = LGR point r-1 */
P_HD double time_node(const double* tau, int r, double to, double tf) {
  if (r == 0) return to;
  return tau[r - 1] * (tf - to) / 2 + (tf + to) / 2;
}

/* ========================================================================= */
/* Jacobian kernel, DYN role: 16 lanes per node                              */
/*   lane 0 centre | 1 mass | 2-4 position | 5-7 velocity | 8-11 quaternion | */
/*   12 to | 13 tf | 14 dyn_pos entries | 15 idle                             */
/*   lanes 0-6 also evaluate the quaternion kinematics (centre, q x4, u x2)   */
/* reference: con_dynamics.py:292-496 (velocity), :536-632 (quaternion),      */
/*            :155-213 (position)                                             */
/* ========================================================================= */
P_HD void dyn_jac_phase1(const PlanView& P, int scen, const double* x, int sec, int node0, int count, int tid,
                         BlockScratch& sm) {
  const int nl = tid >> 4, lane = tid & 15;
  if (nl >= count) return;
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  const int n = si[GS_N], ua = si[GS_UA], xa = si[GS_XA], flags = si[GS_FLAGS];
  const int j = node0 + nl;   /* LGR node index inside the section */
  const int row = xa + 1 + j; /* state row */
  const Units un = scen_units(P, scen);
  const double dx = un.dx;
  const bool air = flags & GSF_AIR, air_fd = flags & GSF_AIR_FD, hold = flags & GSF_HOLD;
  (void)n;

  /* velocity dynamics */
  bool active = (lane <= 4) || (lane >= 8 && lane <= 11) || ((lane >= 5 && lane <= 7) && air_fd) ||
                ((lane == 12 || lane == 13) && air_fd);
  if (active) {
    double v[11]; /* mass, pos[3], vel[3], quat[4] */
    v[0] = x[row];
    for (int k = 0; k < 3; k++) v[1 + k] = x[P.off_pos + 3 * row + k];
    for (int k = 0; k < 3; k++) v[4 + k] = x[P.off_vel + 3 * row + k];
    for (int k = 0; k < 4; k++) v[7 + k] = x[P.off_quat + 4 * row + k];
    if (lane != 0) {
      const int pidx = (lane <= 11) ? lane - 1 : 11;
#pragma unroll
      for (int w = 0; w < 11; w++) {
        const bool in_protocol = (w < 4) || (w >= 7) || air_fd;
        if (!in_protocol) continue;
        if (w < pidx) v[w] = residue(v[w], dx);
        else if (w == pidx) v[w] = v[w] + dx;
      }
    }
    const double to0 = x[P.off_t + sec], tf0 = x[P.off_t + sec + 1];
    const double to = (lane == 12) ? to0 + dx : to0;
    const double tf = (lane == 13) ? tf0 + dx : tf0;
    const SecParam sp = sec_param(P, scen, sec);
    Vec3 f;
    Quat q = q4(v[7], v[8], v[9], v[10]);
    if (air) {
      const double tn = time_node(P.tau_pool + si[GS_TAU_OFF], j + 1, to, tf);
      f = rhs_velocity_air(v[0], v3(v[1], v[2], v[3]), v3(v[4], v[5], v[6]), q, tn, sp, un, scen_tables(P, scen));
    } else {
      f = rhs_velocity_noair(v[0], v3(v[1], v[2], v[3]), q, sp, un);
    }
    sm.f[tid][0] = f.x;
    sm.f[tid][1] = f.y;
    sm.f[tid][2] = f.z;
  }

  /* quaternion kinematics: state already carries one residue from the velocity pass */
  if (!hold && lane <= 6) {
    double qv[4], uv[2];
    for (int k = 0; k < 4; k++) qv[k] = residue(x[P.off_quat + 4 * row + k], dx);
    for (int k = 0; k < 2; k++) uv[k] = x[P.off_u + 2 * (ua + j) + k];
    if (lane >= 1 && lane <= 4) {
      const int k = lane - 1;
      for (int w = 0; w < k; w++) qv[w] = residue(qv[w], dx);
      qv[k] = qv[k] + dx;
    } else if (lane >= 5) {
      const int k = lane - 5;
      for (int w = 0; w < 4; w++) qv[w] = residue(qv[w], dx);
      for (int w = 0; w < k; w++) uv[w] = residue(uv[w], dx);
      uv[k] = uv[k] + dx;
    }
    Quat d = rhs_quaternion(q4(qv[0], qv[1], qv[2], qv[3]), uv[0], uv[1], un.u);
    sm.q[tid][0] = d.w;
    sm.q[tid][1] = d.x;
    sm.q[tid][2] = d.y;
    sm.q[tid][3] = d.z;
  }
}

P_HD void dyn_jac_phase2(const PlanView& P, int scen, const double* x, double* vals, int sec, int node0, int count,
                         int tid, const BlockScratch& sm) {
  const int nl = tid >> 4, lane = tid & 15;
  if (nl >= count) return;
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  const int64_t* sj = P.sec_i64 + sec * GS_I64_COLS;
  const int n = si[GS_N], xa = si[GS_XA], flags = si[GS_FLAGS];
  const int j = node0 + nl;
  const int row = xa + 1 + j;
  const Units un = scen_units(P, scen);
  const double dx = un.dx, ut = un.t;
  const bool air_fd = flags & GSF_AIR_FD, hold = flags & GSF_HOLD;
  const double to = x[P.off_t + sec], tf = x[P.off_t + sec + 1];
  const double dt = tf - to;
  const int c = tid & ~15; /* centre lane of this node */
  const double* D = P.d_pool + si[GS_D_OFF];
  const double d_diag = D[(long long)j * (n + 1) + (j + 1)];
  const long long n3 = 3LL * n, n4 = 4LL * n, nn1 = (long long)n * (n + 1);

  /* ---- velocity dynamics: -(f_p - f_c)/dx*(tf-to)*unit_t/2 (con_dynamics.py:372) ---- */
  if (lane >= 1 && lane <= 11 && !(lane >= 5 && lane <= 7 && !air_fd)) {
    double rh[3];
    for (int k = 0; k < 3; k++) rh[k] = -(sm.f[tid][k] - sm.f[c][k]) / dx * dt * ut / 2.0;
    if (lane == 1) {
      for (int k = 0; k < 3; k++) vals[sj[GS_JV_MASS] + 3LL * j + k] = rh[k];
    } else if (lane <= 4) {
      const int kk = lane - 2;
      for (int k = 0; k < 3; k++) vals[sj[GS_JV_POS] + kk * n3 + 3LL * j + k] = rh[k];
    } else if (lane <= 7) {
      const int kk = lane - 5; /* submat_vel[3j+ki, 3(j+1)+kk] += rh[ki]  (:415-416) */
      for (int ki = 0; ki < 3; ki++) {
        const double base = (ki == kk) ? d_diag : 0.0;
        vals[sj[GS_JV_VEL] + (ki * 3 + kk) * nn1 + (long long)j * (n + 1) + (j + 1)] = base + rh[ki];
      }
    } else {
      const int kk = lane - 8;
      for (int k = 0; k < 3; k++) vals[sj[GS_JV_QUAT] + kk * n3 + 3LL * j + k] = rh[k];
    }
  }
  if (lane == 12) {
    if (air_fd) { /* :454-465 */
      const double to_p = to + dx;
      for (int k = 0; k < 3; k++)
        vals[sj[GS_JV_T] + 3LL * j + k] = -(sm.f[tid][k] * (tf - to_p) - sm.f[c][k] * dt) / dx * ut / 2.0;
    } else { /* :478-480 */
      for (int k = 0; k < 3; k++) {
        const double rh_to = sm.f[c][k] * ut / 2.0;
        vals[sj[GS_JV_T] + 3LL * j + k] = rh_to;
        vals[sj[GS_JV_T] + n3 + 3LL * j + k] = -rh_to;
      }
    }
  }
  if (lane == 13 && air_fd) { /* :466-477 */
    const double tf_p = tf + dx;
    for (int k = 0; k < 3; k++)
      vals[sj[GS_JV_T] + n3 + 3LL * j + k] = -(sm.f[tid][k] * (tf_p - to) - sm.f[c][k] * dt) / dx * ut / 2.0;
  }
  /* ---- position dynamics (analytic, depends on x through vel and t): :180-195 ---- */
  if (lane == 14) {
    const double rh_vel = -un.vel * dt * ut / 2.0 / un.pos;
    for (int k = 0; k < 3; k++) {
      vals[sj[GS_JP_VEL] + 3LL * j + k] = rh_vel;
      const double rh_to = x[P.off_vel + 3 * row + k] * un.vel * ut / 2.0 / un.pos;
      vals[sj[GS_JP_T] + 3LL * j + k] = rh_to;
      vals[sj[GS_JP_T] + n3 + 3LL * j + k] = -rh_to;
    }
  }
  /* ---- quaternion kinematics: :580-625 ---- */
  if (!hold) {
    if (lane >= 1 && lane <= 6) {
      double rh[4];
      for (int a = 0; a < 4; a++) rh[a] = -(sm.q[tid][a] - sm.q[c][a]) / dx * dt * ut / 2.0;
      if (lane <= 4) {
        const int kk = lane - 1; /* submat_quat[4j+a, 4(j+1)+kk] += rh[a] */
        for (int a = 0; a < 4; a++) {
          const double base = (a == kk) ? d_diag : 0.0;
          vals[sj[GS_JQ_QUAT] + (4LL * j + a) * (4LL * (n + 1)) + 4LL * (j + 1) + kk] = base + rh[a];
        }
      } else {
        const int kk = lane - 5;
        for (int a = 0; a < 4; a++) vals[sj[GS_JQ_U] + kk * n4 + 4LL * j + a] = rh[a];
      }
    } else if (lane == 0) {
      for (int a = 0; a < 4; a++) {
        const double rh_to = sm.q[c][a] * ut / 2.0;
        vals[sj[GS_JQ_T] + 4LL * j + a] = rh_to;
        vals[sj[GS_JQ_T] + n4 + 4LL * j + a] = -rh_to;
      }
    }
  }
}

/* ========================================================================= */
/* Residual kernel, DYN role: up to 32 nodes of one section per block.        */
/*   phase 1: threads 0..count-1 evaluate the right-hand sides of their node  */
/*   phase 2: all threads sweep the (node, state column) items: D.X - rhs     */
/* reference: con_dynamics.py:34-63, 116-152, 216-289, 499-533                */
/* ========================================================================= */
P_HD void dyn_res_phase1(const PlanView& P, int scen, const double* x, int sec, int node0, int count, int tid,
                         BlockScratch& sm) {
  if (tid >= count) return;
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  const int ua = si[GS_UA], xa = si[GS_XA], flags = si[GS_FLAGS];
  const int j = node0 + tid, row = xa + 1 + j;
  const Units un = scen_units(P, scen);
  const SecParam sp = sec_param(P, scen, sec);
  const double to = x[P.off_t + sec], tf = x[P.off_t + sec + 1];
  const double m = x[row];
  const Vec3 p = v3(x[P.off_pos + 3 * row], x[P.off_pos + 3 * row + 1], x[P.off_pos + 3 * row + 2]);
  const Vec3 v = v3(x[P.off_vel + 3 * row], x[P.off_vel + 3 * row + 1], x[P.off_vel + 3 * row + 2]);
  const Quat q = q4(x[P.off_quat + 4 * row], x[P.off_quat + 4 * row + 1], x[P.off_quat + 4 * row + 2],
                    x[P.off_quat + 4 * row + 3]);
  Vec3 f;
  if (flags & GSF_AIR) {
    const double tn = time_node(P.tau_pool + si[GS_TAU_OFF], j + 1, to, tf);
    f = rhs_velocity_air(m, p, v, q, tn, sp, un, scen_tables(P, scen));
  } else {
    f = rhs_velocity_noair(m, p, q, sp, un);
  }
  sm.f[tid][0] = f.x;
  sm.f[tid][1] = f.y;
  sm.f[tid][2] = f.z;
  if (!(flags & GSF_HOLD)) {
    Quat d = rhs_quaternion(q, x[P.off_u + 2 * (ua + j)], x[P.off_u + 2 * (ua + j) + 1], un.u);
    sm.q[tid][0] = d.w;
    sm.q[tid][1] = d.x;
    sm.q[tid][2] = d.y;
    sm.q[tid][3] = d.z;
  }
}

/* D.X for one (row of D, state column): acc = fma(D[j][m], X[xa+m], acc), m ascending */
P_HD double dx_dot(const double* Drow, const double* xcol, int stride, int n1) {
  double acc = 0.0;
  for (int m = 0; m < n1; m++) acc = gm_fma(Drow[m], xcol[(long long)m * stride], acc);
  return acc;
}

P_HD void dyn_res_phase2(const PlanView& P, int scen, const double* x, double* g, int sec, int node0, int count,
                         int tid, int nthreads, const BlockScratch& sm) {
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  const int n = si[GS_N], xa = si[GS_XA], flags = si[GS_FLAGS];
  const Units un = scen_units(P, scen);
  const double ut = un.t;
  const double to = x[P.off_t + sec], tf = x[P.off_t + sec + 1];
  const double dt = tf - to;
  const double* D = P.d_pool + si[GS_D_OFF];
  for (int item = tid; item < count * 11; item += nthreads) {
    const int nl = item / 11, col = item - nl * 11;
    const int j = node0 + nl, row = xa + 1 + j;
    const double* Drow = D + (long long)j * (n + 1);
    if (col == 0) { /* mass: con_dynamics.py:53-61 */
      double r;
      if (flags & GSF_ENGINE_ON) {
        const double lh = dx_dot(Drow, x + xa, 1, n + 1);
        const double rh = -sec_param(P, scen, sec).massflow / un.mass * dt * ut / 2.0;
        r = lh - rh;
      } else {
        r = x[row] - x[xa];
      }
      g[si[GS_R_MASS] + j] = r;
    } else if (col <= 3) { /* position: :146-150 */
      const int k = col - 1;
      const double lh = dx_dot(Drow, x + P.off_pos + 3 * xa + k, 3, n + 1);
      const double rh = x[P.off_vel + 3 * row + k] * un.vel * dt * ut / 2.0 / un.pos;
      g[si[GS_R_POS] + 3 * j + k] = lh - rh;
    } else if (col <= 6) { /* velocity: :256-287 */
      const int k = col - 4;
      const double lh = dx_dot(Drow, x + P.off_vel + 3 * xa + k, 3, n + 1);
      const double rh = sm.f[nl][k] * dt * ut / 2.0;
      g[si[GS_R_VEL] + 3 * j + k] = lh - rh;
    } else { /* quaternion: :520-531 */
      const int k = col - 7;
      double r;
      if (flags & GSF_HOLD) {
        r = x[P.off_quat + 4 * row + k] - x[P.off_quat + 4 * xa + k];
      } else {
        const double lh = dx_dot(Drow, x + P.off_quat + 4 * xa + k, 4, n + 1);
        const double rh = sm.q[nl][k] * dt * ut / 2.0;
        r = lh - rh;
      }
      g[si[GS_R_QUAT] + 4 * j + k] = r;
    }
  }
}

/* ========================================================================= */
/* Aero inequality jobs (con_aero.py:48-248, 311-756)                         */
/*   Jacobian: 16 lanes per constraint row                                    */
/*   lane 0 centre | 1-3 position | 4-6 velocity | 7-10 quaternion | 11 to | 12 tf */
/* ========================================================================= */
P_HD double aero_value(const PlanView& P, int scen, int kind, const double* v /*pos3 vel3 quat4*/, double t_e,
                       double limit) {
  const Units un = scen_units(P, scen);
  Vec3 pos = v3(v[0] * un.pos, v[1] * un.pos, v[2] * un.pos);
  Vec3 vel = v3(v[3] * un.vel, v[4] * un.vel, v[5] * un.vel);
  double t = t_e * un.t;
  return aero_quantity(kind, pos, vel, q4(v[6], v[7], v[8], v[9]), t, scen_tables(P, scen)) / limit;
}

P_HD void aero_load(const PlanView& P, const double* x, int row, double* v) {
  for (int k = 0; k < 3; k++) v[k] = x[P.off_pos + 3 * row + k];
  for (int k = 0; k < 3; k++) v[3 + k] = x[P.off_vel + 3 * row + k];
  for (int k = 0; k < 4; k++) v[6 + k] = x[P.off_quat + 4 * row + k];
}

P_HD void aero_jac_phase1(const PlanView& P, int scen, const double* x, int job, int r0, int count, int tid,
                          BlockScratch& sm) {
  const int rl = tid >> 4, lane = tid & 15;
  if (rl >= count) return;
  const int32_t* ai = P.aero_i32 + job * GA_I32_COLS;
  const int kind = ai[GA_KIND], sec = ai[GA_SECTION];
  const bool has_quat = kind != 1;
  if (lane > 12 || (!has_quat && lane >= 7 && lane <= 10)) return;
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  const int r = r0 + rl, row = si[GS_XA] + r;
  const double dx = P.un.dx;
  double v[10];
  aero_load(P, x, row, v);
  double to = x[P.off_t + sec], tf = x[P.off_t + sec + 1];
  if (P.rc_aero) { /* residue left by the groups that ran before */
    for (int k = 0; k < 3; k++) v[k] = residue_n(v[k], dx, P.rc_aero[P.off_pos + 3 * row + k]);
    for (int k = 0; k < 3; k++) v[3 + k] = residue_n(v[3 + k], dx, P.rc_aero[P.off_vel + 3 * row + k]);
    for (int k = 0; k < 4; k++) v[6 + k] = residue_n(v[6 + k], dx, P.rc_aero[P.off_quat + 4 * row + k]);
    to = residue_n(to, dx, P.rc_aero[P.off_t + sec]);
    tf = residue_n(tf, dx, P.rc_aero[P.off_t + sec + 1]);
  }
  if (lane != 0) { /* gradient works on a copy: columns leave residue inside the copy only */
    const int pidx = (lane <= 10) ? lane - 1 : 10;
    for (int w = 0; w < 10; w++) {
      if (!has_quat && w >= 6) continue;
      if (w < pidx) v[w] = residue(v[w], dx);
      else if (w == pidx) v[w] = v[w] + dx;
    }
  }
  if (lane == 11) to = to + dx;
  if (lane == 12) tf = tf + dx;
  const double tn = time_node(P.tau_pool + si[GS_TAU_OFF], r, to, tf);
  sm.f[tid][0] = aero_value(P, scen, kind, v, tn, P.aero_f64[job * GA_F64_COLS + GA_LIMIT]);
}

P_HD void aero_jac_phase2(const PlanView& P, int scen, double* vals, int job, int r0, int count, int tid,
                          const BlockScratch& sm) {
  (void)scen;
  const int rl = tid >> 4, lane = tid & 15;
  if (rl >= count || lane == 0 || lane > 12) return;
  const int32_t* ai = P.aero_i32 + job * GA_I32_COLS;
  const int64_t* aj = P.aero_i64 + job * GA_I64_COLS;
  const int kind = ai[GA_KIND], nk = ai[GA_NK];
  if (kind == 1 && lane >= 7 && lane <= 10) return;
  const int r = r0 + rl;
  const double dx = P.un.dx;
  const double gval = -((sm.f[tid][0] - sm.f[tid & ~15][0]) / dx); /* -dfdx (con_aero.py:439-461) */
  if (lane <= 3) vals[aj[GA_J_POS] + (long long)(lane - 1) * nk + r] = gval;
  else if (lane <= 6) vals[aj[GA_J_VEL] + (long long)(lane - 4) * nk + r] = gval;
  else if (lane <= 10) vals[aj[GA_J_QUAT] + (long long)(lane - 7) * nk + r] = gval;
  else if (lane == 11) vals[aj[GA_J_T] + r] = gval;
  else vals[aj[GA_J_T] + nk + r] = gval;
}

/* residual kernel: one thread per aero row, pristine inputs: 1 - f (con_aero.py:89-248) */
P_HD void aero_res(const PlanView& P, int scen, const double* x, double* g, int job, int r) {
  const int32_t* ai = P.aero_i32 + job * GA_I32_COLS;
  const int sec = ai[GA_SECTION];
  const int32_t* si = P.sec_i32 + sec * GS_I32_COLS;
  double v[10];
  aero_load(P, x, si[GS_XA] + r, v);
  const double tn = time_node(P.tau_pool + si[GS_TAU_OFF], r, x[P.off_t + sec], x[P.off_t + sec + 1]);
  g[ai[GA_ROW0] + r] = 1.0 - aero_value(P, scen, ai[GA_KIND], v, tn, P.aero_f64[job * GA_F64_COLS + GA_LIMIT]);
}

/* ========================================================================= */
/* Event-point jobs: waypoint LLH / IIP / antenna rows, terminal orbit,       */
/* perigee user constraint.  16 lanes per job in the Jacobian kernel.         */
/* ========================================================================= */
struct EvtOut {
  double v[3];
};

/* evaluate the leaf of an event job: pe/ve non-dimensional state row, t_e non-dimensional */
P_HD EvtOut evt_leaf(const PlanView& P, int scen, int type, const double* ef, const double* pe, const double* ve,
                     double t_e) {
  const Units un = scen_units(P, scen);
  EvtOut o;
  o.v[0] = o.v[1] = o.v[2] = 0.0;
  Vec3 pos = v3(pe[0] * un.pos, pe[1] * un.pos, pe[2] * un.pos);
  Vec3 vel = v3(ve[0] * un.vel, ve[1] * un.vel, ve[2] * un.vel);
  if (type == GE_LLH) { /* con_waypoint.py:565-566 */
    Vec3 llh = eci2geodetic_deg(pos, t_e * un.t);
    o.v[0] = llh.x; o.v[1] = llh.y; o.v[2] = llh.z;
  } else if (type == GE_IIP) { /* :214-218 */
    Vec3 llh = iip_from_eci_deg(pos, vel, t_e * un.t);
    o.v[0] = llh.x; o.v[1] = llh.y; o.v[2] = llh.z;
  } else if (type == GE_ANT) { /* :45-51 */
    o.v[0] = sin_elevation(pos, t_e * un.t, v3(ef[GE_A0], ef[GE_A1], ef[GE_A2]));
  } else if (type == GE_TERM) { /* con_init_terminal_knot.py:362-370 */
    o.v[0] = (orbit_energy(pos, vel) / ef[GE_A0]) - 1.0;
    o.v[1] = (angular_momentum(pos, vel) / ef[GE_A1]) - 1.0;
    o.v[2] = inclination_rad(pos, vel) - ef[GE_A2];
  } else { /* GE_USER_PERIGEE: example/user_constraints.py:133-137 */
    double a, e;
    orbital_a_e(pos, vel, &a, &e);
    o.v[0] = (a * (1.0 - e) / 6378137.0) - 1.0;
  }
  return o;
}

P_HD double evt_form_value(int form, double v, double ref, double den) {
  switch (form) {
    case GEF_DIFF_OVER_DEN: return (v - ref) / den;
    case GEF_NEG_DIFF_OVER_DEN: return -(v - ref) / den;
    case GEF_RATIO_M1: return (v / ref) - 1.0;
    case GEF_NEG_RATIO_P1: return -(v / ref) + 1.0;
    case GEF_REF_MINUS_OVER_DEN: return (ref - v) / den;
    default: return v - ref;
  }
}
P_HD double evt_form_grad(int form, double gfd, double ref, double den) {
  switch (form) {
    case GEF_DIFF_OVER_DEN: return gfd / den;
    case GEF_NEG_DIFF_OVER_DEN: return -gfd / den;
    case GEF_RATIO_M1: return gfd / ref;
    case GEF_NEG_RATIO_P1: return -gfd / ref;
    case GEF_REF_MINUS_OVER_DEN: return -gfd / den;
    default: return gfd;
  }
}

/* residual kernel: one thread per event job, pristine inputs */
P_HD void evt_res(const PlanView& P, int scen, const double* x, double* g, int job) {
  const int32_t* ei = P.evt_i32 + job * GE_I32_COLS;
  const double* ef = P.evt_f64 + job * GE_F64_COLS;
  const int type = ei[GE_TYPE], srow = ei[GE_SROW];
  const double t_e = (ei[GE_TIDX] >= 0) ? x[P.off_t + ei[GE_TIDX]] : 0.0;
  EvtOut o = evt_leaf(P, scen, type, ef, x + P.off_pos + 3 * srow, x + P.off_vel + 3 * srow, t_e);
  if (type == GE_TERM) {
    for (int r = 0; r < ei[GE_NROW]; r++) g[ei[GE_ROW] + r] = o.v[r];
  } else if (type == GE_USER_PERIGEE) {
    g[ei[GE_ROW]] = o.v[0];
  } else {
    g[ei[GE_ROW]] = evt_form_value(ei[GE_FORM], o.v[ei[GE_COMP]], ef[GE_REF], ef[GE_DEN]);
  }
}

/* Jacobian kernel lane maps (perturbation order of the reference):
 *   LLH / ANT : 0 centre | 1-3 pos j | 4 t                       (con_waypoint.py:54-66, 570-578)
 *   IIP       : 0 centre | 1 p0 2 v0 3 p1 4 v1 5 p2 6 v2 | 7 t   (:225-236, interleaved)
 *   TERM      : 0 centre | 1-3 pos j | 4-6 vel j                  (con_init_terminal_knot.py:391-399)
 *   PERIGEE   : 0 base | 1-6 FD of p0 p1 p2 v0 v1 v2 | 7-12 background after k restores
 *               (jac_fd.py:54-60 restricted to the six variables the function reads) */
P_HD int evt_n_lanes(int type) {
  return type == GE_IIP ? 8 : (type == GE_TERM ? 7 : (type == GE_USER_PERIGEE ? 13 : 5));
}

P_HD void evt_jac_phase1(const PlanView& P, int scen, const double* x, int job, int tid, BlockScratch& sm) {
  const int lane = tid & 15;
  const int32_t* ei = P.evt_i32 + job * GE_I32_COLS;
  const double* ef = P.evt_f64 + job * GE_F64_COLS;
  const int type = ei[GE_TYPE], srow = ei[GE_SROW];
  if (lane >= evt_n_lanes(type)) return;
  const double dx = P.un.dx;
  double s[6]; /* pos[3], vel[3] base state = r^rc(x) */
  for (int k = 0; k < 3; k++) s[k] = residue_n(x[P.off_pos + 3 * srow + k], dx, ei[GE_RC0 + k]);
  for (int k = 0; k < 3; k++) s[3 + k] = residue_n(x[P.off_vel + 3 * srow + k], dx, ei[GE_RC0 + 3 + k]);
  double t_e = (ei[GE_TIDX] >= 0) ? residue_n(x[P.off_t + ei[GE_TIDX]], dx, ei[GE_RC0 + 6]) : 0.0;
  /* order[] = sequence in which the reference perturbs the six state variables */
  int order[6], nvar;
  if (type == GE_IIP) {
    order[0] = 0; order[1] = 3; order[2] = 1; order[3] = 4; order[4] = 2; order[5] = 5;
    nvar = 6;
  } else if (type == GE_TERM || type == GE_USER_PERIGEE) {
    for (int k = 0; k < 6; k++) order[k] = k;
    nvar = 6;
  } else {
    for (int k = 0; k < 3; k++) order[k] = k;
    nvar = 3;
  }
  if (lane >= 1) {
    int n_restored, perturbed = -1;
    if (lane <= nvar) { /* FD lane for variable order[lane-1] */
      n_restored = lane - 1;
      perturbed = order[lane - 1];
    } else if (type == GE_USER_PERIGEE) { /* background lanes 7..12: k = lane-6 restores done */
      n_restored = lane - 6;
    } else { /* the t lane: every variable already restored */
      n_restored = nvar;
      t_e = t_e + dx;
    }
    for (int k = 0; k < n_restored; k++) s[order[k]] = residue(s[order[k]], dx);
    if (perturbed >= 0) s[perturbed] = s[perturbed] + dx;
  }
  EvtOut o = evt_leaf(P, scen, type, ef, s, s + 3, t_e);
  sm.f[tid][0] = o.v[0];
  sm.f[tid][1] = o.v[1];
  sm.f[tid][2] = o.v[2];
}

P_HD void evt_jac_phase2(const PlanView& P, double* vals, int job, int tid, const BlockScratch& sm) {
  const int lane = tid & 15;
  const int32_t* ei = P.evt_i32 + job * GE_I32_COLS;
  const int64_t* ej = P.evt_i64 + job * GE_I64_COLS;
  const double* ef = P.evt_f64 + job * GE_F64_COLS;
  const int type = ei[GE_TYPE];
  if (lane == 0 || lane >= evt_n_lanes(type)) return;
  const double dx = P.un.dx;
  const int c = tid & ~15;
  if (type == GE_TERM) { /* (f_p - f_c)/dx, COO order: column-major over the nRow rows */
    const int nrow = ei[GE_NROW];
    const int64_t base = (lane <= 3) ? ej[GE_J_POS] + (long long)(lane - 1) * nrow
                                     : ej[GE_J_VEL] + (long long)(lane - 4) * nrow;
    for (int r = 0; r < nrow; r++) vals[base + r] = (sm.f[tid][r] - sm.f[c][r]) / dx;
    return;
  }
  if (type == GE_USER_PERIGEE) { /* aux tail: fd[6] then background[6] */
    vals[ej[GE_J_POS] + (lane - 1)] = (sm.f[tid][0] - sm.f[c][0]) / dx;
    return;
  }
  const int comp = ei[GE_COMP], form = ei[GE_FORM];
  const double gfd = (sm.f[tid][comp] - sm.f[c][comp]) / dx;
  const double val = evt_form_grad(form, gfd, ef[GE_REF], ef[GE_DEN]);
  if (type == GE_IIP) {
    if (lane == 7) vals[ej[GE_J_T]] = val;
    else if (lane & 1) vals[ej[GE_J_POS] + (lane - 1) / 2] = val;
    else vals[ej[GE_J_VEL] + (lane - 2) / 2] = val;
  } else {
    if (lane == 4) vals[ej[GE_J_T]] = val;
    else vals[ej[GE_J_POS] + (lane - 1)] = val;
  }
}

/* ========================================================================= */
/* Linear rows (init / time / knot / rate / stage mass / kick / time order)   */
/* ========================================================================= */
P_HD void lin_res(const PlanView& P, int scen, const double* x, double* g, int k) {
  const int32_t* li = P.lin_i32 + k * GL_I32_COLS;
  const double* lf = P.lin_f64 + k * GL_F64_COLS;
  const double a = (li[GL_IDX_PLUS] >= 0) ? x[li[GL_IDX_PLUS]] * lf[GL_SCALE_PLUS] : 0.0;
  const double b = (li[GL_IDX_MINUS] >= 0) ? x[li[GL_IDX_MINUS]] * lf[GL_SCALE_MINUS] : 0.0;
  const double c = P.lin_const_scen ? P.lin_const_scen[(long long)scen * P.n_lin + k] : lf[GL_CONST];
  g[li[GL_ROW]] = (a - b) + c;
}

/* ========================================================================= */
/* Block dispatch: one function per kernel and phase                          */
/* ========================================================================= */
P_HD void jac_block_phase1(const PlanView& P, int scen, const int32_t* bt, const double* x, int tid,
                           BlockScratch& sm) {
  switch (bt[BT_ROLE]) {
    case BR_DYN: dyn_jac_phase1(P, scen, x, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, sm); break;
    case BR_AERO: aero_jac_phase1(P, scen, x, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, sm); break;
    case BR_EVT:
      if ((tid >> 4) < bt[BT_COUNT]) evt_jac_phase1(P, scen, x, bt[BT_START] + (tid >> 4), tid, sm);
      break;
    default: break;
  }
}
P_HD void jac_block_phase2(const PlanView& P, int scen, const int32_t* bt, const double* x, double* vals, int tid,
                           const BlockScratch& sm) {
  switch (bt[BT_ROLE]) {
    case BR_DYN: dyn_jac_phase2(P, scen, x, vals, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, sm); break;
    case BR_AERO: aero_jac_phase2(P, scen, vals, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, sm); break;
    case BR_EVT:
      if ((tid >> 4) < bt[BT_COUNT]) evt_jac_phase2(P, vals, bt[BT_START] + (tid >> 4), tid, sm);
      break;
    default: break;
  }
}
P_HD void res_block_phase1(const PlanView& P, int scen, const int32_t* bt, const double* x, double* g, int tid,
                           BlockScratch& sm) {
  switch (bt[BT_ROLE]) {
    case BR_DYN: dyn_res_phase1(P, scen, x, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, sm); break;
    case BR_AERO:
      if (tid < bt[BT_COUNT]) aero_res(P, scen, x, g, bt[BT_JOB], bt[BT_START] + tid);
      break;
    case BR_EVT:
      if (tid < bt[BT_COUNT]) evt_res(P, scen, x, g, bt[BT_START] + tid);
      break;
    case BR_LIN:
      if (tid < bt[BT_COUNT]) lin_res(P, scen, x, g, bt[BT_START] + tid);
      if (bt[BT_START] == 0 && tid == 0) /* objective: cost_gradient.py:29-34 */
        g[0] = P.payload_mode ? -x[0] : x[P.off_t + P.S];
      break;
    default: break;
  }
}
P_HD void res_block_phase2(const PlanView& P, int scen, const int32_t* bt, const double* x, double* g, int tid,
                           int nthreads, const BlockScratch& sm) {
  if (bt[BT_ROLE] == BR_DYN) dyn_res_phase2(P, scen, x, g, bt[BT_JOB], bt[BT_START], bt[BT_COUNT], tid, nthreads, sm);
}

#endif /* GELATO_B200_JOBS_H_ */
