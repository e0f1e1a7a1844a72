/* plan_host.h -- host-side helpers shared by the CUDA library and the test
 * emulator: PlanView scalar fields from a GelatoPlanDesc and the block tables. */
#ifndef GELATO_B200_PLAN_HOST_H_
#define GELATO_B200_PLAN_HOST_H_

#include <algorithm>
#include <vector>

#include "jobs.h"

/* fills every non-pointer field of the view; pointer fields are left untouched */
static inline void planview_scalars(const GelatoPlanDesc* d, PlanView& v) {
  const int S = d->n_sections, N = d->n_nodes, M = N + S;
  v.S = S; v.N = N; v.M = M;
  v.n_vars = M + 3 * M + 3 * M + 4 * M + 2 * N + S + 1;
  v.n_rows = d->n_rows;
  v.n_vals = d->n_vals;
  v.payload_mode = d->payload_mode;
  v.off_pos = M; v.off_vel = 4 * M; v.off_quat = 7 * M; v.off_u = 11 * M; v.off_t = 11 * M + 2 * N;
  v.un.mass = d->unit_mass; v.un.pos = d->unit_pos; v.un.vel = d->unit_vel; v.un.u = d->unit_u;
  v.un.t = d->unit_t; v.un.dx = d->dx;
  v.n_wind = d->n_wind; v.n_ca = d->n_ca;
  v.n_lin = d->n_lin; v.n_aero = d->n_aero; v.n_evt = d->n_evt;
}

static inline void push_block(std::vector<int32_t>& t, int role, int job, int start, int count) {
  t.push_back(role);
  t.push_back(job);
  t.push_back(start);
  t.push_back(count);
}

/* jb: Jacobian kernel blocks, rb: residual kernel blocks (BT_COLS ints each) */
static inline void build_block_tables(const GelatoPlanDesc* d, std::vector<int32_t>& jb, std::vector<int32_t>& rb) {
  for (int s = 0; s < d->n_sections; s++) {
    const int n = d->sec_i32[s * GS_I32_COLS + GS_N];
    for (int a = 0; a < n; a += GB_DYN_JAC_NODES) push_block(jb, BR_DYN, s, a, std::min(GB_DYN_JAC_NODES, n - a));
    for (int a = 0; a < n; a += GB_DYN_RES_NODES) push_block(rb, BR_DYN, s, a, std::min(GB_DYN_RES_NODES, n - a));
  }
  for (int k = 0; k < d->n_aero; k++) {
    const int nk = d->aero_i32[k * GA_I32_COLS + GA_NK];
    for (int a = 0; a < nk; a += GB_ROWS16) push_block(jb, BR_AERO, k, a, std::min(GB_ROWS16, nk - a));
    for (int a = 0; a < nk; a += GB_THREADS) push_block(rb, BR_AERO, k, a, std::min(GB_THREADS, nk - a));
  }
  for (int a = 0; a < d->n_evt; a += GB_ROWS16) push_block(jb, BR_EVT, 0, a, std::min(GB_ROWS16, d->n_evt - a));
  for (int a = 0; a < d->n_evt; a += GB_THREADS) push_block(rb, BR_EVT, 0, a, std::min(GB_THREADS, d->n_evt - a));
  for (int a = 0; a < d->n_lin; a += GB_THREADS) push_block(rb, BR_LIN, 0, a, std::min(GB_THREADS, d->n_lin - a));
}

#endif
