/* plan_host.h -- host-side helpers shared by the CUDA library and the test
 * emulator: PlanView scalar fields from a GelatoPlanDesc, the node / row lists
 * and the block tables of the two kernels. */
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

/* Bounds of every index table of a plan description; returns NULL or what is wrong.  The kernels trust
 * these tables, so a description that fails here is refused before anything is uploaded. */
static inline const char* validate_desc(const GelatoPlanDesc* d) {
  if (!d) return "null description";
  const long long S = d->n_sections, N = d->n_nodes, M = N + S;
  if (S <= 0 || N <= 0) return "empty problem";
  if (!(d->dx > 0.0)) return "dx must be positive";
  if (d->n_rows <= 0 || d->n_vals < 0) return "bad n_rows / n_vals";
  if (!d->sec_i32 || !d->sec_i64 || !d->sec_f64 || !d->d_pool || !d->tau_pool) return "missing section tables";
  if (!d->vals_template && d->n_vals > 0) return "missing vals_template";
  if (d->n_wind < 2 || !d->wind || d->n_ca < 2 || !d->ca) return "wind / CA tables need at least two rows";
  const long long n_vars = 11 * M + 2 * N + S + 1;
  long long nodes = 0;
  for (long long s = 0; s < S; s++) {
    const int32_t* si = d->sec_i32 + s * GS_I32_COLS;
    const long long n = si[GS_N];
    if (n < 1) return "a section has no nodes";
    if (si[GS_UA] != nodes) return "sections must tile the control rows in order";
    if (si[GS_XA] < 0 || si[GS_XA] + n + 1 > M) return "state rows of a section out of range";
    if (si[GS_D_OFF] < 0 || si[GS_D_OFF] + n * (n + 1) > d->d_pool_len) return "D block outside the pool";
    if (si[GS_TAU_OFF] < 0 || si[GS_TAU_OFF] + n > d->tau_pool_len) return "tau outside the pool";
    for (int c = GS_R_MASS; c <= GS_R_QUAT; c++)
      if (si[c] < 1 || si[c] >= d->n_rows) return "residual row offset of a section out of range";
    nodes += n;
  }
  if (nodes != N) return "section node counts do not add up to n_nodes";
  if (d->n_lin < 0 || d->n_aero < 0 || d->n_evt < 0) return "negative table length";
  for (long long k = 0; k < d->n_lin; k++) {
    const int32_t* li = d->lin_i32 + k * GL_I32_COLS;
    if (li[GL_ROW] < 1 || li[GL_ROW] >= d->n_rows) return "linear row out of range";
    if (li[GL_IDX_PLUS] < -1 || li[GL_IDX_PLUS] >= n_vars || li[GL_IDX_MINUS] < -1 || li[GL_IDX_MINUS] >= n_vars)
      return "linear row reads outside x";
  }
  for (long long k = 0; k < d->n_aero; k++) {
    const int32_t* ai = d->aero_i32 + k * GA_I32_COLS;
    if (ai[GA_KIND] < 0 || ai[GA_KIND] > 2) return "unknown aero kind";
    if (ai[GA_SECTION] < 0 || ai[GA_SECTION] >= S) return "aero job section out of range";
    const long long n = d->sec_i32[ai[GA_SECTION] * GS_I32_COLS + GS_N];
    if (ai[GA_NK] < 1 || ai[GA_NK] > n + 1) return "aero job has more rows than its section has state nodes";
    if (ai[GA_ROW0] < 1 || ai[GA_ROW0] + ai[GA_NK] > d->n_rows) return "aero rows out of range";
  }
  for (long long k = 0; k < d->n_evt; k++) {
    const int32_t* ei = d->evt_i32 + k * GE_I32_COLS;
    if (ei[GE_TYPE] < GE_LLH || ei[GE_TYPE] > GE_USER_PERIGEE) return "unknown event job type";
    if (ei[GE_SROW] < 0 || ei[GE_SROW] >= M) return "event job state row out of range";
    if (ei[GE_TIDX] < -1 || ei[GE_TIDX] > S) return "event job time index out of range";
    if (ei[GE_ROW] < 1 || ei[GE_ROW] + ei[GE_NROW] > d->n_rows || ei[GE_NROW] < 1 || ei[GE_NROW] > 3)
      return "event job rows out of range";
    if (ei[GE_COMP] < 0 || ei[GE_COMP] > 2) return "event job component out of range";
  }
  return nullptr;
}

struct HostTables {
  std::vector<int32_t> jac_blocks, res_blocks; /* BT_COLS ints per block */
  std::vector<NodeRec> node_rec;               /* [N] natural order */
  std::vector<NodeRec> jac_rec;                /* [N] air-FD nodes, then vacuum nodes, then fallback nodes */
  std::vector<AeroRec> aero_rows;              /* one per aero constraint row */
};

static inline void push_block(std::vector<int32_t>& t, int role, int job, int start, int count) {
  t.push_back(role);
  t.push_back(job);
  t.push_back(start);
  t.push_back(count);
}
static inline void push_chunks(std::vector<int32_t>& t, int role, int first, int total, int per_block) {
  for (int a = 0; a < total; a += per_block) push_block(t, role, 0, first + a, std::min(per_block, total - a));
}

static inline void build_host_tables(const GelatoPlanDesc* d, HostTables& h) {
  const int S = d->n_sections, N = d->n_nodes;
  h.node_rec.assign(N, NodeRec());
  std::vector<int32_t> air, vac, gen;
  for (int s = 0; s < S; s++) {
    const int32_t* si = d->sec_i32 + s * GS_I32_COLS;
    const int flags = si[GS_FLAGS];
    for (int j = 0; j < si[GS_N]; j++) {
      const int g = si[GS_UA] + j;
      NodeRec& q = h.node_rec[g];
      q.sec = s; q.j = j; q.row = si[GS_XA] + 1 + j; q.ua = si[GS_UA];
      q.n = si[GS_N]; q.flags = flags; q.d_off = si[GS_D_OFF]; q.tau_off = si[GS_TAU_OFF];
      if (flags & GSF_AIR_FD) air.push_back(g);
      else if (!(flags & GSF_AIR)) vac.push_back(g);
      else gen.push_back(g);
    }
  }
  for (const std::vector<int32_t>* list : {&air, &vac, &gen})
    for (int32_t g : *list) h.jac_rec.push_back(h.node_rec[g]);
  for (int k = 0; k < d->n_aero; k++) {
    const int32_t* ai = d->aero_i32 + k * GA_I32_COLS;
    const int32_t* si = d->sec_i32 + ai[GA_SECTION] * GS_I32_COLS;
    for (int r = 0; r < ai[GA_NK]; r++) {
      AeroRec q;
      q.job = k; q.r = r; q.sec = ai[GA_SECTION]; q.row = si[GS_XA] + r;
      q.kind = ai[GA_KIND]; q.nk = ai[GA_NK]; q.tau_off = si[GS_TAU_OFF]; q.row0 = ai[GA_ROW0];
      h.aero_rows.push_back(q);
    }
  }
  const int n_aero_rows = (int)h.aero_rows.size();

  std::vector<int32_t>& jb = h.jac_blocks;
  push_chunks(jb, BR_DYN_AIR, 0, (int)air.size(), GD_NODES);
  push_chunks(jb, BR_DYN_NOAIR, (int)air.size(), (int)vac.size(), GN_NODES);
  push_chunks(jb, BR_DYN_GEN, (int)(air.size() + vac.size()), (int)gen.size(), GG_NODES);
  push_chunks(jb, BR_AERO, 0, n_aero_rows, GJ_NODES);
  push_chunks(jb, BR_EVT, 0, d->n_evt, GJ_EVT);

  std::vector<int32_t>& rb = h.res_blocks;
  push_chunks(rb, BR_DYN, 0, N, GR_NODES);
  push_chunks(rb, BR_AERO, 0, n_aero_rows, GR_THREADS);
  push_chunks(rb, BR_EVT, 0, d->n_evt, GR_THREADS);
  push_chunks(rb, BR_LIN, 0, d->n_lin, GR_THREADS);
}

#endif
