// emu.cpp -- host emulator of the two fused CUDA kernels (TEST HARNESS ONLY).
//
// Compiles gelato_b200/csrc/jobs.h (the per-thread job functions the kernels
// run) with g++ and steps through every block and thread serially: phase 1 for
// all threads of a block, then phase 2 -- the same order the block barrier
// enforces on the GPU (the Jacobian kernel has GJ_PHASES phases, the residual kernel two).  It lets the CPU-only test tier check the plan compiler
// and the kernels' index arithmetic against the oracle bit for bit.  It is not
// part of the product: libgelato_b200.so has no CPU path and bench.py never
// loads this file.
#include <cstring>
#include <vector>

#include "../../gelato_b200/csrc/coord_leaves.h"
#include "../../gelato_b200/csrc/host_pool.h"
#include "../../gelato_b200/csrc/initguess.h"
#include "../../gelato_b200/csrc/output.h"
#include "../../gelato_b200/csrc/plan_host.h"

static PlanView host_view(const GelatoPlanDesc* d) {
  PlanView v;
  memset(&v, 0, sizeof v);
  planview_scalars(d, v);
  v.sec_i32 = d->sec_i32; v.sec_i64 = d->sec_i64; v.sec_f64 = d->sec_f64;
  v.d_pool = d->d_pool; v.tau_pool = d->tau_pool;
  v.wind = d->wind; v.ca = d->ca;
  v.lin_i32 = d->lin_i32; v.lin_f64 = d->lin_f64;
  v.aero_i32 = d->aero_i32; v.aero_i64 = d->aero_i64; v.aero_f64 = d->aero_f64; v.rc_aero = d->rc_aero;
  v.evt_i32 = d->evt_i32; v.evt_i64 = d->evt_i64; v.evt_f64 = d->evt_f64;
  v.packed = 0;
  return v;
}

static void attach_tables(PlanView& v, const HostTables& h) {
  v.node_rec = h.node_rec.data();
  v.jac_rec = h.jac_rec.data();
  v.aero_rows = h.aero_rows.data();
  v.n_aero_rows = (int)h.aero_rows.size();
}

static void apply_scen(PlanView& v, const GelatoScenarioDesc* sc) {
  if (!sc) return;
  if (sc->sec_f64) { v.sec_f64 = sc->sec_f64; v.sec_f64_sstride = (long long)v.S * GS_F64_COLS; }
  if (sc->wind) { v.wind = sc->wind; v.wind_sstride = (long long)v.n_wind * 3; }
  if (sc->unit_mass) v.unit_mass_scen = sc->unit_mass;
  if (sc->lin_const) v.lin_const_scen = sc->lin_const;
}

extern "C" int emu_unfused_check(void) {
  volatile double a = 1.0 + 0x1p-30, b = 1.0 - 0x1p-30, c = -1.0;
  double r = a * b + c;
  return r == 0.0;
}

extern "C" int emu_eval_residuals(const GelatoPlanDesc* d, const GelatoScenarioDesc* sc, const double* x_all,
                                  double* g_all, int n_scen, const int32_t* ids) {
  if (validate_desc(d)) return -1;
  PlanView P = host_view(d);
  apply_scen(P, sc);
  HostTables h;
  build_host_tables(d, h);
  attach_tables(P, h);
  const std::vector<int32_t>& rb = h.res_blocks;
  ResScratch sm;
  for (int scen = 0; scen < n_scen; scen++) {
    const double* x = x_all + (size_t)scen * P.n_vars;
    double* g = g_all + (size_t)scen * P.n_rows;
    for (size_t b = 0; b < rb.size() / BT_COLS; b++) {
      const int32_t* bt = rb.data() + b * BT_COLS;
      memset(&sm, 0xff, sizeof sm);
      const int sid = ids ? ids[scen] : scen;
      for (int tid = 0; tid < GR_THREADS; tid++) res_block_phase0(P, sid, bt, x, tid, sm);
      for (int tid = 0; tid < GR_THREADS; tid++) res_block_phase1(P, sid, bt, x, g, tid, sm);
      for (int tid = 0; tid < GR_THREADS; tid++) res_block_phase2(P, sid, bt, x, g, tid, GR_THREADS, sm);
    }
  }
  return 0;
}

// packed: out_all is [n_scen][n_pack] and nothing is pre-filled;  g_all != NULL: pair evaluation -- objfunc's
// rows come out of the Jacobian blocks (what gelato_eval_pair_* launches).
static int g_range[4] = {0, -1, 0, 0}; /* block first / count (-1: all), vacuum first / count */
static int emu_jacobian(const GelatoPlanDesc* d, const GelatoScenarioDesc* sc, const double* x_all, double* g_all,
                        double* out_all, int n_scen, const int32_t* ids, int packed) {
  if (validate_desc(d)) return -1;
  PlanView P = host_view(d);
  apply_scen(P, sc);
  HostTables h;
  build_host_tables(d, h);
  attach_tables(P, h);
  PackedLayout L;
  build_packed_layout(d, L);
  P.packed = packed;
  P.n_pack = L.n_pack;
  P.sec_pk = L.sec_pk.data(); P.aero_pk = L.aero_pk.data(); P.evt_pk = L.evt_pk.data();
  const std::vector<int32_t>& jb = h.jac_blocks;
  JacStore store;
  const JacScratch sm = jac_scratch(store);
  const size_t stride = packed ? (size_t)L.n_pack : (size_t)P.n_vals;
  for (int scen = 0; scen < n_scen; scen++) {
    const double* x = x_all + (size_t)scen * P.n_vars;
    double* vals = out_all + scen * stride;
    double* g = g_all ? g_all + (size_t)scen * P.n_rows : nullptr;
    const int sid = ids ? ids[scen] : scen;
    if (!packed) {
      const double* tmpl = (sc && sc->vals_template) ? sc->vals_template + (size_t)sid * P.n_vals : d->vals_template;
      memcpy(vals, tmpl, (size_t)P.n_vals * sizeof(double));
    }
    const size_t nb_all = g ? jb.size() / BT_COLS : (size_t)h.n_jac_main;  /* the linear-row blocks: pair evaluations only */
    const size_t b_lo = g_range[1] < 0 ? 0 : (size_t)g_range[0], nb = g_range[1] < 0 ? nb_all : (size_t)(g_range[0] + g_range[1]);
    const int v_lo = g_range[1] < 0 ? 0 : g_range[2], v_hi = g_range[1] < 0 ? h.n_vac : g_range[2] + g_range[3];
    for (size_t b = b_lo; b < nb; b++) {
      const int32_t* bt = jb.data() + b * BT_COLS;
      /* poison the scratch so a phase that reads what no thread wrote shows up as NaN */
      memset(&store, 0xff, sizeof store);
      for (int phase = 0; phase < GJ_PHASES; phase++)
        for (int tid = 0; tid < GJ_THREADS; tid++) jac_block_phase<JR_ALL>(P, sid, bt, x, vals, g, tid, phase, sm);
    }
    /* vacuum dynamics nodes: GV_PARTS threads each (kernel k_jacobian_noair), here part after part */
    for (int k = v_lo; k < v_hi; k++) dyn_noair_node(P, sid, x, vals, g, jac_node(P, h.vac_first + k));
  }
  return 0;
}

extern "C" int emu_eval_jacobian(const GelatoPlanDesc* d, const GelatoScenarioDesc* sc, const double* x_all,
                                 double* vals_all, int n_scen, const int32_t* ids) {
  return emu_jacobian(d, sc, x_all, nullptr, vals_all, n_scen, ids, 0);
}

// one problem sharded over ranks: the pair evaluation restricted to a block range and a vacuum-node range
extern "C" int emu_eval_pair_range(const GelatoPlanDesc* d, const GelatoScenarioDesc* sc, const double* x_all, double* g_all,
                                   double* out_all, int n_scen, int b0, int nb, int v0, int nv) {
  g_range[0] = b0; g_range[1] = nb; g_range[2] = v0; g_range[3] = nv;
  const int rc = emu_jacobian(d, sc, x_all, g_all, out_all, n_scen, nullptr, 1);
  g_range[1] = -1;
  return rc;
}
extern "C" void emu_block_counts(const GelatoPlanDesc* d, int* n_blocks_pair, int* n_vac) {
  HostTables h;
  build_host_tables(d, h);
  *n_blocks_pair = (int)(h.jac_blocks.size() / BT_COLS);
  *n_vac = h.n_vac;
}

extern "C" int emu_eval_pair(const GelatoPlanDesc* d, const GelatoScenarioDesc* sc, const double* x_all, double* g_all,
                             double* out_all, int n_scen, const int32_t* ids, int packed) {
  return emu_jacobian(d, sc, x_all, g_all, out_all, n_scen, ids, packed);
}

// packed layout of a plan: returns n_pack; with non-NULL arrays also the map of the n_xdep COO slots
extern "C" long long emu_packed_map(const GelatoPlanDesc* d, int64_t* full_slot, int64_t* src, double* sgn) {
  if (validate_desc(d)) return -1;
  PackedLayout L;
  build_packed_layout(d, L);
  if (full_slot && src && sgn)
    for (size_t i = 0; i < L.full_slot.size(); i++) {
      full_slot[i] = L.full_slot[i];
      src[i] = L.src[i];
      sgn[i] = L.sgn[i];
    }
  return L.n_pack;
}
extern "C" long long emu_n_xdep(const GelatoPlanDesc* d) {
  PackedLayout L;
  build_packed_layout(d, L);
  return (long long)L.full_slot.size();
}

// the result-table kernel's per-thread function (output.h), one call per node
extern "C" void emu_output_rows(int n, const double* mass, const double* pos, const double* vel, const double* quat,
                                const double* t, const double* thrust_vac, const double* air_area,
                                const double* nozzle_area, const double* wind, int n_wind, const double* ca, int n_ca,
                                double lat0, double lon0, double* out) {
  Tables tb;
  tb.wind = wind; tb.n_wind = n_wind; tb.ca = ca; tb.n_ca = n_ca;
  for (int i = 0; i < n; i++)
    output_row(mass[i], v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), v3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]),
               q4(quat[4 * i], quat[4 * i + 1], quat[4 * i + 2], quat[4 * i + 3]), t[i], thrust_vac[i], air_area[i],
               nozzle_area[i], tb, lat0, lon0, out + (size_t)i * GO_COLS);
}

// update mode's host side (host_pool.h): packed x-dependent slots -> the caller's Jacobian buffers
extern "C" void emu_scatter_parallel(const int64_t* idx, long long n_idx, const double* packed, double* vals,
                                     long long n_vals, int s0, int s1, int threads) {
  gelato_host::scatter_parallel(idx, n_idx, packed, vals, n_vals, s0, s1, threads);
}

extern "C" int emu_pool_workers(void) { return gelato_host::HostPool::instance().workers(); }

// the forward-simulation kernel's per-thread function (initguess.h), one call per scenario
extern "C" void emu_rocket_simulation(int n, const double* x_init, const double* events, const int32_t* zlt, int n_ev,
                                      const double* u_table, int n_u, const double* wind, int n_wind, const double* ca,
                                      int n_ca, const int64_t* ss, double t_init, const double* t_out, int n_out, double dt,
                                      double* x_out, double* u_out) {
  for (int s = 0; s < n; s++) {
    Tables tb;
    tb.wind = wind + s * ss[2]; tb.n_wind = n_wind;
    tb.ca = ca + s * ss[3]; tb.n_ca = n_ca;
    rocket_simulation_thread(x_init + s * ss[0], events + s * ss[1], zlt, n_ev, u_table, n_u, tb, t_init, t_out, n_out, dt,
                             x_out + (size_t)s * n_out * 11, u_out ? u_out + (size_t)s * n_out * 3 : nullptr);
  }
}

// the coordinate leaf kernel's per-thread function (coord_leaves.h), one call per item
extern "C" void emu_coord_leaf(int fn, int n, const double* a, int sa, const double* b, int sb, const double* t, double* out) {
  const int so = coord_leaf_n_out(fn);
  for (int i = 0; i < n; i++) {
    double av[GC_IN] = {0.0, 0.0, 0.0, 0.0}, bv[GC_IN] = {0.0, 0.0, 0.0, 0.0}, ov[GC_OUT];
    for (int k = 0; k < sa; k++) av[k] = a[(size_t)i * sa + k];
    for (int k = 0; k < sb; k++) bv[k] = b[(size_t)i * sb + k];
    coord_leaf(fn, av, bv, t ? t[i] : 0.0, ov);
    for (int k = 0; k < so; k++) out[(size_t)i * so + k] = ov[k];
  }
}
