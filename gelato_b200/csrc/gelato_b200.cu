// gelato_b200.cu -- CUDA kernels (sm_100a) and the C ABI declared in
// include/gelato_b200.h.
//
// Two fused kernels cover the whole NLP callback path:
//   k_residuals : everything `objfunc` evaluates (Trajectory_Optimization.py:194-242)
//   k_jacobian  : every x-dependent Jacobian value `sens` produces (:245-312)
// Each is ONE launch per call; blocks take their role from a block table built
// at plan creation (dynamics sections in chunks of nodes, aero rows, event
// rows, linear rows), one copy of the block list per scenario of a batched solve.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (see DESIGN.md H1;
// gelato_selftest_unfused() verifies the flag at run time).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "host_pool.h"
#include "plan_host.h"

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// One Jacobian evaluation = ONE launch: every role's blocks, heavy (air dynamics nodes, aero rows: the pos_part /
// rotq_part code, ~1 kFLOP per column) and light (vacuum dynamics, fallback, event rows) interleaved in the block
// table so that FP64-issue-bound and latency-bound blocks share each SM.  (Round 2 first split the roles into two
// kernels on two streams: they ran one after the other -- the first kernel fills every SM -- and the instruction
// cache hit rate did not move; profiles/r02a_ab_probe.txt.)  The role-subset instantiations below exist for
// measurements only (gelato_launch_kernel_dev).
#ifndef GJ_MIN_BLOCKS
#define GJ_MIN_BLOCKS 2 /* 448 threads x 2 blocks: 28 warps per SM, up to 72 registers */
#endif
// The blocks of the first wave ask L2 for the whole decision-vector batch (one 128-byte line per thread, no
// register, no wait): a block reads ~25 scattered lines of x before it can start, and on an L2 that does not hold
// them every later wave would pay the DRAM latency again.  0 switches it off.
#ifndef GJ_PREFETCH_BLOCKS
#define GJ_PREFETCH_BLOCKS 296
#endif
__device__ __forceinline__ void prefetch_batch_l2(const double* x_all, size_t n_doubles, int blocks, int threads) {
#if GJ_PREFETCH_BLOCKS > 0
  if ((int)blockIdx.x < blocks) {
    const size_t lines = (n_doubles * sizeof(double) + 127) / 128;
    for (size_t l = (size_t)blockIdx.x * threads + threadIdx.x; l < lines; l += (size_t)blocks * threads)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(x_all) + l * 128));
  }
#endif
}
// Measurement build only (-DGJ_CLOCKS): cycles every warp spends in each phase and at each barrier, summed per
// (role, warp) over the launch; read back with gelato_debug_clocks (tools/phase_clocks.py).
#ifdef GJ_CLOCKS
__device__ unsigned long long gj_clk_sum[8][16][6];  // [role][warp][phase0, wait0, phase2, wait2, phase3, blocks]
#define GJ_CLK_DECL long long clk_[6] = {0, 0, 0, 0, 0, 0}
#define GJ_CLK(i) clk_[i] = clock64()
#define GJ_CLK_FLUSH(role, two)                                                                         \
  if ((threadIdx.x & 31) == 0) {                                                                        \
    unsigned long long* c = gj_clk_sum[(role) & 7][threadIdx.x >> 5];                                   \
    atomicAdd(c + 0, (unsigned long long)(clk_[1] - clk_[0]));                                              \
    atomicAdd(c + 1, (unsigned long long)(clk_[2] - clk_[1]));                                              \
    atomicAdd(c + 2, (unsigned long long)((two) ? 0 : clk_[3] - clk_[2]));                                  \
    atomicAdd(c + 3, (unsigned long long)((two) ? 0 : clk_[4] - clk_[3]));                                  \
    atomicAdd(c + 4, (unsigned long long)(clk_[5] - clk_[4]));                                              \
    atomicAdd(c + 5, 1ull);                                                                             \
  }
#else
#define GJ_CLK_DECL
#define GJ_CLK(i)
#define GJ_CLK_FLUSH(role, two)
#endif
template <int ROLES>
__device__ __forceinline__ void jacobian_body(const PlanView& P, const int32_t* __restrict__ block_table, const int n_scen,
                                              const int32_t* __restrict__ scen_ids, const double* __restrict__ x_all,
                                              double* __restrict__ out_all, double* __restrict__ g_all) {
  __shared__ JacStore store;
  const JacScratch sm = jac_scratch(store);
  prefetch_batch_l2(x_all, (size_t)n_scen * P.n_vars, GJ_PREFETCH_BLOCKS, GJ_THREADS);
  // block-major launch order: block b of every scenario before block b+1 of any
  const int scen = blockIdx.x % n_scen;
  const int32_t* bt = block_table + (size_t)(blockIdx.x / n_scen) * BT_COLS;
  const double* x = x_all + (size_t)scen * P.n_vars;
  // COO output: vals[n_scen][n_vals] (constants pre-filled); packed output: [n_scen][n_pack]
  double* out = out_all + (size_t)scen * (size_t)(P.packed ? P.n_pack : P.n_vals);
  // pair evaluation: every block also writes objfunc's rows of its nodes / constraint rows / events
  double* g = g_all ? g_all + (size_t)scen * P.n_rows : nullptr;
  // which scenario's parameter blocks this batch slot uses (a coalesced subset of the configured scenarios)
  const int sid = scen_ids ? scen_ids[scen] : scen;
  const bool two_phase = jac_role_two_phase(bt[BT_ROLE]);
  GJ_CLK_DECL;
  GJ_CLK(0);
  jac_block_phase<ROLES>(P, sid, bt, x, out, g, threadIdx.x, 0, sm);
  GJ_CLK(1);
  __syncthreads();
  GJ_CLK(2);
  if (!two_phase) {
    jac_block_phase<ROLES>(P, sid, bt, x, out, g, threadIdx.x, 2, sm);
    GJ_CLK(3);
    __syncthreads();
  }
  GJ_CLK(4);
  jac_block_phase<ROLES>(P, sid, bt, x, out, g, threadIdx.x, 3, sm);
  GJ_CLK(5);
  GJ_CLK_FLUSH(bt[BT_ROLE], two_phase);
}
__global__ void __launch_bounds__(GJ_THREADS, GJ_MIN_BLOCKS)
k_jacobian(const __grid_constant__ PlanView P, const int32_t* __restrict__ block_table, const int n_scen, const int32_t* __restrict__ scen_ids,
           const double* __restrict__ x_all, double* __restrict__ out_all, double* __restrict__ g_all) {
  jacobian_body<JR_ALL>(P, block_table, n_scen, scen_ids, x_all, out_all, g_all);
}
__global__ void __launch_bounds__(GJ_THREADS, GJ_MIN_BLOCKS)
k_jacobian_heavy(const __grid_constant__ PlanView P, const int32_t* __restrict__ block_table, const int n_scen,
                 const int32_t* __restrict__ scen_ids, const double* __restrict__ x_all, double* __restrict__ out_all,
                 double* __restrict__ g_all) {
  jacobian_body<JR_HEAVY>(P, block_table, n_scen, scen_ids, x_all, out_all, g_all);
}
__global__ void __launch_bounds__(GJ_THREADS, GJ_MIN_BLOCKS)
k_jacobian_light(const __grid_constant__ PlanView P, const int32_t* __restrict__ block_table, const int n_scen,
                 const int32_t* __restrict__ scen_ids, const double* __restrict__ x_all, double* __restrict__ out_all,
                 double* __restrict__ g_all) {
  jacobian_body<JR_LIGHT>(P, block_table, n_scen, scen_ids, x_all, out_all, g_all);
}

// Vacuum dynamics nodes: GV_PARTS threads per node, a warp per part (jobs.h: dyn_noair_part); no shared memory, no
// barrier.  A block is GV_PARTS warps = 32 nodes.
#define GV_THREADS (32 * GV_PARTS)
#ifndef GV_MIN_BLOCKS
#define GV_MIN_BLOCKS 5 /* the 158 registers ptxas would take leave 12 warps per SM; measured best of 1, 4, 5, 6, 8 */
#endif
__global__ void __launch_bounds__(GV_THREADS, GV_MIN_BLOCKS)
k_jacobian_noair(const __grid_constant__ PlanView P, const int first, const int count, const int n_scen,
                 const int32_t* __restrict__ scen_ids, const double* __restrict__ x_all, double* __restrict__ out_all,
                 double* __restrict__ g_all) {
  const int per = (count + 31) / 32;  // blocks per scenario
  const int scen = blockIdx.x / per;
  const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int k = (blockIdx.x - scen * per) * 32 + lane;
  __shared__ double grav[32][NPV * 3];  // the node's five gravity vectors, shared by the four parts
  const double* x = x_all + (size_t)scen * P.n_vars;
  double* out = out_all + (size_t)scen * (size_t)(P.packed ? P.n_pack : P.n_vals);
  double* g = g_all ? g_all + (size_t)scen * P.n_rows : nullptr;
  const int sid = scen_ids ? scen_ids[scen] : scen;
  const bool live = k < count;
  NodeRef nr;
  if (live) {
    nr = jac_node(P, first + k);
    for (int pv = part; pv < NPV; pv += GV_PARTS) dyn_noair_gravity(P, sid, x, nr, pv, grav[lane] + 3 * pv);
  }
  __syncthreads();
  if (live) dyn_noair_part(P, sid, x, out, g, nr, part, grav[lane]);
}

#ifndef GR_MIN_BLOCKS
#define GR_MIN_BLOCKS 8 /* 64 registers; measured best of 5, 8, 10 (profiles/r01g_ab.txt) */
#endif
__global__ void __launch_bounds__(GR_THREADS, GR_MIN_BLOCKS)
k_residuals(const __grid_constant__ PlanView P, const int32_t* __restrict__ block_table, const int n_scen,
            const int32_t* __restrict__ scen_ids, const double* __restrict__ x_all, double* __restrict__ g_all) {
  __shared__ ResScratch sm;
  const int scen = blockIdx.x % n_scen;
  const int32_t* bt = block_table + (size_t)(blockIdx.x / n_scen) * BT_COLS;
  const double* x = x_all + (size_t)scen * P.n_vars;
  double* g = g_all + (size_t)scen * P.n_rows;
  const int sid = scen_ids ? scen_ids[scen] : scen;
  const bool dyn = bt[BT_ROLE] == BR_DYN;
  if (dyn) {
    res_block_phase0(P, sid, bt, x, threadIdx.x, sm);
    __syncthreads();
  }
  res_block_phase1(P, sid, bt, x, g, threadIdx.x, sm);
  if (dyn) {
    __syncthreads();
    res_block_phase2(P, sid, bt, x, g, threadIdx.x, GR_THREADS, sm);
  }
}

// packed[scen][i] = vals[scen][idx[i]]: the x-dependent slots, gathered for the PCIe copy of update mode
__global__ void k_pack_xdep(const double* __restrict__ vals_all, const int64_t* __restrict__ idx, long long n_xdep,
                            long long n_vals, double* __restrict__ packed_all) {
  const int scen = blockIdx.y;
  const double* vals = vals_all + (size_t)scen * n_vals;
  double* packed = packed_all + (size_t)scen * n_xdep;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_xdep; i += (long long)gridDim.x * blockDim.x)
    packed[i] = vals[idx[i]];
}

// update mode, zero-copy flavour: the x-dependent slots written straight into the caller's page-locked
// (device-mapped) host buffer over PCIe; runs of consecutive slots coalesce into full-width writes
__global__ void k_scatter_xdep_host(const double* __restrict__ vals_all, const int64_t* __restrict__ idx,
                                    long long n_xdep, long long n_vals, double* __restrict__ host_vals_all) {
  const int scen = blockIdx.y;
  const double* vals = vals_all + (size_t)scen * n_vals;
  double* host = host_vals_all + (size_t)scen * n_vals;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_xdep; i += (long long)gridDim.x * blockDim.x) {
    const long long k = idx[i];
    host[k] = vals[k];
  }
}

// probe: is a*b+c left unfused?  (1 + 2^-30)(1 - 2^-30) - 1 is 0 unfused, -2^-60 fused
__global__ void k_unfused_probe(double a, double b, double c, double* out) { *out = a * b + c; }

// FP64 issue-rate probes
__global__ void k_fp64_fma(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c);
    a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_fp64_muladd(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = __dadd_rn(__dmul_rn(a0, b), c); a1 = __dadd_rn(__dmul_rn(a1, b), c);
    a2 = __dadd_rn(__dmul_rn(a2, b), c); a3 = __dadd_rn(__dmul_rn(a3, b), c);
    a4 = __dadd_rn(__dmul_rn(a4, b), c); a5 = __dadd_rn(__dmul_rn(a5, b), c);
    a6 = __dadd_rn(__dmul_rn(a6, b), c); a7 = __dadd_rn(__dmul_rn(a7, b), c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(GELATO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

struct GelatoPlan {
  int device = 0;
  PlanView view{};  // device pointers
  std::vector<void*> owned;
  int32_t* jac_blocks = nullptr;  // n_jac_main dynamics / aero / event blocks (roles interleaved), then the linear-row blocks
  int n_jac_blocks = 0, n_jac_main = 0;
  int32_t *jac_heavy = nullptr, *jac_light = nullptr;  // the same blocks by role group (measurements)
  int n_jac_heavy = 0, n_jac_light = 0;
  int vac_first = 0, n_vac = 0;  // vacuum dynamics nodes: k_jacobian_noair, one thread per node
  int32_t* res_blocks = nullptr;
  int n_res_blocks = 0;
  // packed output (plan_host.h: build_packed_layout)
  long long n_pack = 0;
  std::vector<int64_t> pk_full, pk_src;
  std::vector<double> pk_sgn;
  std::vector<void*> scen_owned;  // uploads of gelato_plan_set_scenarios (replaced by the next call)
  double *d_packed = nullptr, *h_packed = nullptr;  // staging of the packed host entry points
  size_t cap_packed = 0;
  double *d_px = nullptr, *d_pg = nullptr, *h_px = nullptr;
  int32_t* d_pids = nullptr;
  bool pids_iota = false;
  size_t cap_px = 0;
  double* vals_template = nullptr;       // [n_vals] (or [n_scen][n_vals])
  long long vals_template_sstride = 0;
  int n_scen_cfg = 1;
  cudaStream_t stream = nullptr;
  // staging for the host-buffer entry points
  double *d_x = nullptr, *d_g = nullptr, *d_vals = nullptr;
  double *h_x = nullptr, *h_out = nullptr;  // pinned
  size_t cap_scen = 0;
  // small host-buffer objfunc calls: upload, kernel and download as ONE captured graph launch (the staging pointers
  // and the plan view are baked into it: rebuilt whenever either changes)
  cudaGraphExec_t res_graph = nullptr;
  int res_graph_n = 0;
  double* h_small_out = nullptr;  // pinned staging of the small-problem path of update mode
  size_t cap_small_out = 0;
  // subset batches (gelato_eval_*_ids)
  int32_t* d_ids = nullptr;
  double* d_vals_ids = nullptr;
  size_t cap_ids = 0;
  // update mode
  const int64_t* d_xdep = nullptr;
  std::vector<int64_t> h_xdep;
  long long n_xdep = 0;
  double *d_pack = nullptr, *h_pack = nullptr;  // [cap_pack][n_xdep], h_pack pinned
  size_t cap_pack = 0;
  int host_threads = 0;
  // scattered slots of a page-locked caller buffer: packed + host pool (0) or written from the device (1).
  // With 8+ host threads the pool is 1.6x faster end to end (profiles/r01v_e2e.txt); few cores: zero-copy.
  int update_zero_copy = std::thread::hardware_concurrency() < 8 ? 1 : 0;
  std::vector<std::pair<int64_t, int64_t>> big_runs;  // (first slot, length) of the long runs in xdep_idx
  const int64_t* d_xdep_small = nullptr;              // the slots outside those runs
  std::vector<int64_t> h_small;
  long long n_small = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_kernel = nullptr, ev_copy = nullptr;
  cudaStream_t pair_stream = nullptr;  // residual kernel of a pair evaluation, concurrent with the Jacobian kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // update mode pipelined over slices of the batch: slice k's upload and kernels overlap slice k-1's copies.
  // Slice 0 runs on the streams above, the others on lanes created on first use.
  struct Lane {
    cudaStream_t stream = nullptr, copy_stream = nullptr, pair_stream = nullptr;
    cudaEvent_t ev_kernel = nullptr, ev_copy = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_h2d = nullptr;
    cudaEvent_t ev_g = nullptr, ev_gdone = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
  };
  std::vector<Lane> lanes;
  int update_slices = 0;  // 0 = choose from the batch size
  int32_t* d_iota = nullptr;  // 0, 1, 2 ...: a slice's scenario ids
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

template <typename T>
static int upload(GelatoPlan* p, const T* src, size_t count, const T** dst) {
  *dst = nullptr;
  if (count == 0 || src == nullptr) return GELATO_OK;
  void* d = nullptr;
  CU(cudaMalloc(&d, count * sizeof(T)));
  p->owned.push_back(d);
  CU(cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *dst = static_cast<const T*>(d);
  return GELATO_OK;
}

// One Jacobian evaluation: the block kernel (air dynamics nodes, aero rows, fallback nodes, event rows) on `st` and the
// one-thread-per-node kernel of the vacuum nodes on `aux`, forked from and joined back into `st` with the two events.
// g_dev != NULL: pair evaluation -- both also write objfunc's rows (collocation defects of their nodes, aero / event
// rows at the pristine state) and the linear-row blocks at the end of the block table are launched too.
// packed: output layout.
// (Variants measured and dropped: phase 0 as its own launch with pp | rq | q staged through L2, 38 % slower,
// profiles/r01i_split_ab.txt; the vacuum nodes as blocks of the block kernel, 0.084 ms instead of 0.03,
// profiles/r02_probe.txt.)
// A sub-range of the evaluation (one problem sharded over GPUs): blocks [b0, b0 + nb) of the block table and vacuum
// nodes [v0, v0 + nv) of the vacuum list; nb < 0: everything.
struct JacRange {
  int b0 = 0, nb = -1, v0 = 0, nv = 0;
};
static int launch_jacobian(GelatoPlan* p, const double* x_dev, double* out_dev, double* g_dev, int n_scen, cudaStream_t st,
                           cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join, const int32_t* ids_dev, bool packed,
                           JacRange rg = JacRange()) {
  PlanView v = p->view;
  v.packed = packed ? 1 : 0;
  const bool all = rg.nb < 0;
  const int nb = all ? (g_dev ? p->n_jac_blocks : p->n_jac_main) : rg.nb;
  const int nv = all ? p->n_vac : rg.nv;
  const int v_first = p->vac_first + (all ? 0 : rg.v0);
  const int32_t* table = p->jac_blocks + (size_t)(all ? 0 : rg.b0) * BT_COLS;
  const bool side = nv > 0 && nb > 0 && aux != st;
  if (side) {
    CU(cudaEventRecord(ev_fork, st));
    CU(cudaStreamWaitEvent(aux, ev_fork, 0));
  }
  if (nv > 0) {
    const int per = (nv + 31) / 32;
    k_jacobian_noair<<<(unsigned)per * n_scen, GV_THREADS, 0, side ? aux : st>>>(v, v_first, nv, n_scen, ids_dev, x_dev, out_dev,
                                                                                g_dev);
    p->launches++;
  }
  if (nb > 0) {
    k_jacobian<<<(unsigned)nb * n_scen, GJ_THREADS, 0, st>>>(v, table, n_scen, ids_dev, x_dev, out_dev, g_dev);
    p->launches++;
  }
  CU(cudaGetLastError());
  if (side) {
    CU(cudaEventRecord(ev_join, aux));
    CU(cudaStreamWaitEvent(st, ev_join, 0));
  }
  return GELATO_OK;
}

extern "C" {

const char* gelato_last_error(void) { return g_err.c_str(); }
void gelato_set_error_(const char* msg) { g_err = msg ? msg : ""; }  // for leaf_api.cu

int gelato_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int gelato_plan_create(const GelatoPlanDesc* d, int device, GelatoPlan** out) {
  if (!d || !out) return fail(GELATO_ERR_ARG, "null argument");
  if (const char* why = validate_desc(d)) return fail(GELATO_ERR_ARG, why);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GELATO_ERR_CUDA, "no CUDA device: the B200 kernels are the only evaluation path (no CPU fallback)");
  CU(cudaSetDevice(device));
  GelatoPlan* p = new GelatoPlan();
  p->device = device;
  PlanView& v = p->view;
  planview_scalars(d, v);
  const int S = v.S;
  int rc;
#define UP(field, src, count) if ((rc = upload(p, src, (size_t)(count), &v.field)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  UP(sec_i32, d->sec_i32, (size_t)S * GS_I32_COLS)
  UP(sec_i64, d->sec_i64, (size_t)S * GS_I64_COLS)
  UP(sec_f64, d->sec_f64, (size_t)S * GS_F64_COLS)
  UP(d_pool, d->d_pool, d->d_pool_len)
  UP(tau_pool, d->tau_pool, d->tau_pool_len)
  UP(wind, d->wind, (size_t)d->n_wind * 3)
  UP(ca, d->ca, (size_t)d->n_ca * 2)
  UP(lin_i32, d->lin_i32, (size_t)d->n_lin * GL_I32_COLS)
  UP(lin_f64, d->lin_f64, (size_t)d->n_lin * GL_F64_COLS)
  UP(aero_i32, d->aero_i32, (size_t)d->n_aero * GA_I32_COLS)
  UP(aero_i64, d->aero_i64, (size_t)d->n_aero * GA_I64_COLS)
  UP(aero_f64, d->aero_f64, (size_t)d->n_aero * GA_F64_COLS)
  UP(rc_aero, d->rc_aero, d->rc_aero ? (size_t)v.n_vars : 0)
  UP(evt_i32, d->evt_i32, (size_t)d->n_evt * GE_I32_COLS)
  UP(evt_i64, d->evt_i64, (size_t)d->n_evt * GE_I64_COLS)
  UP(evt_f64, d->evt_f64, (size_t)d->n_evt * GE_F64_COLS)
  if (d->xdep_idx && d->n_xdep > 0) {
    for (int64_t i = 0; i < d->n_xdep; i++)
      if (d->xdep_idx[i] < 0 || d->xdep_idx[i] >= d->n_vals || (i > 0 && d->xdep_idx[i] <= d->xdep_idx[i - 1])) {
        gelato_plan_destroy(p);
        return fail(GELATO_ERR_ARG, "xdep_idx must be strictly ascending and inside [0, n_vals)");
      }
    p->h_xdep.assign(d->xdep_idx, d->xdep_idx + d->n_xdep);
    p->n_xdep = d->n_xdep;
    const int64_t* dx = nullptr;
    if ((rc = upload(p, d->xdep_idx, (size_t)d->n_xdep, &dx)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
    p->d_xdep = dx;
    // split for the zero-copy route: long runs of consecutive slots go through the copy engine (one strided
    // 2-D copy per run covers every scenario), the scattered rest is written by a kernel
    std::vector<int64_t> small;
    for (int64_t i = 0; i < d->n_xdep;) {
      int64_t j = i + 1;
      while (j < d->n_xdep && d->xdep_idx[j] == d->xdep_idx[j - 1] + 1) j++;
      if (j - i >= 512) p->big_runs.push_back({d->xdep_idx[i], j - i});
      else small.insert(small.end(), d->xdep_idx + i, d->xdep_idx + j);
      i = j;
    }
    p->n_small = (long long)small.size();
    p->h_small = small;
    if ((rc = upload(p, small.data(), small.size(), &dx)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
    p->d_xdep_small = dx;
  }
  const double* tmpl = nullptr;
  if ((rc = upload(p, d->vals_template, (size_t)d->n_vals, &tmpl)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  p->vals_template = const_cast<double*>(tmpl);
#undef UP

  // block tables
  HostTables ht;
  build_host_tables(d, ht);
  const int32_t* tb = nullptr;
  p->n_jac_blocks = (int)(ht.jac_blocks.size() / BT_COLS);
  p->n_res_blocks = (int)(ht.res_blocks.size() / BT_COLS);
  p->n_jac_main = ht.n_jac_main;
  p->vac_first = ht.vac_first;
  p->n_vac = ht.n_vac;
  p->n_jac_heavy = (int)(ht.jac_heavy.size() / BT_COLS);
  p->n_jac_light = (int)(ht.jac_light.size() / BT_COLS);
  if ((rc = upload(p, ht.jac_heavy.data(), ht.jac_heavy.size(), &tb)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  p->jac_heavy = const_cast<int32_t*>(tb);
  if ((rc = upload(p, ht.jac_light.data(), ht.jac_light.size(), &tb)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  p->jac_light = const_cast<int32_t*>(tb);
  {
    PackedLayout L;
    build_packed_layout(d, L);
    if (d->xdep_idx && d->n_xdep > 0 &&
        ((size_t)d->n_xdep != L.full_slot.size() || !std::equal(L.full_slot.begin(), L.full_slot.end(), d->xdep_idx))) {
      gelato_plan_destroy(p);
      return fail(GELATO_ERR_ARG, "xdep_idx is not the set of slots the Jacobian kernel writes for this plan");
    }
    p->n_pack = L.n_pack;
    v.n_pack = L.n_pack;
    if ((rc = upload(p, L.sec_pk.data(), L.sec_pk.size(), &v.sec_pk)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
    if ((rc = upload(p, L.aero_pk.data(), L.aero_pk.size(), &v.aero_pk)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
    if ((rc = upload(p, L.evt_pk.data(), L.evt_pk.size(), &v.evt_pk)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
    p->pk_full.swap(L.full_slot);
    p->pk_src.swap(L.src);
    p->pk_sgn.swap(L.sgn);
  }
  if ((rc = upload(p, ht.jac_blocks.data(), ht.jac_blocks.size(), &tb)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  p->jac_blocks = const_cast<int32_t*>(tb);
  if ((rc = upload(p, ht.res_blocks.data(), ht.res_blocks.size(), &tb)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  p->res_blocks = const_cast<int32_t*>(tb);
  if ((rc = upload(p, ht.node_rec.data(), ht.node_rec.size(), &v.node_rec)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  if ((rc = upload(p, ht.jac_rec.data(), ht.jac_rec.size(), &v.jac_rec)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  if ((rc = upload(p, ht.aero_rows.data(), ht.aero_rows.size(), &v.aero_rows)) != GELATO_OK) { gelato_plan_destroy(p); return rc; }
  v.n_aero_rows = (int)ht.aero_rows.size();

  cudaError_t ce = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&p->ev_kernel, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&p->ev_copy, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&p->pair_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreate(&p->ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&p->ev1);
  if (ce != cudaSuccess) {
    gelato_plan_destroy(p);
    return fail(GELATO_ERR_CUDA, std::string("stream / event creation: ") + cudaGetErrorString(ce));
  }
  *out = p;
  return GELATO_OK;
}

int gelato_plan_set_scenarios(GelatoPlan* p, const GelatoScenarioDesc* sc) {
  if (!p || !sc || sc->n_scen <= 0) return fail(GELATO_ERR_ARG, "bad scenario descriptor");
  CU(cudaSetDevice(p->device));
  CU(cudaDeviceSynchronize());  // nothing in flight may still read the blocks being replaced
  PlanView& v = p->view;
  // the blocks of an earlier call are superseded as a whole
  for (void* d : p->scen_owned) cudaFree(d);
  p->scen_owned.clear();
  auto up = [&](const double* src, size_t count, const double** dst) -> int {
    void* d = nullptr;
    CU(cudaMalloc(&d, count * sizeof(double)));
    p->scen_owned.push_back(d);
    CU(cudaMemcpy(d, src, count * sizeof(double), cudaMemcpyHostToDevice));
    *dst = static_cast<const double*>(d);
    return GELATO_OK;
  };
  int rc;
  const double* dptr;
  if (sc->sec_f64) {
    if ((rc = up(sc->sec_f64, (size_t)sc->n_scen * v.S * GS_F64_COLS, &dptr)) != GELATO_OK) return rc;
    v.sec_f64 = dptr;
    v.sec_f64_sstride = (long long)v.S * GS_F64_COLS;
  }
  if (sc->wind) {
    if ((rc = up(sc->wind, (size_t)sc->n_scen * v.n_wind * 3, &dptr)) != GELATO_OK) return rc;
    v.wind = dptr;
    v.wind_sstride = (long long)v.n_wind * 3;
  }
  if (sc->unit_mass) {
    if ((rc = up(sc->unit_mass, (size_t)sc->n_scen, &dptr)) != GELATO_OK) return rc;
    v.unit_mass_scen = dptr;
  }
  if (sc->lin_const) {
    if ((rc = up(sc->lin_const, (size_t)sc->n_scen * v.n_lin, &dptr)) != GELATO_OK) return rc;
    v.lin_const_scen = dptr;
  }
  if (sc->vals_template) {
    if ((rc = up(sc->vals_template, (size_t)sc->n_scen * v.n_vals, &dptr)) != GELATO_OK) return rc;
    p->vals_template = const_cast<double*>(dptr);
    p->vals_template_sstride = v.n_vals;
  }
  p->n_scen_cfg = sc->n_scen;
  // the staging Jacobian buffer holds the constants of the PREVIOUS scenarios: have it rebuilt on the next call
  p->cap_scen = 0;
  p->cap_ids = 0;
  if (p->res_graph) cudaGraphExecDestroy(p->res_graph);
  p->res_graph = nullptr;
  p->res_graph_n = 0;
  return GELATO_OK;
}

int gelato_plan_destroy(GelatoPlan* p) {
  if (!p) return GELATO_OK;
  cudaSetDevice(p->device);
  if (p->res_graph) cudaGraphExecDestroy(p->res_graph);
  if (p->h_small_out) cudaFreeHost(p->h_small_out);
  for (void* d : p->owned) cudaFree(d);
  for (void* d : p->scen_owned) cudaFree(d);
  if (p->d_packed) cudaFree(p->d_packed);
  if (p->h_packed) cudaFreeHost(p->h_packed);
  if (p->d_px) cudaFree(p->d_px);
  if (p->d_pg) cudaFree(p->d_pg);
  if (p->d_pids) cudaFree(p->d_pids);
  if (p->h_px) cudaFreeHost(p->h_px);
  if (p->d_x) cudaFree(p->d_x);
  if (p->d_g) cudaFree(p->d_g);
  if (p->d_vals) cudaFree(p->d_vals);
  if (p->h_x) cudaFreeHost(p->h_x);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->d_pack) cudaFree(p->d_pack);
  if (p->h_pack) cudaFreeHost(p->h_pack);
  if (p->d_ids) cudaFree(p->d_ids);
  if (p->d_vals_ids) cudaFree(p->d_vals_ids);
  for (size_t k = 0; k < p->lanes.size(); k++) {
    GelatoPlan::Lane& L = p->lanes[k];
    for (cudaEvent_t e : L.chunk_ev) cudaEventDestroy(e);
    if (k == 0) continue;  // lane 0 borrows the plan's own streams and events
    for (cudaEvent_t e : {L.ev_kernel, L.ev_copy, L.ev_fork, L.ev_join})
      if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {L.stream, L.copy_stream, L.pair_stream})
      if (st) cudaStreamDestroy(st);
  }
  for (GelatoPlan::Lane& L : p->lanes)
    for (cudaEvent_t e : {L.ev_h2d, L.ev_g, L.ev_gdone})
      if (e) cudaEventDestroy(e);
  if (p->d_iota) cudaFree(p->d_iota);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->ev_kernel) cudaEventDestroy(p->ev_kernel);
  if (p->ev_copy) cudaEventDestroy(p->ev_copy);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  if (p->pair_stream) cudaStreamDestroy(p->pair_stream);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return GELATO_OK;
}

int32_t gelato_plan_n_vars(const GelatoPlan* p) { return p ? p->view.n_vars : 0; }
int32_t gelato_plan_n_rows(const GelatoPlan* p) { return p ? p->view.n_rows : 0; }
int64_t gelato_plan_n_vals(const GelatoPlan* p) { return p ? p->view.n_vals : 0; }
int64_t gelato_plan_launch_count(const GelatoPlan* p) { return p ? p->launches : 0; }
int64_t gelato_plan_n_xdep(const GelatoPlan* p) { return p ? p->n_xdep : 0; }
int32_t gelato_plan_n_blocks(const GelatoPlan* p, int which) {
  if (!p) return 0;
  switch (which) {
    case 0: return p->n_res_blocks;
    case 1: return p->n_jac_main;
    case 2: return p->n_jac_heavy;
    case 3: return p->n_jac_light;
    case 5: return p->n_vac;
    default: return p->n_jac_blocks;
  }
}

static int check_scen(GelatoPlan* p, int n_scen) {
  if (!p) return fail(GELATO_ERR_ARG, "null plan");
  if (n_scen <= 0) return fail(GELATO_ERR_ARG, "n_scen must be positive");
  const bool per_scen = p->view.sec_f64_sstride || p->view.wind_sstride || p->view.unit_mass_scen ||
                        p->view.lin_const_scen || p->vals_template_sstride;
  if (per_scen && n_scen > p->n_scen_cfg)
    return fail(GELATO_ERR_ARG, "n_scen exceeds the configured scenario blocks");
  const long long most = std::max(p->n_jac_blocks, p->n_res_blocks);
  if (most * n_scen > 0x7fffffffLL) return fail(GELATO_ERR_ARG, "blocks x n_scen exceeds the grid limit");
  return GELATO_OK;
}

int gelato_eval_residuals_dev(GelatoPlan* p, const double* x_dev, double* g_dev, int32_t n_scen, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  k_residuals<<<(unsigned)p->n_res_blocks * n_scen, GR_THREADS, 0, st>>>(p->view, p->res_blocks, n_scen, nullptr, x_dev, g_dev);
  p->launches++;
  CU(cudaGetLastError());
  return GELATO_OK;
}

int gelato_eval_pair_dev(GelatoPlan* p, const double* x_dev, double* g_dev, double* vals_dev, int32_t n_scen, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  // objfunc's rows come out of the Jacobian launch: the dynamics blocks hold the right-hand side of the pristine
  // x as their centre column and subtract it from the D.X products; aero and event blocks carry one more column
  // at the pristine state; the linear rows are blocks of their own (no second pass over the physics)
  return launch_jacobian(p, x_dev, vals_dev, g_dev, n_scen, st, p->pair_stream, p->ev_fork, p->ev_join, nullptr, false);
}

int gelato_eval_pair_packed_dev(GelatoPlan* p, const double* x_dev, double* g_dev, double* packed_dev, int32_t n_scen,
                                void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  return launch_jacobian(p, x_dev, packed_dev, g_dev, n_scen, st, p->pair_stream, p->ev_fork, p->ev_join, nullptr, true);
}

// ONE problem sharded over GPUs: the pair evaluation restricted to blocks
// [block_first, block_first + block_count) of the block table and vacuum nodes [vac_first, vac_first + vac_count).
// Every output slot and residual row belongs to exactly one block or vacuum node, so ranks with disjoint ranges that
// cover everything fill disjoint parts of the packed vector and of g.
int gelato_eval_pair_packed_range_dev(GelatoPlan* p, const double* x_dev, double* g_dev, double* packed_dev, int32_t n_scen,
                                      int32_t block_first, int32_t block_count, int32_t vac_first, int32_t vac_count,
                                      void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!g_dev || !packed_dev) return fail(GELATO_ERR_ARG, "null buffer");
  if (block_first < 0 || block_count < 0 || block_first + block_count > p->n_jac_blocks || vac_first < 0 || vac_count < 0 ||
      vac_first + vac_count > p->n_vac)
    return fail(GELATO_ERR_ARG, "block or vacuum-node range outside the plan");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  JacRange rg;
  rg.b0 = block_first; rg.nb = block_count; rg.v0 = vac_first; rg.nv = vac_count;
  return launch_jacobian(p, x_dev, packed_dev, g_dev, n_scen, st, p->pair_stream, p->ev_fork, p->ev_join, nullptr, true, rg);
}

int gelato_fill_template(GelatoPlan* p, double* vals_dev, int32_t n_scen, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  if (p->vals_template_sstride) {
    CU(cudaMemcpyAsync(vals_dev, p->vals_template, (size_t)n_scen * p->view.n_vals * sizeof(double),
                       cudaMemcpyDeviceToDevice, st));
  } else {
    for (int s = 0; s < n_scen; s++)
      CU(cudaMemcpyAsync(vals_dev + (size_t)s * p->view.n_vals, p->vals_template,
                         (size_t)p->view.n_vals * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  return GELATO_OK;
}

int gelato_eval_jacobian_dev(GelatoPlan* p, const double* x_dev, double* vals_dev, int32_t n_scen, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  // the constants and D entries of vals_dev were put there once by gelato_fill_template;
  // the kernels rewrite every x-dependent slot and never touch the rest
  return launch_jacobian(p, x_dev, vals_dev, nullptr, n_scen, st, p->pair_stream, p->ev_fork, p->ev_join, nullptr, false);
}

static void drop_graphs(GelatoPlan* p) {
  if (p->res_graph) cudaGraphExecDestroy(p->res_graph);
  p->res_graph = nullptr;
  p->res_graph_n = 0;
}

static int ensure_staging(GelatoPlan* p, size_t n_scen) {
  if (n_scen <= p->cap_scen) return GELATO_OK;
  CU(cudaSetDevice(p->device));
  drop_graphs(p);
  if (p->d_x) cudaFree(p->d_x);
  if (p->d_g) cudaFree(p->d_g);
  if (p->d_vals) cudaFree(p->d_vals);
  if (p->h_x) cudaFreeHost(p->h_x);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->d_iota) cudaFree(p->d_iota);
  p->d_x = p->d_g = p->d_vals = p->h_x = p->h_out = nullptr;
  p->d_iota = nullptr;
  p->cap_scen = 0;
  const PlanView& v = p->view;
  const size_t nout = std::max<size_t>((size_t)v.n_rows, (size_t)v.n_vals);
  CU(cudaMalloc(&p->d_x, n_scen * v.n_vars * sizeof(double)));
  CU(cudaMalloc(&p->d_g, n_scen * v.n_rows * sizeof(double)));
  CU(cudaMalloc(&p->d_vals, n_scen * (size_t)v.n_vals * sizeof(double)));
  CU(cudaMallocHost(&p->h_x, n_scen * v.n_vars * sizeof(double)));
  CU(cudaMallocHost(&p->h_out, n_scen * nout * sizeof(double)));
  CU(cudaMalloc(&p->d_iota, n_scen * sizeof(int32_t)));
  std::vector<int32_t> iota(n_scen);
  for (size_t k = 0; k < n_scen; k++) iota[k] = (int32_t)k;
  CU(cudaMemcpy(p->d_iota, iota.data(), n_scen * sizeof(int32_t), cudaMemcpyHostToDevice));
  p->cap_scen = n_scen;
  // constants of the Jacobian: written once, the kernel only rewrites x-dependent slots
  int rc = gelato_fill_template(p, p->d_vals, (int)n_scen, p->stream);
  if (rc) return rc;
  CU(cudaStreamSynchronize(p->stream));  // the other lanes' kernels write into it too
  return GELATO_OK;
}

// page-locked host memory can be DMA'd directly; pageable memory goes through the plan's staging buffers
static bool is_pinned(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static int eval_host(GelatoPlan* p, int which, const double* x, double* out, int32_t n_scen,
                     const int32_t* ids = nullptr) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!x || !out) return fail(GELATO_ERR_ARG, "null buffer");
  if ((rc = ensure_staging(p, n_scen))) return rc;
  const PlanView& v = p->view;
  const size_t nx = (size_t)n_scen * v.n_vars;
  const size_t no = (size_t)n_scen * (which == 0 ? (size_t)v.n_rows : (size_t)v.n_vals);
  double* d_out = which == 0 ? p->d_g : p->d_vals;
  if (ids) {  // batch slot k evaluates configured scenario ids[k]
    for (int k = 0; k < n_scen; k++)
      if (ids[k] < 0 || ids[k] >= p->n_scen_cfg) return fail(GELATO_ERR_ARG, "scenario id outside the configured scenarios");
    if ((size_t)n_scen > p->cap_ids) {
      if (p->d_ids) cudaFree(p->d_ids);
      if (p->d_vals_ids) cudaFree(p->d_vals_ids);
      p->d_ids = nullptr;
      p->d_vals_ids = nullptr;
      p->cap_ids = 0;
      CU(cudaMalloc(&p->d_ids, (size_t)n_scen * sizeof(int32_t)));
      CU(cudaMalloc(&p->d_vals_ids, (size_t)n_scen * v.n_vals * sizeof(double)));
      p->cap_ids = n_scen;
    }
    CU(cudaMemcpyAsync(p->d_ids, ids, (size_t)n_scen * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream));
    if (which == 1) {  // constants of each slot's scenario
      d_out = p->d_vals_ids;
      for (int k = 0; k < n_scen; k++)
        CU(cudaMemcpyAsync(d_out + (size_t)k * v.n_vals, p->vals_template + (size_t)ids[k] * p->vals_template_sstride,
                           (size_t)v.n_vals * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    }
  }
  // a small objfunc call (one NLP, a few scenarios) is launch-latency bound: upload + kernel + download go out as one
  // graph launch through the pinned staging buffers (GELATO_B200_NO_GRAPH=1 keeps the three separate submissions)
  static const bool use_graph = !getenv("GELATO_B200_NO_GRAPH");
  if (which == 0 && !ids && use_graph && (nx + no) * sizeof(double) <= (size_t)64 * 1024) {
    if (p->res_graph && p->res_graph_n != n_scen) drop_graphs(p);
    if (!p->res_graph) {
      cudaGraph_t graph = nullptr;
      CU(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
      cudaMemcpyAsync(p->d_x, p->h_x, nx * sizeof(double), cudaMemcpyHostToDevice, p->stream);
      k_residuals<<<(unsigned)p->n_res_blocks * n_scen, GR_THREADS, 0, p->stream>>>(p->view, p->res_blocks, n_scen, nullptr, p->d_x,
                                                                                p->d_g);
      cudaMemcpyAsync(p->h_out, p->d_g, no * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
      CU(cudaStreamEndCapture(p->stream, &graph));
      cudaError_t ge = cudaGraphInstantiate(&p->res_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ge != cudaSuccess) {
        p->res_graph = nullptr;
        return fail(GELATO_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ge));
      }
      p->res_graph_n = n_scen;
    }
    memcpy(p->h_x, x, nx * sizeof(double));
    CU(cudaGraphLaunch(p->res_graph, p->stream));
    p->launches++;
    CU(cudaStreamSynchronize(p->stream));
    memcpy(out, p->h_out, no * sizeof(double));
    return GELATO_OK;
  }
  const double* hx = x;
  if (!is_pinned(x)) {
    memcpy(p->h_x, x, nx * sizeof(double));
    hx = p->h_x;
  }
  CU(cudaMemcpyAsync(p->d_x, hx, nx * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  if (!ids) {
    if (which == 0) rc = gelato_eval_residuals_dev(p, p->d_x, p->d_g, n_scen, p->stream);
    else rc = gelato_eval_jacobian_dev(p, p->d_x, p->d_vals, n_scen, p->stream);
    if (rc) return rc;
  } else {
    if (which == 0) {
      k_residuals<<<(unsigned)p->n_res_blocks * n_scen, GR_THREADS, 0, p->stream>>>(p->view, p->res_blocks, n_scen, p->d_ids,
                                                                                p->d_x, p->d_g);
      p->launches++;
    } else if ((rc = launch_jacobian(p, p->d_x, d_out, nullptr, n_scen, p->stream, p->pair_stream, p->ev_fork, p->ev_join,
                                     p->d_ids, false))) {
      return rc;
    }
    CU(cudaGetLastError());
  }
  const bool direct = is_pinned(out);
  CU(cudaMemcpyAsync(direct ? out : p->h_out, d_out, no * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CU(cudaStreamSynchronize(p->stream));
  if (!direct) memcpy(out, p->h_out, no * sizeof(double));
  return GELATO_OK;
}

int gelato_eval_residuals(GelatoPlan* p, const double* x, double* g, int32_t n_scen) {
  return eval_host(p, 0, x, g, n_scen);
}

int gelato_eval_jacobian(GelatoPlan* p, const double* x, double* vals, int32_t n_scen) {
  return eval_host(p, 1, x, vals, n_scen);
}

int gelato_eval_residuals_ids(GelatoPlan* p, const double* x, double* g, int32_t n_scen, const int32_t* scen_ids) {
  if (!scen_ids) return fail(GELATO_ERR_ARG, "null scen_ids");
  return eval_host(p, 0, x, g, n_scen, scen_ids);
}

int gelato_eval_jacobian_ids(GelatoPlan* p, const double* x, double* vals, int32_t n_scen, const int32_t* scen_ids) {
  if (!scen_ids) return fail(GELATO_ERR_ARG, "null scen_ids");
  return eval_host(p, 1, x, vals, n_scen, scen_ids);
}

int gelato_pack_xdep_dev(GelatoPlan* p, const double* vals_dev, double* packed_dev, int32_t n_scen, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!p->d_xdep) return fail(GELATO_ERR_ARG, "the plan was created without xdep_idx");
  if (n_scen > 65535) return fail(GELATO_ERR_ARG, "n_scen > 65535 in update mode (gridDim.y)");
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  const int threads = 256;
  const int bx = (int)std::min<long long>((p->n_xdep + threads - 1) / threads, 4096);
  k_pack_xdep<<<dim3(bx, n_scen), threads, 0, st>>>(vals_dev, p->d_xdep, p->n_xdep, p->view.n_vals, packed_dev);
  p->launches++;
  CU(cudaGetLastError());
  return GELATO_OK;
}

int gelato_jacobian_template(GelatoPlan* p, double* vals, int32_t n_scen) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!vals) return fail(GELATO_ERR_ARG, "null buffer");
  CU(cudaSetDevice(p->device));
  const size_t nv = (size_t)p->view.n_vals;
  if (p->vals_template_sstride) {
    CU(cudaMemcpy(vals, p->vals_template, (size_t)n_scen * nv * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    CU(cudaMemcpy(vals, p->vals_template, nv * sizeof(double), cudaMemcpyDeviceToHost));
    for (int s = 1; s < n_scen; s++) memcpy(vals + (size_t)s * nv, vals, nv * sizeof(double));
  }
  return GELATO_OK;
}

int gelato_set_update_zero_copy(GelatoPlan* p, int32_t on) {
  if (!p) return fail(GELATO_ERR_ARG, "null plan");
  p->update_zero_copy = on ? 1 : 0;
  return GELATO_OK;
}

int gelato_set_host_threads(GelatoPlan* p, int32_t n) {
  if (!p || n < 0) return fail(GELATO_ERR_ARG, "bad thread count");
  p->host_threads = n;
  return GELATO_OK;
}

using gelato_host::scatter_parallel;

static int ensure_lanes(GelatoPlan* p, int n) {
  if (p->lanes.empty()) {
    GelatoPlan::Lane L;
    L.stream = p->stream; L.copy_stream = p->copy_stream; L.pair_stream = p->pair_stream;
    L.ev_kernel = p->ev_kernel; L.ev_copy = p->ev_copy; L.ev_fork = p->ev_fork; L.ev_join = p->ev_join;
    p->lanes.push_back(L);
  }
  while ((int)p->lanes.size() < n) {
    p->lanes.emplace_back();
    GelatoPlan::Lane& L = p->lanes.back();  // registered first: gelato_plan_destroy releases whatever exists
    for (cudaStream_t* st : {&L.stream, &L.copy_stream, &L.pair_stream}) CU(cudaStreamCreateWithFlags(st, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&L.ev_kernel, &L.ev_copy, &L.ev_fork, &L.ev_join}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  for (GelatoPlan::Lane& L : p->lanes)
    for (cudaEvent_t* e : {&L.ev_h2d, &L.ev_g, &L.ev_gdone})
      if (!*e) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  return GELATO_OK;
}

// shared body of gelato_eval_jacobian_update / gelato_eval_pair_update; g == NULL: Jacobian only.
// The batch is cut into slices of scenarios, each on its own lane of streams: slice k's upload and kernels
// overlap slice k-1's device->host traffic (PCIe is full duplex), and the host threads scatter slice k-1's
// packed slots meanwhile.  Per slice: upload x -> { residual kernel -> copy g | Jacobian kernel -> transfers }.
static int ensure_small_out(GelatoPlan* p, size_t n_doubles) {
  if (n_doubles <= p->cap_small_out) return GELATO_OK;
  if (p->h_small_out) cudaFreeHost(p->h_small_out);
  p->h_small_out = nullptr;
  p->cap_small_out = 0;
  CU(cudaMallocHost(&p->h_small_out, n_doubles * sizeof(double)));
  p->cap_small_out = n_doubles;
  return GELATO_OK;
}

static int eval_update(GelatoPlan* p, const double* x, double* g, double* vals, int32_t n_scen) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!x || !vals) return fail(GELATO_ERR_ARG, "null buffer");
  if (!p->d_xdep) return fail(GELATO_ERR_ARG, "the plan was created without xdep_idx");
  if (n_scen > 65535) return fail(GELATO_ERR_ARG, "n_scen > 65535 in update mode (gridDim.y)");
  if ((rc = ensure_staging(p, n_scen))) return rc;
  const PlanView& v = p->view;
  if ((size_t)n_scen > p->cap_pack) {
    if (p->d_pack) cudaFree(p->d_pack);
    if (p->h_pack) cudaFreeHost(p->h_pack);
    p->d_pack = p->h_pack = nullptr;
    p->cap_pack = 0;
    CU(cudaMalloc(&p->d_pack, (size_t)n_scen * p->n_xdep * sizeof(double)));
    CU(cudaMallocHost(&p->h_pack, (size_t)n_scen * p->n_xdep * sizeof(double)));
    p->cap_pack = n_scen;
  }
  const double* hx = x;
  if (!is_pinned(x)) {
    memcpy(p->h_x, x, (size_t)n_scen * v.n_vars * sizeof(double));
    hx = p->h_x;
  }
  // A small Jacobian (one NLP of a few hundred nodes) is cheaper to bring down whole than through the gather / strided
  // copies / host scatter below: the staging copy on the device always holds the complete value vector (constants
  // filled once, x-dependent slots rewritten by every launch), so one contiguous copy refreshes the caller's buffer
  // to the same bits.
  if ((size_t)n_scen * (size_t)v.n_vals * sizeof(double) <= (size_t)512 * 1024) {
    CU(cudaMemcpyAsync(p->d_x, hx, (size_t)n_scen * v.n_vars * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if ((rc = launch_jacobian(p, p->d_x, p->d_vals, g ? p->d_g : nullptr, n_scen, p->stream, p->pair_stream, p->ev_fork,
                              p->ev_join, nullptr, false)))
      return rc;
    const size_t nv = (size_t)n_scen * (size_t)v.n_vals, ng = (size_t)n_scen * v.n_rows;
    const bool v_direct = is_pinned(vals), gd = g && is_pinned(g);
    double* v_dst = vals;
    double* g_dst = g;
    if (!v_direct || (g && !gd)) {  // pageable: through one pinned staging block laid out [vals | g]
      if ((rc = ensure_small_out(p, nv + ng))) return rc;
      if (!v_direct) v_dst = p->h_small_out;
      if (g && !gd) g_dst = p->h_small_out + nv;
    }
    CU(cudaMemcpyAsync(v_dst, p->d_vals, nv * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (g) CU(cudaMemcpyAsync(g_dst, p->d_g, ng * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    if (v_dst != vals) memcpy(vals, v_dst, nv * sizeof(double));
    if (g && g_dst != g) memcpy(g, g_dst, ng * sizeof(double));
    return GELATO_OK;
  }
  const bool g_direct = g && is_pinned(g);
  double* const g_dst = g_direct ? g : p->h_out;

  // Which slots travel packed (gathered on the device, copied, scattered by the host pool) and which go
  // straight to their place: a page-locked `vals` takes the long runs of consecutive slots through the copy
  // engine (one strided 2-D copy per run covers a slice's scenarios) and, in zero-copy mode, the scattered
  // rest through a kernel that stores into the buffer's device mapping.
  const bool direct = is_pinned(vals);
  const int64_t* pack_idx_dev = direct ? p->d_xdep_small : p->d_xdep;
  const int64_t* pack_idx_host = direct ? p->h_small.data() : p->h_xdep.data();
  long long n_pack = direct ? p->n_small : p->n_xdep;
  double* dev_view = nullptr;  // the device's address of the caller's page-locked buffer
  if (direct && p->update_zero_copy && n_pack > 0) {
    if (cudaHostGetDevicePointer((void**)&dev_view, vals, 0) != cudaSuccess) {
      cudaGetLastError();  // registered without a device mapping: the scattered slots travel packed instead
      dev_view = nullptr;
    }
  }
  const bool zero_copy = dev_view != nullptr;

  int threads = p->host_threads > 0 ? p->host_threads : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if ((long long)n_scen * n_pack < (1LL << 18)) threads = 1;
  threads = std::min(threads, (int)n_scen);
  int n_slices = p->update_slices > 0 ? p->update_slices : n_scen / 16;  // measured: profiles/r01v_e2e.txt
  n_slices = std::max(1, std::min(n_slices, std::min((int)n_scen, 8)));
  // one slice: the packed slots come down in chunks so the scatter of chunk k overlaps the copy of chunk k+1
  const int n_chunks = (n_slices == 1 && n_scen >= 4 * threads && threads > 1) ? 4 : 1;
  if ((rc = ensure_lanes(p, n_slices))) return rc;
  for (int k = 0; k < n_slices; k++)
    while ((int)p->lanes[k].chunk_ev.size() < n_chunks) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p->lanes[k].chunk_ev.push_back(e);
    }
  auto slice_lo = [&](int k) { return (int)((long long)n_scen * k / n_slices); };
  auto chunk_lo = [&](int k, int c) {
    const int s0 = slice_lo(k), s1 = slice_lo(k + 1);
    return s0 + (int)((long long)(s1 - s0) * c / n_chunks);
  };

  for (int k = 0; k < n_slices; k++) {
    GelatoPlan::Lane& L = p->lanes[k];
    const int s0 = slice_lo(k), ns = slice_lo(k + 1) - s0;
    const int32_t* ids = n_slices > 1 ? p->d_iota + s0 : nullptr;  // the slice's scenario parameter blocks
    double* const dx = p->d_x + (size_t)s0 * v.n_vars;
    double* const dg = p->d_g + (size_t)s0 * v.n_rows;
    double* const dvals = p->d_vals + (size_t)s0 * v.n_vals;
    double* const hvals = vals + (size_t)s0 * v.n_vals;
    if (k > 0) CU(cudaStreamWaitEvent(L.stream, p->lanes[k - 1].ev_h2d, 0));  // uploads one after the other
    CU(cudaMemcpyAsync(dx, hx + (size_t)s0 * v.n_vars, (size_t)ns * v.n_vars * sizeof(double), cudaMemcpyHostToDevice, L.stream));
    CU(cudaEventRecord(L.ev_h2d, L.stream));
    // g != NULL: pair evaluation, the residual rows come out of the same launches
    if ((rc = launch_jacobian(p, dx, dvals, g ? dg : nullptr, ns, L.stream, L.pair_stream, L.ev_fork, L.ev_join, ids, false)))
      return rc;
    if (g) {  // its copy travels on the side stream, next to the Jacobian's transfers
      CU(cudaEventRecord(L.ev_g, L.stream));
      CU(cudaStreamWaitEvent(L.pair_stream, L.ev_g, 0));
      CU(cudaMemcpyAsync(g_dst + (size_t)s0 * v.n_rows, dg, (size_t)ns * v.n_rows * sizeof(double), cudaMemcpyDeviceToHost,
                         L.pair_stream));
      CU(cudaEventRecord(L.ev_gdone, L.pair_stream));
    }

    // the long runs, on the lane's copy stream once `after` has happened
    auto issue_runs = [&](cudaEvent_t after) -> int {
      CU(cudaStreamWaitEvent(L.copy_stream, after, 0));
      const size_t pitch = (size_t)v.n_vals * sizeof(double);
      for (const auto& run : p->big_runs)
        CU(cudaMemcpy2DAsync(hvals + run.first, pitch, dvals + run.first, pitch, (size_t)run.second * sizeof(double),
                             (size_t)ns, cudaMemcpyDeviceToHost, L.copy_stream));
      CU(cudaEventRecord(L.ev_copy, L.copy_stream));
      return GELATO_OK;
    };
    const int gpu_threads = 256;
    const int bx = (int)std::max<long long>(1, std::min<long long>((n_pack + gpu_threads - 1) / gpu_threads, 4096));
    if (direct && (zero_copy || n_pack == 0)) {
      CU(cudaEventRecord(L.ev_kernel, L.stream));
      if (zero_copy) {
        k_scatter_xdep_host<<<dim3(bx, ns), gpu_threads, 0, L.stream>>>(dvals, pack_idx_dev, n_pack, v.n_vals,
                                                                       dev_view + (size_t)s0 * v.n_vals);
        p->launches++;
        CU(cudaGetLastError());
      }
      if ((rc = issue_runs(L.ev_kernel))) return rc;
    } else if (n_pack > 0) {
      k_pack_xdep<<<dim3(bx, ns), gpu_threads, 0, L.stream>>>(dvals, pack_idx_dev, n_pack, v.n_vals,
                                                             p->d_pack + (size_t)s0 * n_pack);
      p->launches++;
      CU(cudaGetLastError());
      for (int c = 0; c < n_chunks; c++) {
        const size_t off = (size_t)chunk_lo(k, c) * n_pack, cnt = (size_t)(chunk_lo(k, c + 1) - chunk_lo(k, c)) * n_pack;
        CU(cudaMemcpyAsync(p->h_pack + off, p->d_pack + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, L.stream));
        CU(cudaEventRecord(L.chunk_ev[c], L.stream));
      }
      // the packed slots cross PCIe first: the host threads scatter them while the long runs follow
      if (direct && (rc = issue_runs(L.chunk_ev[n_chunks - 1]))) return rc;
    }
  }
  if (!zero_copy && n_pack > 0)
    for (int k = 0; k < n_slices; k++)
      for (int c = 0; c < n_chunks; c++) {
        CU(cudaEventSynchronize(p->lanes[k].chunk_ev[c]));
        scatter_parallel(pack_idx_host, n_pack, p->h_pack, vals, v.n_vals, chunk_lo(k, c), chunk_lo(k, c + 1), threads);
      }
  for (int k = 0; k < n_slices; k++) {
    GelatoPlan::Lane& L = p->lanes[k];
    if (direct) CU(cudaStreamWaitEvent(L.stream, L.ev_copy, 0));
    if (g) CU(cudaStreamWaitEvent(L.stream, L.ev_gdone, 0));
    CU(cudaStreamSynchronize(L.stream));
  }
  if (g && !g_direct) memcpy(g, p->h_out, (size_t)n_scen * v.n_rows * sizeof(double));
  return GELATO_OK;
}

// Packed host entry points: x[n_scen][n_vars] up, the pair (or Jacobian-only) evaluation, then g[n_scen][n_rows] and
// packed[n_scen][n_pack] down as CONTIGUOUS copies -- no gather kernel, no host scatter.  Pipelined over slices of
// the batch like update mode: slice k's upload and kernels overlap slice k-1's device->host copies (PCIe is full
// duplex).  Page-locked caller buffers (gelato_host_alloc) are DMA'd directly, others staged.
static int eval_packed(GelatoPlan* p, const double* x, double* g, double* packed, int32_t n_scen, const int32_t* scen_ids) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!x || !packed) return fail(GELATO_ERR_ARG, "null buffer");
  const PlanView& v = p->view;
  if (scen_ids)
    for (int k = 0; k < n_scen; k++)
      if (scen_ids[k] < 0 || scen_ids[k] >= p->n_scen_cfg) return fail(GELATO_ERR_ARG, "scenario id outside the configured scenarios");
  CU(cudaSetDevice(p->device));
  const size_t np = (size_t)p->n_pack;
  if ((size_t)n_scen > p->cap_packed) {
    if (p->d_packed) cudaFree(p->d_packed);
    if (p->h_packed) cudaFreeHost(p->h_packed);
    p->d_packed = p->h_packed = nullptr;
    p->cap_packed = 0;
    CU(cudaMalloc(&p->d_packed, (size_t)n_scen * np * sizeof(double)));
    CU(cudaMallocHost(&p->h_packed, (size_t)n_scen * np * sizeof(double)));
    p->cap_packed = n_scen;
  }
  // device x / g / ids: shared with the COO entry points when those were used, else allocated here
  if ((size_t)n_scen > p->cap_px) {
    for (void* d : {(void*)p->d_px, (void*)p->d_pg, (void*)p->d_pids}) if (d) cudaFree(d);
    if (p->h_px) cudaFreeHost(p->h_px);
    p->d_px = p->d_pg = nullptr; p->d_pids = nullptr; p->h_px = nullptr;
    p->cap_px = 0;
    CU(cudaMalloc(&p->d_px, (size_t)n_scen * v.n_vars * sizeof(double)));
    CU(cudaMalloc(&p->d_pg, (size_t)n_scen * v.n_rows * sizeof(double)));
    CU(cudaMalloc(&p->d_pids, (size_t)n_scen * sizeof(int32_t)));
    CU(cudaMallocHost(&p->h_px, (size_t)n_scen * ((size_t)v.n_vars + (size_t)v.n_rows) * sizeof(double)));
    p->cap_px = n_scen;
    p->pids_iota = false;
  }
  const double* hx = x;
  if (!is_pinned(x)) {
    memcpy(p->h_px, x, (size_t)n_scen * v.n_vars * sizeof(double));
    hx = p->h_px;
  }
  double* const h_g_stage = p->h_px + (size_t)n_scen * v.n_vars;
  const bool g_direct = g && is_pinned(g), pk_direct = is_pinned(packed);
  double* const g_dst = g_direct ? g : h_g_stage;
  double* const pk_dst = pk_direct ? packed : p->h_packed;

  int n_slices = p->update_slices > 0 ? p->update_slices : n_scen / 16;
  n_slices = std::max(1, std::min(n_slices, std::min((int)n_scen, 8)));
  if ((rc = ensure_lanes(p, n_slices))) return rc;
  // scenario ids of the batch slots: the caller's subset, or 0, 1, 2 ... when the batch is cut into slices
  const bool need_ids = scen_ids || n_slices > 1;
  if (scen_ids) {
    CU(cudaMemcpyAsync(p->d_pids, scen_ids, (size_t)n_scen * sizeof(int32_t), cudaMemcpyHostToDevice, p->lanes[0].stream));
    p->pids_iota = false;
  } else if (need_ids && !p->pids_iota) {
    std::vector<int32_t> iota(p->cap_px);
    for (size_t k = 0; k < iota.size(); k++) iota[k] = (int32_t)k;
    CU(cudaMemcpy(p->d_pids, iota.data(), iota.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    p->pids_iota = true;
  }
  auto slice_lo = [&](int k) { return (int)((long long)n_scen * k / n_slices); };
  for (int k = 0; k < n_slices; k++) {
    GelatoPlan::Lane& L = p->lanes[k];
    const int s0 = slice_lo(k), ns = slice_lo(k + 1) - s0;
    double* const dx = p->d_px + (size_t)s0 * v.n_vars;
    double* const dg = p->d_pg + (size_t)s0 * v.n_rows;
    double* const dpk = p->d_packed + (size_t)s0 * np;
    if (k > 0) CU(cudaStreamWaitEvent(L.stream, p->lanes[k - 1].ev_h2d, 0));  // uploads one after the other (and after the ids)
    CU(cudaMemcpyAsync(dx, hx + (size_t)s0 * v.n_vars, (size_t)ns * v.n_vars * sizeof(double), cudaMemcpyHostToDevice, L.stream));
    CU(cudaEventRecord(L.ev_h2d, L.stream));
    if ((rc = launch_jacobian(p, dx, dpk, g ? dg : nullptr, ns, L.stream, L.pair_stream, L.ev_fork, L.ev_join,
                              need_ids ? p->d_pids + s0 : nullptr, true)))
      return rc;
    // both results on the lane's copy stream, after everything of the slice has been computed
    CU(cudaEventRecord(L.ev_kernel, L.stream));
    CU(cudaStreamWaitEvent(L.copy_stream, L.ev_kernel, 0));
    if (g)
      CU(cudaMemcpyAsync(g_dst + (size_t)s0 * v.n_rows, dg, (size_t)ns * v.n_rows * sizeof(double), cudaMemcpyDeviceToHost,
                         L.copy_stream));
    CU(cudaMemcpyAsync(pk_dst + (size_t)s0 * np, dpk, (size_t)ns * np * sizeof(double), cudaMemcpyDeviceToHost, L.copy_stream));
    CU(cudaEventRecord(L.ev_copy, L.copy_stream));
  }
  for (int k = 0; k < n_slices; k++) CU(cudaEventSynchronize(p->lanes[k].ev_copy));
  if (g && !g_direct) memcpy(g, h_g_stage, (size_t)n_scen * v.n_rows * sizeof(double));
  if (!pk_direct) memcpy(packed, p->h_packed, (size_t)n_scen * np * sizeof(double));
  return GELATO_OK;
}

int gelato_eval_pair_packed(GelatoPlan* p, const double* x, double* g, double* packed, int32_t n_scen) {
  if (!g) return fail(GELATO_ERR_ARG, "null buffer");
  return eval_packed(p, x, g, packed, n_scen, nullptr);
}

int gelato_eval_jacobian_packed(GelatoPlan* p, const double* x, double* packed, int32_t n_scen) {
  return eval_packed(p, x, nullptr, packed, n_scen, nullptr);
}

int gelato_eval_pair_packed_ids(GelatoPlan* p, const double* x, double* g, double* packed, int32_t n_scen,
                                const int32_t* scen_ids) {
  if (!g || !scen_ids) return fail(GELATO_ERR_ARG, "null buffer");
  return eval_packed(p, x, g, packed, n_scen, scen_ids);
}

int64_t gelato_plan_n_pack(const GelatoPlan* p) { return p ? p->n_pack : 0; }

int gelato_plan_packed_map(const GelatoPlan* p, int64_t* full_slot, int64_t* src, double* sgn) {
  if (!p || !full_slot || !src || !sgn) return fail(GELATO_ERR_ARG, "null argument");
  std::copy(p->pk_full.begin(), p->pk_full.end(), full_slot);
  std::copy(p->pk_src.begin(), p->pk_src.end(), src);
  std::copy(p->pk_sgn.begin(), p->pk_sgn.end(), sgn);
  return GELATO_OK;
}

int gelato_set_update_slices(GelatoPlan* p, int32_t n) {
  if (!p || n < 0) return fail(GELATO_ERR_ARG, "bad slice count");
  p->update_slices = n;
  return GELATO_OK;
}

int gelato_eval_jacobian_update(GelatoPlan* p, const double* x, double* vals, int32_t n_scen) {
  return eval_update(p, x, nullptr, vals, n_scen);
}

int gelato_eval_pair_update(GelatoPlan* p, const double* x, double* g, double* vals, int32_t n_scen) {
  if (!g) return fail(GELATO_ERR_ARG, "null buffer");
  return eval_update(p, x, g, vals, n_scen);
}

// Measurement only: device times (ms) of the transfer pieces of update mode, each alone, for a page-locked
// `vals` after a gelato_eval_*_update call: [0] the 2-D copies of the long runs, [1] the zero-copy kernel of the
// scattered slots, [2] both at once, [3] pack kernel + one contiguous copy of the scattered slots (no host
// scatter), [4] one contiguous copy of as many bytes as all x-dependent slots, [5] the residual copy.
int gelato_probe_update(GelatoPlan* p, double* vals, int32_t n_scen, int reps, float* out_ms) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (!vals || !out_ms || reps <= 0 || !is_pinned(vals)) return fail(GELATO_ERR_ARG, "page-locked vals required");
  if (!p->d_xdep || (size_t)n_scen > p->cap_pack) return fail(GELATO_ERR_ARG, "call gelato_eval_jacobian_update first");
  const PlanView& v = p->view;
  double* dev_view = nullptr;
  CU(cudaHostGetDevicePointer((void**)&dev_view, vals, 0));
  const size_t pitch = (size_t)v.n_vals * sizeof(double);
  const int threads = 256;
  const int bx = (int)std::min<long long>((p->n_small + threads - 1) / threads, 4096);
  auto copies = [&]() {
    for (const auto& run : p->big_runs)
      cudaMemcpy2DAsync(vals + run.first, pitch, p->d_vals + run.first, pitch, (size_t)run.second * sizeof(double),
                        (size_t)n_scen, cudaMemcpyDeviceToHost, p->copy_stream);
  };
  auto zero_copy = [&]() {
    if (p->n_small > 0)
      k_scatter_xdep_host<<<dim3(bx, n_scen), threads, 0, p->stream>>>(p->d_vals, p->d_xdep_small, p->n_small, v.n_vals, dev_view);
  };
  for (int piece = 0; piece < 6; piece++) {
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(p->ev0, p->stream));
    CU(cudaStreamWaitEvent(p->copy_stream, p->ev0, 0));
    for (int r = 0; r < reps; r++) {
      if (piece == 0 || piece == 2) copies();
      if (piece == 1 || piece == 2) zero_copy();
      if (piece == 3 && p->n_small > 0) {
        k_pack_xdep<<<dim3(bx, n_scen), threads, 0, p->stream>>>(p->d_vals, p->d_xdep_small, p->n_small, v.n_vals, p->d_pack);
        cudaMemcpyAsync(p->h_pack, p->d_pack, (size_t)n_scen * p->n_small * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
      }
      if (piece == 4)
        cudaMemcpyAsync(p->h_pack, p->d_pack, (size_t)n_scen * p->n_xdep * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
      if (piece == 5)
        cudaMemcpyAsync(p->h_out, p->d_g, (size_t)n_scen * v.n_rows * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
    }
    CU(cudaEventRecord(p->ev_copy, p->copy_stream));
    CU(cudaStreamWaitEvent(p->stream, p->ev_copy, 0));
    CU(cudaEventRecord(p->ev1, p->stream));
    CU(cudaEventSynchronize(p->ev1));
    CU(cudaGetLastError());
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    out_ms[piece] = ms / reps;
  }
  return GELATO_OK;
}

int gelato_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(GELATO_ERR_ARG, "null");
  CU(cudaMallocHost(out, bytes));
  return GELATO_OK;
}

int gelato_host_free(void* ptr) {
  if (ptr) CU(cudaFreeHost(ptr));
  return GELATO_OK;
}

// Measurement helper: enqueue exactly ONE kernel on `stream` (0 residual kernel | 2 heavy Jacobian kernel |
// 3 light Jacobian kernel | 4 the residual kernel's non-dynamics blocks; out_dev = g for 0 and 4), COO or packed
// output, g_dev != NULL: with the pair evaluation's defect rows -- so that a benchmark can bracket it with its own
// CUDA events.
int gelato_launch_kernel_dev(GelatoPlan* p, int which, const double* x_dev, double* out_dev, double* g_dev, int32_t n_scen,
                             int32_t packed, void* stream) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  CU(cudaSetDevice(p->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : p->stream;
  PlanView v = p->view;
  v.packed = packed ? 1 : 0;
  if (which == 0) {
    k_residuals<<<(unsigned)p->n_res_blocks * n_scen, GR_THREADS, 0, st>>>(v, p->res_blocks, n_scen, nullptr, x_dev, out_dev);
  } else if (which == 1) {
    return launch_jacobian(p, x_dev, out_dev, g_dev, n_scen, st, p->pair_stream, p->ev_fork, p->ev_join, nullptr, packed != 0);
  } else if (which == 2) {
    if (p->n_jac_heavy == 0) return fail(GELATO_ERR_ARG, "the plan has no heavy blocks");
    k_jacobian_heavy<<<(unsigned)p->n_jac_heavy * n_scen, GJ_THREADS, 0, st>>>(v, p->jac_heavy, n_scen, nullptr, x_dev, out_dev, g_dev);
  } else if (which == 3) {
    if (p->n_jac_light == 0) return fail(GELATO_ERR_ARG, "the plan has no light blocks");
    k_jacobian_light<<<(unsigned)p->n_jac_light * n_scen, GJ_THREADS, 0, st>>>(v, p->jac_light, n_scen, nullptr, x_dev, out_dev, g_dev);
  } else if (which == 5) {
    if (p->n_vac == 0) return fail(GELATO_ERR_ARG, "the plan has no vacuum nodes");
    const int per = (p->n_vac + 31) / 32;
    k_jacobian_noair<<<(unsigned)per * n_scen, GV_THREADS, 0, st>>>(v, p->vac_first, p->n_vac, n_scen, nullptr, x_dev, out_dev, g_dev);
  } else if (which == 6) {  // the block kernel alone, as an evaluation launches it
    const int nb = g_dev ? p->n_jac_blocks : p->n_jac_main;
    if (nb == 0) return fail(GELATO_ERR_ARG, "the plan has no Jacobian blocks");
    k_jacobian<<<(unsigned)nb * n_scen, GJ_THREADS, 0, st>>>(v, p->jac_blocks, n_scen, nullptr, x_dev, out_dev, g_dev);
  } else {
    return fail(GELATO_ERR_ARG, "which must be 0, 1, 2, 3, 5 or 6");
  }
  p->launches++;
  CU(cudaGetLastError());
  return GELATO_OK;
}

int gelato_time_kernel(GelatoPlan* p, int which, const double* x_dev, double* out_dev, int32_t n_scen, int reps,
                       float* avg_ms) {
  int rc = check_scen(p, n_scen);
  if (rc) return rc;
  if (reps <= 0 || !avg_ms) return fail(GELATO_ERR_ARG, "bad reps");
  CU(cudaSetDevice(p->device));
  CU(cudaStreamSynchronize(p->stream));
  CU(cudaEventRecord(p->ev0, p->stream));
  for (int i = 0; i < reps; i++) {
    if (which == 0) {
      k_residuals<<<(unsigned)p->n_res_blocks * n_scen, GR_THREADS, 0, p->stream>>>(p->view, p->res_blocks, n_scen, nullptr, x_dev, out_dev);
    } else if (which == 1) {  // the whole Jacobian evaluation (heavy and light kernels)
      if ((rc = launch_jacobian(p, x_dev, out_dev, nullptr, n_scen, p->stream, p->pair_stream, p->ev_fork, p->ev_join, nullptr,
                                false)))
        return rc;
      continue;
    } else if (which == 2) {  // the heavy roles' blocks alone (air dynamics + aero rows)
      if (p->n_jac_heavy > 0)
        k_jacobian_heavy<<<(unsigned)p->n_jac_heavy * n_scen, GJ_THREADS, 0, p->stream>>>(p->view, p->jac_heavy, n_scen, nullptr,
                                                                                        x_dev, out_dev, nullptr);
    } else {  // the light roles' blocks alone
      if (p->n_jac_light > 0)
        k_jacobian_light<<<(unsigned)p->n_jac_light * n_scen, GJ_THREADS, 0, p->stream>>>(p->view, p->jac_light, n_scen, nullptr,
                                                                                        x_dev, out_dev, nullptr);
    }
    p->launches++;
  }
  CU(cudaEventRecord(p->ev1, p->stream));
  CU(cudaEventSynchronize(p->ev1));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
  *avg_ms = ms / reps;
  return GELATO_OK;
}

int gelato_selftest_unfused(int device, int* ok) {
  if (!ok) return fail(GELATO_ERR_ARG, "null");
  CU(cudaSetDevice(device));
  double* d = nullptr;
  CU(cudaMalloc(&d, sizeof(double)));
  k_unfused_probe<<<1, 1>>>(1.0 + 0x1p-30, 1.0 - 0x1p-30, -1.0, d);
  double h = 1.0;
  CU(cudaMemcpy(&h, d, sizeof(double), cudaMemcpyDeviceToHost));
  cudaFree(d);
  *ok = (h == 0.0);
  return GELATO_OK;
}

int gelato_fp64_peak(int device, double* tflops_fma, double* tflops_nofma) {
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  CU(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float ms;
  for (int pass = 0; pass < 2; pass++) {
    k_fp64_fma<<<blocks, threads>>>(d, iters);  // warm-up on pass 0
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(e0));
    k_fp64_fma<<<blocks, threads>>>(d, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (tflops_fma) *tflops_fma = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    k_fp64_muladd<<<blocks, threads>>>(d, iters);
    CU(cudaDeviceSynchronize());
    CU(cudaEventRecord(e0));
    k_fp64_muladd<<<blocks, threads>>>(d, iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (tflops_nofma) *tflops_nofma = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return GELATO_OK;
}

#ifdef GJ_CLOCKS
// measurement build: copy (and clear) the per-(role, warp) phase cycle sums; out[8][16][6]
int gelato_debug_clocks(unsigned long long* out) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out, gj_clk_sum, sizeof(unsigned long long) * 8 * 16 * 6));
  static unsigned long long zero[8 * 16 * 6];
  CU(cudaMemcpyToSymbol(gj_clk_sum, zero, sizeof zero));
  return GELATO_OK;
}
#endif

}  // extern "C"
