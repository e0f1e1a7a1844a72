#!/usr/bin/env python
"""Benchmark of the NLP-callback hot path (BASELINE.json metric: FD-Jacobian +
residual leaf evaluations per second).

Workload (BASELINE.json configs[1]): the shipped example refined to ~1 000 LGR
nodes (x15 -> N = 990 in 53 sections of <= 20 nodes).  One STEP = one `objfunc`
+ one `sens` (residual kernel + FD-Jacobian kernel) over a batch of 128 dispersed
launch scenarios of that problem per GPU (8 GPUs = the 1 024 scenarios of configs[3]) (mass / thrust / wind perturbations, each
scenario its own decision vector).  An "eval" is one physics-leaf evaluation at
one (node x perturbation column): see CompiledPlan.eval_counts / DESIGN.md.

  value   device-resident: x already in HBM, outputs stay in HBM, CUDA events; the two kernels
          of a step are launched as one pair (residual kernel on a side stream).
  e2e     the same step through the host-buffer C-ABI calls: x from page-locked
          host memory -> device, both kernels, residual vector and Jacobian values
          back into host buffers, wall clock.  The Jacobian uses update mode
          (gelato_eval_jacobian_update: the x-dependent slots are packed on the
          device, copied, and scattered into the batch's persistent host buffer);
          e2e.full_copy is the same step copying the FULL value vector each call.
Multi-GPU: scenarios are independent NLPs; each rank owns `--scenarios` of them
(weak scaling), no data-path collective (DESIGN.md "multi-GPU").

`--impl reference` times the CPU path instead: the reference's own C++ physics
(oracle/_ref: /root/reference/src compiled against oracle/ref_shim in the build
container; falls back to the bit-identical oracle restatement if that file is
absent) under oracle/nlp.py, the numpy port of the reference's lib/con_*.py
(the reference's Python files themselves do not travel to the GPU box), one
scenario per host core in parallel.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fd_jacobian_plus_residual_evals_per_s"
UNIT = "evals/s"
INPUTS = os.path.join(ROOT, "tests", "golden", "example_inputs.json")
USER_EVENT = "IIP_END"

# algorithmic FP64 work per leaf evaluation (DESIGN.md "roofline"): elementary ops (+ - * / sqrt)
# counted 1 each, sin/cos/exp = 40, atan2/acos/asin = 60, pow = 120
FLOPS = {"air": 1000.0, "noair": 70.0, "quat": 20.0, "aero": 800.0, "evt": 400.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gelato", choices=["gelato", "reference"])
    ap.add_argument("--scenarios", type=int, default=128,
                    help="dispersed scenarios per GPU in one step (128 x 8 GPUs = the 1 024 scenarios of BASELINE.json configs[3])")
    ap.add_argument("--factor", type=int, default=15, help="mesh refinement of the example (15 -> 990 nodes)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "pool", "zero-copy"],
                    help="scattered Jacobian slots of the end-to-end leg: packed + host thread pool, or written by the "
                         "device into the mapped buffer; auto = pool when this rank has 8+ host threads to itself")
    ap.add_argument("--e2e-slices", type=int, default=0, help="pipeline slices of the end-to-end leg (0 = library default)")
    return ap.parse_args()


def load_workload(factor, n_scen_total, first, count, coord=None):
    """Plans and decision vectors of scenarios [first, first+count)."""
    from gelato_b200 import plan as gplan
    from gelato_b200 import problem, scenarios

    inp = problem.load_inputs_json(INPUTS)
    scen = scenarios.disperse(inp, n_scen_total, seed=20260117)
    plans, xs, probs = [], [], []
    for k in range(first, first + count):
        p, u, c, x0 = problem.problem_from_inputs(scen[k], coord=coord, factor=factor, max_nodes=20)
        rng = np.random.default_rng(1000 + k)
        x = {key: v + (0.0 if key == "t" else 1e-3) * rng.uniform(-1.0, 1.0, v.shape) for key, v in x0.items()}
        plans.append(gplan.CompiledPlan(p, u, c, user_eq=gplan.PerigeeAtEvent(USER_EVENT), coord=coord))
        xs.append(problem.xdict_to_vector(x))
        probs.append((p, u, c, x))
    return plans, np.stack(xs), probs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, grid_blocks):
    """DRAM bytes (read + write) per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py), if it was taken on this launch geometry."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(path))[kernel]
        if int(e["grid"].strip("()").split(",")[0]) == int(grid_blocks):
            return e["dram_bytes_per_launch"], "profiles/" + e["source"]
    except (OSError, KeyError, ValueError):
        pass
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# CPU legs (oracle port of the reference's callbacks)
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    factor, n_total, k, reps = args
    from oracle import leaves, nlp, user_builtin

    plans, X, probs = load_workload(factor, n_total, k, 1)
    p, u, c, x = probs[0]
    flav = cpu_flavour()
    L = leaves.get(flav)
    O = nlp.OracleNLP(p, u, c, flav, "numpy", user_eq=user_builtin.perigee_ratio_at(L, USER_EVENT))
    t0 = time.perf_counter()
    for _ in range(reps):
        xa = {key: v.copy() for key, v in x.items()}
        f, _ = O.objfunc(xa)
        O.sens(xa)
    return time.perf_counter() - t0


def cpu_flavour():
    """Physics leaves of the CPU arm: the reference's own C++ (oracle/_ref, built where the
    reference tree is mounted and shipped with the snapshot) when present, else the
    bit-identical hand-written restatement on glibc."""
    from oracle import leaves

    return "ref" if leaves.ref_available() else "libm"


CPU_DESC = {"ref": "the reference's own C++ leaves (oracle/_ref) under oracle/nlp.py, the numpy port of lib/con_*.py",
            "libm": "oracle/nlp.py port on the oracle's libm leaves"}


def cpu_baseline(factor, evals_per_scen, budget_s):
    """One host core, one scenario of the workload, objfunc+sens repeated for ~budget_s."""
    t1 = _cpu_worker((factor, 1, 0, 1))  # also warms imports / builds
    reps = int(max(2, min(50, budget_s / max(t1, 1e-3))))
    dt = _cpu_worker((factor, 1, 0, reps))
    return {"value": evals_per_scen * reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d x (objfunc+sens) of 1 scenario of the workload (N=%d nodes), %s, "
                      "%.2f s per pair" % (reps, 66 * factor, CPU_DESC[cpu_flavour()], dt / reps)}


def run_reference(args):
    """The reference's CPU path on all host cores: one scenario per core per step."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gelato_b200 import plan as gplan  # counts only; nothing is evaluated by the engine here

    cores = os.cpu_count() or 1
    plans, _, _ = load_workload(args.factor, 1, 0, 1)
    ec = plans[0].eval_counts()
    evals_per_scen = ec["objfunc"] + ec["sens"]
    del gplan
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(max(1, min(args.warmup, 1))):
            pool.map(_cpu_worker, [(args.factor, cores, k, 1) for k in range(cores)])
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_cpu_worker, [(args.factor, cores, k, 1) for k in range(cores)])
        dt = time.perf_counter() - t0
    value = evals_per_scen * cores * steps / dt
    sample = ("each step = objfunc+sens of %d dispersed scenarios of the workload (one per host core, spawn pool), "
              "%s; %d timed steps" % (cores, CPU_DESC[cpu_flavour()], steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "example_x%d_N%d_split20" % (args.factor, 66 * args.factor),
                   "scenarios_per_step": cores, "evals_per_scenario_step": evals_per_scen},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gelato(args):
    import torch
    import torch.distributed as dist

    from gelato_b200 import engine
    from gelato_b200 import scenarios as gscen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.scenarios
    own = gscen.partition(B * world, world, rank)
    plans, X, probs = load_workload(args.factor, B * world, own.start, len(own))
    P = plans[0]
    E = engine.Engine(P, device=local, scenario_plans=plans)
    # scatter threads of update mode: the ranks of one node share the host cores (and its memory bandwidth)
    host_threads = max(1, min(16, (os.cpu_count() or 1) // world))
    E.set_host_threads(host_threads)
    zero_copy = args.e2e_mode == "zero-copy" or (args.e2e_mode == "auto" and host_threads < 8)
    E.set_update_zero_copy(zero_copy)
    E.set_update_slices(args.e2e_slices)
    ec = P.eval_counts()
    evals_step_rank = (ec["objfunc"] + ec["sens"]) * B
    # a dedicated non-default stream: the C ABI reads stream 0 as "the plan's own stream", and the
    # CUDA events below must sit on the stream the kernels are launched on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    st = tstream.cuda_stream
    assert st != 0

    xd = torch.from_numpy(X).cuda()
    gd = torch.empty((B, P.n_rows), dtype=torch.float64, device="cuda")
    vd = torch.empty((B, P.n_vals), dtype=torch.float64, device="cuda")
    E.fill_template(vd.data_ptr(), B, st)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(evs=None):
        if evs:
            evs[0].record()
        E.eval_residuals_dev(xd.data_ptr(), gd.data_ptr(), B, st)
        if evs:
            evs[1].record()
        E.eval_jacobian_dev(xd.data_ptr(), vd.data_ptr(), B, st)
        if evs:
            evs[2].record()

    for _ in range(max(args.warmup, 3)):
        step_dev()
        flush.zero_()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = E.launches
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    for k in range(args.steps):
        step_dev(evs[k])
        flush.zero_()  # L2 flush between timed steps (outside the per-step event brackets)
    barrier()
    serial_ms = sum(e[0].elapsed_time(e[2]) for e in evs)
    res_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    jac_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps

    # ---- the headline: objfunc + sens of the same decision vectors as ONE pair evaluation (the residual kernel
    #      on a side stream next to the Jacobian kernel); same work, same results as the two calls above ----
    def step_pair(ev=None):
        if ev:
            ev[0].record()
        E.eval_pair_dev(xd.data_ptr(), gd.data_ptr(), vd.data_ptr(), B, st)
        if ev:
            ev[1].record()

    g_serial, v_serial = gd.clone(), vd.clone()
    for _ in range(max(args.warmup, 3)):
        step_pair()
        flush.zero_()
    evp = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    barrier()
    launches0 = E.launches
    for k in range(args.steps):
        step_pair(evp[k])
        flush.zero_()
    barrier()
    dev_ms = sum(e[0].elapsed_time(e[1]) for e in evp)
    launches = E.launches - launches0
    assert torch.equal(g_serial, gd) and torch.equal(v_serial, vd), "pair and separate evaluations disagree"

    # ---- end to end through the host-buffer C ABI (what objfunc / sens call) ----
    px, pg, pv = engine.PinnedArray(X.size), engine.PinnedArray(B * P.n_rows), engine.PinnedArray(B * P.n_vals)
    px.array[:] = X.ravel()

    def step_e2e_full():
        E.eval_residuals(px.array, B, out=pg.array)
        E.eval_jacobian(px.array, B, out=pv.array)

    def step_e2e_separate():
        # update mode: the batch's host Jacobian buffer lives across calls (as in a batched solve), so
        # only the x-dependent slots cross PCIe; the buffer ends up identical to the full copy
        E.eval_residuals(px.array, B, out=pg.array)
        E.eval_jacobian_update(px.array, pv.array, B)

    def step_e2e():
        # the same as one pair call: x uploaded once, the residual kernel and its copy overlap the Jacobian's
        E.eval_pair_update(px.array, pg.array, pv.array, B)

    def timed(step):
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        barrier()
        return time.perf_counter() - t0

    e2e_full_s = timed(step_e2e_full)
    assert np.array_equal(pv.array.reshape(B, -1), vd.cpu().numpy()), "host and device paths disagree"
    pv.array[:] = np.nan
    E.jacobian_template(pv.array, B)
    e2e_sep_s = timed(step_e2e_separate)
    pg.array[:] = np.nan
    e2e_s = timed(step_e2e)
    clocks = sampler.stop()
    assert np.array_equal(pg.array.reshape(B, -1), gd.cpu().numpy()), "host and device paths disagree"
    assert np.array_equal(pv.array.reshape(B, -1), vd.cpu().numpy()), "host and device paths disagree"

    t = torch.tensor([dev_ms, e2e_s * 1e3, jac_ms, res_ms, e2e_full_s * 1e3, serial_ms, e2e_sep_s * 1e3], dtype=torch.float64,
                     device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, jac_ms, res_ms, e2e_full_ms, serial_ms, e2e_sep_ms = [float(v) for v in t.cpu()]

    if rank == 0:
        peak, peak_src = measured_peaks()
        n_xdep = int(P.n_xdep)
        jac_bytes = B * (P.n_vars + n_xdep) * 8.0
        n_hold = P.N - ec["free_nodes"]
        del n_hold
        flops_jac = B * (FLOPS["air"] * 14 * ec["air_fd_nodes"] + FLOPS["noair"] * 9 * (P.N - ec["air_fd_nodes"])
                         + FLOPS["quat"] * 7 * ec["free_nodes"] + FLOPS["aero"] * ec["aero_jac_evals"]
                         + FLOPS["evt"] * ec["evt_jac_evals"])
        traffic, traffic_src = ncu_traffic("k_jacobian", E.n_jac_blocks * B)
        try:
            fma_tf, nofma_tf = engine.fp64_peak(local)
        except Exception:
            fma_tf = nofma_tf = None
        line = {
            "metric": METRIC, "value": evals_step_rank * world * args.steps / (dev_ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "example_x%d_N%d_split20" % (args.factor, P.N), "scenarios_per_gpu": B,
                       "nodes": P.N, "sections": P.S, "n_vars": P.n_vars, "n_rows": P.n_rows, "n_vals": int(P.n_vals),
                       "evals_per_scenario_step": ec["objfunc"] + ec["sens"], "parallelism": "scenarios x%d" % world,
                       "step": "objfunc + sens of one batch as one pair evaluation (gelato_eval_pair_dev: residual "
                               "kernel on a side stream next to the Jacobian kernel)",
                       "l2": "flushed between timed steps (256 MiB write); per-step CUDA events on the launch stream"},
            "clocks": clocks,
            "e2e": {"value": evals_step_rank * world * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(X.size * 8),
                    "d2h_bytes_per_step": int(B * (P.n_rows + n_xdep) * 8),
                    "mode": "page-locked host Jacobian buffer kept across calls; only the %d x-dependent of %d slots per "
                            "scenario cross PCIe (long runs by strided copies into place, scattered slots %s); pipelined "
                            "over slices of 16 scenarios"
                            % (n_xdep, int(P.n_vals), "written by the device into the mapped buffer (zero-copy)" if zero_copy
                               else "packed and scattered by %d pooled host threads" % host_threads),
                    "separate_calls": {"value": evals_step_rank * world * args.steps / (e2e_sep_ms * 1e-3),
                                       "ms_per_step": e2e_sep_ms / args.steps},
                    "full_copy": {"value": evals_step_rank * world * args.steps / (e2e_full_ms * 1e-3),
                                  "ms_per_step": e2e_full_ms / args.steps,
                                  "d2h_bytes_per_step": int(B * (P.n_rows + P.n_vals) * 8)}},
            "gpu_launches": int(launches),
            "kernels": {"k_residuals_ms": res_ms, "k_jacobian_ms": jac_ms,
                        "note": "each kernel timed alone, back to back on one stream",
                        "separate_calls_ms_per_step": serial_ms / args.steps,
                        "separate_calls_value": evals_step_rank * world * args.steps / (serial_ms * 1e-3)},
            "roofline": {"kernel": "k_jacobian", "bound": "hbm", "achieved": jac_bytes / (jac_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": jac_bytes / (jac_ms * 1e-3) / 1e9 / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": jac_bytes,
                         "note": "the kernel is FP64-ALU bound, not HBM bound (DESIGN.md section 5): `fp64` is the "
                                 "binding roofline, against the measured DMUL+DADD issue peak (fused multiply-add is "
                                 "off by the bit-parity contract)",
                         "fp64": {"achieved_tflops": flops_jac / (jac_ms * 1e-3) / 1e12,
                                  "peak_tflops_dfma": fma_tf, "peak_tflops_dmul_dadd": nofma_tf,
                                  "frac_of_unfused_peak": (flops_jac / (jac_ms * 1e-3) / 1e12 / nofma_tf) if nofma_tf else None,
                                  "algorithmic_flops_per_launch": flops_jac}},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.factor, ec["objfunc"] + ec["sens"], args.cpu_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gelato(args)


if __name__ == "__main__":
    main()
