#!/usr/bin/env python
"""SURVEY.md 8(d), last item of "Reference CPU path timed beside it": the leaf-only rate of
`dynamics_velocity` on ONE batch of 1e5 nodes, so that the Python overhead of the reference's constraint
layer and the cost of the physics are separated.  CPU: the reference's own C++ leaf (oracle/_ref, one core,
as the reference runs it) and the oracle's restatement; GPU: the batch leaf kernel behind
gelato_b200.lib.dynamics_c (host buffers in and out, and the kernel alone).  Run on the GPU box:
    python tests/scripts/leaf_rate.py > gpurun_out/leaf_rate.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import leaves  # noqa: E402


def states(n, seed):
    rng = np.random.default_rng(seed)
    lat, lon = rng.uniform(-1.4, 1.4, n), rng.uniform(-np.pi, np.pi, n)
    r = 6378137.0 + rng.uniform(0.0, 90e3, n)
    pos = np.stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)], axis=1)
    vel = rng.normal(size=(n, 3)) * rng.uniform(1.0, 8000.0, (n, 1))
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return pos, vel, q, rng.uniform(0.0, 900.0, n)


def best(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    inp = helpers.example_inputs()
    wind, ca = np.asarray(inp["wind_table"], dtype=float), np.asarray(inp["ca_table"], dtype=float)
    pos, vel, q, t = states(n, 1)
    units = np.array([27442.0, 6378137.0, 1000.0])
    mass = np.random.default_rng(2).uniform(0.05, 1.0, n)
    param = np.array([420000.0, 140.9, 2.21, 0.0, 0.68])
    args = (mass, pos / units[1], vel / units[2], q, t, param, wind, ca, units)
    print("dynamics_velocity, one call on %d nodes (air path, pybind_dynamics.cpp:30-71)" % n)
    flavours = (["ref"] if leaves.ref_available() else []) + ["libm", "gmath"]
    outs = {}
    for fl in flavours:
        O = leaves.get(fl).dynamics_c
        dt = best(lambda: outs.__setitem__(fl, O.dynamics_velocity(*args)), 3)
        what = {"ref": "reference C++ (oracle/_ref), 1 core", "libm": "oracle restatement, libm, 1 core",
                "gmath": "oracle restatement, gmath, 1 core"}[fl]
        print("  CPU %-40s %8.1f ms  %7.3f M evals/s" % (what, dt * 1e3, n / dt / 1e6))
    try:
        import torch
        from gelato_b200.lib import dynamics_c
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device")
    except Exception as e:  # the CPU half is still worth printing
        print("  GPU skipped:", e)
        return
    got = dynamics_c.dynamics_velocity(*args)
    assert np.array_equal(got, outs["gmath"]), "GPU leaf differs from the gmath oracle"
    dt = best(lambda: dynamics_c.dynamics_velocity(*args), 10)
    print("  GPU %-40s %8.3f ms  %7.1f M evals/s" % ("gelato_leaf_dynamics_velocity, host buffers", dt * 1e3, n / dt / 1e6))
    # the fused Jacobian kernel for comparison: evaluations per second of the bench workload are in bench.json
    print("  (bit-identical to the gmath oracle; the fused kernels' rate is `value` of bench.py)")


if __name__ == "__main__":
    main()
