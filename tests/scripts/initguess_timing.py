"""Time the batched forward-simulation initial guess (gelato_init_rocket_simulation) at the reference's step
(dt = 0.005 s over the shipped example's 630 s schedule) and the oracle's Python restatement of the reference loop
on a bounded span, on the same machine.  Prints one JSON line.  Usage: python tests/scripts/initguess_timing.py [n ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import helpers  # noqa: E402
from gelato_b200 import initialize  # noqa: E402
from oracle import initguess, leaves  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1, 4736]
    p, u, c, _ = helpers.example_problem()
    x_init = np.concatenate(([c["init"]["mass"]], c["init"]["position"], c["init"]["velocity"], c["init"]["quaternion"]))
    t_nodes, t_x = initialize.mesh_times(p)
    _, u_table = initialize.rate_table(p, t_nodes)
    dt = 0.005
    out = {"dt": dt, "t_final": float(t_x[-1]), "steps": int(np.ceil((t_x[-1] - t_nodes[0]) / dt)), "gpu": []}
    initialize.rocket_simulation(x_init, u_table, p, t_nodes[0], t_x[:5], 0.5)  # context, module load
    for n in sizes:
        x0 = np.stack([x_init * np.r_[1.0 + 1e-4 * (k % 50), np.ones(10)] for k in range(n)])
        t0 = time.perf_counter()
        x_out, _ = initialize.rocket_simulation_batch(x0, u_table, [p] * n, t_nodes[0], t_x, dt)
        el = time.perf_counter() - t0
        out["gpu"].append({"scenarios": n, "seconds": el, "scenarios_per_s": n / el,
                           "final_mass": float(x_out[0, -1, 0]), "finite": bool(np.isfinite(x_out).all())})
    # the reference's loop (the oracle restates it line by line on the reference-faithful leaves): a bounded span
    F = initguess.ForwardSimulation(leaves.get("libm"), "numpy")
    span = 2.0
    t0 = time.perf_counter()
    F.rocket_simulation(x_init.copy(), u_table, p, t_nodes[0], np.array([t_nodes[0] + span]), dt)
    el = time.perf_counter() - t0
    out["cpu_python_loop"] = {"simulated_s": span, "seconds": el,
                              "whole_schedule_s_extrapolated": el * (t_x[-1] - t_nodes[0]) / span, "cores": 1}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
