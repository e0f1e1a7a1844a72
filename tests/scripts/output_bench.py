#!/usr/bin/env python
"""Result tables of many solved scenarios: one gelato_leaf_output_table launch vs the CPU oracle's C++ loop.
(The reference's own output_result, a Python loop with ~30 leaf calls per row, took 40 ms per 78-row table in
the build container; it cannot travel to the GPU box.)   python tests/scripts/output_bench.py [n_scenarios]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from gelato_b200 import output as gout  # noqa: E402
from oracle import output as oout  # noqa: E402
from test_output_table import _times  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    p, u, c, x0 = helpers.example_problem()
    sols = []
    for k in range(n):
        x = helpers.perturbed(x0, seed=k)
        tx, tu = _times(x, p, u)
        sols.append((x, u, tx, tu, p))
    gout.output_tables(sols[:2])
    t0 = time.perf_counter()
    tabs = gout.output_tables(sols)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    for s in sols[:64]:
        oout.output_result(*s, "libm")
    t_cpu = (time.perf_counter() - t0) / 64 * n
    print(json.dumps({"scenarios": n, "rows": int(sum(len(s[2]) for s in sols)), "gpu_one_launch_s": t_gpu,
                      "cpu_oracle_cpp_loop_s": t_cpu, "reference_python_loop_s_estimate": 0.040 * n,
                      "finite": bool(np.isfinite(tabs[-1]["M"]).all())}))


if __name__ == "__main__":
    main()
