#!/usr/bin/env python
"""Device times of the pieces of a pair evaluation at the bench workload (measurement only; run on the GPU box):
each kernel alone (with and without the defect rows), the Jacobian evaluation, the pair, CUDA events, L2 flushed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from gelato_b200 import engine

    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    plans, X, probs = bench.load_workload("example", 15, B, 0, B)
    P = plans[0]
    E = engine.Engine(P, scenario_plans=plans)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    xd = torch.from_numpy(X).cuda()
    gd = torch.empty((B, P.n_rows), dtype=torch.float64, device="cuda")
    pd = torch.empty((B, E.n_pack), dtype=torch.float64, device="cuda")
    vd = torch.empty((B, P.n_vals), dtype=torch.float64, device="cuda")
    E.fill_template(vd.data_ptr(), B, st)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=20, do_flush=True):
        for _ in range(3):
            fn()
            flush.zero_()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for a, b in evs:
            a.record()
            fn()
            b.record()
            if do_flush:
                flush.zero_()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / reps

    x, g, p_, v = xd.data_ptr(), gd.data_ptr(), pd.data_ptr(), vd.data_ptr()
    out = {
        "scenarios": B, "blocks_heavy": E.n_jac_heavy, "blocks_light": E.n_jac_light,
        "heavy_packed": timed(lambda: E.launch_kernel_dev(2, x, p_, B, True, st)),
        "heavy_packed_g": timed(lambda: E.launch_kernel_dev(2, x, p_, B, True, st, g)),
        "heavy_coo": timed(lambda: E.launch_kernel_dev(2, x, v, B, False, st)),
        "light_packed": timed(lambda: E.launch_kernel_dev(3, x, p_, B, True, st)),
        "light_packed_g": timed(lambda: E.launch_kernel_dev(3, x, p_, B, True, st, g)),
        "jac_packed": timed(lambda: E.launch_kernel_dev(1, x, p_, B, True, st)),
        "blocks_only": timed(lambda: E.launch_kernel_dev(6, x, p_, B, True, st)),
        "blocks_only_g": timed(lambda: E.launch_kernel_dev(6, x, p_, B, True, st, g)),
        "noair": timed(lambda: E.launch_kernel_dev(5, x, p_, B, True, st)),
        "noair_g": timed(lambda: E.launch_kernel_dev(5, x, p_, B, True, st, g)),
        "res_full": timed(lambda: E.launch_kernel_dev(0, x, g, B, False, st)),
        "serial_heavy_light_g": timed(lambda: (E.launch_kernel_dev(2, x, p_, B, True, st, g), E.launch_kernel_dev(3, x, p_, B, True, st, g))),
        "jacobian_coo": timed(lambda: E.eval_jacobian_dev(x, v, B, st)),
        "pair_packed": timed(lambda: E.eval_pair_packed_dev(x, g, p_, B, st)),
        "pair_coo": timed(lambda: E.eval_pair_dev(x, g, v, B, st)),
        "pair_packed_noflush": timed(lambda: E.eval_pair_packed_dev(x, g, p_, B, st), do_flush=False),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
