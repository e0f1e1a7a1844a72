#!/usr/bin/env python
"""Per-rank breakdown of the end-to-end step when N ranks run at once (torchrun, one rank per GPU):
which host-side resource the ranks share.  Each line is the MAX over ranks of the mean call time."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gelato_b200 import engine  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = 64
    plans, X, _ = bench.load_workload(15, B * world, rank * B, B)
    P = plans[0]
    E = engine.Engine(P, device=local, scenario_plans=plans)
    px, pg, pv = engine.PinnedArray(X.size), engine.PinnedArray(B * P.n_rows), engine.PinnedArray(B * P.n_vals)
    px.array[:] = X.ravel()
    E.jacobian_template(pv.array, B)
    hbuf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    dbuf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def t(fn, reps=10):
        fn()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps * 1e3
        v = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.cpu())

    rows = [
        ("H2D 64 MiB pinned copy", lambda: dbuf.copy_(hbuf, non_blocking=True)),
        ("D2H 64 MiB pinned copy", lambda: hbuf.copy_(dbuf, non_blocking=True)),
        ("residuals host call", lambda: E.eval_residuals(px.array, B, out=pg.array)),
        ("jacobian update, zero-copy", lambda: E.eval_jacobian_update(px.array, pv.array, B)),
    ]
    out = [(name, t(fn)) for name, fn in rows]
    E.set_update_zero_copy(False)
    E.set_host_threads(max(1, min(16, (os.cpu_count() or 1) // world)))
    out.append(("jacobian update, host scatter", t(lambda: E.eval_jacobian_update(px.array, pv.array, B))))
    out.append(("jacobian full copy", t(lambda: E.eval_jacobian(px.array, B, out=pv.array), 5)))
    if rank == 0:
        print("ranks=%d cores=%d" % (world, os.cpu_count()))
        for name, ms in out:
            print("  %-32s %.3f ms" % (name, ms))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
