#!/usr/bin/env python
"""Where the end-to-end time of one batched step goes (run on the GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gelato_b200 import engine  # noqa: E402


def t(fn, reps=10):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    plans, X, _ = bench.load_workload(15, B, 0, B)
    P = plans[0]
    E = engine.Engine(P, scenario_plans=plans)
    E.set_update_slices(1)
    px, pg, pv = engine.PinnedArray(X.size), engine.PinnedArray(B * P.n_rows), engine.PinnedArray(B * P.n_vals)
    px.array[:] = X.ravel()
    E.jacobian_template(pv.array, B)
    print("B=%d n_vals=%d n_xdep=%d cores=%d" % (B, P.n_vals, P.n_xdep, os.cpu_count()))
    print("residuals host call      %.3f ms" % t(lambda: E.eval_residuals(px.array, B, out=pg.array)))
    print("jacobian full copy       %.3f ms" % t(lambda: E.eval_jacobian(px.array, B, out=pv.array)))
    pageable = np.empty(B * P.n_vals)
    E.jacobian_template(pageable, B)
    for th in (4, 16):
        E.set_host_threads(th)
        print("update, pageable buffer (pack all + host scatter), %2d thr   %.3f ms"
              % (th, t(lambda: E.eval_jacobian_update(px.array, pageable, B))))
    E.set_update_zero_copy(False)
    for th in (1, 4, 8, 16):
        E.set_host_threads(th)
        print("update, pinned: 2-D copies + packed scattered slots, %2d thr  %.3f ms"
              % (th, t(lambda: E.eval_jacobian_update(px.array, pv.array, B))))
    E.set_update_zero_copy(True)
    E.set_update_slices(1)
    print("update, pinned: 2-D copies + zero-copy scattered slots       %.3f ms"
          % t(lambda: E.eval_jacobian_update(px.array, pv.array, B)))
    want = E.eval_jacobian(px.array, B).copy()
    assert np.array_equal(pv.array.reshape(B, -1), want), "zero-copy update differs from the full copy"
    print("transfer pieces alone (device ms):", {k: round(v, 3) for k, v in E.probe_update(pv.array, B).items()})
    print("pair update, zero-copy   %.3f ms" % t(lambda: E.eval_pair_update(px.array, pg.array, pv.array, B)))
    E.set_update_zero_copy(False)
    E.set_update_slices(1)
    for th in (4, 8, 16, 32):
        E.set_host_threads(th)
        print("pair update, packed + pool of %2d threads, 1 slice   %.3f ms"
              % (th, t(lambda: E.eval_pair_update(px.array, pg.array, pv.array, B), 20)))
    for sl in (2, 3, 4, 6, 8):
        E.set_update_slices(sl)
        for th in (8, 16):
            E.set_host_threads(th)
            print("pair update, packed + pool of %2d threads, %d slices  %.3f ms"
                  % (th, sl, t(lambda: E.eval_pair_update(px.array, pg.array, pv.array, B), 20)))
    E.set_update_zero_copy(True)
    for sl in (1, 2, 4):
        E.set_update_slices(sl)
        print("pair update, zero-copy, %d slices  %.3f ms" % (sl, t(lambda: E.eval_pair_update(px.array, pg.array, pv.array, B), 20)))
    E.set_update_zero_copy(False)
    E.set_update_slices(0)
    E.set_host_threads(0)
    print("pair update, defaults    %.3f ms" % t(lambda: E.eval_pair_update(px.array, pg.array, pv.array, B), 20))
    assert np.array_equal(pv.array.reshape(B, -1), want), "packed update differs from the full copy"
    E.set_update_zero_copy(False)
    idx = P.xdep_index()
    packed = np.random.rand(B, idx.size)
    big = pv.array.reshape(B, -1)

    def np_scatter():
        big[:, idx] = packed
    print("numpy fancy scatter      %.3f ms" % t(np_scatter, 3))


if __name__ == "__main__":
    main()
