#!/usr/bin/env python
"""Cycles per (role, warp) and phase of k_jacobian from a -DGJ_CLOCKS measurement build (GPU box):
   python -c "from gelato_b200 import engine; engine.build_library(out='/tmp/libclk.so', extra=['-DGJ_CLOCKS'])"
   GELATO_B200_LIB=/tmp/libclk.so python tools/phase_clocks.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ROLES = {0: "DYN_AIR", 1: "DYN_NOAIR", 2: "DYN_GEN", 3: "AERO", 4: "EVT", 5: "LIN"}


def main():
    import torch

    from gelato_b200 import engine

    B = 128
    plans, X, probs = bench.load_workload("example", 15, B, 0, B)
    P = plans[0]
    E = engine.Engine(P, scenario_plans=plans)
    L = engine.load_library()
    st = torch.cuda.current_stream().cuda_stream
    xd = torch.from_numpy(X).cuda()
    gd = torch.empty((B, P.n_rows), dtype=torch.float64, device="cuda")
    pd = torch.empty((B, E.n_pack), dtype=torch.float64, device="cuda")
    buf = np.zeros((8, 16, 6), dtype=np.uint64)
    ptr = buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong))
    for _ in range(3):
        E.launch_kernel_dev(6, xd.data_ptr(), pd.data_ptr(), B, True, st, gd.data_ptr())
    L.gelato_debug_clocks(ptr)
    reps = 5
    for _ in range(reps):
        E.launch_kernel_dev(6, xd.data_ptr(), pd.data_ptr(), B, True, st, gd.data_ptr())
    assert L.gelato_debug_clocks(ptr) == 0
    print("mean cycles per block, by role and warp: phase0 | barrier wait | phase2 | barrier wait | phase3   (blocks per launch)")
    for r in range(8):
        if buf[r, :, 5].sum() == 0:
            continue
        print("role", r, ROLES.get(r, "?"))
        for w in range(16):
            n = buf[r, w, 5]
            if n == 0:
                continue
            c = buf[r, w, :5] / n
            print("  warp %2d: %8.0f | %8.0f | %8.0f | %8.0f | %8.0f   total %8.0f  (%d)" % (
                w, c[0], c[1], c[2], c[3], c[4], c.sum(), n // reps))


if __name__ == "__main__":
    main()
