#!/usr/bin/env python
"""How the cache state left by the L2 flush changes the event-timed duration of k_jacobian (measurement only; GPU box).
Modes: write (256 MiB memset: L2 left full of dirty lines), write+read (memset, then a 256 MiB read: L2 cold AND clean),
read (a 256 MiB read only), none, rotate (three input/output sets used in turn, no flush: working set 270 MB > L2)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from gelato_b200 import engine

    B = 128
    plans, X, probs = bench.load_workload("example", 15, B, 0, B)
    P = plans[0]
    E = engine.Engine(P, scenario_plans=plans)
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    st = ts.cuda_stream
    sets = []
    for _ in range(3):
        sets.append((torch.from_numpy(X).cuda(), torch.empty((B, P.n_rows), dtype=torch.float64, device="cuda"),
                     torch.empty((B, E.n_pack), dtype=torch.float64, device="cuda")))
    wbuf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rbuf = torch.zeros(32 * 1024 * 1024, dtype=torch.float64, device="cuda")
    tiny = torch.zeros(32, device="cuda")

    def flush(mode):
        if mode in ("write", "write+read"):
            wbuf.zero_()
        if mode in ("read", "write+read"):
            rbuf.sum()

    def timed(fn, mode, reps=30):
        for k in range(4):
            fn(k)
            flush(mode)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(evs):
            a.record()
            fn(k)
            b.record()
            flush(mode)
        torch.cuda.synchronize()
        ts_ = sorted(a.elapsed_time(b) for a, b in evs)
        return {"mean_ms": sum(ts_) / reps, "median_ms": ts_[reps // 2], "min_ms": ts_[0]}

    def kern(which, with_g):
        def fn(k, rot=False):
            x, g, p = sets[k % 3 if rot else 0]
            E.launch_kernel_dev(which, x.data_ptr(), p.data_ptr(), B, True, st, g.data_ptr() if with_g else None)
        return fn

    out = {"event_pair_around_a_tiny_kernel": timed(lambda k: tiny.zero_(), "none")}
    blk = kern(6, True)
    for mode in ("write", "write+read", "read", "none"):
        out["k_jacobian/" + mode] = timed(blk, mode)
    out["k_jacobian/rotate"] = timed(lambda k: blk(k, True), "none")
    pair = lambda k, rot=False: E.eval_pair_packed_dev(sets[k % 3 if rot else 0][0].data_ptr(), sets[k % 3 if rot else 0][1].data_ptr(),  # noqa: E731
                                                       sets[k % 3 if rot else 0][2].data_ptr(), B, st)
    for mode in ("write", "write+read", "read", "none"):
        out["pair/" + mode] = timed(pair, mode)
    out["pair/rotate"] = timed(lambda k: pair(k, True), "none")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
