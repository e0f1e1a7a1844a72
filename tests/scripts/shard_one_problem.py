#!/usr/bin/env python
"""ONE large NLP over N GPUs: `torchrun --nproc-per-node N tests/scripts/shard_one_problem.py
[--factor F]`.  Every rank evaluates its share of the pair kernel's blocks, one NCCL all_gather puts the whole
(g, packed) on every rank; checked bit for bit against the single-GPU evaluation on rank 0 and timed with CUDA
events (max over ranks).  One JSON line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gelato_b200 import batch, engine  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--factor", type=int, default=1500)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    plans, X, _ = bench.load_workload("example", a.factor, 1, 0, 1)
    P = plans[0]
    E = engine.Engine(P, device=local)
    x = np.ascontiguousarray(X[0])
    sp = batch.ShardedProblem(batch.EngineRangeEvaluator(E), x, rank=rank, world_size=world)
    xd = torch.from_numpy(x).cuda()
    g, pk = sp.pair(xd)
    torch.cuda.synchronize()
    g_want, pk_want = E.eval_pair_packed(x, 1)
    same = bool(np.array_equal(g.cpu().numpy(), g_want.ravel()) and np.array_equal(pk.cpu().numpy(), pk_want.ravel()))

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.cpu())

    ws = torch.empty(E.n_rows + E.n_pack, dtype=torch.float64, device="cuda")
    ev = sp.ev

    def one_gpu():
        cur = torch.cuda.current_stream()
        ev.stream.wait_stream(cur)
        E.eval_pair_packed_dev(xd.data_ptr(), ws.data_ptr(), ws[E.n_rows:].data_ptr(), 1, ev.stream.cuda_stream)
        cur.wait_stream(ev.stream)

    whole = timed(one_gpu)
    share = timed(lambda: sp.ev.pair_range(sp.x, sp.buf[: sp.n_rows], sp.buf[sp.n_rows:], sp.blocks, sp.vacuum))
    sharded = timed(lambda: sp.pair(xd))
    ok = torch.tensor([1 if same else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "one NLP sharded over GPUs by block ranges + one all_gather", "n_gpus": world,
                          "nodes": int(P.N), "n_vars": int(P.n_vars), "n_rows": int(P.n_rows), "n_pack": int(E.n_pack),
                          "blocks": int(sp.ev.n_blocks), "vacuum_nodes": int(sp.ev.n_vac),
                          "bit_identical_on_every_rank": bool(int(ok.cpu())),
                          "ms_one_gpu_whole": whole, "ms_own_share_kernel_only": share, "ms_sharded_pair": sharded,
                          "gathered_bytes_per_rank": int(sp.pad * 8 * world),
                          "speedup_vs_one_gpu": whole / sharded}))
    E.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
