"""world_size-2 test of the multi-GPU path's host logic on CPU (gloo): scenario
partition, per-rank batched evaluation (the host emulator stands in for the
engine) and the final gather.  The sharded result must equal the single-process
result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers
from gelato_b200 import scenarios


def test_partition_covers_all_scenarios():
    for n, w in ((1024, 8), (10, 4), (3, 8), (7, 2)):
        parts = [scenarios.partition(n, w, r) for r in range(w)]
        assert [i for p in parts for i in p] == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_scen, out_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    import emu_binding
    from gelato_b200 import batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sb = batch.ShardedBatch(helpers.example_inputs(), n_scen, rank=rank, world_size=world, user_event=helpers.USER_EVENT,
                            engine_factory=lambda base, plans: emu_binding.Emulator(base, scenario_plans=plans))
    X = np.stack(sb.x0)
    allrows = sb.gather(sb.summaries(X))
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), allrows)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    import emu_binding
    from gelato_b200 import batch

    n_scen = 5  # uneven split: 3 + 2
    mp.spawn(_worker, args=(2, _free_port(), n_scen, str(tmp_path)), nprocs=2, join=True)
    sb = batch.ShardedBatch(helpers.example_inputs(), n_scen, user_event=helpers.USER_EVENT,
                            engine_factory=lambda base, plans: emu_binding.Emulator(base, scenario_plans=plans))
    want = sb.summaries(np.stack(sb.x0))
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == (n_scen, 3)
        assert np.array_equal(got, want)
    assert len(set(want[:, 1].tolist())) == n_scen  # the scenarios really differ
