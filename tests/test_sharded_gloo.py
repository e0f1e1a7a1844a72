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


# ---- ONE problem sharded over ranks: block ranges of the pair kernel + one all_gather ----
class EmuRangeEvaluator:
    """The host emulator behind batch.ShardedProblem (CPU tensors; same interface as batch.EngineRangeEvaluator)."""

    def __init__(self, emu):
        import torch

        self.emu, self.device = emu, torch.device("cpu")
        self.n_rows, self.n_pack, self.n_vars = emu.plan.n_rows, emu.n_pack, emu.plan.n_vars
        self.n_blocks, self.n_vac = emu.block_counts()

    def pair_range(self, x, g, packed, blocks, vacuum):
        self.emu.eval_pair_range(x.numpy(), g.numpy(), packed.numpy(), blocks, vacuum)


def _one_problem(variant):
    import emu_binding
    from gelato_b200 import problem
    from oracle import leaves

    Lg = leaves.get("gmath")
    p, u, c, x0 = problem.problem_from_inputs(helpers.variant_inputs(variant), coord=Lg.coordinate_c, factor=2, max_nodes=12)
    P = helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c)
    return emu_binding.Emulator(P), problem.xdict_to_vector(helpers.perturbed(x0))


@pytest.mark.parametrize("variant", ["example", "three_stage"])
def test_block_ranges_partition_the_pair_evaluation(variant):
    """Any cut of the block table and of the vacuum-node list into ranges gives, together, the whole evaluation."""
    emu, x = _one_problem(variant)
    g_want, pk_want = emu.eval_pair(x, packed=True)
    nb, nv = emu.block_counts()
    assert nb > 3
    for cuts in (2, 3, 5):
        g = np.full(emu.plan.n_rows, np.nan)
        pk = np.full(emu.n_pack, np.nan)
        written = np.zeros(g.size + pk.size, dtype=int)
        for r in range(cuts):
            from gelato_b200.batch import split_range

            g1, p1 = np.full(g.size, np.nan), np.full(pk.size, np.nan)
            emu.eval_pair_range(x, g1, p1, split_range(nb, cuts, r), split_range(nv, cuts, r))
            m = ~np.isnan(np.concatenate([g1, p1]))
            written += m
            g[m[: g.size]] = g1[m[: g.size]]
            pk[m[g.size:]] = p1[m[g.size:]]
        assert (written == 1).all()
        assert np.array_equal(g, g_want) and np.array_equal(pk, pk_want)


def _problem_worker(rank, world, port, variant, out_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from gelato_b200 import batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu, x = _one_problem(variant)
    sp = batch.ShardedProblem(EmuRangeEvaluator(emu), x, rank=rank, world_size=world)
    assert sum(sp.sizes) == sp.n_out and 0 < sp.sizes[rank] < sp.n_out
    for k in range(2):  # the second call at another point: the parts found at the probe point hold
        g, pk = sp.pair(x * (1.0 + 1e-3 * k))
        np.save(os.path.join(out_dir, "g%d_%d.npy" % (k, rank)), g.numpy())
        np.save(os.path.join(out_dir, "pk%d_%d.npy" % (k, rank)), pk.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_one_problem_over_two_ranks_equals_one_rank(tmp_path):
    import torch.multiprocessing as mp

    variant = "example"
    mp.spawn(_problem_worker, args=(2, _free_port(), variant, str(tmp_path)), nprocs=2, join=True)
    emu, x = _one_problem(variant)
    for k in range(2):
        g_want, pk_want = emu.eval_pair(x * (1.0 + 1e-3 * k), packed=True)
        for r in range(2):
            assert np.array_equal(np.load(os.path.join(str(tmp_path), "g%d_%d.npy" % (k, r))), g_want)
            assert np.array_equal(np.load(os.path.join(str(tmp_path), "pk%d_%d.npy" % (k, r))), pk_want)


def test_sharded_problem_refuses_parts_that_do_not_cover_the_outputs():
    """The parts are found by evaluation and CHECKED: an evaluator whose ranges leave output slots unwritten (here:
    the last block dropped) must be refused at construction, not produce a vector with holes."""
    from gelato_b200 import batch

    emu, x = _one_problem("example")
    ev = EmuRangeEvaluator(emu)
    ev.n_blocks -= 1  # a rank layout that forgets one block
    with pytest.raises(RuntimeError, match="written by no rank"):
        batch.ShardedProblem(ev, x)
    ok = batch.ShardedProblem(EmuRangeEvaluator(emu), x)  # the full range on one rank: the whole evaluation
    g, pk = ok.pair(x)
    g_want, pk_want = emu.eval_pair(x, packed=True)
    assert np.array_equal(g.numpy(), g_want) and np.array_equal(pk.numpy(), pk_want)
    assert [batch.split_range(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
