"""gelato_b200.server: concurrent solver workers, one coalesced launch per callback kind.
CPU tier: the host emulator of the kernels stands in for the engine; GPU tier: the CUDA engine."""
import threading

import numpy as np
import pytest

import helpers
from gelato_b200 import problem, scenarios, server
from oracle import leaves


def _emu_factory(base, plans):
    import emu_binding

    return emu_binding.EmuEngine(base, scenario_plans=plans)


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def setup(request):
    Lg = leaves.get("gmath")
    scen = scenarios.disperse(helpers.example_inputs(), 5, seed=77)
    plans, xs, oracles = [], [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        xs.append(x0)
        oracles.append(helpers.oracle_nlp(p, u, c, "gmath", "seqfma"))
    srv = server.CoalescingServer(plans, engine_factory=_emu_factory if request.param == "emu" else None, max_wait_s=0.05)
    yield srv, plans, xs, oracles
    srv.close()


def test_concurrent_workers_get_their_own_scenario_bitwise_and_launches_are_coalesced(setup):
    srv, plans, xs, oracles = setup
    n_iter = 4
    got = [[] for _ in plans]
    barrier = threading.Barrier(len(plans))

    def worker(k):
        cb = srv.client(k)
        x = helpers.perturbed(xs[k], seed=k)
        barrier.wait()
        for it in range(n_iter):
            f, fail = cb.objfunc(x)
            s, _ = cb.sens(x, f)
            got[k].append((helpers.copy_x(x), f, s))
            x = {n: v + 1e-4 * (it + 1) * np.cos(np.arange(v.size) + k) for n, v in x.items()}  # next "iterate"
        cb.done()

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(plans))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
        assert not t.is_alive()
    for k in range(len(plans)):
        for x, f, s in got[k]:
            xa = helpers.copy_x(x)
            fo, _ = oracles[k].objfunc(xa)
            so, _ = oracles[k].sens(xa)
            helpers.assert_funcs_equal(fo, f)
            helpers.assert_sens_equal(so, s)
    calls = 2 * n_iter * len(plans)
    assert srv.stats["calls"] == calls
    assert srv.stats["launches"] < calls and srv.stats["largest_batch"] >= 2, srv.stats


def test_subset_batches_with_repeated_and_reordered_scenarios(setup):
    srv, plans, xs, oracles = setup
    E = srv.engine
    ids = [3, 0, 3, 4]
    X = np.stack([problem.xdict_to_vector(helpers.perturbed(xs[k], seed=10 + i)) for i, k in enumerate(ids)])
    G = np.asarray(E.eval_residuals(X, n_scen=4, scen_ids=ids)).reshape(4, -1)
    V = np.asarray(E.eval_jacobian(X, n_scen=4, scen_ids=ids)).reshape(4, -1)
    for i, k in enumerate(ids):
        x = problem.vector_to_xdict(X[i].copy(), plans[k].M, plans[k].N, plans[k].S)
        fo, _ = oracles[k].objfunc(helpers.copy_x(x))
        so, _ = oracles[k].sens(helpers.copy_x(x))
        helpers.assert_funcs_equal(fo, plans[k].split_residuals(G[i]))
        helpers.assert_sens_equal(so, plans[k].split_jacobian(V[i], key_order=list(x.keys())))
