"""Scenario-batched evaluation, sharded over the GPUs of one node.

Dispersed scenarios are independent NLPs: rank r owns a
contiguous block of them, evaluates the block as ONE kernel launch per callback
(scenario index on blockIdx.y) and nothing is exchanged while solving.  The only
collective is the final gather of per-scenario summaries (objective, worst
equality defect, worst inequality violation), a few doubles per scenario.

One process per GPU, `torch.distributed` for the plumbing (NCCL on GPUs; the
CPU test tier drives the same code over gloo with the host emulator standing in
for the engine).
"""
import numpy as np

from . import plan as gplan
from . import problem, scenarios


def default_engine_factory(device):
    from . import engine

    def make(base_plan, plans):
        return engine.Engine(base_plan, device=device, scenario_plans=plans)

    return make


class ShardedBatch:
    def __init__(self, inputs, n_scen, rank=0, world_size=1, factor=1, max_nodes=20, seed=20260117, user_event=None,
                 engine_factory=None, coord=None):
        self.rank, self.world_size, self.n_scen = rank, world_size, n_scen
        self.own = scenarios.partition(n_scen, world_size, rank)
        scen = scenarios.disperse(inputs, n_scen, seed=seed)
        self.plans, self.x0 = [], []
        for k in self.own:
            p, u, c, x0 = problem.problem_from_inputs(scen[k], coord=coord, factor=factor, max_nodes=max_nodes)
            ue = gplan.PerigeeAtEvent(user_event) if user_event else None
            self.plans.append(gplan.CompiledPlan(p, u, c, user_eq=ue, coord=coord))
            self.x0.append(problem.xdict_to_vector(x0))
        self.engine = None
        if len(self.plans):
            make = engine_factory or default_engine_factory(rank)
            self.engine = make(self.plans[0], self.plans)

    @property
    def n_local(self):
        return len(self.plans)

    def residuals(self, X):
        """X[n_local, n_vars] -> g[n_local, n_rows]"""
        return np.asarray(self.engine.eval_residuals(X, self.n_local)).reshape(self.n_local, -1)

    def jacobian_values(self, X):
        return np.asarray(self.engine.eval_jacobian(X, self.n_local)).reshape(self.n_local, -1)

    def summaries(self, X):
        """[n_local, 3]: objective, max |equality row|, max inequality violation."""
        out = np.zeros((self.n_local, 3))
        if self.n_local == 0:
            return out
        G = self.residuals(X)
        P = self.plans[0]
        eq = np.zeros(P.n_rows, dtype=bool)
        ineq = np.zeros(P.n_rows, dtype=bool)
        for key, gr in P.group_rows.items():
            if gr is not None:
                (eq if key.startswith("eqcon") else ineq)[gr[0]: gr[0] + gr[1]] = True
        out[:, 0] = G[:, 0]
        out[:, 1] = np.abs(G[:, eq]).max(axis=1)
        out[:, 2] = np.maximum(0.0, -G[:, ineq]).max(axis=1) if ineq.any() else 0.0
        return out

    def gather(self, local):
        """All ranks' per-scenario rows in scenario order (the one collective of the path)."""
        import torch
        import torch.distributed as dist

        if self.world_size == 1:
            return np.asarray(local)
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        width = local.shape[1]
        sizes = [len(scenarios.partition(self.n_scen, self.world_size, r)) for r in range(self.world_size)]
        pad = max(sizes)
        buf = torch.zeros((pad, width), dtype=torch.float64, device=dev)
        buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
        parts = [torch.empty_like(buf) for _ in range(self.world_size)]
        dist.all_gather(parts, buf)
        return np.concatenate([parts[r][: sizes[r]].cpu().numpy() for r in range(self.world_size)], axis=0)
