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


# ---------------------------------------------------------------------------------------------------------------
# ONE large NLP sharded over the GPUs
# ---------------------------------------------------------------------------------------------------------------
_HOLE = 0x7FF8_6E0C_A7ED_0B20  # a quiet NaN with a payload no kernel produces: "this rank did not write here"


def split_range(n, world_size, rank):
    """[lo, hi): rank's contiguous share of n items, sizes differing by at most one."""
    q, r = divmod(n, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class EngineRangeEvaluator:
    """The CUDA engine behind ShardedProblem: device tensors in, gelato_eval_pair_packed_range_dev on torch's
    current stream."""

    def __init__(self, eng):
        import torch

        self.eng, self.torch = eng, torch
        self.device = torch.device("cuda", eng.device)
        self.n_rows, self.n_pack, self.n_vars = eng.n_rows, eng.n_pack, eng.n_vars
        self.n_blocks, self.n_vac = eng.n_jac_blocks_pair, eng.n_vacuum_nodes
        # a stream of its own: torch's default stream has handle 0, which the C ABI reads as "the plan's stream"
        self.stream = torch.cuda.Stream(self.device)

    def pair_range(self, x, g, packed, blocks, vacuum):
        """Ordered after the work queued on torch's current stream (x, the fill of g / packed) and before what is
        queued on it next (the compaction, the all_gather)."""
        cur = self.torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        self.eng.eval_pair_packed_range_dev(x.data_ptr(), g.data_ptr(), packed.data_ptr(), blocks, vacuum, 1,
                                            self.stream.cuda_stream)
        cur.wait_stream(self.stream)


class ShardedProblem:
    """objfunc + sens of ONE problem with its sections' work spread over the ranks.

    The reference walks the sections one after the other (lib/con_dynamics.py:320, the FD columns inside at
    :362-372); here the unit of work is a block of the pair kernel's block table (a chunk of dynamics nodes, aero
    rows, event rows or linear rows) or a vacuum dynamics node.  The block table interleaves heavy and light blocks
    (plan_host.h), so a contiguous share of it is a balanced share.  Rank r evaluates blocks split_range(n_blocks)
    and vacuum nodes split_range(n_vac); each packed Jacobian value and each residual row is written by exactly
    one block or node, so the ranks' results are disjoint parts of (g, packed).  The one exchange of the path:
    every rank's part, compacted, in ONE all_gather (NCCL on GPUs, gloo in the CPU tier), after which every rank
    holds the whole g and packed vector -- bit for bit what a single GPU computes.

    Which slots a rank owns is found once, by evaluation: the buffers are filled with a NaN payload no kernel
    produces, the rank's share is run, and what changed is its part (the set does not depend on x).  The parts
    are checked to be disjoint and to cover everything."""

    def __init__(self, evaluator, x_probe, rank=0, world_size=1):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.ev, self.rank, self.world_size = evaluator, rank, world_size
        self.device = evaluator.device
        self.blocks = split_range(evaluator.n_blocks, world_size, rank)
        self.vacuum = split_range(evaluator.n_vac, world_size, rank)
        self.n_rows, self.n_pack = evaluator.n_rows, evaluator.n_pack
        self.n_out = self.n_rows + self.n_pack
        # g and packed side by side in one buffer: one index set, one message
        self.buf = torch.empty(self.n_out, dtype=torch.float64, device=self.device)
        self.x = torch.empty(evaluator.n_vars, dtype=torch.float64, device=self.device)
        self._find_parts(x_probe)

    def _run(self, x):
        self.x.copy_(torch_from(self.torch, x, self.device))
        self.ev.pair_range(self.x, self.buf[: self.n_rows], self.buf[self.n_rows:], self.blocks, self.vacuum)

    def _find_parts(self, x_probe):
        torch, dist = self.torch, self.dist
        bits = self.buf.view(torch.int64)
        bits.fill_(_HOLE)
        self._run(x_probe)
        own = torch.nonzero(bits != _HOLE).flatten()
        if self.world_size == 1:
            self.parts, self.sizes = [own], [int(own.numel())]
        else:
            n = torch.tensor([own.numel()], dtype=torch.int64, device=self.device)
            ns = [torch.zeros_like(n) for _ in range(self.world_size)]
            dist.all_gather(ns, n)
            self.sizes = [int(t.item()) for t in ns]
            pad = max(self.sizes)
            mine = torch.zeros(pad, dtype=torch.int64, device=self.device)
            mine[: own.numel()] = own
            got = [torch.empty_like(mine) for _ in range(self.world_size)]
            dist.all_gather(got, mine)
            self.parts = [got[r][: self.sizes[r]] for r in range(self.world_size)]
        self.own = own
        self.pad = max(self.sizes)
        count = torch.zeros(self.n_out, dtype=torch.int64, device=self.device)
        for part in self.parts:
            count.index_add_(0, part, torch.ones_like(part))
        if not bool((count == 1).all()):
            raise RuntimeError("ShardedProblem: %d output slots written by no rank, %d by more than one"
                               % (int((count == 0).sum()), int((count > 1).sum())))
        self.send = torch.zeros(self.pad, dtype=torch.float64, device=self.device)
        self.recv = torch.empty(self.world_size * self.pad, dtype=torch.float64, device=self.device)
        # position of every output slot inside the gathered message
        src = torch.empty(self.n_out, dtype=torch.int64, device=self.device)
        for r, part in enumerate(self.parts):
            src[part] = r * self.pad + torch.arange(part.numel(), dtype=torch.int64, device=self.device)
        self.src = src

    def pair(self, x):
        """(g[n_rows], packed[n_pack]) of the whole problem, on every rank (views of one buffer that the next call
        overwrites)."""
        torch, dist = self.torch, self.dist
        self._run(x)
        if self.world_size == 1:
            return self.buf[: self.n_rows], self.buf[self.n_rows:]
        torch.index_select(self.buf, 0, self.own, out=self.send[: self.own.numel()])
        dist.all_gather_into_tensor(self.recv, self.send)
        torch.index_select(self.recv, 0, self.src, out=self.buf)
        return self.buf[: self.n_rows], self.buf[self.n_rows:]


def torch_from(torch, x, device):
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(device)
