"""Per-GPU coalescing server for batches of concurrent solves.

A dispersion study runs many independent NLP solves (one per scenario: perturbed masses,
thrust, wind).  Each solver asks for `objfunc` / `sens` at its own pace.  The server owns ONE
engine holding every scenario's parameter blocks and turns whatever requests are pending at a
given moment into one batched launch per callback kind (C ABI gelato_eval_*_ids: batch slot k
is evaluated with the blocks of scenario ids[k]), so the GPU sees batches even though every solver
sees the reference's plain `objfunc(xdict)` / `sens(xdict, funcs)` signatures
(/root/reference/Trajectory_Optimization.py:194, 245).

    server = CoalescingServer(plans)                 # plans: one CompiledPlan per scenario
    cb = server.client(k)                             # in solver worker k (a thread)
    sol = IPOPT(options)(register(cb.objfunc, cb.sens, x0, condition), sens=cb.sens)
    cb.done()
    server.close()

Workers are threads: the engine call releases the GIL, and so does a native solver between
callbacks.  Nothing is exchanged between scenarios; across GPUs, run one server per rank
(gelato_b200/batch.py partitions the scenarios).
"""
import threading
import time

import numpy as np

from . import problem
from .plan import VAR_ORDER


class _Request:
    __slots__ = ("kind", "k", "x", "result", "error", "ready")

    def __init__(self, kind, k, x):
        self.kind, self.k, self.x = kind, k, x
        self.result = self.error = None
        self.ready = threading.Event()


class ScenarioCallbacks:
    """The reference's two callbacks for scenario k, served by a CoalescingServer."""

    def __init__(self, server, k):
        self.server, self.k = server, k
        self.plan = server.plans[k]

    def _x(self, xdict):
        return np.concatenate([np.asarray(xdict[n], dtype=np.float64).ravel() for n in VAR_ORDER])

    def objfunc(self, xdict):
        g = self.server._submit("f", self.k, self._x(xdict))
        return self.plan.split_residuals(g), False

    def sens(self, xdict, funcs=None):
        vals = self.server._submit("j", self.k, self._x(xdict))
        return self.plan.split_jacobian(vals, key_order=[n for n in xdict.keys() if n in VAR_ORDER]), False

    def done(self):
        """This solver will not call again (lets the server stop waiting for it when it batches)."""
        self.server._retire(self.k)


class CoalescingServer:
    def __init__(self, plans, device=0, engine_factory=None, max_wait_s=2e-4):
        """plans: one CompiledPlan per scenario (same structure).  engine_factory(base_plan, plans) ->
        engine (default: the CUDA engine on `device`).  max_wait_s: how long a pending request may wait
        for others to join its batch once not every active solver is waiting."""
        from . import engine as _engine

        self.plans = list(plans)
        make = engine_factory or (lambda base, ps: _engine.Engine(base, device=device, scenario_plans=ps))
        self.engine = make(self.plans[0], self.plans)
        self.max_wait_s = float(max_wait_s)
        self._lock = threading.Condition()
        self._pending = []
        self._active = set()
        self._stop = False
        self.stats = {"calls": 0, "launches": 0, "largest_batch": 0}
        self._thread = threading.Thread(target=self._serve, name="gelato-coalescer", daemon=True)
        self._thread.start()

    # ---- client side -----------------------------------------------------
    def client(self, k):
        with self._lock:
            self._active.add(k)
        return ScenarioCallbacks(self, k)

    def _retire(self, k):
        with self._lock:
            self._active.discard(k)
            self._lock.notify_all()

    def _submit(self, kind, k, x):
        req = _Request(kind, k, x)
        with self._lock:
            if self._stop:
                raise RuntimeError("the coalescing server is closed")
            self._pending.append(req)
            self.stats["calls"] += 1
            self._lock.notify_all()
        req.ready.wait()
        if req.error is not None:
            raise req.error
        return req.result

    # ---- server thread ---------------------------------------------------
    def _take_batch(self):
        """Block until there is work, give stragglers max_wait_s to join, return the requests of ONE kind."""
        with self._lock:
            while not self._pending and not self._stop:
                self._lock.wait()
            if self._stop and not self._pending:
                return None
            deadline = time.monotonic() + self.max_wait_s
            while len(self._pending) < len(self._active):
                left = deadline - time.monotonic()
                if left <= 0 or self._stop:
                    break
                self._lock.wait(left)
            kind = self._pending[0].kind
            batch = [r for r in self._pending if r.kind == kind]
            self._pending = [r for r in self._pending if r.kind != kind]
            return batch

    def _serve(self):
        while True:
            batch = self._take_batch()
            if batch is None:
                return
            try:
                X = np.stack([r.x for r in batch])
                ids = np.array([r.k for r in batch], dtype=np.int32)
                fn = self.engine.eval_residuals if batch[0].kind == "f" else self.engine.eval_jacobian
                out = np.asarray(fn(X, n_scen=len(batch), scen_ids=ids)).reshape(len(batch), -1)
                self.stats["launches"] += 1
                self.stats["largest_batch"] = max(self.stats["largest_batch"], len(batch))
                for r, row in zip(batch, out):
                    r.result = np.array(row, copy=True)
            except Exception as e:  # deliver the failure to every waiting solver
                for r in batch:
                    r.error = e
            for r in batch:
                r.ready.set()

    def close(self):
        """Stop the server: pending requests are served, the thread is joined WITHOUT a timeout (a launch in
        flight keeps using the plan handle until it returns), then the engine is released."""
        with self._lock:
            self._stop = True
            self._lock.notify_all()
        self._thread.join()
        with self._lock:
            late, self._pending = self._pending, []
        for r in late:  # submitted after the stop flag: fail them instead of leaving their solvers waiting
            r.error = RuntimeError("the coalescing server was closed")
            r.ready.set()
        self.engine.close()


def solve_batch(plans, x0s, conditions, solver_factory, device=0, engine_factory=None, max_workers=None):
    """Run one solve per scenario concurrently on one GPU: worker thread k registers the problem with the
    reference's registration block (nlpshim.register) on its coalesced callbacks and calls
    solver_factory()(optProb, sens=cb.sens).  Returns (solutions in scenario order, server statistics)."""
    from . import nlpshim

    server = CoalescingServer(plans, device=device, engine_factory=engine_factory)
    sols = [None] * len(plans)
    errors = []

    def work(k):
        cb = server.client(k)
        try:
            prob = nlpshim.register(cb.objfunc, cb.sens, problem.vector_to_xdict(np.array(x0s[k], dtype=np.float64),
                                                                                plans[k].M, plans[k].N, plans[k].S),
                                    conditions[k])
            sols[k] = solver_factory()(prob, sens=cb.sens)
        except Exception as e:
            errors.append((k, e))
        finally:
            cb.done()

    n_workers = min(len(plans), max_workers or len(plans))
    pending = list(range(len(plans)))
    threads = []
    lock = threading.Lock()

    def runner():
        while True:
            with lock:
                if not pending:
                    return
                k = pending.pop(0)
            work(k)

    for _ in range(n_workers):
        t = threading.Thread(target=runner)
        t.start()
        threads.append(t)
    for t in threads:
        t.join()
    stats = dict(server.stats)
    server.close()
    if errors:
        raise errors[0][1]
    return sols, stats
