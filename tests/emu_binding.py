"""ctypes front-end of tests/emu (the host emulator of the CUDA kernels).

TEST HARNESS ONLY: lets the CPU test tier run the kernels' per-thread code
(gelato_b200/csrc/jobs.h) through the same GelatoPlanDesc the GPU library gets.
"""
import ctypes
import os
import subprocess

import numpy as np

from gelato_b200 import engine

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
BUILD = os.path.join(_HERE, "_build")
LIB = os.path.join(BUILD, "libgelato_emu.so")
_pd = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    src = os.path.join(_HERE, "emu", "emu.cpp")
    csrc = os.path.join(ROOT, "gelato_b200", "csrc")
    deps = [src, os.path.join(ROOT, "include", "gelato_b200.h")] + [
        os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".inc"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(f) for f in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-pthread", "-o", LIB, src])
    return LIB


class Emulator:
    def __init__(self, plan, scenario_plans=None):
        build()
        self.L = ctypes.CDLL(LIB)
        assert self.L.emu_unfused_check() == 1
        self.plan = plan
        self.desc, self._keep = engine.make_desc(plan)
        self.sc = None
        if scenario_plans is not None:
            self.sc, keep = engine.make_scenario_desc(scenario_plans)
            self._keep += keep
        for fn in (self.L.emu_eval_residuals, self.L.emu_eval_jacobian):
            fn.argtypes = [ctypes.POINTER(engine.PlanDesc), ctypes.POINTER(engine.ScenarioDesc), _pd, _pd, ctypes.c_int,
                           ctypes.POINTER(ctypes.c_int32)]
        self.L.emu_eval_pair.argtypes = [ctypes.POINTER(engine.PlanDesc), ctypes.POINTER(engine.ScenarioDesc), _pd, _pd, _pd,
                                         ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        _pi64 = ctypes.POINTER(ctypes.c_int64)
        self.L.emu_packed_map.argtypes = [ctypes.POINTER(engine.PlanDesc), _pi64, _pi64, _pd]
        self.L.emu_packed_map.restype = ctypes.c_longlong
        self.L.emu_n_xdep.argtypes = [ctypes.POINTER(engine.PlanDesc)]
        self.L.emu_n_xdep.restype = ctypes.c_longlong
        self.n_pack = int(self.L.emu_packed_map(ctypes.byref(self.desc), None, None, None))

    def _sc(self):
        return ctypes.byref(self.sc) if self.sc is not None else None

    @staticmethod
    def _ids(scen_ids):
        if scen_ids is None:
            return None, None
        ids = np.ascontiguousarray(scen_ids, dtype=np.int32)
        return ids, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    def eval_residuals(self, x, n_scen=1, scen_ids=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.full(n_scen * self.plan.n_rows, np.nan)
        ids, pids = self._ids(scen_ids)
        rc = self.L.emu_eval_residuals(ctypes.byref(self.desc), self._sc(), x.ctypes.data_as(_pd), g.ctypes.data_as(_pd), n_scen, pids)
        assert rc == 0, "the plan description failed validate_desc (plan_host.h)"
        return g if n_scen == 1 else g.reshape(n_scen, -1)

    def eval_jacobian(self, x, n_scen=1, scen_ids=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = np.full(n_scen * self.plan.n_vals, np.nan)
        ids, pids = self._ids(scen_ids)
        rc = self.L.emu_eval_jacobian(ctypes.byref(self.desc), self._sc(), x.ctypes.data_as(_pd), v.ctypes.data_as(_pd), n_scen, pids)
        assert rc == 0, "the plan description failed validate_desc (plan_host.h)"
        return v if n_scen == 1 else v.reshape(n_scen, -1)


    def packed_map(self):
        """(full_slot, src, sgn): vals[full_slot] = sgn * packed[src] (plan_host.h: build_packed_layout)."""
        n = int(self.L.emu_n_xdep(ctypes.byref(self.desc)))
        full, src, sgn = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64), np.empty(n)
        _pi64 = ctypes.POINTER(ctypes.c_int64)
        self.L.emu_packed_map(ctypes.byref(self.desc), full.ctypes.data_as(_pi64), src.ctypes.data_as(_pi64),
                              sgn.ctypes.data_as(_pd))
        return full, src, sgn

    def eval_pair(self, x, n_scen=1, scen_ids=None, packed=False):
        """(g, vals or packed) of a pair evaluation: defects from the Jacobian blocks' centre columns."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        g = np.full(n_scen * self.plan.n_rows, np.nan)
        width = self.n_pack if packed else self.plan.n_vals
        v = np.full(n_scen * width, np.nan)
        ids, pids = self._ids(scen_ids)
        rc = self.L.emu_eval_pair(ctypes.byref(self.desc), self._sc(), x.ctypes.data_as(_pd), g.ctypes.data_as(_pd),
                                  v.ctypes.data_as(_pd), n_scen, pids, 1 if packed else 0)
        assert rc == 0
        if n_scen == 1:
            return g, v
        return g.reshape(n_scen, -1), v.reshape(n_scen, -1)


    # ---- one problem sharded over ranks (gelato_b200/batch.py: ShardedProblem) ----
    def block_counts(self):
        """(blocks of a pair evaluation, vacuum dynamics nodes): the two ranges a rank's share is cut from."""
        nb, nv = ctypes.c_int(0), ctypes.c_int(0)
        self.L.emu_block_counts.argtypes = [ctypes.POINTER(engine.PlanDesc), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        self.L.emu_block_counts(ctypes.byref(self.desc), ctypes.byref(nb), ctypes.byref(nv))
        return nb.value, nv.value

    def eval_pair_range(self, x, g, packed, blocks, vacuum):
        """The packed pair evaluation restricted to blocks [b0, b1) and vacuum nodes [v0, v1), written into the
        caller's g[n_rows] / packed[n_pack] (what gelato_eval_pair_packed_range_dev launches)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert g.dtype == np.float64 and g.flags.c_contiguous and g.size == self.plan.n_rows
        assert packed.dtype == np.float64 and packed.flags.c_contiguous and packed.size == self.n_pack
        self.L.emu_eval_pair_range.argtypes = [ctypes.POINTER(engine.PlanDesc), ctypes.POINTER(engine.ScenarioDesc), _pd, _pd, _pd,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        rc = self.L.emu_eval_pair_range(ctypes.byref(self.desc), self._sc(), x.ctypes.data_as(_pd), g.ctypes.data_as(_pd),
                                        packed.ctypes.data_as(_pd), 1, blocks[0], blocks[1] - blocks[0], vacuum[0],
                                        vacuum[1] - vacuum[0])
        assert rc == 0


class EmuEngine(Emulator):
    """The emulator behind the Engine interface GelatoProblem uses (CPU test tier only)."""

    def __init__(self, plan, scenario_plans=None):
        super().__init__(plan, scenario_plans=scenario_plans)
        self.launches = self.calls = 0

    def eval_residuals(self, x, n_scen=1, out=None, scen_ids=None):
        self.launches += 1
        self.calls += 1
        return super().eval_residuals(x, n_scen, scen_ids)

    def eval_jacobian(self, x, n_scen=1, out=None, scen_ids=None):
        self.launches += 1
        self.calls += 1
        return super().eval_jacobian(x, n_scen, scen_ids)

    # update mode as the drop-in's default `sens` uses it: one output buffer kept across calls
    class _HostArray:
        def __init__(self, n):
            self.array = np.empty(int(n))

        def free(self):
            self.array = None

    def alloc_output(self, n):
        return EmuEngine._HostArray(n)

    def jacobian_template(self, out, n_scen=1):
        assert n_scen == 1
        out[:] = self.plan.vals_template
        return out

    def eval_jacobian_update(self, x, out, n_scen=1):
        assert n_scen == 1
        self.launches += 1
        self.calls += 1
        v = super().eval_jacobian(x, 1, None)
        idx = self.plan.xdep_index()
        out[idx] = v[idx]  # only the x-dependent slots move, as on the device
        return out

    def eval_pair_update(self, x, g_out, vals_out, n_scen=1):
        assert n_scen == 1
        self.launches += 1
        self.calls += 1
        g, v = super().eval_pair(x, 1, None, packed=False)
        g_out[:] = g
        idx = self.plan.xdep_index()
        vals_out[idx] = v[idx]
        return g_out.reshape(1, -1), vals_out.reshape(1, -1)

    def close(self):
        pass
