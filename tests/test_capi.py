"""The C-ABI library: loads, exports every symbol include/gelato_b200.h declares,
matches the ctypes struct layout, and refuses to evaluate without a CUDA device
(there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from gelato_b200 import engine

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "gelato_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(gelato_[a-z0-9_]+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(engine.LIB_PATH), "run __graft_entry__.build() first"
    L = engine.load_library()
    declared = _declared_functions()
    assert len(declared) >= 16
    assert sorted(set(declared)) == sorted(engine.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None, name


def test_header_compiles_as_plain_c(tmp_path):
    """No C++ / torch types at the boundary: the header is valid C99."""
    src = tmp_path / "t.c"
    src.write_text('#include "gelato_b200.h"\nint main(void){GelatoPlanDesc d; GelatoScenarioDesc s; (void)d; (void)s; '
                   "return sizeof(d) > 0 ? 0 : 1;}\n")
    import subprocess

    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


def test_struct_layout_matches_header(tmp_path):
    import subprocess

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gelato_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", '
                   "sizeof(GelatoPlanDesc), offsetof(GelatoPlanDesc, vals_template), offsetof(GelatoPlanDesc, unit_mass), "
                   "sizeof(GelatoScenarioDesc), offsetof(GelatoPlanDesc, rc_aero)); return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(engine.PlanDesc)
    assert out[1] == engine.PlanDesc.vals_template.offset
    assert out[2] == engine.PlanDesc.unit_mass.offset
    assert out[3] == ctypes.sizeof(engine.ScenarioDesc)
    assert out[4] == engine.PlanDesc.rc_aero.offset


def test_no_cpu_fallback_without_device():
    """Without a GPU plan creation must fail loudly, not fall back to the host."""
    L = engine.load_library()
    if L.gelato_device_count() > 0:
        pytest.skip("a CUDA device is present")
    p, u, c, x0 = helpers.example_problem()
    P = helpers.compiled_plan(p, u, c)
    with pytest.raises(engine.GelatoError, match="no CUDA device"):
        engine.Engine(P)


def test_enum_constants_match_header():
    """gelato_b200/plan.py mirrors the header's table-column enums."""
    from gelato_b200 import plan as gp

    src = open(HEADER).read()
    for name, val in (("GS_I32_COLS", gp.GS_I32_COLS), ("GS_I64_COLS", gp.GS_I64_COLS), ("GS_F64_COLS", gp.GS_F64_COLS),
                      ("GL_I32_COLS", gp.GL_I32_COLS), ("GL_F64_COLS", gp.GL_F64_COLS), ("GA_I32_COLS", gp.GA_I32_COLS),
                      ("GA_I64_COLS", gp.GA_I64_COLS), ("GE_I64_COLS", gp.GE_I64_COLS), ("GE_F64_COLS", gp.GE_F64_COLS)):
        assert name in src
    prog = '#include <stdio.h>\n#include "gelato_b200.h"\nint main(void){printf("%d %d %d %d %d %d %d %d %d %d\\n", GS_I32_COLS, GS_I64_COLS, GS_F64_COLS, GL_I32_COLS, GL_F64_COLS, GA_I32_COLS, GA_I64_COLS, GE_I32_COLS, GE_I64_COLS, GE_F64_COLS); return 0;}\n'
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "e.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), os.path.join(d, "e.c"), "-o", os.path.join(d, "e")])
        got = [int(v) for v in subprocess.check_output([os.path.join(d, "e")]).split()]
    assert got == [gp.GS_I32_COLS, gp.GS_I64_COLS, gp.GS_F64_COLS, gp.GL_I32_COLS, gp.GL_F64_COLS, gp.GA_I32_COLS,
                   gp.GA_I64_COLS, gp.GE_I32_COLS, gp.GE_I64_COLS, gp.GE_F64_COLS]
    assert np.dtype("i4").itemsize == 4


def test_plan_create_refuses_malformed_descriptions():
    """The kernels trust the plan's index tables, so gelato_plan_create checks their bounds first
    (before it even looks for a device): a corrupted description is GELATO_ERR_ARG with a reason."""
    L = engine.load_library()
    p, u, c, x0 = helpers.example_problem()
    P = helpers.compiled_plan(p, u, c)

    def create(mutate):
        desc, keep = engine.make_desc(P)
        mutate(desc, keep)
        h = ctypes.c_void_p()
        rc = L.gelato_plan_create(ctypes.byref(desc), 0, ctypes.byref(h))
        msg = L.gelato_last_error().decode()
        if rc == 0:
            L.gelato_plan_destroy(h)
        return rc, msg

    def bad_section(desc, keep):
        a = np.array(P.sec_i32, dtype=np.int32, copy=True)
        a[3, 2] = 10 ** 6  # GS_XA far outside the state rows
        keep.append(a)
        desc.sec_i32 = a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    def bad_lin(desc, keep):
        a = np.array(P.lin_i32, dtype=np.int32, copy=True)
        a[0, 1] = P.n_vars + 5
        keep.append(a)
        desc.lin_i32 = a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    def bad_dx(desc, keep):
        desc.dx = 0.0

    def bad_evt(desc, keep):
        a = np.array(P.evt_i32, dtype=np.int32, copy=True)
        a[0, 0] = 99
        keep.append(a)
        desc.evt_i32 = a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

    for mutate, word in ((bad_section, "state rows"), (bad_lin, "outside x"), (bad_dx, "dx"), (bad_evt, "event job type")):
        rc, msg = create(mutate)
        assert rc == -1 and word in msg, (rc, msg)
    rc, msg = create(lambda d, k: None)  # the untouched description passes validation
    assert rc == 0 or "no CUDA device" in msg, (rc, msg)


def test_scenario_plans_must_agree_on_everything_that_is_not_carried_per_scenario():
    """make_scenario_desc refuses a batch whose plans differ in a table every scenario shares (they would silently
    be evaluated with the base plan's copy)."""
    import copy

    import pytest

    import bench
    from gelato_b200 import engine
    plans, _, _ = bench.load_workload("example", 1, 2, 0, 2)
    engine.make_scenario_desc(plans)
    bad = copy.copy(plans[1])
    bad.ca = plans[1].ca * 1.01
    with pytest.raises(ValueError, match="`ca`"):
        engine.make_scenario_desc([plans[0], bad])
    bad = copy.copy(plans[1])
    bad.units = (plans[1].units[0],) + (plans[1].units[1] * 2.0,) + tuple(plans[1].units[2:])
    with pytest.raises(ValueError, match="unit"):
        engine.make_scenario_desc([plans[0], bad])


def test_forward_simulation_entry_point_rejects_bad_arguments():
    """gelato_init_rocket_simulation checks its arguments before it touches the device: n <= 0, null buffers, empty
    tables, descending output times, strides shorter than a table (no GPU needed: every case fails first)."""
    import ctypes

    from gelato_b200 import engine
    L = engine.load_library()
    pd, pi32, pi64 = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64)
    x0, ev, zlt = np.zeros(11), np.zeros((2, 6)), np.zeros(2, dtype=np.int32)
    ut, wind, ca = np.zeros((3, 4)), np.zeros((2, 3)), np.zeros((2, 2))
    tout, out, strides = np.array([0.0, 1.0]), np.zeros((1, 2, 11)), np.zeros(4, dtype=np.int64)

    def call(n=1, x0=x0, tout=tout, strides=strides, dt=0.1, n_ev=2):
        P = lambda a, t=pd: a.ctypes.data_as(t) if a is not None else None  # noqa: E731
        return L.gelato_init_rocket_simulation(0, n, P(x0), P(ev), P(zlt, pi32), n_ev, P(ut), 3, P(wind), 2, P(ca), 2,
                                               P(strides, pi64), 0.0, P(tout), tout.size, dt, P(out), None)

    assert call(n=0) == -1
    assert call(x0=None) == -1 and b"null" in L.gelato_last_error()
    assert call(dt=0.0) == -1
    assert call(n_ev=0) == -1
    assert call(tout=np.array([1.0, 0.5])) == -1 and b"ascend" in L.gelato_last_error()
    assert call(strides=np.array([5, 0, 0, 0], dtype=np.int64)) == -1 and b"stride" in L.gelato_last_error()
