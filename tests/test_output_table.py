"""The result table (reference: output_result.py:37-263).
  * oracle/output.py (libm flavour) reproduces the fixture produced by the reference's OWN output_result
    function BIT FOR BIT, every column, strings included;
  * the kernel's per-node function (output.h), stepped on the host, equals the oracle's gmath flavour bit
    for bit; the GPU tier checks the same through the C ABI, plus the fixture within 1e-9 relative;
  * many trajectories in one launch give the rows of each."""
import ctypes
import os

import numpy as np
import pytest

import emu_binding
import helpers
from gelato_b200 import output as gout
from gelato_b200 import problem
from oracle import leaves
from oracle import output as oout

FIX = np.load(os.path.join(helpers.GOLDEN, "example_output_result.npz"))


def _times(x, p, u):
    """tx_res / tu_res as the reference driver builds them (Trajectory_Optimization.py:477-492)."""
    tu, tx, ps = np.array([]), np.array([]), p["ps_params"]
    for i in range(p["num_sections"]):
        to, tf = x["t"][i], x["t"][i + 1]
        tu = np.hstack((tu, (ps.tau(i) * (tf - to) / 2 + (tf + to) / 2) * u["t"]))
        tx = np.hstack((tx, (np.hstack((-1.0, ps.tau(i))) * (tf - to) / 2 + (tf + to) / 2) * u["t"]))
    return tx, tu


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind in "USO" or b.dtype.kind in "USO":
        return [str(v) for v in a] == [str(v) for v in b]
    return bool(((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))).all())


@pytest.mark.parametrize("name", ["x0", "x1"])
def test_oracle_reproduces_the_reference_output_result_bitwise(name):
    p, u, c, x0 = helpers.example_problem()
    x = x0 if name == "x0" else helpers.perturbed(x0)
    tx, tu = _times(x, p, u)
    tab = oout.output_result(x, u, tx, tu, p, "libm")
    assert list(FIX[name + "/columns"]) == list(tab.keys()) == gout.COLUMNS
    for col in tab:
        assert _same(tab[col], FIX["%s/%s" % (name, col)]), col
    assert np.isnan(tab["lat_IIP"]).any() and not np.isnan(tab["lat_IIP"]).all()  # both IIP outcomes occur


def _emu_fn():
    emu_binding.build()
    L = ctypes.CDLL(emu_binding.LIB)
    L.emu_output_rows.restype = None
    return L.emu_output_rows


@pytest.mark.parametrize("engine", ["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def test_kernel_rows_match_the_oracle_bitwise_and_the_reference_fixture(engine):
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c, factor=3)
    x = helpers.perturbed(x0)
    tx, tu = _times(x, p, u)
    want = oout.output_result(x, u, tx, tu, p, "gmath")
    got = gout.output_result(x, u, tx, tu, p, as_frame=False, _fn=_emu_fn() if engine == "emu" else None)
    assert list(got.keys()) == gout.COLUMNS
    for col in got:
        assert _same(got[col], want[col]), col
    # against the reference's own table (libm): same strings, numbers within 1e-9 of the column's scale
    p, u, c, x0 = helpers.example_problem()
    tx, tu = _times(x0, p, u)
    got = gout.output_result(x0, u, tx, tu, p, as_frame=False, _fn=_emu_fn() if engine == "emu" else None)
    for col in got:
        ref = FIX["x0/" + col]
        if ref.dtype.kind in "US" or col in ("stage", "section"):
            assert _same(got[col], ref), col
        else:
            scale = max(1.0, float(np.nanmax(np.abs(ref))))
            np.testing.assert_allclose(got[col], ref, rtol=1e-9, atol=1e-9 * scale, err_msg=col)
    frame = gout.output_result(x0, u, tx, tu, p, _fn=_emu_fn() if engine == "emu" else None)
    assert list(frame.columns) == gout.COLUMNS and len(frame) == len(tx)


@pytest.mark.gpu
def test_many_trajectories_in_one_launch():
    p, u, c, x0 = helpers.example_problem()
    sols = []
    for k in range(5):
        x = helpers.perturbed(x0, seed=k)
        tx, tu = _times(x, p, u)
        sols.append((x, u, tx, tu, p))
    tabs = gout.output_tables(sols)
    for (x, u_, tx, tu, p_), tab in zip(sols, tabs):
        one = gout.output_result(x, u_, tx, tu, p_, as_frame=False)
        for col in one:
            assert _same(tab[col], one[col]), col
