"""CPU tier: the kernels' per-thread code (gelato_b200/csrc/jobs.h + physics.h),
stepped through by the host emulator in tests/emu with the plan the problem
compiler emits, must equal the oracle (gmath leaves, sequential-FMA D.X) BIT FOR
BIT: residual rows, Jacobian sparsity (row/col arrays, dtype, order, explicit
zeros) and Jacobian values -- including the fl(fl(x+dx)-dx) residue the
reference's in-place finite differences leave for later columns and groups.
The same assertions run against the real GPU in test_gpu_parity.py.
"""
import numpy as np
import pytest

import emu_binding
import helpers
from gelato_b200 import plan as gplan
from gelato_b200 import problem
from oracle import leaves


def _setup(variant, factor, max_nodes=12, user=True):
    Lg = leaves.get("gmath")
    inp = helpers.variant_inputs(variant)
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c, factor=factor, max_nodes=max_nodes)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma", user=user)
    P = helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c, user=user)
    return p, u, c, x0, O, P


@pytest.mark.parametrize("variant,factor,user", [
    ("example", 1, True), ("example", 3, True), ("fuel_inclination", 1, True), ("all_aero", 2, True),
    ("waypoints", 1, True), ("neg_area", 1, True), ("three_stage", 1, True), ("iip_orbital", 1, True), ("waypoints", 2, False), ("bare", 1, False),
])
def test_emulated_kernels_match_oracle_bitwise(variant, factor, user):
    p, u, c, x0, O, P = _setup(variant, factor, user=user)
    E = emu_binding.Emulator(P)
    for x in (x0, helpers.perturbed(x0)):
        xv = problem.xdict_to_vector(x)
        xa = helpers.copy_x(x)
        f, _ = O.objfunc(xa)
        helpers.assert_funcs_equal(f, P.split_residuals(E.eval_residuals(xv)))
        s, _ = O.sens(xa)
        vals = E.eval_jacobian(xv)
        assert not np.isnan(vals).any(), "a Jacobian slot was left unwritten"
        helpers.assert_sens_equal(s, P.split_jacobian(vals, key_order=list(x.keys())))


@pytest.mark.parametrize("variant,factor,user", [
    ("example", 1, True), ("example", 3, True), ("all_aero", 2, True), ("waypoints", 1, True), ("neg_area", 1, True),
    ("three_stage", 1, True), ("fuel_inclination", 1, True), ("bare", 1, False),
])
def test_pair_and_packed_outputs_equal_the_separate_kernels(variant, factor, user):
    """A pair evaluation takes objfunc's dynamics rows from the Jacobian blocks (centre column = pristine x);
    the packed output holds only the independent x-dependent values.  Both must reproduce the separate
    residual / Jacobian evaluations bit for bit: g == eval_residuals, and template + map(packed) == eval_jacobian."""
    p, u, c, x0, O, P = _setup(variant, factor, user=user)
    E = emu_binding.Emulator(P)
    full, src, sgn = E.packed_map()
    assert np.array_equal(full, P.xdep_index()), "the C layout and the plan compiler disagree on the x-dependent slots"
    assert src.min() == 0 and src.max() == E.n_pack - 1 and np.unique(src).size == E.n_pack
    assert set(np.unique(sgn)) <= {-1.0, 1.0}
    for x in (x0, helpers.perturbed(x0)):
        xv = problem.xdict_to_vector(x)
        g_ref, v_ref = E.eval_residuals(xv), E.eval_jacobian(xv)
        g, v = E.eval_pair(xv)
        assert np.array_equal(g, g_ref) and np.array_equal(v, v_ref)
        g, pk = E.eval_pair(xv, packed=True)
        assert np.array_equal(g, g_ref) and not np.isnan(pk).any()
        v = P.vals_template.copy()
        v[full] = sgn * pk[src]
        assert np.array_equal(v, v_ref)
        assert np.array_equal(np.signbit(v[full]), np.signbit(v_ref[full]))


def test_large_section_counts():
    """A section larger than one residual block (n = 40 > 32 nodes) and the minimum n = 2."""
    Lg = leaves.get("gmath")
    inp = helpers.example_inputs()
    inp["events"][2]["num_nodes"] = 40
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    P = helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c)
    E = emu_binding.Emulator(P)
    x = helpers.perturbed(x0)
    xa = helpers.copy_x(x)
    f, _ = O.objfunc(xa)
    helpers.assert_funcs_equal(f, P.split_residuals(E.eval_residuals(problem.xdict_to_vector(x))))
    s, _ = O.sens(xa)
    helpers.assert_sens_equal(s, P.split_jacobian(E.eval_jacobian(problem.xdict_to_vector(x)), key_order=list(x.keys())))


def test_residue_sensitive_inputs():
    """Values for which (x + dx) - dx != x (binade crossings, |x| <~ dx)
    planted in every perturbed variable group."""
    p, u, c, x0, O, P = _setup("waypoints", 1)
    E = emu_binding.Emulator(P)
    x = helpers.copy_x(x0)
    specials = np.array([0.9999999999, 0.249999995, 3e-9, 1e-12, 0.0, 0.499999999, -3e-9, 1.0 - 1e-9])
    rng = np.random.default_rng(3)
    for k in ("velocity", "quaternion", "u"):
        idx = rng.choice(x[k].size, size=24, replace=False)
        x[k][idx] = rng.choice(specials, size=24)
    xv = problem.xdict_to_vector(x)
    xa = helpers.copy_x(x)
    s, _ = O.sens(xa)
    assert any(not np.array_equal(xa[k], x[k]) for k in x), "test inputs left no residue"
    helpers.assert_sens_equal(s, P.split_jacobian(E.eval_jacobian(xv), key_order=list(x.keys())))


def test_key_order_of_dense_user_blocks():
    """jac_fd walks xdict in insertion order (jac_fd.py:54); the set-up call of the
    reference uses a different order than pyoptsparse does at run time."""
    p, u, c, x0, O, P = _setup("example", 1)
    E = emu_binding.Emulator(P)
    order = ["t", "u", "mass", "position", "velocity", "quaternion"]
    x = {k: helpers.perturbed(x0)[k] for k in order}
    xa = helpers.copy_x(x)
    s, _ = O.sens(xa)
    got = P.split_jacobian(E.eval_jacobian(problem.xdict_to_vector(x)), key_order=order)
    assert list(got["eqcon_user"].keys()) == order
    for k in order:
        assert np.array_equal(got["eqcon_user"][k], s["eqcon_user"][k]), k


def test_plan_rejects_arbitrary_python_user_constraints():
    p, u, c, x0 = helpers.example_problem()
    with pytest.raises(TypeError):
        gplan.CompiledPlan(p, u, c, user_eq=lambda *a: 0.0)


def test_batched_scenarios_match_per_scenario_oracle():
    """Dispersed scenarios (mass / thrust / wind) evaluated as one batch."""
    from gelato_b200 import scenarios

    Lg = leaves.get("gmath")
    inp = helpers.example_inputs()
    scen = scenarios.disperse(inp, 3, seed=20260117)
    plans, oracles, xs = [], [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        oracles.append(helpers.oracle_nlp(p, u, c, "gmath", "seqfma"))
        xs.append(helpers.perturbed(x0, seed=len(xs)))
    E = emu_binding.Emulator(plans[0], scenario_plans=plans)
    X = np.stack([problem.xdict_to_vector(x) for x in xs])
    G = E.eval_residuals(X, n_scen=3)
    V = E.eval_jacobian(X, n_scen=3)
    for k in range(3):
        xa = helpers.copy_x(xs[k])
        f, _ = oracles[k].objfunc(xa)
        helpers.assert_funcs_equal(f, plans[k].split_residuals(G[k]))
        s, _ = oracles[k].sens(xa)
        helpers.assert_sens_equal(s, plans[k].split_jacobian(V[k], key_order=list(xs[k].keys())))
    assert not np.array_equal(G[0], G[1])


@pytest.mark.parametrize("name", ["x0", "x1"])
def test_emulated_kernels_match_reference_golden_within_fd_noise(name):
    """The kernels' code (stepped on the host) against the fixture made by the
    reference's own Python layer on libm leaves: same sparsity, values within 1e-10
    relative + the reference's own finite-difference noise floor.  The same check runs
    on the GPU (test_gpu_parity.py)."""
    import os

    npz = np.load(os.path.join(helpers.GOLDEN, "example_reference.npz"))
    p, u, c, x0 = helpers.example_problem()
    P = helpers.compiled_plan(p, u, c)
    E = emu_binding.Emulator(P)
    xv = npz["%s/x" % name].copy()
    x = problem.vector_to_xdict(xv, P.M, P.N, P.S)
    f = P.split_residuals(E.eval_residuals(xv))
    for k, v in helpers.flatten_funcs(f).items():
        atol = 1e-11 if "alpha" in k else 1e-13
        np.testing.assert_allclose(v, npz["%s/f/%s" % (name, k)], rtol=1e-10, atol=atol, err_msg=k)
    s = P.split_jacobian(E.eval_jacobian(xv), key_order=list(x.keys()))
    helpers.assert_sens_within_noise(s, npz, name)


def test_emulated_kernels_match_reference_at_the_bench_workload():
    """The same at the workload bench.py times (example x15 in sections of <= 20 nodes, N = 990): fixture
    tests/golden/bench_reference.npz, made by the reference's own Python layer on its own C++ leaves."""
    import os

    npz = np.load(os.path.join(helpers.GOLDEN, "bench_reference.npz"))
    p, u, c, x0 = helpers.example_problem(factor=15, max_nodes=20)
    P = helpers.compiled_plan(p, u, c)
    assert (P.N, P.n_vars, P.n_vals) == (990, 13507, 571254)
    E = emu_binding.Emulator(P)
    xv = npz["x1/x"].copy()
    x = problem.vector_to_xdict(xv, P.M, P.N, P.S)
    f = P.split_residuals(E.eval_residuals(xv))
    for k, v in helpers.flatten_funcs(f).items():
        atol = 1e-11 if "alpha" in k else 1e-13
        np.testing.assert_allclose(v, npz["x1/f/%s" % k], rtol=1e-10, atol=atol, err_msg=k)
    s = P.split_jacobian(E.eval_jacobian(xv), key_order=list(x.keys()))
    helpers.assert_sens_within_noise(s, npz, "x1")


def test_emulated_jacobian_is_as_close_to_the_true_derivative_as_the_reference():
    """Finite differences with dx = 1e-8 carry ~1e-8 |f| of rounding noise, so two correct implementations
    differ slot by slot; what the solver cares about is the distance to the TRUE derivative.  Per Jacobian
    block: the kernels' values are no further from a 4th-order central-difference derivative of the
    (independent, libm-flavoured) oracle residuals than the reference's own values are (x 1.5)."""
    import os

    npz = np.load(os.path.join(helpers.GOLDEN, "example_reference.npz"))
    p, u, c, x0 = helpers.example_problem()
    P = helpers.compiled_plan(p, u, c)
    xv = npz["x1/x"].copy()
    x = problem.vector_to_xdict(xv, P.M, P.N, P.S)
    row0, J, col0 = helpers.true_jacobian(helpers.oracle_nlp(p, u, c, "libm", "numpy").objfunc, x)
    s = P.split_jacobian(emu_binding.Emulator(P).eval_jacobian(xv), key_order=list(x.keys()))
    rep = helpers.assert_as_close_to_truth_as_reference(s, npz, "x1", row0, J, col0)
    assert "eqcon_dyn_vel/position" in rep and "ineqcon_qalpha/position" in rep


@pytest.mark.parametrize("variant,user", [("example", True), ("all_aero", True), ("waypoints", True), ("neg_area", True),
                                          ("fuel_inclination", True), ("bare", False)])
def test_xdep_index_is_exactly_what_the_jacobian_kernel_writes(variant, user):
    """CompiledPlan.xdep_index() drives the sparse device->host update of the batched
    path: it must list exactly the slots the kernel code writes (the emulator starts
    from the template; a template of NaN exposes every slot it touches)."""
    p, u, c, x0, O, P = _setup(variant, 2, user=user)
    keep = P.vals_template
    try:
        P.vals_template = np.full_like(keep, np.nan)
        vals = emu_binding.Emulator(P).eval_jacobian(problem.xdict_to_vector(helpers.perturbed(x0)))
    finally:
        P.vals_template = keep
    written = np.flatnonzero(~np.isnan(vals))
    assert np.array_equal(written, P.xdep_index())
    assert P.n_xdep == written.size


USER_CASES = [
    # (equality built-in, inequality built-in)
    ([("apogee_radius", 6378137.0, 1.03), ("inclination_deg", 42.2, 1.0)], None),
    ([("orbit_energy", -3.0e7, 1.0)], [("eccentricity", 0.05, 0.5), ("semi_major_axis", 6578137.0, 1.0), ("angular_momentum", 5.2e10, 1.0)]),
    (None, [("perigee_radius", 6378137.0, 1.0)]),
]


def _user_setup(eq_rows, ineq_rows, flavour, dot, coord):
    from oracle import nlp, user_builtin

    L = leaves.get(flavour)
    p, u, c, x0 = helpers.example_problem(coord=coord, factor=2, max_nodes=12)
    ue = user_builtin.orbit_rows_at(L, helpers.USER_EVENT, eq_rows) if eq_rows else None
    ui = user_builtin.orbit_rows_at(L, "SECO", ineq_rows) if ineq_rows else None
    O = nlp.OracleNLP(p, u, c, flavour, dot, user_eq=ue, user_ineq=ui)
    P = gplan.CompiledPlan(p, u, c, user_eq=gplan.OrbitAtEvent(helpers.USER_EVENT, eq_rows) if eq_rows else None,
                           user_ineq=gplan.OrbitAtEvent("SECO", ineq_rows) if ineq_rows else None, coord=coord)
    return p, u, c, x0, O, P


@pytest.mark.parametrize("eq_rows,ineq_rows", USER_CASES)
def test_user_constraint_registry_matches_jac_fd_on_the_oracle(eq_rows, ineq_rows):
    """Built-in user constraints (OrbitAtEvent: apogee / perigee radius, inclination, energy, angular momentum,
    a, e at a named event; scalar or vector valued; equality and inequality group) against `jac_fd` over the
    same function written in Python on the oracle's leaves: values and the dense per-variable blocks, bit for bit,
    residue of the perturb / restore protocol included; pair and packed outputs as well."""
    Lg = leaves.get("gmath")
    p, u, c, x0, O, P = _user_setup(eq_rows, ineq_rows, "gmath", "seqfma", Lg.coordinate_c)
    E = emu_binding.Emulator(P)
    full, src, sgn = E.packed_map()
    assert np.array_equal(full, P.xdep_index())
    for x in (x0, helpers.perturbed(x0)):
        xv = problem.xdict_to_vector(x)
        xa = helpers.copy_x(x)
        f, _ = O.objfunc(xa)
        g = E.eval_residuals(xv)
        helpers.assert_funcs_equal(f, P.split_residuals(g))
        s, _ = O.sens(xa)
        vals = E.eval_jacobian(xv)
        helpers.assert_sens_equal(s, P.split_jacobian(vals, key_order=list(x.keys())))
        g2, pk = E.eval_pair(xv, packed=True)
        v2 = P.vals_template.copy()
        v2[full] = sgn * pk[src]
        assert np.array_equal(g2, g) and np.array_equal(v2, vals)
    for key, rows in (("eqcon_user", eq_rows), ("ineqcon_user", ineq_rows)):
        if rows:
            blk = P.split_jacobian(vals, key_order=list(x.keys()))[key]
            assert all(b.shape == (len(rows), P.sizes[k]) for k, b in blk.items())


def test_user_constraint_registry_matches_the_reference_jac_fd():
    """Live against the reference's own lib/jac_fd.py (where /root/reference exists): its dense forward differences
    of the Python user function on the reference-faithful (libm) leaves equal the oracle's."""
    import refharness

    if not refharness.available():
        pytest.skip("/root/reference not present on this machine")
    import importlib

    L = leaves.get("libm")
    refharness.load(L)
    ref_jac_fd = importlib.import_module("lib.jac_fd").jac_fd
    for eq_rows, ineq_rows in USER_CASES:
        p, u, c, x0, O, P = _user_setup(eq_rows, ineq_rows, "libm", "numpy", None)
        for fn in (O.user_eq, O.user_ineq):
            if fn is None:
                continue
            xa, xb = helpers.copy_x(helpers.perturbed(x0)), helpers.copy_x(helpers.perturbed(x0))
            ja = ref_jac_fd(fn, xa, p, u, c)
            jb = O.jac_fd(fn, xb)
            assert list(ja.keys()) == list(jb.keys())
            for k in ja:
                assert np.array_equal(ja[k], jb[k]), k
                assert np.array_equal(xa[k], xb[k]), k


def test_fused_pair_callbacks_return_the_same_dictionaries():
    """GelatoProblem(fuse_pair=True): objfunc runs the pair evaluation and a sens call at the same decision vector is
    answered from the kept Jacobian; at another vector (or twice in a row) sens evaluates.  Same funcs / funcsSens as
    the separate callbacks, and the launch counts show which path ran."""
    from gelato_b200 import callbacks
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    mk = lambda fuse: callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,  # noqa: E731
                                              engine_factory=emu_binding.EmuEngine, fuse_pair=fuse)
    A, B = mk(False), mk(True)
    assert B.fuse_pair
    xa, xb = helpers.perturbed(x0, seed=1), helpers.perturbed(x0, seed=2)
    fa, _ = A.objfunc(xa)
    sa, _ = A.sens(xa, fa)
    n0 = B.engine.calls
    fb, _ = B.objfunc(xa)
    sb, _ = B.sens(xa, fb)
    assert B.engine.calls - n0 == 1  # the pair evaluation only
    helpers.assert_funcs_equal(fa, fb)
    helpers.assert_sens_equal(sa, sb)
    # a sens at another point evaluates; the kept Jacobian is not reused afterwards
    sa2, _ = A.sens(xb)
    sb2, _ = B.sens(xb)
    assert B.engine.calls - n0 == 2
    helpers.assert_sens_equal(sa2, sb2)
    sb3, _ = B.sens(xa)  # same x as the pair, but the buffer has moved on
    assert B.engine.calls - n0 == 3
    helpers.assert_sens_equal(A.sens(xa)[0], sb3)
