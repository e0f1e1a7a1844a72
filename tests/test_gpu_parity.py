"""GPU tier (run with -m gpu on the B200 box): the CUDA engine, called through
the C ABI exactly as the drop-in objfunc / sens call it, against

  * the oracle (gmath leaves, sequential-FMA D.X): BIT-EXACT residual rows,
    Jacobian sparsity and Jacobian values, on the shipped example, refined meshes
    and constraint variants;
  * the committed golden fixtures: bit-exact for the gmath flavour; for the
    fixture produced by the reference's own Python layer on libm leaves the
    north-star tolerance 1e-10 relative (to the term scale for residuals; FD
    slots get the finite-difference noise allowance of DESIGN.md H1);
  * size-independent properties at the full benchmark sizes (1e4 .. 1e5 nodes).
"""
import os

import numpy as np
import pytest

import helpers
from gelato_b200 import callbacks, engine, problem, scenarios
from oracle import leaves

pytestmark = pytest.mark.gpu


def _problem(variant="example", factor=1, max_nodes=12, user=True):
    Lg = leaves.get("gmath")
    inp = helpers.variant_inputs(variant)
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c, factor=factor, max_nodes=max_nodes)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT) if user else None,
                                   coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma", user=user)
    return prob, O, x0


def test_library_is_the_cuda_build():
    L = engine.load_library()
    assert L.gelato_device_count() >= 1
    prob, O, x0 = _problem()
    assert prob.engine.launches == 0
    prob.objfunc(x0)
    prob.sens(x0)
    # objfunc: the residual kernel; sens: the block kernel and the vacuum-node kernel (a Jacobian this small comes
    # down whole; larger ones add the gather of the scattered x-dependent slots, test_update_mode_*)
    assert prob.engine.launches == 3 and prob.engine.calls == 2


@pytest.mark.parametrize("variant,factor,user", [
    ("example", 1, True), ("example", 4, True), ("example", 15, True), ("fuel_inclination", 1, True),
    ("all_aero", 2, True), ("waypoints", 1, True), ("neg_area", 1, True), ("three_stage", 2, True), ("iip_orbital", 1, True), ("waypoints", 3, False), ("bare", 1, False),
])
def test_gpu_matches_oracle_bitwise(variant, factor, user):
    prob, O, x0 = _problem(variant, factor, max_nodes=20 if factor == 15 else 12, user=user)
    for x in (x0, helpers.perturbed(x0)):
        xa = helpers.copy_x(x)
        f, fail = prob.objfunc(x)
        assert fail is False
        fo, _ = O.objfunc(xa)
        helpers.assert_funcs_equal(fo, f)
        s, fail = prob.sens(x, f)
        so, _ = O.sens(xa)
        helpers.assert_sens_equal(so, s)
        for k in x:  # the drop-in never mutates the caller's xdict
            assert np.array_equal(x[k], (x0 if x is x0 else x)[k])
    prob.close()


@pytest.mark.parametrize("name", ["x0", "x1"])
def test_gpu_matches_gmath_golden_bitwise(name):
    npz = np.load(os.path.join(helpers.GOLDEN, "example_gmath.npz"))
    prob, O, x0 = _problem()
    P = prob.plan
    x = problem.vector_to_xdict(npz["%s/x" % name].copy(), P.M, P.N, P.S)
    f, _ = prob.objfunc(x)
    for k, v in helpers.flatten_funcs(f).items():
        assert np.array_equal(v, npz["%s/f/%s" % (name, k)]), k
    s, _ = prob.sens(x, f)
    for k, (r, c, d, shape) in helpers.flatten_sens(s).items():
        assert np.array_equal(d, npz["%s/j/%s/data" % (name, k)]), k
        if r is not None:
            assert np.array_equal(r, npz["%s/j/%s/rows" % (name, k)]), k
            assert np.array_equal(c, npz["%s/j/%s/cols" % (name, k)]), k
    prob.close()


@pytest.mark.parametrize("name", ["x0", "x1"])
def test_gpu_matches_reference_golden_within_tolerance(name):
    """Against the fixture produced by the reference's own Python layer (libm physics,
    BLAS D.X).  Residuals: 1e-10 relative to the term scale (the defect itself tends to
    0 at convergence).  Jacobian: identical sparsity (rows, cols, shape, order, explicit
    zeros); values within 1e-10 relative + the reference's OWN finite-difference noise
    floor for that block -- the fixture's `jn` arrays: how far the reference's value
    moves when its inputs move by one ulp (make_golden.py).  Blocks without finite
    differences (D entries, constants, analytic terms) have a zero floor and must meet
    1e-10 relative outright.  (With dx = 1e-8 a last-bit difference in any elementary
    function moves a quotient by ~1e-8 |f| x conditioning: DESIGN.md "H1".)"""
    npz = np.load(os.path.join(helpers.GOLDEN, "example_reference.npz"))
    p, u, c, x0 = helpers.example_problem()  # host set-up through libm, like the reference
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT))
    P = prob.plan
    x = problem.vector_to_xdict(npz["%s/x" % name].copy(), P.M, P.N, P.S)
    f, _ = prob.objfunc(x)
    for k, v in helpers.flatten_funcs(f).items():
        ref = npz["%s/f/%s" % (name, k)]
        atol = 1e-11 if "alpha" in k else 1e-13  # acos near 1 for the angle-of-attack rows; terms are O(1e-3..1)
        np.testing.assert_allclose(v, ref, rtol=1e-10, atol=atol, err_msg=k)
    s, _ = prob.sens(x, f)
    helpers.assert_sens_within_noise(s, npz, name)
    prob.close()


def test_gpu_matches_reference_at_the_bench_workload():
    """The configuration bench.py times (example x15, N = 990) against the reference's own result on it
    (tests/golden/bench_reference.npz): sparsity identical, residuals 1e-10, Jacobian 1e-10 + per-row FD noise."""
    npz = np.load(os.path.join(helpers.GOLDEN, "bench_reference.npz"))
    p, u, c, x0 = helpers.example_problem(factor=15, max_nodes=20)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT))
    P = prob.plan
    x = problem.vector_to_xdict(npz["x1/x"].copy(), P.M, P.N, P.S)
    f, _ = prob.objfunc(x)
    for k, v in helpers.flatten_funcs(f).items():
        atol = 1e-11 if "alpha" in k else 1e-13
        np.testing.assert_allclose(v, npz["x1/f/%s" % k], rtol=1e-10, atol=atol, err_msg=k)
    s, _ = prob.sens(x, f)
    helpers.assert_sens_within_noise(s, npz, "x1")
    prob.close()


def test_gpu_jacobian_is_as_close_to_the_true_derivative_as_the_reference():
    """Per Jacobian block, |J_gpu - J_true| <= 1.5 |J_reference - J_true| (+ the FD noise floor), with J_true a
    4th-order central-difference derivative of the libm-flavoured oracle residuals (helpers.true_jacobian)."""
    npz = np.load(os.path.join(helpers.GOLDEN, "example_reference.npz"))
    p, u, c, x0 = helpers.example_problem()
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT))
    P = prob.plan
    x = problem.vector_to_xdict(npz["x1/x"].copy(), P.M, P.N, P.S)
    row0, J, col0 = helpers.true_jacobian(helpers.oracle_nlp(p, u, c, "libm", "numpy").objfunc, x)
    s, _ = prob.sens(x, None)
    helpers.assert_as_close_to_truth_as_reference(s, npz, "x1", row0, J, col0)
    prob.close()


def test_batched_scenarios_bitwise_and_independent_of_batching():
    Lg = leaves.get("gmath")
    inp = helpers.example_inputs()
    scen = scenarios.disperse(inp, 5, seed=20260117)
    plans, oracles, xs = [], [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        oracles.append(helpers.oracle_nlp(p, u, c, "gmath", "seqfma"))
        xs.append(helpers.perturbed(x0, seed=len(xs)))
    E = engine.Engine(plans[0], scenario_plans=plans)
    X = np.stack([problem.xdict_to_vector(x) for x in xs])
    G = E.eval_residuals(X, n_scen=5)
    V = E.eval_jacobian(X, n_scen=5)
    for k in range(5):
        xa = helpers.copy_x(xs[k])
        f, _ = oracles[k].objfunc(xa)
        helpers.assert_funcs_equal(f, plans[k].split_residuals(G[k]))
        s, _ = oracles[k].sens(xa)
        helpers.assert_sens_equal(s, plans[k].split_jacobian(V[k], key_order=list(xs[k].keys())))
    # a smaller batch gives the same bits for the scenarios it contains
    assert np.array_equal(E.eval_jacobian(X[:2], n_scen=2), V[:2])
    E.close()


def test_device_resident_entry_points_and_pinned_buffers():
    import torch

    prob, O, x0 = _problem("example", 4)
    E, P = prob.engine, prob.plan
    xv = problem.xdict_to_vector(helpers.perturbed(x0))
    g_host = E.eval_residuals(xv).copy()
    v_host = E.eval_jacobian(xv).copy()
    xd = torch.from_numpy(xv).cuda()
    gd = torch.full((P.n_rows,), float("nan"), dtype=torch.float64, device="cuda")
    vd = torch.full((P.n_vals,), float("nan"), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    E.fill_template(vd.data_ptr(), 1, st)
    E.eval_residuals_dev(xd.data_ptr(), gd.data_ptr(), 1, st)
    E.eval_jacobian_dev(xd.data_ptr(), vd.data_ptr(), 1, st)
    torch.cuda.synchronize()
    assert np.array_equal(gd.cpu().numpy(), g_host)
    assert np.array_equal(vd.cpu().numpy(), v_host)
    # page-locked caller buffers take the direct-DMA path
    px, pv = engine.PinnedArray(P.n_vars), engine.PinnedArray(P.n_vals)
    px.array[:] = xv
    E.eval_jacobian(px.array, out=pv.array)
    assert np.array_equal(pv.array, v_host)
    px.free()
    pv.free()
    prob.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_update_mode_equals_full_jacobian(pinned):
    """gelato_jacobian_template + gelato_eval_jacobian_update (only the x-dependent slots cross
    PCIe) leave the host buffer bit-identical to the full-copy call, call after call."""
    Lg = leaves.get("gmath")
    inp = helpers.example_inputs()
    scen = scenarios.disperse(inp, 6, seed=3)
    plans, xs = [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c, factor=2)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        xs.append(x0)
    E = engine.Engine(plans[0], scenario_plans=plans)
    E.set_host_threads(3)
    P = plans[0]
    assert E.L.gelato_plan_n_xdep(E.h) == P.n_xdep and 0 < P.n_xdep < P.n_vals
    holder = engine.PinnedArray(6 * P.n_vals) if pinned else None
    buf = holder.array if pinned else np.empty(6 * P.n_vals)
    buf[:] = np.nan
    E.jacobian_template(buf, 6)
    assert np.array_equal(buf.reshape(6, -1), np.stack([pl.vals_template for pl in plans]))
    for seed in (1, 2, 3):
        X = np.stack([problem.xdict_to_vector(helpers.perturbed(x, seed=seed + 10 * k)) for k, x in enumerate(xs)])
        want = E.eval_jacobian(X, n_scen=6).copy()
        got = E.eval_jacobian_update(X, buf, n_scen=6)
        assert np.array_equal(got, want)
    # a smaller batch through the same plan, and the single-scenario path
    one = np.empty(P.n_vals)
    E.jacobian_template(one, 1)
    assert np.array_equal(E.eval_jacobian_update(X[0], one, 1), want[0])
    if holder is not None:
        holder.free()
    E.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_sliced_update_pipeline_is_invisible(pinned):
    """Update mode pipelined over 1, 2, 3, 5 slices of a batch of DISPERSED scenarios (every slice must use its own
    scenarios' parameter blocks), packed or zero-copy: the host buffers are the same bits as the plain calls."""
    Lg = leaves.get("gmath")
    inp = helpers.example_inputs()
    n = 7
    scen = scenarios.disperse(inp, n, seed=5)
    plans, xs = [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c, factor=2)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        xs.append(x0)
    E = engine.Engine(plans[0], scenario_plans=plans)
    P = plans[0]
    X = np.stack([problem.xdict_to_vector(helpers.perturbed(x, seed=3 + k)) for k, x in enumerate(xs)])
    g_want, v_want = E.eval_residuals(X, n).copy(), E.eval_jacobian(X, n).copy()
    assert len({v_want[k].tobytes() for k in range(n)}) == n  # the scenarios do differ
    hg = engine.PinnedArray(n * P.n_rows) if pinned else None
    hv = engine.PinnedArray(n * P.n_vals) if pinned else None
    g = hg.array if pinned else np.empty(n * P.n_rows)
    v = hv.array if pinned else np.empty(n * P.n_vals)
    E.jacobian_template(v, n)
    for zero_copy in (False, True):
        E.set_update_zero_copy(zero_copy)
        for slices, threads in ((1, 1), (2, 3), (3, 2), (5, 4), (0, 0)):
            E.set_update_slices(slices)
            E.set_host_threads(threads)
            g[:] = np.nan
            v.reshape(n, -1)[:, P.xdep_index()] = np.nan
            G, V = E.eval_pair_update(X, g, v, n)
            assert np.array_equal(G, g_want) and np.array_equal(V, v_want), (zero_copy, slices)
            v.reshape(n, -1)[:, P.xdep_index()] = np.nan
            assert np.array_equal(E.eval_jacobian_update(X, v, n), v_want), (zero_copy, slices)
    # fewer scenarios than slices asked for
    E.set_update_slices(8)
    v.reshape(n, -1)[:3, P.xdep_index()] = np.nan
    assert np.array_equal(E.eval_jacobian_update(X[:3], v[:3 * P.n_vals], 3), v_want[:3])
    if pinned:
        hg.free()
        hv.free()
    E.close()


def test_pair_evaluation_equals_the_two_calls():
    import torch

    prob, O, x0 = _problem("example", 4)
    E, P = prob.engine, prob.plan
    X = np.stack([problem.xdict_to_vector(helpers.perturbed(x0, seed=s)) for s in (1, 2, 3)])
    g_want = E.eval_residuals(X, 3).copy()
    v_want = E.eval_jacobian(X, 3).copy()
    for pinned in (True, False):
        hg = engine.PinnedArray(3 * P.n_rows) if pinned else None
        hv = engine.PinnedArray(3 * P.n_vals) if pinned else None
        g = hg.array if pinned else np.empty(3 * P.n_rows)
        v = hv.array if pinned else np.empty(3 * P.n_vals)
        g[:] = np.nan
        E.jacobian_template(v, 3)
        G, V = E.eval_pair_update(X, g, v, 3)
        assert np.array_equal(G, g_want) and np.array_equal(V, v_want)
        if pinned:
            hg.free()
            hv.free()
    xd = torch.from_numpy(X).cuda()
    gd = torch.full((3, P.n_rows), float("nan"), dtype=torch.float64, device="cuda")
    vd = torch.empty((3, P.n_vals), dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())  # the fills above run on torch's current stream
    with torch.cuda.stream(s):
        E.fill_template(vd.data_ptr(), 3, s.cuda_stream)
        E.eval_pair_dev(xd.data_ptr(), gd.data_ptr(), vd.data_ptr(), 3, s.cuda_stream)
    s.synchronize()
    assert np.array_equal(gd.cpu().numpy(), g_want) and np.array_equal(vd.cpu().numpy(), v_want)
    prob.close()


@pytest.mark.parametrize("variant,factor,n", [("example", 4, 3), ("all_aero", 2, 40), ("neg_area", 1, 2), ("three_stage", 1, 17)])
def test_packed_pair_evaluation_reproduces_the_two_calls(variant, factor, n):
    """gelato_eval_pair_packed: objfunc's rows from the Jacobian kernels' centre columns, and only the independent
    x-dependent Jacobian values, contiguous.  g must equal gelato_eval_residuals and template + map(packed) must
    equal gelato_eval_jacobian, bit for bit -- through pageable and page-locked buffers, sliced pipeline included
    (n = 40 scenarios -> 2 slices), and on the device-resident entry point."""
    import torch

    prob, O, x0 = _problem(variant, factor)
    E, P = prob.engine, prob.plan
    X = np.stack([problem.xdict_to_vector(helpers.perturbed(x0, seed=s)) for s in range(n)])
    g_want = E.eval_residuals(X, n).reshape(n, -1).copy()
    v_want = E.eval_jacobian(X, n).reshape(n, -1).copy()
    full, src, sgn = E.packed_map()
    assert np.array_equal(full, P.xdep_index()) and E.n_pack == np.unique(src).size < full.size

    def expand(pk):
        v = np.tile(P.vals_template, (n, 1))
        v[:, full] = sgn * pk.reshape(n, -1)[:, src]
        return v

    G, PK = E.eval_pair_packed(X, n)
    assert np.array_equal(G.reshape(n, -1), g_want) and np.array_equal(expand(PK), v_want)
    assert np.array_equal(expand(E.eval_jacobian_packed(X, n)), v_want)
    hx, hg, hp = engine.PinnedArray(X.size), engine.PinnedArray(n * P.n_rows), engine.PinnedArray(n * E.n_pack)
    hx.array[:] = X.ravel()
    hg.array[:] = np.nan
    hp.array[:] = np.nan
    E.eval_pair_packed(hx.array, n, g_out=hg.array, packed_out=hp.array)
    assert np.array_equal(hg.array.reshape(n, -1), g_want) and np.array_equal(expand(hp.array), v_want)
    for a in (hx, hg, hp):
        a.free()
    xd = torch.from_numpy(X).cuda()
    gd = torch.full((n, P.n_rows), float("nan"), dtype=torch.float64, device="cuda")
    pd = torch.full((n, E.n_pack), float("nan"), dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())  # the fills above run on torch's current stream
    with torch.cuda.stream(s):
        E.eval_pair_packed_dev(xd.data_ptr(), gd.data_ptr(), pd.data_ptr(), n, s.cuda_stream)
    s.synchronize()
    assert np.array_equal(gd.cpu().numpy(), g_want) and np.array_equal(expand(pd.cpu().numpy()), v_want)
    prob.close()


def test_replacing_the_scenarios_refreshes_the_staged_constants():
    """gelato_plan_set_scenarios after an evaluation: the staged Jacobian buffer must not keep the previous
    scenarios' constant slots."""
    Lg = leaves.get("gmath")
    scen = scenarios.disperse(helpers.example_inputs(), 4, seed=3)
    plans, xs = [], []
    for si in scen:
        p, u, c, x0 = problem.problem_from_inputs(si, coord=Lg.coordinate_c)
        plans.append(helpers.compiled_plan(p, u, c, coord=Lg.coordinate_c))
        xs.append(problem.xdict_to_vector(helpers.perturbed(x0, seed=len(xs))))
    X = np.stack(xs)
    Ea = engine.Engine(plans[0], scenario_plans=plans[:2])
    va = Ea.eval_jacobian(X[:2], 2).copy()
    Eb = engine.Engine(plans[0], scenario_plans=plans[2:])
    want = Eb.eval_jacobian(X[2:], 2).copy()
    assert not np.array_equal(va, want)
    import ctypes
    sc, keep = engine.make_scenario_desc(plans[2:])
    assert Ea.L.gelato_plan_set_scenarios(Ea.h, ctypes.byref(sc)) == 0
    assert np.array_equal(Ea.eval_jacobian(X[2:], 2), want)
    Ea.close()
    Eb.close()


@pytest.mark.parametrize("eq_rows,ineq_rows", [
    ([("apogee_radius", 6378137.0, 1.03), ("inclination_deg", 42.2, 1.0)], None),
    ([("orbit_energy", -3.0e7, 1.0)], [("eccentricity", 0.05, 0.5), ("semi_major_axis", 6578137.0, 1.0), ("angular_momentum", 5.2e10, 1.0)]),
])
def test_user_constraint_registry_on_the_gpu(eq_rows, ineq_rows):
    """Built-in user constraints (callbacks.OrbitAtEvent; scalar and vector valued, equality and inequality group):
    the CUDA callbacks against `jac_fd` over the same function in Python on the oracle's leaves, bit for bit."""
    from oracle import nlp, user_builtin

    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c, factor=2, max_nodes=12)
    ue = user_builtin.orbit_rows_at(Lg, helpers.USER_EVENT, eq_rows) if eq_rows else None
    ui = user_builtin.orbit_rows_at(Lg, "SECO", ineq_rows) if ineq_rows else None
    O = nlp.OracleNLP(p, u, c, "gmath", "seqfma", user_eq=ue, user_ineq=ui)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.OrbitAtEvent(helpers.USER_EVENT, eq_rows) if eq_rows else None,
                                   user_ineq=callbacks.OrbitAtEvent("SECO", ineq_rows) if ineq_rows else None, coord=Lg.coordinate_c)
    for x in (x0, helpers.perturbed(x0)):
        xa = helpers.copy_x(x)
        f, _ = O.objfunc(xa)
        helpers.assert_funcs_equal(f, prob.objfunc(x)[0])
        s, _ = O.sens(xa)
        helpers.assert_sens_equal(s, prob.sens(x)[0])
    prob.close()


def test_reuse_output_sens_equals_fresh_sens():
    """The drop-in `sens` keeps one page-locked Jacobian buffer across calls by default (only the x-dependent values
    cross PCIe; the COO arrays are views valid until the next call); reuse_output=False returns fresh arrays like the
    reference.  Same values either way, and the views do change with the next call."""
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c, factor=3, max_nodes=12)
    kw = dict(user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c)
    fresh = callbacks.GelatoProblem(p, u, c, reuse_output=False, **kw)
    reuse = callbacks.GelatoProblem(p, u, c, **kw)
    assert reuse.reuse_output and not fresh.reuse_output
    first = None
    for seed in (1, 2):
        x = helpers.perturbed(x0, seed=seed)
        sr = reuse.sens(x)[0]
        helpers.assert_sens_equal(fresh.sens(x)[0], sr)
        if first is None:
            first = sr["eqcon_dyn_vel"]["mass"]["coo"][2]
            kept = first.copy()
    assert not np.array_equal(first, kept)  # a view of the reused buffer: the second call rewrote it
    fresh.close()
    reuse.close()


def test_repeated_calls_are_deterministic_and_template_survives():
    prob, O, x0 = _problem("example", 2)
    E = prob.engine
    xa = problem.xdict_to_vector(x0)
    xb = problem.xdict_to_vector(helpers.perturbed(x0))
    va = E.eval_jacobian(xa).copy()
    vb = E.eval_jacobian(xb).copy()
    assert not np.array_equal(va, vb)
    assert np.array_equal(E.eval_jacobian(xa), va)  # x-dependent slots fully rewritten, constants untouched
    assert np.array_equal(E.eval_jacobian(xb), vb)
    prob.close()


@pytest.mark.parametrize("variant,factor", [("example", 150), ("example", 1500), ("three_stage", 62)])
def test_full_size_properties(variant, factor):
    """BASELINE.json configs[4] sizes (N = 9 900 and 99 000 nodes) and configs[2] (three stages, 250
    sections, N = 4 836): the oracle is too slow there, so check
    properties that do not depend on the size:
      * section locality: every section's rows equal the rows of the same section
        evaluated inside a small problem (sampled sections vs the oracle);
      * linear groups satisfy g = J x + c with the emitted COO blocks;
      * the Jacobian of a batch equals the Jacobian of each member."""
    Lg = leaves.get("gmath")
    inp = helpers.variant_inputs(variant)
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c, factor=factor, max_nodes=20)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c)
    P = prob.plan
    assert P.N == sum(e["num_nodes"] for e in inp["events"][:-1]) * factor
    x = helpers.perturbed(x0)
    f, _ = prob.objfunc(x)
    s, _ = prob.sens(x, f)
    # (1) oracle on the dynamics rows of sampled sections (the oracle's leaves evaluate any row subset)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    rng = np.random.default_rng(5)
    ps = p["ps_params"]
    mass, pos = x["mass"], x["position"].reshape(-1, 3)
    vel, quat = x["velocity"].reshape(-1, 3), x["quaternion"].reshape(-1, 4)
    units = O._units3()
    for i in rng.choice(P.S, size=40, replace=False):
        ua, ub, xa, xb, n = ps.get_index(i)
        to, tf = x["t"][i], x["t"][i + 1]
        tn = ps.time_nodes(i, to, tf)
        lh = O._dot(ps.D(i), vel[xa:xb])
        rhs = O._rhs_vel(i, mass[xa + 1:xb], pos[xa + 1:xb], vel[xa + 1:xb], quat[xa + 1:xb], tn[1:], units)
        want = (lh - rhs * (tf - to) * u["t"] / 2.0).ravel()
        assert np.array_equal(f["eqcon_dyn_vel"][3 * ua:3 * ub], want), int(i)
    # (2) linear groups: g = J x + c, where c = g(0)
    zero = {k: np.zeros_like(v) for k, v in x.items()}
    f0, _ = prob.objfunc(zero)
    for key in ("eqcon_knot", "eqcon_rate", "ineqcon_time", "ineqcon_kick", "eqcon_time"):
        acc = np.array(f0[key], dtype=np.float64, copy=True)
        for var, blk in s[key].items():
            r, cidx, d = blk["coo"]
            np.add.at(acc, r, d * x[var][cidx])
        np.testing.assert_allclose(acc, f[key], rtol=0, atol=1e-12, err_msg=key)
    # (3) every value finite, every slot written
    v = prob.engine.eval_jacobian(problem.xdict_to_vector(x))
    assert np.isfinite(v).all()
    prob.close()


def test_small_objfunc_calls_through_the_captured_graph_stay_exact():
    """A small host-buffer objfunc call is one captured graph launch (upload, k_residuals, download).  The graph bakes
    in the batch size and the staging pointers: alternating batch sizes, growing the staging buffers and going back
    must keep returning the bits of the separate device-pointer path."""
    import torch
    prob, O, x0 = _problem()
    E, P = prob.engine, prob.plan
    X = np.stack([problem.xdict_to_vector(helpers.perturbed(x0, seed=k)) for k in range(5)])

    def device_path(n):
        xd = torch.from_numpy(X[:n].copy()).cuda()
        gd = torch.empty((n, P.n_rows), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        E.eval_residuals_dev(xd.data_ptr(), gd.data_ptr(), n, None)
        torch.cuda.synchronize()
        return gd.cpu().numpy()

    launches0 = E.launches
    for n in (1, 1, 2, 1, 5, 2, 1):
        g = np.asarray(E.eval_residuals(X[:n].ravel(), n)).reshape(n, -1)
        assert np.array_equal(g, device_path(n)), n
    assert E.launches - launches0 == 14  # one kernel per call on either path
    fo, _ = O.objfunc(helpers.perturbed(x0, seed=0))
    f, _ = prob.objfunc(helpers.perturbed(x0, seed=0))
    helpers.assert_funcs_equal(fo, f)
    prob.close()


def test_fused_pair_callbacks_on_the_gpu():
    """GelatoProblem(fuse_pair=True) on the device: objfunc runs the pair evaluation, the sens that follows at the
    same decision vector launches nothing; both equal the oracle bit for bit, also after the buffer has moved on."""
    Lg = leaves.get("gmath")
    p, u, c, x0 = helpers.example_problem(coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                                   fuse_pair=True)
    assert prob.fuse_pair
    for seed in (1, 2):
        x = helpers.perturbed(x0, seed=seed)
        n0 = prob.engine.calls
        f, _ = prob.objfunc(x)
        s, _ = prob.sens(x, f)
        assert prob.engine.calls - n0 == 1
        fo, _ = O.objfunc(helpers.copy_x(x))
        so, _ = O.sens(helpers.copy_x(x))
        helpers.assert_funcs_equal(fo, f)
        helpers.assert_sens_equal(so, s)
    xb = helpers.perturbed(x0, seed=3)
    s, _ = prob.sens(xb)  # no objfunc before it: evaluates
    helpers.assert_sens_equal(O.sens(helpers.copy_x(xb))[0], s)
    s, _ = prob.sens(helpers.perturbed(x0, seed=2))  # the pair's point, but the buffer has moved on: evaluates
    helpers.assert_sens_equal(O.sens(helpers.copy_x(helpers.perturbed(x0, seed=2)))[0], s)
    prob.close()


@pytest.mark.parametrize("variant,factor", [("example", 4), ("three_stage", 2)])
def test_block_ranges_of_one_problem_on_the_gpu(variant, factor):
    """SURVEY.md 8(e)-2, one problem sharded over GPUs: gelato_eval_pair_packed_range_dev over disjoint covering
    ranges of the block table and of the vacuum nodes writes every residual row and packed value exactly once and
    reproduces gelato_eval_pair_packed bit for bit; batch.ShardedProblem on one rank is that evaluation."""
    import torch

    from gelato_b200 import batch

    prob, O, x0 = _problem(variant, factor)
    E = prob.engine
    x = problem.xdict_to_vector(helpers.perturbed(x0))
    g_want, pk_want = E.eval_pair_packed(x, 1)
    ev = batch.EngineRangeEvaluator(E)
    assert ev.n_blocks > 4
    xd = torch.from_numpy(x).cuda()
    for cuts in (1, 2, 5):
        hits = torch.zeros(E.n_rows + E.n_pack, dtype=torch.int64, device="cuda")
        out = torch.full((E.n_rows + E.n_pack,), float("nan"), dtype=torch.float64, device="cuda")
        for r in range(cuts):
            part = torch.full_like(out, float("nan"))
            ev.pair_range(xd, part[: E.n_rows], part[E.n_rows:], batch.split_range(ev.n_blocks, cuts, r),
                          batch.split_range(ev.n_vac, cuts, r))
            m = ~torch.isnan(part)
            hits += m
            out[m] = part[m]
        torch.cuda.synchronize()
        assert bool((hits == 1).all())
        assert np.array_equal(out[: E.n_rows].cpu().numpy(), g_want.ravel())
        assert np.array_equal(out[E.n_rows:].cpu().numpy(), pk_want.ravel())
    sp = batch.ShardedProblem(ev, x)
    g, pk = sp.pair(x)
    assert np.array_equal(g.cpu().numpy(), g_want.ravel()) and np.array_equal(pk.cpu().numpy(), pk_want.ravel())
    with pytest.raises(engine.GelatoError):
        E.eval_pair_packed_range_dev(xd.data_ptr(), out.data_ptr(), out.data_ptr(), (0, ev.n_blocks + 1), (0, 0))
    prob.close()


def test_recorded_solution_passes_the_termination_test_through_the_cuda_callbacks():
    """tests/test_solver_loop.py's warm-started solve with the CUDA engine behind the callbacks: from the recorded
    solution of the shipped example (tests/golden/example_solution.npz) two short penalty levels, then the
    termination test of the original problem -- same end point, bit for bit, as with the oracle's callbacks on this
    host; payload and the 13 event times within 1e-6 relative of the record."""
    from gelato_b200 import nlpshim, redsqp

    Lg = leaves.get("gmath")
    p, u, c, x0 = problem.problem_from_inputs(helpers.variant_inputs("example"), coord=Lg.coordinate_c, factor=1, max_nodes=20)
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "example_solution.npz"))
    xs = problem.vector_to_xdict(rec["x"], p["M"], p["N"], p["num_sections"])
    ends = []
    for objfunc, sens in ((lambda x: O.objfunc(x), lambda x, f=None: O.sens(x)), (prob.objfunc, prob.sens)):
        opt = nlpshim.register(objfunc, sens, helpers.copy_x(xs), c)
        sol = redsqp.ReducedSQP({"phase1_evals": 0, "penalty0": float(rec["multiplier"]) / 10.0 ** 0.5, "penalty_levels": 2,
                                 "level_iter": 50, "max_iter": 300})(opt, sens=sens)
        assert sol.constr_violation <= 1e-8 and len(sol.penalty_levels) == 2
        assert abs(sol.xStar["mass"][0] * u["mass"] / float(rec["payload_kg"]) - 1.0) <= 5e-6
        assert np.allclose(sol.xStar["t"] * u["t"], rec["event_times_s"], rtol=1e-5, atol=1e-5)
        ends.append((problem.xdict_to_vector(sol.xStar), sol.status, sol.nit))
    assert np.array_equal(ends[0][0], ends[1][0]) and ends[0][1:] == ends[1][1:]
    assert prob.engine.launches > 100
    prob.close()


def test_batched_solves_in_worker_processes_plumbing():
    """solve_batch.solve_dispersed_processes (the solve leg of bench.py): two dispersed scenarios, a worker process
    each with its own engine on GPU 0, a handful of solver iterations -- the record a rank reports, not a solve."""
    from gelato_b200 import solve_batch

    res = solve_batch.solve_dispersed_processes(helpers.example_inputs(), 2, 1, 0, device=0, iters=3, processes=2)
    assert res["scenarios"] == 2 and res["worker_processes"] == 2 and len(res["statuses"]) == 2
    assert res["converged"] == 0 and res["solves_per_hour"] == 0.0 and res["runs_per_hour"] > 0.0
    assert res["launches"] > 20 and res["userObjCalls_mean"] > 10 and res["userSensCalls_mean"] >= 2
    assert all(2.0e4 < p < 3.5e4 for p in res["payload_kg"]), res["payload_kg"]
