"""the constraint Jacobian as one CSR matrix whose data is a gather of the kernel's flat value
vector (`plan.csr_map`, `GelatoProblem.jacobian_csr`), against the matrix scipy assembles from the ORACLE's `sens`
dictionaries the way a pyoptsparse-style driver would (stack the groups, sort the COO entries)."""
import numpy as np
import pytest
import scipy.sparse as sp

import helpers
from gelato_b200 import callbacks, nlpshim, problem
from gelato_b200.plan import GROUPS, VAR_ORDER
from oracle import leaves


def _emu_factory(plan):
    import emu_binding

    return emu_binding.EmuEngine(plan)


def _assemble(sens, funcs, sizes, wrt):
    col0 = dict(zip(VAR_ORDER, np.concatenate(([0], np.cumsum([sizes[k] for k in VAR_ORDER])[:-1]))))
    rows, cols, data, r0 = [], [], [], 0
    for key in GROUPS:
        if sens.get(key) is None:
            continue
        for var, blk in sens[key].items():
            if wrt is not None and var not in wrt[key]:
                continue
            if isinstance(blk, dict):
                r, c, d = blk["coo"]
            else:
                d = np.asarray(blk, dtype=float).ravel()
                r, c = np.zeros(d.size, dtype=int), np.arange(d.size)
            rows.append(np.asarray(r) + r0)
            cols.append(np.asarray(c) + col0[var])
            data.append(np.asarray(d, dtype=float))
        r0 += np.size(funcs[key])
    n = sum(sizes[k] for k in VAR_ORDER)
    return sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, n))


@pytest.mark.parametrize("variant", ["example", "waypoints", "all_aero", "iip_orbital"])
@pytest.mark.parametrize("use_wrt", [False, True])
def test_csr_matrix_equals_the_assembled_dictionaries(variant, use_wrt):
    Lg = leaves.get("gmath")
    inp = helpers.variant_inputs(variant)
    p, u, c, x0 = problem.problem_from_inputs(inp, coord=Lg.coordinate_c)
    x = helpers.perturbed(x0)
    O = helpers.oracle_nlp(p, u, c, "gmath", "seqfma")
    prob = callbacks.GelatoProblem(p, u, c, user_eq=callbacks.PerigeeAtEvent(helpers.USER_EVENT), coord=Lg.coordinate_c,
                                   engine_factory=_emu_factory)
    wrt = dict(nlpshim.WRT) if use_wrt else None
    J = prob.jacobian_csr(x, wrt=wrt)
    xa = helpers.copy_x(x)
    fo, _ = O.objfunc(xa)
    so, _ = O.sens(xa)
    want = _assemble(so, fo, prob.plan.sizes, wrt)
    assert J.shape == want.shape and J.has_sorted_indices
    # same structure (explicit zeros included: the sparsity never changes between calls) and the same bits
    order = np.lexsort((want.col, want.row))
    assert np.array_equal(J.indices, want.col[order])
    assert np.array_equal(np.repeat(np.arange(J.shape[0]), np.diff(J.indptr)), want.row[order])
    assert np.array_equal(J.data.view(np.uint64), want.data[order].view(np.uint64))
    # a second evaluation reuses the compiled structure
    x2 = helpers.perturbed(x0, seed=11)
    J2 = prob.jacobian_csr(x2, wrt=wrt)
    assert J2.indices is not None and np.array_equal(J2.indptr, J.indptr) and not np.array_equal(J2.data, J.data)
    prob.close()
