"""LGR mesh parameters (gelato_b200/psparams.py) against the reference's PSparams
(golden fixture produced from /root/reference/lib/SectionParameters.py) and against
closed-form identities."""
import os

import numpy as np
import pytest

import helpers
from gelato_b200.psparams import PSparams, lgr_diff_matrix, lgr_nodes


def test_matches_reference_golden_bitwise():
    npz = np.load(os.path.join(helpers.GOLDEN, "psparams.npz"))
    for n in range(2, 25):
        ps = PSparams([n])
        assert np.array_equal(ps.tau(0), npz["tau_%d" % n]), n
        assert np.array_equal(ps.D(0), npz["D_%d" % n]), n


@pytest.mark.parametrize("n", [2, 3, 5, 8, 16, 20, 32])
def test_differentiation_identities(n):
    tau = lgr_nodes(n)
    assert tau[-1] == 1.0 and np.all(np.diff(tau) > 0) and tau[0] > -1.0
    D = lgr_diff_matrix(n, tau)
    assert D.shape == (n, n + 1)
    support = np.hstack((-1.0, tau))
    assert np.max(np.abs(D @ np.ones(n + 1))) < 1e-11
    for k in range(1, min(n, 6) + 1):  # exact for polynomials of degree <= n
        np.testing.assert_allclose(D @ support**k, k * tau ** (k - 1), rtol=0, atol=1e-10)


def test_index_arithmetic():
    ps = PSparams([5, 5, 16, 8, 2])
    assert ps.get_index(0) == (0, 5, 0, 6, 5)
    assert ps.get_index(2) == (10, 26, 12, 29, 16)
    assert ps.index_start_x(4) == 38 and ps.num_u() == 36 and ps.num_x() == 41
    t = ps.time_nodes(1, 0.25, 0.5)
    assert t[0] == 0.25 and t[-1] == 0.5 and t.size == 6
    with pytest.raises(ValueError):
        ps.tau(5)
