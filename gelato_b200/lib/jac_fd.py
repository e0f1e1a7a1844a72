"""/root/reference/lib/jac_fd.py:29-62 -- dense forward-difference Jacobian of a constraint function.

`con` is either one of this package's user-constraint functions (then the finite differences are
the ones the Jacobian kernel computed: GE_USER_PERIGEE jobs, jobs.h) or an arbitrary Python
callable.  The latter cannot run on the GPU; it is differenced on the host with the reference's
protocol -- key order of `xdict`, `x += dx`, evaluate, `(g_p - g_base) / dx`, `x -= dx` in place --
because that is user code, not part of the accelerated path."""
import numpy as np

from . import con_user


def jac_fd(con, xdict, pdict, unitdict, condition):
    if con is con_user.equality_user:
        return con_user.equality_jac_user(xdict, pdict, unitdict, condition)
    if con is con_user.inequality_user:
        return con_user.inequality_jac_user(xdict, pdict, unitdict, condition)
    dx = pdict["dx"]
    base = con(xdict, pdict, unitdict, condition)
    n_rows = len(base) if hasattr(base, "__len__") else 1
    out = {}
    for key, val in xdict.items():
        block = np.zeros((n_rows, val.size))
        for i in range(val.size):
            xdict[key][i] += dx
            block[:, i] = (con(xdict, pdict, unitdict, condition) - base) / dx
            xdict[key][i] -= dx
        out[key] = block
    return out
