"""/root/reference/lib/cost_gradient.py:29-47 -- objective and its constant gradient.  Both are a
single element of x and a unit vector; no kernel is involved (the fused residual kernel also
writes the objective into g[0] for the objfunc path)."""
import numpy as np


def cost_6DoF(xdict, condition):
    if condition["OptimizationMode"] == "Payload":
        return -xdict["mass"][0]
    return xdict["t"][-1]


def cost_jac(xdict, condition):
    if condition["OptimizationMode"] == "Payload":
        g = np.zeros(xdict["mass"].size)
        g[0] = -1.0
        return {"mass": g}
    g = np.zeros(xdict["t"].size)
    g[-1] = 1.0
    return {"t": g}
