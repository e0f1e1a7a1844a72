"""/root/reference/lib/cost_gradient.py:29-47 -- objective and its constant gradient.

The objective is one signed element of the decision vector: minus the initial mass in
"Payload" mode (maximise lift-off mass), the final time otherwise; its gradient is the
matching signed unit vector.  No kernel is involved here (the fused residual kernel writes the
same value into g[0] for the objfunc path)."""
import numpy as np

# OptimizationMode == "Payload" -> (variable group, element, sign)
_OBJECTIVE = {True: ("mass", 0, -1.0), False: ("t", -1, 1.0)}


def _objective(condition):
    return _OBJECTIVE[condition["OptimizationMode"] == "Payload"]


def cost_6DoF(xdict, condition):
    group, element, sign = _objective(condition)
    return sign * xdict[group][element]


def cost_jac(xdict, condition):
    group, element, sign = _objective(condition)
    grad = np.zeros(xdict[group].size)
    grad[element] = sign
    return {group: grad}
