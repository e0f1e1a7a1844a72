"""/root/reference/lib/con_init_terminal_knot.py:41-452 -- initial state, event times, knots, terminal orbit."""
from . import _jacobian, _value

equality_init = _value("eqcon_init")  # :41
equality_jac_init = _jacobian("eqcon_init")  # :60
equality_time = _value("eqcon_time")  # :124
equality_jac_time = _jacobian("eqcon_time")  # :148
equality_knot_LGR = _value("eqcon_knot")  # :174
equality_jac_knot_LGR = _jacobian("eqcon_knot")  # :248
equality_6DoF_LGR_terminal = _value("eqcon_terminal")  # :329
equality_jac_6DoF_LGR_terminal = _jacobian("eqcon_terminal")  # :375
inequality_time = _value("ineqcon_time")  # :408
inequality_jac_time = _jacobian("ineqcon_time")  # :424
