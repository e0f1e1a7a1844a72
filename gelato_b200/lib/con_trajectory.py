"""/root/reference/lib/con_trajectory.py:34-347 -- stage propellant, kick-turn sign, angular-rate patterns."""
from . import _jacobian, _length, _value

inequality_mass = _value("ineqcon_mass")  # :34 (a Python list, as in the reference)
inequality_jac_mass = _jacobian("ineqcon_mass")  # :64
inequality_kickturn = _value("ineqcon_kick")  # :106
inequality_jac_kickturn = _jacobian("ineqcon_kick")  # :127
equality_6DoF_rate = _value("eqcon_rate")  # :160
equality_length_6DoF_rate = _length("eqcon_rate")  # :210
equality_jac_6DoF_rate = _jacobian("eqcon_rate")  # :249
