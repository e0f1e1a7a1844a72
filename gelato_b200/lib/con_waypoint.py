"""/root/reference/lib/con_waypoint.py:70-944 -- antenna elevation, IIP and position waypoints."""
from . import _jacobian, _value

inequality_antenna = _value("ineqcon_antenna")  # :70
inequality_jac_antenna = _jacobian("ineqcon_antenna")  # :108
equality_IIP = _value("eqcon_iip")  # :164
equality_jac_IIP = _jacobian("eqcon_iip")  # :243
inequality_IIP = _value("ineqcon_iip")  # :330
inequality_jac_IIP = _jacobian("ineqcon_iip")  # :384
equality_posLLH = _value("eqcon_pos")  # :507
equality_jac_posLLH = _jacobian("eqcon_pos")  # :610
inequality_posLLH = _value("ineqcon_pos")  # :717
inequality_jac_posLLH = _jacobian("ineqcon_pos")  # :786
