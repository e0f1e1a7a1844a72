"""/root/reference/lib/con_aero.py:89-756 -- angle-of-attack, dynamic-pressure and Q-alpha limits."""
from . import _jacobian, _length, _value

inequality_max_alpha = _value("ineqcon_alpha")  # :89
inequality_max_q = _value("ineqcon_q")  # :147
inequality_max_qalpha = _value("ineqcon_qalpha")  # :196
inequality_length_max_alpha = _length("ineqcon_alpha")  # :251
inequality_length_max_q = _length("ineqcon_q")  # :271
inequality_length_max_qalpha = _length("ineqcon_qalpha")  # :291
inequality_jac_max_alpha = _jacobian("ineqcon_alpha")  # :374
inequality_jac_max_q = _jacobian("ineqcon_q")  # :518
inequality_jac_max_qalpha = _jacobian("ineqcon_qalpha")  # :661
