"""/root/reference/lib/con_dynamics.py:34-632 -- collocation defects and their Jacobians."""
from . import _jacobian, _value

equality_dynamics_mass = _value("eqcon_dyn_mass")  # :34
equality_jac_dynamics_mass = _jacobian("eqcon_dyn_mass")  # :66
equality_dynamics_position = _value("eqcon_dyn_pos")  # :116
equality_jac_dynamics_position = _jacobian("eqcon_dyn_pos")  # :155
equality_dynamics_velocity = _value("eqcon_dyn_vel")  # :216
equality_jac_dynamics_velocity = _jacobian("eqcon_dyn_vel")  # :292
equality_dynamics_quaternion = _value("eqcon_dyn_quat")  # :499
equality_jac_dynamics_quaternion = _jacobian("eqcon_dyn_quat")  # :536
