"""/root/reference/lib/con_user.py:33-42 -- user constraints.  Only registered built-ins run on the
GPU (gelato_b200.lib.configure(user_eq=PerigeeAtEvent(...))); without one the group is absent (None),
like a user_constraints.py that returns None."""
from . import _jacobian, _value

equality_user = _value("eqcon_user")
inequality_user = _value("ineqcon_user")
equality_jac_user = _jacobian("eqcon_user")  # :33
inequality_jac_user = _jacobian("ineqcon_user")  # :39
