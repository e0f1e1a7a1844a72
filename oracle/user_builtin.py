"""Oracle-side user constraints (TEST INFRASTRUCTURE ONLY).

The reference lets users plug arbitrary Python constraints in through
`user_constraints.py` (/root/reference/lib/con_user.py:33-42).  The shipped
example constrains the perigee radius at a named event
(/root/reference/example/user_constraints.py:120-139): a*(1-e)/6378137 - 1 from
`orbital_elements` of the state at the event's first node.  This is the
oracle's statement of that function, written against oracle/leaves.py.
"""


def perigee_ratio_at(leaves, event_name):
    crd = leaves.coordinate_c

    def equality_user(xdict, pdict, unitdict, condition):
        index = pdict["event_index"][event_name]
        a2 = pdict["ps_params"].index_start_u(index) + index
        pos = xdict["position"][a2 * 3 : (a2 + 1) * 3] * unitdict["position"]
        vel = xdict["velocity"][a2 * 3 : (a2 + 1) * 3] * unitdict["velocity"]
        elem = crd.orbital_elements(pos, vel)
        return (elem[0] * (1.0 - elem[1]) / 6378137.0) - 1.0

    return equality_user
