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


def orbit_rows_at(leaves, event_name, rows):
    """The user function a gelato_b200.plan.OrbitAtEvent stands for, written the way a GELATO user would write it
    in user_constraints.py: rows of quantity / scale - offset from `orbital_elements` (a, e, inclination ...),
    `orbit_energy` and `angular_momentum` of the state at the event's first node.  One row returns a scalar."""
    crd = leaves.coordinate_c

    def value(q, pos, vel):
        if q == "orbit_energy":
            return crd.orbit_energy(pos, vel)
        if q == "angular_momentum":
            return crd.angular_momentum(pos, vel)
        elem = crd.orbital_elements(pos, vel)
        return {"perigee_radius": lambda: elem[0] * (1.0 - elem[1]), "apogee_radius": lambda: elem[0] * (1.0 + elem[1]),
                "semi_major_axis": lambda: elem[0], "eccentricity": lambda: elem[1], "inclination_deg": lambda: elem[2]}[q]()

    def user(xdict, pdict, unitdict, condition):
        import numpy as np

        index = pdict["event_index"][event_name]
        a2 = pdict["ps_params"].index_start_u(index) + index
        pos = xdict["position"][a2 * 3 : (a2 + 1) * 3] * unitdict["position"]
        vel = xdict["velocity"][a2 * 3 : (a2 + 1) * 3] * unitdict["velocity"]
        out = [(value(q, pos, vel) / scale) - offset for q, scale, offset in rows]
        return out[0] if len(out) == 1 else np.array(out)

    return user
