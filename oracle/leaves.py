"""ctypes front-end of the CPU oracle's physics leaves (TEST INFRASTRUCTURE ONLY).

Presents oracle/_build/liboracle_{libm,gmath}.so under the module and function
names of the reference's five pybind11 extension modules
(/root/reference/src/pybind_dynamics.cpp:108-114, pybind_utils.cpp:28-48,
pybind_coordinate.cpp:28-78, pybind_IIP.cpp:53-57,
pybind_USStandardAtmosphere.cpp:28-35) with the same by-value semantics
(copy in, fresh ndarray out), so that the reference's own Python layer can be
driven on top of it in this container (tests/golden/make_golden.py) and so that
oracle/nlp.py reads like the reference.

Nothing under gelato_b200/ may import this module.
"""
import ctypes
import os
import subprocess
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_D = ctypes.c_double
_I = ctypes.c_int
_P = ctypes.POINTER(ctypes.c_double)


_REF_DIR = os.path.join(_HERE, "_ref")
REF_LIB = os.path.join(_REF_DIR, "libref_leaves.so")
REF_SRC = os.environ.get("GELATO_REFERENCE_SRC", "/root/reference/src")


def build(force=False):
    """Compile both oracle flavours with the committed Makefile and, where the reference
    tree is present (the build container; never the GPU box), the reference's own C++
    sources against oracle/ref_shim into oracle/_ref/ (flavour "ref")."""
    libs = [os.path.join(_BUILD, "liboracle_%s.so" % f) for f in ("libm", "gmath")]
    # make decides from the time stamps (gmath.h is a dependency of the gmath flavour): a no-op when up to date
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    if os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REF_SRC=" + REF_SRC] + (["-B"] if force else []))
    return libs


def ref_available():
    """True when oracle/_ref/libref_leaves.so (the shim-compiled reference C++) exists."""
    return os.path.exists(REF_LIB)


def _a(x, shape=None):
    arr = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None:
        arr = arr.reshape(shape)
    return arr


def _p(arr):
    return arr.ctypes.data_as(_P)


class OracleLeaves:
    """One flavour of the leaf library: "libm" / "gmath" = the hand-written restatement
    (oracle_leaves.cpp) on glibc / gmath elementary functions; "ref" = the reference's own
    C++ sources compiled against oracle/ref_shim (oracle/_ref/, glibc)."""

    def __init__(self, flavour="libm"):
        assert flavour in ("libm", "gmath", "ref")
        build()
        self.flavour = flavour
        if flavour == "ref":
            if not ref_available():
                raise FileNotFoundError("%s: build it where /root/reference exists (make -C oracle ref)" % REF_LIB)
            self.lib = ctypes.CDLL(REF_LIB)
        else:
            self.lib = ctypes.CDLL(os.path.join(_BUILD, "liboracle_%s.so" % flavour))
        L = self.lib
        for name in (
            "o_geopotential_altitude o_airtemperature_at o_airpressure_at o_airdensity_at "
            "o_speed_of_sound"
        ).split():
            getattr(L, name).restype = _D
            getattr(L, name).argtypes = [_D]
        for name in "o_angular_momentum o_inclination_cosine o_inclination_rad o_orbit_energy".split():
            getattr(L, name).restype = _D
            getattr(L, name).argtypes = [_P, _P]
        for name in "o_angular_momentum_from_altitude o_orbit_energy_from_altitude".split():
            getattr(L, name).restype = _D
            getattr(L, name).argtypes = [_D, _D]
        L.o_distance_vincenty.restype = _D
        L.o_distance_vincenty.argtypes = [_D, _D, _D, _D]
        L.o_haversine.restype = _D
        L.o_haversine.argtypes = [_D, _D, _D, _D, _D]
        L.o_interp.restype = _D
        L.o_interp.argtypes = [_D, _P, _P, _I]
        assert L.oracle_flavour() == {"libm": 0, "gmath": 1, "ref": 2}[flavour]
        assert L.oracle_unfused_check() == 1, "oracle built with FP contraction"
        self._make_modules()

    # -- helpers -----------------------------------------------------------
    def _v(self, fn, n_out, *args):
        out = np.empty(n_out)
        fn(*args, _p(out))
        return out

    def _make_modules(self):
        L = self.lib
        ns = types.SimpleNamespace

        # USStandardAtmosphere_c
        self.USStandardAtmosphere_c = ns(
            geopotential_altitude=lambda z: L.o_geopotential_altitude(_D(z)),
            airtemperature_at=lambda h: L.o_airtemperature_at(_D(h)),
            airpressure_at=lambda h: L.o_airpressure_at(_D(h)),
            airdensity_at=lambda h: L.o_airdensity_at(_D(h)),
            speed_of_sound=lambda h: L.o_speed_of_sound(_D(h)),
        )

        # coordinate_c
        def v3(x):
            return _a(x, (3,))

        def v4(x):
            return _a(x, (4,))

        def normalize(v):
            v = _a(v).ravel()
            out = np.empty(v.size)
            L.o_normalize(_p(v), _I(v.size), _p(out))
            return out

        self.coordinate_c = ns(
            quatmult=lambda q, p: self._v(L.o_quatmult, 4, _p(v4(q)), _p(v4(p))),
            conj=lambda q: self._v(L.o_conj, 4, _p(v4(q))),
            normalize=normalize,
            quatrot=lambda q, v: self._v(L.o_quatrot, 3, _p(v4(q)), _p(v3(v))),
            ecef2geodetic=lambda x, y, z: self._v(L.o_ecef2geodetic, 3, _D(x), _D(y), _D(z)),
            geodetic2ecef=lambda lat, lon, alt: self._v(L.o_geodetic2ecef, 3, _D(lat), _D(lon), _D(alt)),
            ecef2eci=lambda a, t: self._v(L.o_ecef2eci, 3, _p(v3(a)), _D(t)),
            eci2ecef=lambda a, t: self._v(L.o_eci2ecef, 3, _p(v3(a)), _D(t)),
            vel_ecef2eci=lambda v, p, t: self._v(L.o_vel_ecef2eci, 3, _p(v3(v)), _p(v3(p)), _D(t)),
            vel_eci2ecef=lambda v, p, t: self._v(L.o_vel_eci2ecef, 3, _p(v3(v)), _p(v3(p)), _D(t)),
            quat_eci2ecef=lambda t: self._v(L.o_quat_eci2ecef, 4, _D(t)),
            quat_ecef2eci=lambda t: self._v(L.o_quat_ecef2eci, 4, _D(t)),
            quat_ecef2nedg=lambda p: self._v(L.o_quat_ecef2nedg, 4, _p(v3(p))),
            quat_nedg2ecef=lambda p: self._v(L.o_quat_nedg2ecef, 4, _p(v3(p))),
            quat_eci2nedg=lambda p, t: self._v(L.o_quat_eci2nedg, 4, _p(v3(p)), _D(t)),
            quat_nedg2eci=lambda p, t: self._v(L.o_quat_nedg2eci, 4, _p(v3(p)), _D(t)),
            quat_from_euler=lambda az, el, ro: self._v(L.o_quat_from_euler, 4, _D(az), _D(el), _D(ro)),
            gravity=lambda p: self._v(L.o_gravity, 3, _p(v3(p))),
            eci2geodetic=lambda p, t: self._v(L.o_eci2geodetic, 3, _p(v3(p)), _D(t)),
            orbital_elements=lambda p, v: self._v(L.o_orbital_elements, 6, _p(v3(p)), _p(v3(v))),
            euler_from_quat=lambda q: self._v(L.o_euler_from_quat, 3, _p(v4(q))),
            quat_nedg2body=lambda q, p, t: self._v(L.o_quat_nedg2body, 4, _p(v4(q)), _p(v3(p)), _D(t)),
            distance_vincenty=lambda a, b, c, d: L.o_distance_vincenty(_D(a), _D(b), _D(c), _D(d)),
            angular_momentum_vec=lambda p, v: self._v(L.o_angular_momentum_vec, 3, _p(v3(p)), _p(v3(v))),
            angular_momentum=lambda p, v: L.o_angular_momentum(_p(v3(p)), _p(v3(v))),
            inclination_rad=lambda p, v: L.o_inclination_rad(_p(v3(p)), _p(v3(v))),
            inclination_cosine=lambda p, v: L.o_inclination_cosine(_p(v3(p)), _p(v3(v))),
            orbit_energy=lambda p, v: L.o_orbit_energy(_p(v3(p)), _p(v3(v))),
            laplace_vector=lambda p, v: self._v(L.o_laplace_vector, 3, _p(v3(p)), _p(v3(v))),
            dcm_from_quat=lambda q: self._v(L.o_dcm_from_quat, 9, _p(v4(q))).reshape(3, 3),
            quat_from_dcm=lambda C: self._v(L.o_quat_from_dcm, 4, _p(_a(C).reshape(9))),
            euler_from_dcm=lambda C: self._v(L.o_euler_from_dcm, 3, _p(_a(C).reshape(9))),
            dcm_from_thrustvector=lambda p, th: self._v(L.o_dcm_from_thrustvector, 9, _p(v3(p)), _p(v3(th))).reshape(3, 3),
            angular_momentum_from_altitude=lambda ha, hp: L.o_angular_momentum_from_altitude(_D(ha), _D(hp)),
            orbit_energy_from_altitude=lambda ha, hp: L.o_orbit_energy_from_altitude(_D(ha), _D(hp)),
        )

        # utils_c
        def wind_ned(alt, wind):
            w = _a(wind)
            return self._v(L.o_wind_ned, 3, _D(alt), _p(w), _I(w.shape[0]))

        def _arr3(fn, with_quat):
            def f(pos, vel, *rest):
                pos = _a(pos).reshape(-1, 3)
                vel = _a(vel).reshape(-1, 3)
                n = pos.shape[0]
                if with_quat:
                    quat, t, wind = rest
                    quat = _a(quat).reshape(-1, 4)
                else:
                    t, wind = rest
                t = _a(t).reshape(-1)
                w = _a(wind)
                out = np.empty(n)
                if with_quat:
                    fn(_p(pos), _p(vel), _p(quat), _p(t), _I(n), _p(w), _I(w.shape[0]), _p(out))
                else:
                    fn(_p(pos), _p(vel), _p(t), _I(n), _p(w), _I(w.shape[0]), _p(out))
                return out

            return f

        aoa_arr = _arr3(L.o_angle_of_attack_all_array_rad, True)
        q_arr = _arr3(L.o_dynamic_pressure_array_pa, False)
        qa_arr = _arr3(L.o_q_alpha_array_pa_rad, True)

        def interp(x, xp, yp):
            xp = _a(xp).ravel()
            yp = _a(yp).ravel()
            return L.o_interp(_D(x), _p(xp), _p(yp), _I(xp.size))

        def aoa_ab(pos, vel, quat, t, wind):
            w = _a(wind)
            return self._v(L.o_angle_of_attack_ab_rad, 2, _p(_a(pos, (3,))), _p(_a(vel, (3,))), _p(_a(quat, (4,))), _D(t),
                           _p(w), _I(w.shape[0]))

        self.utils_c = ns(
            angle_of_attack_ab_rad=aoa_ab,
            interp=interp,
            wind_ned=wind_ned,
            haversine=lambda lon1, lat1, lon2, lat2, r: L.o_haversine(_D(lon1), _D(lat1), _D(lon2), _D(lat2), _D(r)),
            angle_of_attack_all_array_rad=aoa_arr,
            dynamic_pressure_array_pa=q_arr,
            q_alpha_array_pa_rad=qa_arr,
            angle_of_attack_all_rad=lambda p, v, q, t, w: float(aoa_arr(p, v, q, [t], w)[0]),
            dynamic_pressure_pa=lambda p, v, t, w: float(q_arr(p, v, [t], w)[0]),
            q_alpha_pa_rad=lambda p, v, q, t, w: float(qa_arr(p, v, q, [t], w)[0]),
        )

        # IIP_c
        def posLLH_IIP_FAA(posECEF, velECEF, fill_na=True, n_iter=5):
            return self._v(L.o_posLLH_IIP_FAA, 3, _p(v3(posECEF)), _p(v3(velECEF)), _I(1 if fill_na else 0))

        self.IIP_c = ns(posLLH_IIP_FAA=posLLH_IIP_FAA)

        # dynamics_c
        def dynamics_velocity(mass_e, pos, vel, quat, t, param, wind, ca, units):
            mass_e = _a(mass_e).ravel()
            n = mass_e.size
            pos = _a(pos).reshape(n, 3)
            vel = _a(vel).reshape(n, 3)
            quat = _a(quat).reshape(n, 4)
            t = _a(t).ravel()
            param = _a(param).ravel()
            wind = _a(wind)
            ca = _a(ca)
            units = _a(units).ravel()
            out = np.empty((n, 3))
            L.o_dynamics_velocity(_p(mass_e), _p(pos), _p(vel), _p(quat), _p(t), _I(n), _p(param), _p(wind),
                                  _I(wind.shape[0]), _p(ca), _I(ca.shape[0]), _p(units), _p(out))
            return out

        def dynamics_velocity_NoAir(mass_e, pos, quat, param, units):
            mass_e = _a(mass_e).ravel()
            n = mass_e.size
            pos = _a(pos).reshape(n, 3)
            quat = _a(quat).reshape(n, 4)
            param = _a(param).ravel()
            units = _a(units).ravel()
            out = np.empty((n, 3))
            L.o_dynamics_velocity_NoAir(_p(mass_e), _p(pos), _p(quat), _I(n), _p(param), _p(units), _p(out))
            return out

        def dynamics_quaternion(quat, u_e, unit_u):
            quat = _a(quat).reshape(-1, 4)
            n = quat.shape[0]
            u_e = _a(u_e).reshape(n, 2)
            out = np.empty((n, 4))
            L.o_dynamics_quaternion(_p(quat), _p(u_e), _D(unit_u), _I(n), _p(out))
            return out

        self.dynamics_c = ns(
            dynamics_velocity=dynamics_velocity,
            dynamics_velocity_NoAir=dynamics_velocity_NoAir,
            dynamics_quaternion=dynamics_quaternion,
        )

    def modules(self):
        """{name: namespace} for the five reference extension modules."""
        return {
            name: getattr(self, name)
            for name in ("dynamics_c", "utils_c", "coordinate_c", "IIP_c", "USStandardAtmosphere_c")
        }


_cache = {}


def get(flavour="libm"):
    if flavour not in _cache:
        _cache[flavour] = OracleLeaves(flavour)
    return _cache[flavour]
