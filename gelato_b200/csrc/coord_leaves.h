/* coord_leaves.h -- the reference's coordinate_c / utils_c leaf functions that are NOT on the NLP path, one function
 * code each, for the batch leaf entry point gelato_leaf_coordinate (one thread per item).  They serve the callers
 * around the solve (launch state, initial guess, result tables, user constraints).
 *
 * Reference: /root/reference/src/pybind_coordinate.cpp:28-78 (names and argument lists), wrapper_coordinate.hpp:50-265,
 * Coordinate.cpp:41-245, Earth.cpp:49-154 -- same operations in the same order as the oracle's restatement
 * (oracle/oracle_leaves.cpp), built from the device functions the kernels use.  host+device: tests/emu steps it.
 *
 * Item layout: a[GC_IN] and b[GC_IN] input vectors (unused entries ignored), t one scalar, out[GC_OUT].
 * Matrices (the four DCM helpers) travel row-major, nine values. */
#ifndef GELATO_B200_COORD_LEAVES_H_
#define GELATO_B200_COORD_LEAVES_H_

#include "../../include/gelato_b200.h"
#include "initguess.h"
#include "output.h"

#define GC_IN 9
#define GC_OUT 9
/* function codes GC_*: include/gelato_b200.h */

P_HD int coord_leaf_n_out(int fn) {
  switch (fn) {
    case GC_QUATMULT: case GC_CONJ: case GC_NORMALIZE4: case GC_QUAT_ECI2ECEF: case GC_QUAT_ECEF2ECI:
    case GC_QUAT_ECEF2NEDG: case GC_QUAT_NEDG2ECEF: case GC_QUAT_ECI2NEDG: case GC_QUAT_NEDG2ECI:
    case GC_QUAT_FROM_EULER: case GC_QUAT_NEDG2BODY: case GC_QUAT_FROM_DCM:
      return 4;
    case GC_DCM_FROM_QUAT: case GC_DCM_FROM_THRUSTVECTOR: return 9;
    case GC_ORBITAL_ELEMENTS: return 6;
    case GC_DISTANCE_VINCENTY: case GC_ANGMOM: case GC_INCLINATION_RAD: case GC_INCLINATION_COS: case GC_ORBIT_ENERGY:
    case GC_ANGMOM_FROM_ALT: case GC_ENERGY_FROM_ALT: case GC_HAVERSINE:
      return 1;
    default: return 3;
  }
}

/* Earth.cpp:63-71 geodetic2ecef (radians) */
P_HD Vec3 geodetic2ecef_rad(double lat, double lon, double alt) {
  double s0, c0, s1, c1;
  gm_sincos(lat, &s0, &c0);
  gm_sincos(lon, &s1, &c1);
  const double N = P_RA / gm_sqrt(1.0 - P_E2 * s0 * s0);
  return v3((N + alt) * c0 * c1, (N + alt) * c0 * s1, (N * (1.0 - P_E2) + alt) * s0);
}

/* Coordinate.cpp:85-98 quat_ecef2ned from an ECEF position */
P_HD Quat quat_ecef2ned_pos(Vec3 pos_ecef) {
  const Geodetic g = ecef2geodetic<2>(pos_ecef);
  return quat_ecef2ned_ll<>(g.lat, g.lon);
}

/* Coordinate.cpp:112-123 quat_from_euler (degrees): AngleAxis(az, Z) * AngleAxis(el, Y) * AngleAxis(ro, X) */
P_HD Quat quat_from_euler_deg(double az_deg, double el_deg, double ro_deg) {
  const double az = az_deg * P_PI / 180.0, el = el_deg * P_PI / 180.0, ro = ro_deg * P_PI / 180.0;
  double sz, cz, sy, cy, sx, cx;
  gm_sincos(0.5 * az, &sz, &cz);
  gm_sincos(0.5 * el, &sy, &cy);
  gm_sincos(0.5 * ro, &sx, &cx);
  return eigen_quat_prod(eigen_quat_prod(q4(cz, 0.0, 0.0, sz), q4(cy, 0.0, sy, 0.0)), q4(cx, sx, 0.0, 0.0));
}

/* Eigen's normalized(): v / sqrt(squaredNorm) when the squared norm is positive, else v */
P_HD Vec3 eigen_normalized3(Vec3 a) {
  const double z = dot3(a, a);
  if (z > 0.0) {
    const double n = gm_sqrt(z);
    return v3(a.x / n, a.y / n, a.z / n);
  }
  return a;
}

P_HD void coord_leaf(int fn, const double* a, const double* b, double t, double* out) {
  const Vec3 a3 = v3(a[0], a[1], a[2]), b3 = v3(b[0], b[1], b[2]);
  const Quat a4 = q4(a[0], a[1], a[2], a[3]), b4 = q4(b[0], b[1], b[2], b[3]);
  Vec3 r3 = v3(0.0, 0.0, 0.0);
  Quat r4 = q4(0.0, 0.0, 0.0, 0.0);
  double s, c;
  switch (fn) {
    case GC_QUATMULT: r4 = quatmult(a4, b4); break;
    case GC_CONJ: r4 = quatconj(a4); break;
    case GC_NORMALIZE3:
    case GC_NORMALIZE4: {
      const int n = fn == GC_NORMALIZE3 ? 3 : 4;
      for (int i = 0; i < n; i++) out[i] = a[i];
      init_normalize(out, n);
      return;
    }
    case GC_QUATROT: r3 = quatrot(a4, b3); break;
    case GC_ECEF2GEODETIC: {
      const Geodetic g = ecef2geodetic<3>(a3);
      r3 = v3(g.lat * 180.0 / P_PI, g.lon * 180.0 / P_PI, g.alt);
      break;
    }
    case GC_GEODETIC2ECEF: r3 = geodetic2ecef_rad(a[0] * P_PI / 180.0, a[1] * P_PI / 180.0, a[2]); break;
    case GC_ECEF2ECI: gm_sincos(P_OMEGA * t, &s, &c); r3 = rot_ecef2eci(a3, c, s); break;
    case GC_ECI2ECEF: gm_sincos(P_OMEGA * t, &s, &c); r3 = rot_eci2ecef(a3, c, s); break;
    case GC_VEL_ECEF2ECI: gm_sincos(P_OMEGA * t, &s, &c); r3 = vel_ecef2eci_cs(a3, b3, c, s); break;
    case GC_VEL_ECI2ECEF: gm_sincos(P_OMEGA * t, &s, &c); r3 = vel_eci2ecef_cs(a3, b3, c, s); break;
    case GC_QUAT_ECI2ECEF:
    case GC_QUAT_ECEF2ECI:
      gm_sincos(P_OMEGA * t / 2.0, &s, &c);
      r4 = q4(c, 0.0, 0.0, s);
      if (fn == GC_QUAT_ECEF2ECI) r4 = quatconj(r4);
      break;
    case GC_QUAT_ECEF2NEDG: r4 = quat_ecef2ned_pos(a3); break;
    case GC_QUAT_NEDG2ECEF: r4 = quatconj(quat_ecef2ned_pos(a3)); break;
    case GC_QUAT_ECI2NEDG: r4 = quat_eci2ned(a3, t); break;
    case GC_QUAT_NEDG2ECI: r4 = quatconj(quat_eci2ned(a3, t)); break;
    case GC_QUAT_FROM_EULER: r4 = quat_from_euler_deg(a[0], a[1], a[2]); break;
    case GC_EULER_FROM_QUAT: {
      const Vec3 e = euler_from_quat(a4);
      r3 = v3(e.x * 180.0 / P_PI, e.y * 180.0 / P_PI, e.z * 180.0 / P_PI);
      break;
    }
    case GC_QUAT_NEDG2BODY: r4 = quatmult(quatconj(quat_eci2ned(b3, t)), a4); break;
    case GC_ORBITAL_ELEMENTS:
      orbital_elements6(a3, b3, out);
      for (int i = 2; i < 6; i++) out[i] = out[i] * 180.0 / P_PI;
      return;
    case GC_DISTANCE_VINCENTY:
      out[0] = distance_vincenty(a[0] * P_PI / 180.0, a[1] * P_PI / 180.0, a[2] * P_PI / 180.0, a[3] * P_PI / 180.0);
      return;
    case GC_ANGMOM_VEC: r3 = cross3(a3, b3); break;
    case GC_ANGMOM: out[0] = angular_momentum(a3, b3); return;
    case GC_INCLINATION_RAD: out[0] = inclination_rad(a3, b3); return;
    case GC_INCLINATION_COS: {
      const Vec3 h = cross3(a3, b3);
      out[0] = h.z / norm3(h);
      return;
    }
    case GC_ORBIT_ENERGY: out[0] = orbit_energy(a3, b3); return;
    case GC_ANGMOM_FROM_ALT: {
      const double ra = P_RA + a[0], rp = P_RA + a[1];
      const double sma = (ra + rp) / 2.0;
      out[0] = rp * gm_sqrt(P_MU * (2.0 / rp - 1.0 / sma));
      return;
    }
    case GC_ENERGY_FROM_ALT: {
      const double ra = P_RA + a[0], rp = P_RA + a[1];
      out[0] = -P_MU / 2.0 / ((ra + rp) / 2.0);
      return;
    }
    case GC_DCM_FROM_QUAT: { /* wrapper_coordinate.hpp:80-94 */
      const double* q = a;
      out[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
      out[1] = 2 * (q[1] * q[2] + q[0] * q[3]);
      out[2] = 2 * (q[1] * q[3] - q[0] * q[2]);
      out[3] = 2 * (q[1] * q[2] - q[0] * q[3]);
      out[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
      out[5] = 2 * (q[2] * q[3] + q[0] * q[1]);
      out[6] = 2 * (q[1] * q[3] + q[0] * q[2]);
      out[7] = 2 * (q[2] * q[3] - q[0] * q[1]);
      out[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
      return;
    }
    case GC_QUAT_FROM_DCM: { /* wrapper_coordinate.hpp:96-103 */
      const double q0 = 0.5 * gm_sqrt(1 + a[0] + a[4] + a[8]);
      r4 = q4(q0, (a[5] - a[7]) / (4 * q0), (a[6] - a[2]) / (4 * q0), (a[1] - a[3]) / (4 * q0));
      break;
    }
    case GC_EULER_FROM_DCM: { /* wrapper :182-186, Coordinate.cpp:147-164: the angles of C^T, degrees */
      const Vec3 e = euler_from_matrix(a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]);
      r3 = v3(e.x * 180.0 / P_PI, e.y * 180.0 / P_PI, e.z * 180.0 / P_PI);
      break;
    }
    case GC_DCM_FROM_THRUSTVECTOR: { /* wrapper :188-191 (pos_eci, thrustvec_eci), Coordinate.cpp:176-191 */
      const Vec3 xb = eigen_normalized3(b3), pn = eigen_normalized3(a3);
      const Vec3 yb = (1.0 - dot3(xb, pn) < 1.0e-10) ? eigen_normalized3(cross3(v3(0.0, 0.0, 1.0), xb))
                                                     : eigen_normalized3(cross3(xb, pn));
      const Vec3 zb = cross3(xb, yb);
      out[0] = xb.x; out[1] = xb.y; out[2] = xb.z;
      out[3] = yb.x; out[4] = yb.y; out[5] = yb.z;
      out[6] = zb.x; out[7] = zb.y; out[8] = zb.z;
      return;
    }
    case GC_LAPLACE_VECTOR: { /* wrapper_coordinate.hpp:238-244: v x h - mu r / |r| */
      const Vec3 h = cross3(a3, b3), vh = cross3(b3, h);
      const double rn = norm3(a3);
      r3 = v3(vh.x - P_MU * a3.x / rn, vh.y - P_MU * a3.y / rn, vh.z - P_MU * a3.z / rn);
      break;
    }
    case GC_HAVERSINE: { /* wrapper_utils.hpp:37-49: degrees in, radius t */
      const double lon1 = a[0] * P_PI / 180.0, lat1 = a[1] * P_PI / 180.0, lon2 = a[2] * P_PI / 180.0, lat2 = a[3] * P_PI / 180.0;
      const double dlon = lon2 - lon1, dlat = lat2 - lat1;
      const double h = gm_pow(gm_sin(dlat / 2), 2.0) + gm_cos(lat1) * gm_cos(lat2) * gm_pow(gm_sin(dlon / 2), 2.0);
      out[0] = 2 * t * gm_asin(gm_sqrt(h));
      return;
    }
    default: return;
  }
  if (coord_leaf_n_out(fn) == 4) {
    out[0] = r4.w; out[1] = r4.x; out[2] = r4.y; out[3] = r4.z;
  } else {
    out[0] = r3.x; out[1] = r3.y; out[2] = r3.z;
  }
}

#endif /* GELATO_B200_COORD_LEAVES_H_ */
