/* output.h -- the derived quantities of one state node of a solved trajectory, for the batched
 * result-table kernel (gelato_leaf_output_table): what the loop body of the reference's
 * output_result computes (/root/reference/output_result.py:130-261) through ~30 leaf calls per node.
 * One thread per (scenario, node); same device functions as the NLP kernels (physics.h, gmath.h).
 *
 * numpy pieces of that loop are evaluated the way numpy evaluates them on 3-vectors: linalg.norm =
 * sqrt(x.dot(x)) with the dot product accumulated by fused multiply-add in index order, interp =
 * slope * (x - xp[j]) + fp[j], math.degrees(x) = x * (180 / pi), x ** 2 = x * x.
 */
#ifndef GELATO_B200_OUTPUT_H_
#define GELATO_B200_OUTPUT_H_

#include <math.h>

#include "physics.h"

#define GO_COLS 34
enum {
  GO_THRUST = 0, GO_LAT, GO_LON, GO_LAT_IIP, GO_LON_IIP, GO_DOWNRANGE, GO_ALTITUDE, GO_APOGEE, GO_PERIGEE,
  GO_INCLINATION, GO_ARG_PERIGEE, GO_ASC_NODE, GO_TRUE_ANOMALY, GO_VGN_X, GO_VGN_Y, GO_VGN_Z, GO_ACCEL_BODY_X,
  GO_AERO_BODY_X, GO_HEADING, GO_PITCH, GO_ROLL, GO_FLIGHTPATH, GO_AZIMUTH, GO_TDIR_X, GO_TDIR_Y, GO_TDIR_Z,
  GO_VEL_GROUND, GO_VEL_AIR, GO_AOA_TOTAL, GO_AOA_PITCH, GO_AOA_YAW, GO_DYNP, GO_Q_ALPHA, GO_MACH
};

P_HD double np_degrees(double x) { return x * (180.0 / P_PI); }
/* Earth.cpp:75-154 distance_vincenty (radians in, metres out) */
P_HD double distance_vincenty(double lat1, double lon1, double lat2, double lon2) {
  if (lat1 == lat2 && lon1 == lon2) return 0.0;
  const double U1 = gm_atan((1.0 - P_F) * gm_tan(lat1));
  const double U2 = gm_atan((1.0 - P_F) * gm_tan(lat2));
  const double diff_lon = lon2 - lon1;
  double sU1, cU1, sU2, cU2;
  gm_sincos(U1, &sU1, &cU1);
  gm_sincos(U2, &sU2, &cU2);
  double sin_sigma = 0.0, cos_sigma = 0.0, sigma = 0.0, sin_alpha = 0.0, cos_alpha = 0.0, cos_2sigma_m = 0.0, coeff;
  double lamda = diff_lon;
  for (int i = 0; i < 100; ++i) {
    double sl, cl;
    gm_sincos(lamda, &sl, &cl);
    const double a = cU2 * sl, b = cU1 * sU2 - sU1 * cU2 * cl;
    sin_sigma = gm_sqrt(a * a + b * b);
    cos_sigma = sU1 * sU2 + cU1 * cU2 * cl;
    sigma = gm_atan2(sin_sigma, cos_sigma);
    sin_alpha = cU1 * cU2 * sl / sin_sigma;
    cos_alpha = gm_sqrt(1.0 - sin_alpha * sin_alpha);
    cos_2sigma_m = cos_sigma - 2.0 * sU1 * sU2 / (cos_alpha * cos_alpha);
    coeff = P_F / 16.0 * (cos_alpha * cos_alpha) * (4.0 + P_F * (4.0 - 3.0 * (cos_alpha * cos_alpha)));
    const double lamda_itr = lamda;
    lamda = diff_lon + (1.0 - coeff) * P_F * sin_alpha *
                           (sigma + coeff * sin_sigma * (cos_2sigma_m + coeff * cos_sigma * (-1.0 + 2.0 * cos_2sigma_m)));
    if (gm_fabs(lamda - lamda_itr) < 1e-12) break;
  }
  const double u_squr = (cos_alpha * cos_alpha) * ((P_RA * P_RA) - (P_RB * P_RB)) / (P_RB * P_RB);
  const double A = 1.0 + u_squr / 16384.0 * (4096.0 + u_squr * (-768.0 + u_squr * (320.0 - 175.0 * u_squr)));
  const double B = u_squr / 1024.0 * (256.0 + u_squr * (-128.0 + u_squr * (74.0 - 47.0 * u_squr)));
  const double delta_sigma =
      B * sin_sigma *
      (cos_2sigma_m + 0.25 * B *
                          (cos_sigma * (-1.0 + 2.0 * (cos_2sigma_m * cos_2sigma_m)) -
                           (1.0 / 6.0) * B * cos_2sigma_m * (-3.0 + 4.0 * (sin_sigma * sin_sigma)) *
                               (-3.0 + 4.0 * (cos_2sigma_m * cos_2sigma_m))));
  return P_RB * A * (sigma - delta_sigma);
}

/* Eigen's eulerAngles(2, 1, 0) of the rotation matrix m (i = 2, j = 1, k = 0, odd) + the reference's range fix-ups
 * (Coordinate.cpp:129-145, :149-163), radians */
P_HD Vec3 euler_from_matrix(double m00, double m01, double m02, double m10, double m11, double m12, double m20, double m21,
                            double m22) {
  double r0 = gm_atan2(m10, m00), r1;
  const double c2 = gm_sqrt(m22 * m22 + m21 * m21);
  if (r0 < 0.0) {
    r0 += P_PI;
    r1 = gm_atan2(-m20, -c2);
  } else {
    r1 = gm_atan2(-m20, c2);
  }
  double s1, c1;
  gm_sincos(r0, &s1, &c1);
  double r2 = gm_atan2(s1 * m02 - c1 * m12, c1 * m11 - s1 * m01);
  if (gm_fabs(r1) > P_PI / 2.0) {
    r1 = (r1 > 0.0) ? P_PI - r1 : -P_PI - r1;
    r0 = P_PI + r0;
    r2 = P_PI + r2;
  }
  r0 = fmod(r0, 2.0 * P_PI);
  if (r0 < 0.0) r0 += 2.0 * P_PI;
  r2 = fmod(r2 + P_PI, 2.0 * P_PI) - P_PI;
  return v3(r0, r1, r2);
}

/* Coordinate.cpp:127-146 euler_from_quat (radians): Eigen toRotationMatrix().eulerAngles(2, 1, 0) + range fix-ups */
P_HD Vec3 euler_from_quat(Quat q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  const double m00 = 1.0 - (tyy + tzz), m10 = txy + twz, m11 = 1.0 - (txx + tzz);
  const double m01 = txy - twz, m02 = txz + twy, m12 = tyz - twx;
  const double m20 = txz - twy, m21 = tyz + twx, m22 = 1.0 - (txx + tyy);
  return euler_from_matrix(m00, m01, m02, m10, m11, m12, m20, m21, m22);
}

/* Coordinate.cpp:75-106 quat_eci2ned(pos_eci, t) */
P_HD Quat quat_eci2ned(Vec3 pos_eci, double t) {
  const double wt = P_OMEGA * t;
  double s, c;
  gm_sincos(wt, &s, &c);
  return quatconj(quat_ned2eci_cs(pos_eci, wt, c, s));
}

/* wrapper_utils.hpp:125-148 angle_of_attack_ab_rad: (pitch-plane, yaw-plane) angles of the air-relative
 * velocity in body axes; dimensional inputs, t in seconds */
P_HD void angle_of_attack_ab(Vec3 pos, Vec3 vel, Quat q, double t, const Tables& tb, double* out) {
  double pp[PP_COLS], rp[RP_COLS];
  pos_part(pos.x, pos.y, pos.z, tb.wind, tb.n_wind, 0, pp);
  rot_part(pos.x, pos.y, pos.z, t, pp[PP_WIND_N], pp[PP_WIND_E], rp);
  const Vec3 vb = quatrot(q, air_velocity(pos, vel, rp));
  if (vb.x < 1e-6) {
    out[0] = 0.0;
    out[1] = 0.0;
  } else {
    out[0] = gm_atan2(vb.z, vb.x);
    out[1] = gm_atan2(vb.y, vb.x);
  }
}

/* one row of the table; quat_raw as stored in x (normalised here like output_result.py:133) */
P_HD void output_row(double mass, Vec3 pos, Vec3 vel, Quat quat_raw, double t, double thrust_vac, double air_area,
                     double nozzle_area, const Tables& tb, double lat0, double lon0, double* out) {
  const double qn = gm_sqrt(((quat_raw.w * quat_raw.w + quat_raw.x * quat_raw.x) + quat_raw.y * quat_raw.y) +
                            quat_raw.z * quat_raw.z);
  const Quat quat = q4(quat_raw.w / qn, quat_raw.x / qn, quat_raw.y / qn, quat_raw.z / qn);
  double s, c;
  gm_sincos(P_OMEGA * t, &s, &c);
  const Vec3 pos_ecef = rot_eci2ecef(pos, c, s);
  const Geodetic g = ecef2geodetic<3>(pos_ecef);
  const double lat_deg = g.lat * 180.0 / P_PI, lon_deg = g.lon * 180.0 / P_PI;
  out[GO_LAT] = lat_deg; out[GO_LON] = lon_deg; out[GO_ALTITUDE] = g.alt;
  out[GO_DOWNRANGE] = distance_vincenty(lat0 * P_PI / 180.0, lon0 * P_PI / 180.0, lat_deg * P_PI / 180.0, lon_deg * P_PI / 180.0);
  double el[6];
  orbital_elements6(pos, vel, el);
  for (int k = 2; k < 6; k++) el[k] = el[k] * 180.0 / P_PI;
  out[GO_APOGEE] = el[0] * (1.0 + el[1]) - 6378137;
  out[GO_PERIGEE] = el[0] * (1.0 - el[1]) - 6378137;
  out[GO_INCLINATION] = el[2]; out[GO_ASC_NODE] = el[3]; out[GO_ARG_PERIGEE] = el[4]; out[GO_TRUE_ANOMALY] = el[5];
  /* atmosphere and wind at the geopotential altitude of the ECEF geodetic height (:151,190,223-224) */
  double pp[PP_COLS];
  pos_part(pos_ecef.x, pos_ecef.y, pos_ecef.z, tb.wind, tb.n_wind, PW_SOUND, pp);
  const Vec3 vel_ground_ecef = vel_eci2ecef_cs(vel, pos, c, s);
  const Quat q_ecef2ned = quat_ecef2ned_ll(g.lat, g.lon);
  const Vec3 vel_ground_ned = quatrot(q_ecef2ned, vel_ground_ecef);
  out[GO_VGN_X] = vel_ground_ned.x; out[GO_VGN_Y] = vel_ground_ned.y; out[GO_VGN_Z] = vel_ground_ned.z;
  const Quat q_eci2ned = quat_eci2ned(pos, t);
  const Vec3 vel_ned = quatrot(q_eci2ned, vel);
  const Vec3 vel_air_ned = sub3(vel_ground_ned, v3(pp[PP_WIND_N], pp[PP_WIND_E], 0.0));
  out[GO_VEL_GROUND] = np_norm3(vel_ground_ecef);
  out[GO_AZIMUTH] = np_degrees(gm_atan2(vel_ned.y, vel_ned.x));
  out[GO_FLIGHTPATH] = np_degrees(gm_asin(-vel_ned.z / np_norm3(vel_ned)));
  const double nva = np_norm3(vel_air_ned);
  const double q = 0.5 * (nva * nva) * pp[PP_RHO];
  out[GO_DYNP] = q;
  /* the two angle-of-attack leaves are whole leaf calls of their own (wrapper_utils.hpp:89-148) */
  const double aoa_all_deg = aero_quantity(0, pos, vel, quat, t, tb) * 180.0 / P_PI;
  double ab[2];
  angle_of_attack_ab(pos, vel, quat, t, tb, ab);
  out[GO_AOA_TOTAL] = aoa_all_deg;
  out[GO_Q_ALPHA] = aoa_all_deg * q;
  out[GO_AOA_PITCH] = ab[0] * 180.0 / P_PI;
  out[GO_AOA_YAW] = ab[1] * 180.0 / P_PI;
  const Vec3 tdir = quatrot(quatconj(quat), v3(1.0, 0.0, 0.0));
  out[GO_TDIR_X] = tdir.x; out[GO_TDIR_Y] = tdir.y; out[GO_TDIR_Z] = tdir.z;
  const Vec3 eul = euler_from_quat(quatmult(quatconj(q_eci2ned), quat)); /* quat_nedg2body, :171-174 */
  out[GO_HEADING] = eul.x * 180.0 / P_PI; out[GO_PITCH] = eul.y * 180.0 / P_PI; out[GO_ROLL] = eul.z * 180.0 / P_PI;
  /* air-relative velocity with the wind of THIS altitude (:229-233) */
  double rp[RP_COLS];
  rot_part(pos.x, pos.y, pos.z, t, pp[PP_WIND_N], pp[PP_WIND_E], rp);
  const Vec3 va = air_velocity(pos, vel, rp);
  const double nv = np_norm3(va);
  const double mach = nv / pp[PP_SOUND];
  out[GO_MACH] = mach;
  const double coeff = np_interp(mach, tb.ca, tb.ca + 1, tb.n_ca, 2);
  out[GO_VEL_AIR] = nv;
  const double k = 0.5 * pp[PP_RHO] * nv;
  const Vec3 aero = v3(k * -va.x * air_area * coeff, k * -va.y * air_area * coeff, k * -va.z * air_area * coeff);
  const Vec3 aero_body = quatrot(quat, aero);
  const double thrust_n = thrust_vac - nozzle_area * pp[PP_PRESS];
  out[GO_THRUST] = thrust_n;
  out[GO_AERO_BODY_X] = aero_body.x;
  out[GO_ACCEL_BODY_X] = (thrust_n + aero_body.x) / mass;
  const Vec3 vel_ecef = vel_ground_ecef;
  const Vec3 iip = iip_faa_deg(pos_ecef, vel_ecef); /* fill_na = False: NaN when there is no impact point */
  if (iip.x == 0.0 && iip.y == 0.0 && iip.z == 0.0) {
    out[GO_LAT_IIP] = out[GO_LON_IIP] = gm_nan();
  } else {
    out[GO_LAT_IIP] = iip.x; out[GO_LON_IIP] = iip.y;
  }
}

#endif /* GELATO_B200_OUTPUT_H_ */
