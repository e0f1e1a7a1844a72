// oracle_leaves.cpp -- CPU restatement of GELATO's native physics leaves.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under gelato_b200/ may include, link or
// call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it (as the checker / CPU baseline).
//
// It restates, operation by operation and WITHOUT Eigen, the reference's
//   src/Air.cpp:47-111, src/Earth.cpp:41-154, src/Coordinate.cpp:41-110,197-245,
//   src/gravity.cpp:11-57, src/iip.cpp:36-150, src/pybind_IIP.cpp:34-51,
//   src/wrapper_coordinate.hpp:50-265, src/wrapper_utils.hpp:37-206,
//   src/pybind_dynamics.cpp:30-106
// (all paths under /root/reference).  Two flavours are built from this file:
//   liboracle_libm.so   elementary functions from glibc  (faithful to the reference)
//   liboracle_gmath.so  elementary functions from gelato_b200/csrc/gmath.h
//                       (-DORACLE_GMATH; the bit-twin of the CUDA path)
//
// Third-party arithmetic the reference gets from Eigen3 (unpinned,
// /root/reference/CMakeLists.txt:13) is restated as: 3-vector norm/dot =
// (a0 + a1) + a2 [Eigen's reduction order is not verifiable in this container],
// cross = Eigen's coefficient formula, Quaterniond product = Eigen's generic
// coefficient formula, normalized() = v / sqrt(squaredNorm).
// PARITY PIN: see oracle/README.md (shim-compiled reference sources + golden).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#ifdef ORACLE_GMATH
#include "../gelato_b200/csrc/gmath.h"
#define O_SIN(x) gm_sin(x)
#define O_COS(x) gm_cos(x)
#define O_TAN(x) gm_tan(x)
#define O_ATAN(x) gm_atan(x)
#define O_ATAN2(y, x) gm_atan2(y, x)
#define O_ASIN(x) gm_asin(x)
#define O_ACOS(x) gm_acos(x)
#define O_EXP(x) gm_exp(x)
#define O_POW(x, y) gm_pow(x, y)
#else
#define O_SIN(x) std::sin(x)
#define O_COS(x) std::cos(x)
#define O_TAN(x) std::tan(x)
#define O_ATAN(x) std::atan(x)
#define O_ATAN2(y, x) std::atan2(y, x)
#define O_ASIN(x) std::asin(x)
#define O_ACOS(x) std::acos(x)
#define O_EXP(x) std::exp(x)
#define O_POW(x, y) std::pow(x, y)
#endif
#define O_SQRT(x) std::sqrt(x)

namespace {

struct V3 {
  double v[3];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct V4 {
  double v[4];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

inline V3 mk3(double a, double b, double c) { return V3{{a, b, c}}; }
inline V4 mk4(double a, double b, double c, double d) { return V4{{a, b, c, d}}; }

// ---- Eigen stand-ins (see header comment) ----
inline double dot3(const V3& a, const V3& b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline double norm3(const V3& a) { return O_SQRT(dot3(a, a)); }
inline V3 cross3(const V3& a, const V3& b) {
  return mk3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline V3 sub3(const V3& a, const V3& b) { return mk3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 add3(const V3& a, const V3& b) { return mk3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 scale3(double s, const V3& a) { return mk3(s * a[0], s * a[1], s * a[2]); }
inline V3 div3(const V3& a, double s) { return mk3(a[0] / s, a[1] / s, a[2] / s); }
inline V3 normalized3(const V3& a) {
  double z = dot3(a, a);
  if (z > 0.0) return div3(a, O_SQRT(z));
  return a;
}
// Eigen::Quaterniond product a*b, coefficients (w,x,y,z)
inline V4 eigen_quat_prod(const V4& a, const V4& b) {
  return mk4(a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
             a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
             a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
             a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]);
}
inline V4 quat_conj(const V4& q) { return mk4(q[0], -q[1], -q[2], -q[3]); }

// ---- Earth (src/Earth.cpp:41-47) ----
const double E_mu = 3.986004418e14;
const double E_omega = 7.2921151467e-5;
const double E_Ra = 6378137.0;
const double E_f = 1.0 / 298.257223563;
const double E_Rb = E_Ra * (1.0 - E_f);
const double E_e2 = (E_Ra * E_Ra - E_Rb * E_Rb) / E_Ra / E_Ra;
const double E_ep2 = (E_Ra * E_Ra - E_Rb * E_Rb) / E_Rb / E_Rb;
inline double pow2(double x) { return x * x; }

// src/Earth.cpp:49-61 (radians)
V3 earth_ecef2geodetic(const V3& p_) {
  double p = O_SQRT(p_[0] * p_[0] + p_[1] * p_[1]);
  double theta = O_ATAN2(p_[2] * E_Ra, p * E_Rb);
  double st = O_SIN(theta), ct = O_COS(theta);
  double lat = O_ATAN2(p_[2] + E_ep2 * E_Rb * (st * st * st), p - E_e2 * E_Ra * (ct * ct * ct));
  double lon = O_ATAN2(p_[1], p_[0]);
  double sl = O_SIN(lat);
  double N = E_Ra / O_SQRT(1.0 - E_e2 * sl * sl);
  double alt = p / O_COS(lat) - N;
  return mk3(lat, lon, alt);
}

// src/Earth.cpp:63-71 (radians)
V3 earth_geodetic2ecef(const V3& g) {
  double s0 = O_SIN(g[0]), c0 = O_COS(g[0]);
  double N = E_Ra / O_SQRT(1.0 - E_e2 * s0 * s0);
  double x = (N + g[2]) * c0 * O_COS(g[1]);
  double y = (N + g[2]) * c0 * O_SIN(g[1]);
  double z = (N * (1.0 - E_e2) + g[2]) * s0;
  return mk3(x, y, z);
}

// src/Earth.cpp:75-154 (radians in, metres out; azimuth dropped by the wrapper)
double earth_distance_vincenty(double lat1, double lon1, double lat2, double lon2) {
  const int itr_limit = 100;
  if (lat1 == lat2 && lon1 == lon2) return 0.0;
  double U1 = O_ATAN((1.0 - E_f) * O_TAN(lat1));
  double U2 = O_ATAN((1.0 - E_f) * O_TAN(lat2));
  double diff_lon = lon2 - lon1;
  double sin_sigma = 0.0, cos_sigma = 0.0, sigma = 0.0, sin_alpha = 0.0, cos_alpha = 0.0;
  double cos_2sigma_m = 0.0, coeff = 0.0;
  double lamda = diff_lon;
  for (int i = 0; i < itr_limit; ++i) {
    sin_sigma = pow2(O_COS(U2) * O_SIN(lamda)) +
                pow2(O_COS(U1) * O_SIN(U2) - O_SIN(U1) * O_COS(U2) * O_COS(lamda));
    sin_sigma = O_SQRT(sin_sigma);
    cos_sigma = O_SIN(U1) * O_SIN(U2) + O_COS(U1) * O_COS(U2) * O_COS(lamda);
    sigma = O_ATAN2(sin_sigma, cos_sigma);
    sin_alpha = O_COS(U1) * O_COS(U2) * O_SIN(lamda) / sin_sigma;
    cos_alpha = O_SQRT(1.0 - pow2(sin_alpha));
    cos_2sigma_m = cos_sigma - 2.0 * O_SIN(U1) * O_SIN(U2) / pow2(cos_alpha);
    coeff = E_f / 16.0 * pow2(cos_alpha) * (4.0 + E_f * (4.0 - 3.0 * pow2(cos_alpha)));
    double lamda_itr = lamda;
    lamda = diff_lon + (1.0 - coeff) * E_f * sin_alpha *
                           (sigma + coeff * sin_sigma *
                                        (cos_2sigma_m + coeff * cos_sigma * (-1.0 + 2.0 * cos_2sigma_m)));
    if (std::abs(lamda - lamda_itr) < 1e-12) break;
  }
  double u_squr = pow2(cos_alpha) * (pow2(E_Ra) - pow2(E_Rb)) / pow2(E_Rb);
  double A = 1.0 + u_squr / 16384.0 * (4096.0 + u_squr * (-768.0 + u_squr * (320.0 - 175.0 * u_squr)));
  double B = u_squr / 1024.0 * (256.0 + u_squr * (-128.0 + u_squr * (74.0 - 47.0 * u_squr)));
  double delta_sigma =
      B * sin_sigma *
      (cos_2sigma_m + 0.25 * B *
                          (cos_sigma * (-1.0 + 2.0 * pow2(cos_2sigma_m)) -
                           (1.0 / 6.0) * B * cos_2sigma_m * (-3.0 + 4.0 * pow2(sin_sigma)) *
                               (-3.0 + 4.0 * pow2(cos_2sigma_m))));
  return E_Rb * A * (sigma - delta_sigma);
}

// ---- Coordinate (src/Coordinate.cpp) ----
V3 c_ecef2eci(const V3& a, double t) {  // :41-49
  double c = O_COS(E_omega * t), s = O_SIN(E_omega * t);
  return mk3(a[0] * c - a[1] * s, a[0] * s + a[1] * c, a[2]);
}
V3 c_eci2ecef(const V3& a, double t) {  // :51-59
  double c = O_COS(E_omega * t), s = O_SIN(E_omega * t);
  return mk3(a[0] * c + a[1] * s, -a[0] * s + a[1] * c, a[2]);
}
V3 c_vel_ecef2eci(const V3& vel_ecef, const V3& pos_ecef, double t) {  // :61-67
  V3 pos_eci = c_ecef2eci(pos_ecef, t);
  V3 vg = c_ecef2eci(vel_ecef, t);
  V3 w = mk3(0.0, 0.0, E_omega);
  return add3(vg, cross3(w, pos_eci));
}
V3 c_vel_eci2ecef(const V3& vel_eci, const V3& pos_eci, double t) {  // :69-73
  V3 w = mk3(0.0, 0.0, E_omega);
  return c_eci2ecef(sub3(vel_eci, cross3(w, pos_eci)), t);
}
V4 c_quat_eci2ecef(double t) {  // :75-79
  return mk4(O_COS(E_omega * t / 2.0), 0.0, 0.0, O_SIN(E_omega * t / 2.0));
}
V4 c_quat_ecef2ned(const V3& pos_ecef) {  // :85-98
  V3 g = earth_ecef2geodetic(pos_ecef);
  double c_hl = O_COS(g[1] / 2.0), s_hl = O_SIN(g[1] / 2.0);
  double c_hp = O_COS(g[0] / 2.0), s_hp = O_SIN(g[0] / 2.0);
  double r2 = O_SQRT(2.0);
  return mk4(c_hl * (c_hp - s_hp) / r2, s_hl * (c_hp + s_hp) / r2, -c_hl * (c_hp + s_hp) / r2,
             s_hl * (c_hp - s_hp) / r2);
}
V4 c_quat_eci2ned(const V3& pos_eci, double t) {  // :104-106
  return eigen_quat_prod(c_quat_eci2ecef(t), c_quat_ecef2ned(c_eci2ecef(pos_eci, t)));
}
V4 c_quat_ned2eci(const V3& pos_eci, double t) { return quat_conj(c_quat_eci2ned(pos_eci, t)); }  // :108-110

// :112-123 -- AngleAxis(az,Z)*AngleAxis(el,Y)*AngleAxis(ro,X) as quaternion products
V4 c_quat_from_euler_deg(double az_deg, double el_deg, double ro_deg) {
  double az = az_deg * M_PI / 180.0, el = el_deg * M_PI / 180.0, ro = ro_deg * M_PI / 180.0;
  V4 qz = mk4(O_COS(0.5 * az), 0.0, 0.0, O_SIN(0.5 * az));
  V4 qy = mk4(O_COS(0.5 * el), 0.0, O_SIN(0.5 * el), 0.0);
  V4 qx = mk4(O_COS(0.5 * ro), O_SIN(0.5 * ro), 0.0, 0.0);
  return eigen_quat_prod(eigen_quat_prod(qz, qy), qx);
}

// :127-146 euler_from_quat: Eigen q.toRotationMatrix().eulerAngles(2, 1, 0) + range fix-ups (radians).
// toRotationMatrix / eulerAngles are restated from Eigen's Quaternion.h / EulerAngles.h (as in ref_shim).
// eulerAngles(2, 1, 0) of a rotation matrix + the reference's range fix-ups (radians)
V3 c_euler_from_matrix(const double m[3][3]) {
  // eulerAngles(2, 1, 0): i = 2, odd = 1, j = 1, k = 0
  const int i = 2, j = 1, k = 0;
  double res[3];
  res[0] = O_ATAN2(m[j][k], m[k][k]);
  const double c2 = O_SQRT(m[i][i] * m[i][i] + m[i][j] * m[i][j]);
  if (res[0] < 0.0) {  // odd && res[0] < 0
    if (res[0] > 0.0) res[0] -= M_PI; else res[0] += M_PI;
    res[1] = O_ATAN2(-m[i][k], -c2);
  } else {
    res[1] = O_ATAN2(-m[i][k], c2);
  }
  const double s1 = O_SIN(res[0]), c1 = O_COS(res[0]);
  res[2] = O_ATAN2(s1 * m[k][i] - c1 * m[j][i], c1 * m[j][j] - s1 * m[k][j]);
  V3 out = mk3(res[0], res[1], res[2]);  // odd: no sign flip
  if (std::fabs(out[1]) > M_PI / 2.0) {
    if (out[1] > 0.0) out[1] = M_PI - out[1];
    else out[1] = -M_PI - out[1];
    out[0] = M_PI + out[0];
    out[2] = M_PI + out[2];
  }
  out[0] = std::fmod(out[0], 2.0 * M_PI);
  if (out[0] < 0.0) out[0] += 2.0 * M_PI;
  out[2] = std::fmod(out[2] + M_PI, 2.0 * M_PI) - M_PI;
  return out;
}
V3 c_euler_from_quat(const V4& q) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  double m[3][3];
  m[0][0] = 1.0 - (tyy + tzz); m[0][1] = txy - twz; m[0][2] = txz + twy;
  m[1][0] = txy + twz; m[1][1] = 1.0 - (txx + tzz); m[1][2] = tyz - twx;
  m[2][0] = txz - twy; m[2][1] = tyz + twx; m[2][2] = 1.0 - (txx + tyy);
  return c_euler_from_matrix(m);
}

// :197-245 (radians)
void c_orbital_elements(const V3& pos, const V3& vel, double out[6]) {
  V3 nr = normalized3(pos);
  V3 c = cross3(pos, vel);
  V3 f = sub3(cross3(vel, c), scale3(E_mu, nr));
  V3 c1 = normalized3(c);
  V3 f1 = normalized3(f);
  double inc = O_ACOS(c1[2]);
  double asc = 0.0, argp = 0.0;
  if (inc > 1.0e-10) {
    asc = O_ATAN2(c1[0], -c1[1]);
    V3 n = mk3(O_COS(asc), O_SIN(asc), 0.0);
    argp = O_ACOS(dot3(n, f1));
    if (f[2] < 0.0) argp *= -1.0;
  } else {
    asc = 0.0;
    if (norm3(f) > 1.0e-10) argp = O_ATAN2(f[1], f[0]);
    else argp = 0.0;
  }
  double p = dot3(c, c) / E_mu;
  double e = norm3(f) / E_mu;
  double a = p / (1.0 - e * e);
  double ta = O_ACOS(dot3(f1, nr));
  if (dot3(vel, pos) < 0.0) ta = 2.0 * M_PI - ta;
  if (asc < 0.0) asc += 2.0 * M_PI;
  if (argp < 0.0) argp += 2.0 * M_PI;
  if (ta < 0.0) ta += 2.0 * M_PI;
  out[0] = a; out[1] = e; out[2] = inc; out[3] = asc; out[4] = argp; out[5] = ta;
}

// ---- gravity (src/gravity.cpp:11-57) ----
V3 gravityECI(const V3& pos) {
  double a = 6378137.0, one_f = 298.257223563, mu = 3.986004418e14;
  double barC20 = -0.484165371736e-3;
  double f = 1.0 / one_f;
  double b = a * (1.0 - f);
  double x = pos[0], y = pos[1], z = pos[2];
  double r = O_SQRT(x * x + y * y + z * z);
  double irx, iry, irz;
  if (r == 0.0) { irx = iry = irz = 0; }
  else { irx = x / r; iry = y / r; irz = z / r; }
  double barP20 = O_SQRT(5.0) * (3.0 * irz * irz - 1.0) * 0.5;
  double barP20d = O_SQRT(5.0) * 3.0 * irz;
  if (r < b) r = b;
  double g_ir = -mu / (r * r) * (1.0 + barC20 * (a / r) * (a / r) * (3.0 * barP20 + irz * barP20d));
  double g_iz = mu / (r * r) * (a / r) * (a / r) * barC20 * barP20d;
  return mk3(g_ir * irx, g_ir * iry, g_ir * irz + g_iz);
}

// ---- Air (src/Air.cpp) ----
const double A_Rstar = 8314.32, A_g0 = 9.80665, A_r0 = 6356766.0;
const double A_hb[11] = {0.0, 11000.0, 20000.0, 32000.0, 47000.0, 51000.0, 71000.0, 86000.0, 91000.0, 110000.0, 120000.0};
const double A_lmb[11] = {-0.0065, 0.0, 0.001, 0.0028, 0.0, -0.0028, -0.002, 0.0, 0.0025, 0.012, 0.012};
const double A_tmb[11] = {288.15, 216.65, 216.65, 228.65, 270.65, 270.65, 214.65, 186.8673, 186.8673, 240.0, 360.0};
const double A_pb[11] = {101325.0, 22632.0, 5474.9, 868.02, 110.91, 66.939, 3.9564, 0.37338, 0.15381, 7.1042e-3, 2.5382e-3};
const double A_mb[11] = {28.9644, 28.9644, 28.9644, 28.9644, 28.9644, 28.9644, 28.9644, 28.9522, 28.89, 27.27, 26.20};
struct AirP { double Hb, Lmb, Tmb, Pb, R; };

double air_geopotential_altitude(double z) {  // :47-54
  if (z < 86000.0) return 1.0 * (A_r0 * z) / (A_r0 + z);
  return z;
}
AirP air_params(double alt) {  // :56-69
  int k = 0;
  for (int i = 0; i < 11; i++) if (alt >= A_hb[i]) k = i;
  AirP p; p.Hb = A_hb[k]; p.Lmb = A_lmb[k]; p.Tmb = A_tmb[k]; p.Pb = A_pb[k]; p.R = A_Rstar / A_mb[k];
  return p;
}
double air_temperature(double h) {  // :71-88
  AirP p = air_params(h);
  if (h <= 91000.0) return p.Tmb + p.Lmb * (h - p.Hb);
  else if (h <= 110000.0) {
    double Tc = 263.1905, A = -76.3232, a = -19942.9;
    return Tc + A * O_SQRT(1.0 - (h - 91000.0) * (h - 91000.0) / a / a);
  } else if (h <= 120000.0) return p.Tmb + p.Lmb * (h - p.Hb);
  else {
    double Tinf = 1000.0;
    double xi = (h - p.Hb) * (A_r0 + p.Hb) / (A_r0 + h);
    return Tinf - (Tinf - p.Tmb) * O_EXP(-0.01875e-3 * xi);
  }
}
double air_pressure(double h) {  // :90-98
  AirP p = air_params(h);
  if (std::abs(p.Lmb) > 1.0e-6)
    return p.Pb * O_POW((p.Tmb + p.Lmb * (h - p.Hb)) / p.Tmb, -A_g0 / p.Lmb / p.R);
  return p.Pb * O_EXP(A_g0 / p.R * (p.Hb - h) / p.Tmb);
}
double air_density(double h) {  // :100-105
  AirP p = air_params(h);
  double T = air_temperature(h);
  double P = air_pressure(h);
  return P / p.R / T;
}
double air_speed_of_sound(double h) {  // :107-111
  AirP p = air_params(h);
  double T = air_temperature(h);
  return O_SQRT(1.4 * p.R * T);
}

// ---- wrapper_coordinate.hpp ----
V4 w_quatmult(const V4& q, const V4& p) {  // :50-57
  return mk4(q[0] * p[0] - q[1] * p[1] - q[2] * p[2] - q[3] * p[3],
             q[0] * p[1] + q[1] * p[0] + q[2] * p[3] - q[3] * p[2],
             q[0] * p[2] - q[1] * p[3] + q[2] * p[0] + q[3] * p[1],
             q[0] * p[3] + q[1] * p[2] - q[2] * p[1] + q[3] * p[0]);
}
V3 w_quatrot(const V4& q, const V3& v) {  // :70-78
  V4 vq = mk4(0, v[0], v[1], v[2]);
  V4 r = w_quatmult(quat_conj(q), w_quatmult(vq, q));
  return mk3(r[1], r[2], r[3]);
}
V3 w_ecef2geodetic_deg(double x, double y, double z) {  // :105-111
  V3 g = earth_ecef2geodetic(mk3(x, y, z));
  g[0] = g[0] * 180.0 / M_PI;
  g[1] = g[1] * 180.0 / M_PI;
  return g;
}
V3 w_eci2geodetic_deg(const V3& pos_eci, double t) {  // :193-199
  V3 g = earth_ecef2geodetic(c_eci2ecef(pos_eci, t));
  g[0] = g[0] * 180.0 / M_PI;
  g[1] = g[1] * 180.0 / M_PI;
  return g;
}

// ---- wrapper_utils.hpp ----
// :51-80.  lower_bound(first element >= x) - 1.  x == xp[0] would index xp[-1]
// in the reference (UB); the oracle defines that case as interval 0.
double u_interp(double x, const double* xp, const double* yp, int n, int stride) {
  if (x < xp[0]) return yp[0];
  if (x > xp[(n - 1) * stride]) return yp[(n - 1) * stride];
  int lo = 0, cnt = n;  // std::lower_bound
  while (cnt > 0) {
    int step = cnt / 2;
    int it = lo + step;
    if (xp[it * stride] < x) { lo = it + 1; cnt -= step + 1; }
    else cnt = step;
  }
  int idx = lo - 1;
  if (idx < 0) idx = 0;
  double x_lower = xp[idx * stride], x_upper = xp[(idx + 1) * stride];
  double y_lower = yp[idx * stride], y_upper = yp[(idx + 1) * stride];
  double alpha = (x - x_lower) / (x_upper - x_lower);
  return y_lower + alpha * (y_upper - y_lower);
}
V3 u_wind_ned(double alt, const double* wind, int nw) {  // :82-87, wind[nw][3] row-major
  double u = u_interp(alt, wind, wind + 1, nw, 3);
  double v = u_interp(alt, wind, wind + 2, nw, 3);
  return mk3(u, v, 0.0);
}
V3 u_vel_air_eci(const V3& pos_eci, const V3& vel_eci, double t, double altitude, const double* wind, int nw) {
  V3 vel_ecef = c_vel_eci2ecef(vel_eci, pos_eci, t);
  V3 vel_wind_ned = u_wind_ned(altitude, wind, nw);
  V3 vel_wind_eci = w_quatrot(c_quat_ned2eci(pos_eci, t), vel_wind_ned);
  return sub3(c_ecef2eci(vel_ecef, t), vel_wind_eci);
}
double u_aoa_all_rad(const V3& pos, const V3& vel, const V4& quat, double t, const double* wind, int nw) {  // :89-111
  V3 thrust_dir = w_quatrot(quat_conj(quat), mk3(1.0, 0.0, 0.0));
  V3 llh = w_ecef2geodetic_deg(pos[0], pos[1], pos[2]);
  double altitude = air_geopotential_altitude(llh[2]);
  V3 va = u_vel_air_eci(pos, vel, t, altitude, wind, nw);
  // normalize(): dynamic vecXd  v / v.norm()
  V3 a = div3(va, norm3(va));
  V3 b = div3(thrust_dir, norm3(thrust_dir));
  double c_alpha = dot3(a, b);
  if (c_alpha > 1.0) return 0.0;
  else if (norm3(va) < 1e-6) return 0.0;
  else return O_ACOS(c_alpha);
}
void u_aoa_ab_rad(const V3& pos, const V3& vel, const V4& quat, double t, const double* wind, int nw, double out[2]) {  // :125-148
  V3 llh = w_ecef2geodetic_deg(pos[0], pos[1], pos[2]);
  double altitude = air_geopotential_altitude(llh[2]);
  V3 va = u_vel_air_eci(pos, vel, t, altitude, wind, nw);
  V3 vb = w_quatrot(quat, va);
  if (vb[0] < 1e-6) { out[0] = 0.0; out[1] = 0.0; return; }
  out[0] = O_ATAN2(vb[2], vb[0]);
  out[1] = O_ATAN2(vb[1], vb[0]);
}
double u_dynamic_pressure_pa(const V3& pos, const V3& vel, double t, const double* wind, int nw) {  // :163-174
  V3 llh = w_ecef2geodetic_deg(pos[0], pos[1], pos[2]);
  double altitude = air_geopotential_altitude(llh[2]);
  double rho = air_density(altitude);
  V3 va = u_vel_air_eci(pos, vel, t, altitude, wind, nw);
  return 0.5 * rho * norm3(va) * norm3(va);
}
double u_q_alpha(const V3& pos, const V3& vel, const V4& quat, double t, const double* wind, int nw) {  // :188-193
  double alpha = u_aoa_all_rad(pos, vel, quat, t, wind, nw);
  double q = u_dynamic_pressure_pa(pos, vel, t, wind, nw);
  return q * alpha;
}

// ---- iip (src/iip.cpp:36-150), radians; zeros = no solution ----
V3 iip_faa(const V3& posECEF, const V3& velECEF) {
  int n_iter = 5;
  double r_k1 = E_Rb;
  V3 p0 = c_ecef2eci(posECEF, 0.0);
  double r0 = norm3(p0);
  if (r0 < r_k1) return mk3(0, 0, 0);
  V3 v0v = c_vel_ecef2eci(velECEF, posECEF, 0.0);
  double v0 = norm3(v0v);
  double eps_cos = (r0 * v0 * v0 / E_mu) - 1.0;
  if (eps_cos >= 1.0) return mk3(0, 0, 0);
  double a_t = r0 / (1 - eps_cos);
  double eps_sin = dot3(p0, v0v) / O_SQRT(E_mu * a_t);
  double eps2 = eps_cos * eps_cos + eps_sin * eps_sin;
  if (O_SQRT(eps2) <= 1.0 && a_t * (1 - O_SQRT(eps2)) - E_Ra >= 0.0) return mk3(0, 0, 0);
  double eps_k_cos = 0, eps_k_sin = 0, d_cos = 0, d_sin = 0;
  double fs = 0, gs = 0, Ek = 0, Fk = 0, Gk = 0, r_k2 = 0, r_k1_tmp = 0;
  for (int i = 0; i < n_iter; i++) {
    eps_k_cos = (a_t - r_k1) / a_t;
    if ((eps2 - eps_k_cos * eps_k_cos) < 0) return mk3(0, 0, 0);
    eps_k_sin = -O_SQRT(eps2 - eps_k_cos * eps_k_cos);
    d_cos = (eps_k_cos * eps_cos + eps_k_sin * eps_sin) / eps2;
    d_sin = (eps_k_sin * eps_cos - eps_k_cos * eps_sin) / eps2;
    fs = (d_cos - eps_cos) / (1 - eps_cos);
    gs = (d_sin + eps_sin - eps_k_sin) * O_SQRT(a_t * a_t * a_t / E_mu);
    Ek = fs * p0[0] + gs * v0v[0];
    Fk = fs * p0[1] + gs * v0v[1];
    Gk = fs * p0[2] + gs * v0v[2];
    r_k2 = E_Ra / O_SQRT((E_e2 / (1 - E_e2)) * (Gk / r_k1) * (Gk / r_k1) + 1);
    r_k1_tmp = r_k1;
    r_k1 = r_k2;
  }
  if (std::abs(r_k1_tmp - r_k2) > 1.0) return mk3(0, 0, 0);
  double delta_eps = O_ATAN2(d_sin, d_cos);
  double time_sec = (delta_eps + eps_sin - eps_k_sin) * O_SQRT(a_t * a_t * a_t / E_mu);
  double phi_tmp = O_ASIN(Gk / r_k2);
  double phi = O_ATAN2(O_TAN(phi_tmp), 1.0 - E_e2);
  double lam = O_ATAN2(Fk, Ek) - E_omega * time_sec;
  return mk3(phi, lam, 0.0);
}

// ---- output_result.py:130-261: the derived quantities of one state node -------------------------
// numpy pieces as numpy evaluates them on this kind of input: numpy.linalg.norm of a 3-vector is
// sqrt(x.dot(x)) with BLAS ddot accumulating by fused multiply-add in index order; numpy.interp uses
// slope * (x - xp[j]) + fp[j]; math.degrees(x) = x * (180 / pi); x ** 2 = x * x.
enum { OUT_COLS = 34 };
inline double np_norm3(const V3& v) { return O_SQRT(__builtin_fma(v[2], v[2], __builtin_fma(v[1], v[1], v[0] * v[0]))); }
inline double np_degrees(double x) { return x * (180.0 / M_PI); }
double np_interp(double x, const double* xp, const double* fp, int n, int stride) {
  if (x <= xp[0]) return fp[0];  // numpy: x < xp[0] -> left = fp[0]; x == xp[0] gives the same value
  if (x >= xp[(n - 1) * stride]) return fp[(n - 1) * stride];
  int j = 0;
  while (j + 1 < n - 1 && xp[(j + 1) * stride] <= x) j++;
  const double slope = (fp[(j + 1) * stride] - fp[j * stride]) / (xp[(j + 1) * stride] - xp[j * stride]);
  return slope * (x - xp[j * stride]) + fp[j * stride];
}
void output_row(double mass, const V3& pos, const V3& vel, const V4& quat_raw, double t, double thrust_vac,
                double air_area, double nozzle_area, const double* wind, int nw, const double* ca, int nca,
                double lat0, double lon0, double* out) {
  // quat = normalize(quat_[i]): dynamic vector v / v.norm()   (:133)
  const double qn = O_SQRT(((quat_raw[0] * quat_raw[0] + quat_raw[1] * quat_raw[1]) + quat_raw[2] * quat_raw[2]) +
                           quat_raw[3] * quat_raw[3]);
  const V4 quat = mk4(quat_raw[0] / qn, quat_raw[1] / qn, quat_raw[2] / qn, quat_raw[3] / qn);
  const V3 llh = w_eci2geodetic_deg(pos, t);  // :150
  const double altitude_m = air_geopotential_altitude(llh[2]);
  out[1] = llh[0]; out[2] = llh[1]; out[6] = llh[2];
  out[5] = earth_distance_vincenty(lat0 * M_PI / 180.0, lon0 * M_PI / 180.0, llh[0] * M_PI / 180.0, llh[1] * M_PI / 180.0);
  double el[6];
  c_orbital_elements(pos, vel, el);  // wrapper converts elements 2..5 to degrees (:201-210)
  for (int k = 2; k < 6; k++) el[k] = el[k] * 180.0 / M_PI;
  out[7] = el[0] * (1.0 + el[1]) - 6378137;
  out[8] = el[0] * (1.0 - el[1]) - 6378137;
  out[9] = el[2]; out[11] = el[3]; out[10] = el[4]; out[12] = el[5];
  const V3 vel_ground_ecef = c_vel_eci2ecef(vel, pos, t);  // :172
  const V3 vel_ground_ned = w_quatrot(c_quat_ecef2ned(c_eci2ecef(pos, t)), vel_ground_ecef);
  out[13] = vel_ground_ned[0]; out[14] = vel_ground_ned[1]; out[15] = vel_ground_ned[2];
  const V3 vel_ned = w_quatrot(c_quat_eci2ned(pos, t), vel);
  const V3 vel_air_ned = sub3(vel_ground_ned, u_wind_ned(altitude_m, wind, nw));
  out[26] = np_norm3(vel_ground_ecef);
  out[22] = np_degrees(O_ATAN2(vel_ned[1], vel_ned[0]));
  out[21] = np_degrees(O_ASIN(-vel_ned[2] / np_norm3(vel_ned)));
  const double nva = np_norm3(vel_air_ned);
  const double q = 0.5 * (nva * nva) * air_density(altitude_m);  // :190
  out[31] = q;
  const double aoa_all_deg = u_aoa_all_rad(pos, vel, quat, t, wind, nw) * 180.0 / M_PI;
  double ab[2];
  u_aoa_ab_rad(pos, vel, quat, t, wind, nw, ab);
  out[28] = aoa_all_deg;
  out[32] = aoa_all_deg * q;
  out[29] = ab[0] * 180.0 / M_PI;
  out[30] = ab[1] * 180.0 / M_PI;
  const V3 thrustdir = w_quatrot(quat_conj(quat), mk3(1.0, 0.0, 0.0));  // :211
  out[23] = thrustdir[0]; out[24] = thrustdir[1]; out[25] = thrustdir[2];
  const V3 eul = c_euler_from_quat(w_quatmult(quat_conj(c_quat_eci2ned(pos, t)), quat));
  out[18] = eul[0] * 180.0 / M_PI; out[19] = eul[1] * 180.0 / M_PI; out[20] = eul[2] * 180.0 / M_PI;
  const double rho = air_density(altitude_m);  // :223
  const double p = air_pressure(altitude_m);
  const V3 pos_ecef = c_eci2ecef(pos, t);
  const V3 vel_ecef = c_vel_eci2ecef(vel, pos, t);
  const V3 vel_air_eci = u_vel_air_eci(pos, vel, t, altitude_m, wind, nw);
  const double nv = np_norm3(vel_air_eci);
  const double mach = nv / air_speed_of_sound(altitude_m);
  out[33] = mach;
  const double coeff = np_interp(mach, ca, ca + 1, nca, 2);
  out[27] = nv;
  const double k = 0.5 * rho * nv;  // 0.5 * rho * norm * -v * area * coeff, left to right (:240-247)
  const V3 aero = mk3(k * -vel_air_eci[0] * air_area * coeff, k * -vel_air_eci[1] * air_area * coeff,
                      k * -vel_air_eci[2] * air_area * coeff);
  const V3 aero_body = w_quatrot(quat, aero);
  const double thrust_n = thrust_vac - nozzle_area * p;
  out[0] = thrust_n;
  out[17] = aero_body[0];
  out[16] = (thrust_n + aero_body[0]) / mass;
  V3 iip = iip_faa(pos_ecef, vel_ecef);  // posLLH_IIP_FAA(pos_ecef, vel_ecef, False): NaN when there is none
  if (iip[0] == 0.0 && iip[1] == 0.0 && iip[2] == 0.0) {
    out[3] = out[4] = std::numeric_limits<double>::quiet_NaN();
  } else {
    out[3] = iip[0] * (180.0 / M_PI);
    out[4] = iip[1] * (180.0 / M_PI);
  }
}

inline V3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
inline V4 ld4(const double* p) { return mk4(p[0], p[1], p[2], p[3]); }
inline void st3(double* p, const V3& v) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }
inline void st4(double* p, const V4& v) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3]; }

}  // namespace

extern "C" {

int oracle_flavour(void) {
#ifdef ORACLE_GMATH
  return 1;
#else
  return 0;
#endif
}

// returns 1 when a*b+c is evaluated unfused (the build contract of both flavours)
int oracle_unfused_check(void) {
  volatile double a = 1.0 + 0x1p-30, b = 1.0 - 0x1p-30, c = -1.0;
  double r = a * b + c;  // exact product 1 - 2^-60 rounds to 1.0 when unfused
  return r == 0.0;
}

// ---- USStandardAtmosphere_c ----
double o_geopotential_altitude(double z) { return air_geopotential_altitude(z); }
double o_airtemperature_at(double h) { return air_temperature(h); }
double o_airpressure_at(double h) { return air_pressure(h); }
double o_airdensity_at(double h) { return air_density(h); }
double o_speed_of_sound(double h) { return air_speed_of_sound(h); }

// ---- coordinate_c ----
void o_quatmult(const double* q, const double* p, double* out) { st4(out, w_quatmult(ld4(q), ld4(p))); }
void o_conj(const double* q, double* out) { st4(out, quat_conj(ld4(q))); }
void o_normalize(const double* v, int n, double* out) {
  // dynamic vecXd: v / v.norm(); sequential sum (see header)
  double s = 0.0;
  for (int i = 0; i < n; i++) s = (i == 0) ? v[0] * v[0] : s + v[i] * v[i];
  double nrm = O_SQRT(s);
  for (int i = 0; i < n; i++) out[i] = v[i] / nrm;
}
void o_quatrot(const double* q, const double* v, double* out) { st3(out, w_quatrot(ld4(q), ld3(v))); }
void o_ecef2geodetic(double x, double y, double z, double* out) { st3(out, w_ecef2geodetic_deg(x, y, z)); }
void o_geodetic2ecef(double lat, double lon, double alt, double* out) {
  st3(out, earth_geodetic2ecef(mk3(lat * M_PI / 180.0, lon * M_PI / 180.0, alt)));
}
void o_ecef2eci(const double* a, double t, double* out) { st3(out, c_ecef2eci(ld3(a), t)); }
void o_eci2ecef(const double* a, double t, double* out) { st3(out, c_eci2ecef(ld3(a), t)); }
void o_vel_ecef2eci(const double* v, const double* p, double t, double* out) { st3(out, c_vel_ecef2eci(ld3(v), ld3(p), t)); }
void o_vel_eci2ecef(const double* v, const double* p, double t, double* out) { st3(out, c_vel_eci2ecef(ld3(v), ld3(p), t)); }
void o_quat_eci2ecef(double t, double* out) { st4(out, c_quat_eci2ecef(t)); }
void o_quat_ecef2eci(double t, double* out) { st4(out, quat_conj(c_quat_eci2ecef(t))); }
void o_quat_ecef2nedg(const double* p, double* out) { st4(out, c_quat_ecef2ned(ld3(p))); }
void o_quat_nedg2ecef(const double* p, double* out) { st4(out, quat_conj(c_quat_ecef2ned(ld3(p)))); }
void o_quat_eci2nedg(const double* p, double t, double* out) { st4(out, c_quat_eci2ned(ld3(p), t)); }
void o_quat_nedg2eci(const double* p, double t, double* out) { st4(out, c_quat_ned2eci(ld3(p), t)); }
void o_quat_from_euler(double az, double el, double ro, double* out) { st4(out, c_quat_from_euler_deg(az, el, ro)); }
void o_gravity(const double* p, double* out) { st3(out, gravityECI(ld3(p))); }
void o_eci2geodetic(const double* p, double t, double* out) { st3(out, w_eci2geodetic_deg(ld3(p), t)); }
void o_euler_from_quat(const double* q, double* out) {  // wrapper_coordinate.hpp:176-180 (degrees)
  V3 e = c_euler_from_quat(ld4(q));
  st3(out, mk3(e[0] * 180.0 / M_PI, e[1] * 180.0 / M_PI, e[2] * 180.0 / M_PI));
}
void o_quat_nedg2body(const double* quat, const double* p, double t, double* out) {  // :171-174
  st4(out, w_quatmult(quat_conj(c_quat_eci2ned(ld3(p), t)), ld4(quat)));
}
void o_orbital_elements(const double* p, const double* v, double* out) {  // wrapper_coordinate.hpp:201-210
  c_orbital_elements(ld3(p), ld3(v), out);
  out[2] = out[2] * 180.0 / M_PI;
  out[3] = out[3] * 180.0 / M_PI;
  out[4] = out[4] * 180.0 / M_PI;
  out[5] = out[5] * 180.0 / M_PI;
}
double o_distance_vincenty(double lat0, double lon0, double lat1, double lon1) {  // :212-222
  return earth_distance_vincenty(lat0 * M_PI / 180.0, lon0 * M_PI / 180.0, lat1 * M_PI / 180.0, lon1 * M_PI / 180.0);
}
void o_angular_momentum_vec(const double* p, const double* v, double* out) { st3(out, cross3(ld3(p), ld3(v))); }
double o_angular_momentum(const double* p, const double* v) { return norm3(cross3(ld3(p), ld3(v))); }  // :228-230
double o_inclination_cosine(const double* p, const double* v) {  // :231-234
  return cross3(ld3(p), ld3(v))[2] / norm3(cross3(ld3(p), ld3(v)));
}
double o_inclination_rad(const double* p, const double* v) { return O_ACOS(o_inclination_cosine(p, v)); }
double o_orbit_energy(const double* p, const double* v) {  // :246-250
  double r = norm3(ld3(p));
  double vv = norm3(ld3(v));
  return 0.5 * vv * vv - E_mu / r;
}
// the four DCM helpers; matrices row-major, C[3 * i + j] = C(i, j)
void o_dcm_from_quat(const double* q, double* C) {  // wrapper_coordinate.hpp:80-94
  C[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3];
  C[1] = 2 * (q[1] * q[2] + q[0] * q[3]);
  C[2] = 2 * (q[1] * q[3] - q[0] * q[2]);
  C[3] = 2 * (q[1] * q[2] - q[0] * q[3]);
  C[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3];
  C[5] = 2 * (q[2] * q[3] + q[0] * q[1]);
  C[6] = 2 * (q[1] * q[3] + q[0] * q[2]);
  C[7] = 2 * (q[2] * q[3] - q[0] * q[1]);
  C[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}
void o_quat_from_dcm(const double* C, double* q) {  // wrapper_coordinate.hpp:96-103
  q[0] = 0.5 * O_SQRT(1 + C[0] + C[4] + C[8]);
  q[1] = (C[5] - C[7]) / (4 * q[0]);
  q[2] = (C[6] - C[2]) / (4 * q[0]);
  q[3] = (C[1] - C[3]) / (4 * q[0]);
}
void o_euler_from_dcm(const double* C, double* out) {  // wrapper :182-186, Coordinate.cpp:147-164: C^T's angles, degrees
  double m[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m[i][j] = C[3 * j + i];
  V3 e = c_euler_from_matrix(m);
  for (int i = 0; i < 3; i++) out[i] = e[i] * 180.0 / M_PI;
}
void o_dcm_from_thrustvector(const double* pos, const double* thrust, double* C) {  // wrapper :188-191, Coordinate.cpp:176-191
  V3 xb = normalized3(ld3(thrust)), pn = normalized3(ld3(pos)), yb;
  if (1.0 - dot3(xb, pn) < 1.0e-10) {
    yb = normalized3(cross3(mk3(0.0, 0.0, 1.0), xb));
  } else {
    yb = normalized3(cross3(xb, pn));
  }
  V3 zb = cross3(xb, yb);
  for (int j = 0; j < 3; j++) {  // columns xb | yb | zb, transposed: rows
    C[j] = xb[j];
    C[3 + j] = yb[j];
    C[6 + j] = zb[j];
  }
}
void o_laplace_vector(const double* p, const double* v, double* out) {  // wrapper_coordinate.hpp:238-244
  V3 r = ld3(p), vel = ld3(v);
  V3 vh = cross3(vel, cross3(r, vel));
  double rn = norm3(r);
  V3 e;
  for (int i = 0; i < 3; i++) e[i] = vh[i] - E_mu * r[i] / rn;
  st3(out, e);
}
double o_haversine(double lon1, double lat1, double lon2, double lat2, double r) {  // wrapper_utils.hpp:37-49
  lon1 = lon1 * M_PI / 180.0;
  lat1 = lat1 * M_PI / 180.0;
  lon2 = lon2 * M_PI / 180.0;
  lat2 = lat2 * M_PI / 180.0;
  double dlon = lon2 - lon1, dlat = lat2 - lat1;
  double a = O_POW(O_SIN(dlat / 2), 2.0) + O_COS(lat1) * O_COS(lat2) * O_POW(O_SIN(dlon / 2), 2.0);
  return 2 * r * O_ASIN(O_SQRT(a));
}
double o_angular_momentum_from_altitude(double ha, double hp) {  // :252-258
  double ra = E_Ra + ha, rp = E_Ra + hp;
  double a = (ra + rp) / 2.0;
  double vp = O_SQRT(E_mu * (2.0 / rp - 1.0 / a));
  return rp * vp;
}
double o_orbit_energy_from_altitude(double ha, double hp) {  // :260-265
  double ra = E_Ra + ha, rp = E_Ra + hp;
  double a = (ra + rp) / 2.0;
  return -E_mu / 2.0 / a;
}

// ---- utils_c ----
double o_interp(double x, const double* xp, const double* yp, int n) { return u_interp(x, xp, yp, n, 1); }
void o_wind_ned(double alt, const double* wind, int nw, double* out) { st3(out, u_wind_ned(alt, wind, nw)); }
void o_angle_of_attack_all_array_rad(const double* pos, const double* vel, const double* quat, const double* t,
                                     int n, const double* wind, int nw, double* out) {
  for (int i = 0; i < n; i++) out[i] = u_aoa_all_rad(ld3(pos + 3 * i), ld3(vel + 3 * i), ld4(quat + 4 * i), t[i], wind, nw);
}
void o_angle_of_attack_ab_rad(const double* pos, const double* vel, const double* quat, double t, const double* wind,
                              int nw, double* out) {
  u_aoa_ab_rad(ld3(pos), ld3(vel), ld4(quat), t, wind, nw, out);
}
void o_dynamic_pressure_array_pa(const double* pos, const double* vel, const double* t, int n, const double* wind,
                                 int nw, double* out) {
  for (int i = 0; i < n; i++) out[i] = u_dynamic_pressure_pa(ld3(pos + 3 * i), ld3(vel + 3 * i), t[i], wind, nw);
}
void o_q_alpha_array_pa_rad(const double* pos, const double* vel, const double* quat, const double* t, int n,
                            const double* wind, int nw, double* out) {
  for (int i = 0; i < n; i++) out[i] = u_q_alpha(ld3(pos + 3 * i), ld3(vel + 3 * i), ld4(quat + 4 * i), t[i], wind, nw);
}

// ---- IIP_c (src/pybind_IIP.cpp:34-51) ----
void o_posLLH_IIP_FAA(const double* posECEF, const double* velECEF, int fill_na, double* out) {
  V3 r = iip_faa(ld3(posECEF), ld3(velECEF));
  if (!fill_na) {
    if (r[0] == 0.0 && r[1] == 0.0 && r[2] == 0.0) {
      double nan = std::numeric_limits<double>::quiet_NaN();
      out[0] = out[1] = out[2] = nan;
      return;
    }
  }
  r[0] *= 180.0 / M_PI;
  r[1] *= 180.0 / M_PI;
  st3(out, r);
}

// ---- dynamics_c (src/pybind_dynamics.cpp:30-106) ----
void o_dynamics_velocity(const double* mass_e, const double* pos_e, const double* vel_e, const double* quat,
                         const double* t, int n, const double* param, const double* wind, int nw,
                         const double* ca, int nc, const double* units, double* out) {
  double thrust_vac = param[0], air_area = param[2], nozzle_area = param[4];
  for (int i = 0; i < n; i++) {
    double mass = mass_e[i] * units[0];
    V3 pos = mk3(pos_e[3 * i] * units[1], pos_e[3 * i + 1] * units[1], pos_e[3 * i + 2] * units[1]);
    V3 vel = mk3(vel_e[3 * i] * units[2], vel_e[3 * i + 1] * units[2], vel_e[3 * i + 2] * units[2]);
    V4 q = ld4(quat + 4 * i);
    V3 llh = w_ecef2geodetic_deg(pos[0], pos[1], pos[2]);
    double altitude = air_geopotential_altitude(llh[2]);
    double rho = air_density(altitude);
    double p = air_pressure(altitude);
    V3 va = u_vel_air_eci(pos, vel, t[i], altitude, wind, nw);
    double mach = norm3(va) / air_speed_of_sound(altitude);
    double cav = u_interp(mach, ca, ca + 1, nc, 2);
    double s = 0.5 * rho * air_area * cav * norm3(va);
    V3 aero = mk3(s * -va[0], s * -va[1], s * -va[2]);
    double thrust = thrust_vac - nozzle_area * p;
    V3 tdir = w_quatrot(quat_conj(q), mk3(1.0, 0.0, 0.0));
    V3 thr = scale3(thrust, tdir);
    V3 g = gravityECI(pos);
    for (int k = 0; k < 3; k++) out[3 * i + k] = ((thr[k] + aero[k]) / mass + g[k]) / units[2];
  }
}

void o_dynamics_velocity_NoAir(const double* mass_e, const double* pos_e, const double* quat, int n,
                               const double* param, const double* units, double* out) {
  double thrust_vac = param[0];
  for (int i = 0; i < n; i++) {
    double mass = mass_e[i] * units[0];
    V3 pos = mk3(pos_e[3 * i] * units[1], pos_e[3 * i + 1] * units[1], pos_e[3 * i + 2] * units[1]);
    V4 q = ld4(quat + 4 * i);
    V3 tdir = w_quatrot(quat_conj(q), mk3(1.0, 0.0, 0.0));
    V3 thr = scale3(thrust_vac, tdir);
    V3 g = gravityECI(pos);
    for (int k = 0; k < 3; k++) out[3 * i + k] = (thr[k] / mass + g[k]) / units[2];
  }
}

// out[n][OUT_COLS]; per-node section parameters thrust_vac / air_area / nozzle_area as output_result.py:139-143 picks them
void o_output_rows(int n, const double* mass, const double* pos, const double* vel, const double* quat, const double* t,
                   const double* thrust_vac, const double* air_area, const double* nozzle_area, const double* wind, int nw,
                   const double* ca, int nca, double lat0, double lon0, double* out) {
  for (int i = 0; i < n; i++)
    output_row(mass[i], ld3(pos + 3 * i), ld3(vel + 3 * i), ld4(quat + 4 * i), t[i], thrust_vac[i], air_area[i],
               nozzle_area[i], wind, nw, ca, nca, lat0, lon0, out + (size_t)i * OUT_COLS);
}
void o_dynamics_quaternion(const double* quat, const double* u_e, double unit_u, int n, double* out) {
  for (int i = 0; i < n; i++) {
    double u0 = u_e[2 * i] * unit_u, u1 = u_e[2 * i + 1] * unit_u;
    // vec4d(0,0,u0,u1) * M_PI / 180.0  (left to right per coefficient)
    V4 om = mk4(0.0 * M_PI / 180.0, 0.0 * M_PI / 180.0, u0 * M_PI / 180.0, u1 * M_PI / 180.0);
    V4 d = w_quatmult(ld4(quat + 4 * i), om);
    for (int k = 0; k < 4; k++) out[4 * i + k] = 0.5 * d[k];
  }
}

}  // extern "C"

// ---- D.X in the CUDA kernels' order: acc = fma(D[k][j], X[j][c], acc), j ascending ----
extern "C" void o_seqfma_matmul(const double* D, const double* X, int n, int m, int w, double* out) {
  for (int k = 0; k < n; k++)
    for (int c = 0; c < w; c++) {
      double acc = 0.0;
      for (int j = 0; j < m; j++) acc = __builtin_fma(D[k * m + j], X[j * w + c], acc);
      out[k * w + c] = acc;
    }
}
