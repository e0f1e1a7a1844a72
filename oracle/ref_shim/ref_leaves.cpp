// ref_leaves.cpp -- the REFERENCE'S OWN C++ physics (/root/reference/src/*.cpp|hpp), compiled
// where it lies against oracle/ref_shim/mini_eigen.h (this image has no Eigen3), behind the
// same extern "C" surface as oracle/oracle_leaves.cpp so oracle/leaves.py can load it as the
// flavour "ref".  TEST INFRASTRUCTURE ONLY: it pins the hand-written oracle restatement
// (tests/test_oracle_vs_refcpp.py) and serves as the CPU baseline of kind "reference".
// Built by `make -C oracle ref` into oracle/_ref/ (git-ignored); no reference source is copied.
#include <limits>

#include "Air.cpp"
#include "Earth.cpp"
#include "Coordinate.cpp"
#include "gravity.cpp"
#include "iip.cpp"
#include "pybind_USStandardAtmosphere.cpp"
#include "pybind_coordinate.cpp"
#include "pybind_utils.cpp"
#include "pybind_IIP.cpp"
#include "pybind_dynamics.cpp"

namespace {
vec3d ld3(const double* p) { return vec3d(p[0], p[1], p[2]); }
vec4d ld4(const double* p) { return vec4d(p[0], p[1], p[2], p[3]); }
template <typename V>
void st(double* out, const V& v, int n) { for (int i = 0; i < n; i++) out[i] = v[i]; }
vecXd ldv(const double* p, int n) { vecXd v(n); for (int i = 0; i < n; i++) v[i] = p[i]; return v; }
matXd ldm(const double* p, int r, int c) {
  matXd m(r, c);
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m(i, j) = p[i * c + j];
  return m;
}
void stm(double* out, const matXd& m) {
  for (int i = 0; i < m.rows(); i++) for (int j = 0; j < m.cols(); j++) out[i * m.cols() + j] = m(i, j);
}
}  // namespace

extern "C" {
int oracle_flavour(void) { return 2; }
int oracle_unfused_check(void) {
  volatile double a = 1.0 + 0x1p-30, b = 1.0 - 0x1p-30, c = -1.0;
  double r = a * b + c;
  return r == 0.0;
}
double o_geopotential_altitude(double z) { return geopotential_altitude(z); }
double o_airtemperature_at(double h) { return airtemperature_at(h); }
double o_airpressure_at(double h) { return airpressure_at(h); }
double o_airdensity_at(double h) { return airdensity_at(h); }
double o_speed_of_sound(double h) { return speed_of_sound(h); }

void o_quatmult(const double* q, const double* p, double* out) { st(out, quatmult(ld4(q), ld4(p)), 4); }
void o_conj(const double* q, double* out) { st(out, conj(ld4(q)), 4); }
void o_normalize(const double* v, int n, double* out) { st(out, normalize(ldv(v, n)), n); }
void o_quatrot(const double* q, const double* v, double* out) { st(out, quatrot(ld4(q), ld3(v)), 3); }
void o_ecef2geodetic(double x, double y, double z, double* out) { st(out, ecef2geodetic(x, y, z), 3); }
void o_geodetic2ecef(double lat, double lon, double alt, double* out) { st(out, geodetic2ecef(lat, lon, alt), 3); }
void o_ecef2eci(const double* a, double t, double* out) { st(out, ecef2eci(ld3(a), t), 3); }
void o_eci2ecef(const double* a, double t, double* out) { st(out, eci2ecef(ld3(a), t), 3); }
void o_vel_ecef2eci(const double* v, const double* p, double t, double* out) { st(out, vel_ecef2eci(ld3(v), ld3(p), t), 3); }
void o_vel_eci2ecef(const double* v, const double* p, double t, double* out) { st(out, vel_eci2ecef(ld3(v), ld3(p), t), 3); }
void o_quat_eci2ecef(double t, double* out) { st(out, quat_eci2ecef(t), 4); }
void o_quat_ecef2eci(double t, double* out) { st(out, quat_ecef2eci(t), 4); }
void o_quat_ecef2nedg(const double* p, double* out) { st(out, quat_ecef2nedg(ld3(p)), 4); }
void o_quat_nedg2ecef(const double* p, double* out) { st(out, quat_nedg2ecef(ld3(p)), 4); }
void o_quat_eci2nedg(const double* p, double t, double* out) { st(out, quat_eci2nedg(ld3(p), t), 4); }
void o_quat_nedg2eci(const double* p, double t, double* out) { st(out, quat_nedg2eci(ld3(p), t), 4); }
void o_quat_from_euler(double az, double el, double ro, double* out) { st(out, quat_from_euler(az, el, ro), 4); }
void o_gravity(const double* p, double* out) { st(out, gravity(ld3(p)), 3); }
void o_eci2geodetic(const double* p, double t, double* out) { st(out, eci2geodetic(ld3(p), t), 3); }
void o_euler_from_quat(const double* q, double* out) { st(out, euler_from_quat(ld4(q)), 3); }
void o_quat_nedg2body(const double* quat, const double* p, double t, double* out) { st(out, quat_nedg2body(ld4(quat), ld3(p), t), 4); }
void o_orbital_elements(const double* p, const double* v, double* out) { st(out, orbital_elements(ld3(p), ld3(v)), 6); }
double o_distance_vincenty(double lat0, double lon0, double lat1, double lon1) {
  return distance_vincenty(lat0, lon0, lat1, lon1);
}
void o_angular_momentum_vec(const double* p, const double* v, double* out) { st(out, angular_momentum_vec(ld3(p), ld3(v)), 3); }
double o_angular_momentum(const double* p, const double* v) { return angular_momentum(ld3(p), ld3(v)); }
double o_inclination_cosine(const double* p, const double* v) { return inclination_cosine(ld3(p), ld3(v)); }
double o_inclination_rad(const double* p, const double* v) { return inclination_rad(ld3(p), ld3(v)); }
double o_orbit_energy(const double* p, const double* v) { return orbit_energy(ld3(p), ld3(v)); }
static Eigen::Matrix3d ldc(const double* C) {
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m(i, j) = C[3 * i + j];
  return m;
}
static void stc(double* C, const Eigen::Matrix3d& m) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = m(i, j);
}
void o_dcm_from_quat(const double* q, double* C) { stc(C, dcm_from_quat(ld4(q))); }
void o_quat_from_dcm(const double* C, double* q) { st(q, quat_from_dcm(ldc(C)), 4); }
void o_euler_from_dcm(const double* C, double* out) { st(out, euler_from_dcm(ldc(C)), 3); }
void o_dcm_from_thrustvector(const double* pos, const double* thrust, double* C) { stc(C, dcm_from_thrustvector(ld3(pos), ld3(thrust))); }
void o_laplace_vector(const double* p, const double* v, double* out) { st(out, laplace_vector(ld3(p), ld3(v)), 3); }
double o_haversine(double lon1, double lat1, double lon2, double lat2, double r) { return haversine(lon1, lat1, lon2, lat2, r); }
double o_angular_momentum_from_altitude(double ha, double hp) { return angular_momentum_from_altitude(ha, hp); }
double o_orbit_energy_from_altitude(double ha, double hp) { return orbit_energy_from_altitude(ha, hp); }

double o_interp(double x, const double* xp, const double* yp, int n) { return interp(x, ldv(xp, n), ldv(yp, n)); }
void o_wind_ned(double alt, const double* wind, int nw, double* out) { st(out, wind_ned(alt, ldm(wind, nw, 3)), 3); }
void o_angle_of_attack_all_array_rad(const double* pos, const double* vel, const double* quat, const double* t, int n,
                                     const double* wind, int nw, double* out) {
  st(out, angle_of_attack_all_array_rad(ldm(pos, n, 3), ldm(vel, n, 3), ldm(quat, n, 4), ldv(t, n), ldm(wind, nw, 3)), n);
}
void o_angle_of_attack_ab_rad(const double* pos, const double* vel, const double* quat, double t, const double* wind,
                              int nw, double* out) {
  st(out, angle_of_attack_ab_rad(ld3(pos), ld3(vel), ld4(quat), t, ldm(wind, nw, 3)), 2);
}
void o_dynamic_pressure_array_pa(const double* pos, const double* vel, const double* t, int n, const double* wind,
                                 int nw, double* out) {
  st(out, dynamic_pressure_array_pa(ldm(pos, n, 3), ldm(vel, n, 3), ldv(t, n), ldm(wind, nw, 3)), n);
}
void o_q_alpha_array_pa_rad(const double* pos, const double* vel, const double* quat, const double* t, int n,
                            const double* wind, int nw, double* out) {
  st(out, q_alpha_array_pa_rad(ldm(pos, n, 3), ldm(vel, n, 3), ldm(quat, n, 4), ldv(t, n), ldm(wind, nw, 3)), n);
}
void o_posLLH_IIP_FAA(const double* posECEF, const double* velECEF, int fill_na, double* out) {
  st(out, posLLH_IIP_FAA_deg(ld3(posECEF), ld3(velECEF), fill_na != 0, 5), 3);
}
void o_dynamics_velocity(const double* mass_e, const double* pos_e, const double* vel_e, const double* quat,
                         const double* t, int n, const double* param, const double* wind, int nw, const double* ca,
                         int nca, const double* units, double* out) {
  stm(out, dynamics_velocity(ldv(mass_e, n), ldm(pos_e, n, 3), ldm(vel_e, n, 3), ldm(quat, n, 4), ldv(t, n),
                             ldv(param, 5), ldm(wind, nw, 3), ldm(ca, nca, 2), ldv(units, 3)));
}
void o_dynamics_velocity_NoAir(const double* mass_e, const double* pos_e, const double* quat, int n,
                               const double* param, const double* units, double* out) {
  stm(out, dynamics_velocity_NoAir(ldv(mass_e, n), ldm(pos_e, n, 3), ldm(quat, n, 4), ldv(param, 5), ldv(units, 3)));
}
void o_dynamics_quaternion(const double* quat, const double* u_e, double unit_u, int n, double* out) {
  stm(out, dynamics_quaternion(ldm(quat, n, 4), ldm(u_e, n, 2), unit_u));
}
}  // extern "C"
