// mini_eigen.h -- the sliver of the Eigen3 API that GELATO's C++ sources use, so that the
// reference's own files (/root/reference/src/*.cpp|hpp) can be compiled WHERE THEY LIE into
// oracle/_ref/ in an image that has no Eigen3 (oracle/Makefile, target `ref`).
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Every arithmetic statement of the
// reference's code runs unchanged; what this header supplies is Eigen's part:
//   * dense fixed / dynamic double matrices with element access, row()/col()/tail(), data(),
//     + - unary- scalar* /scalar, comma initialisation;
//   * norm / squaredNorm / dot as left-to-right sums over the coefficients in (a0 + a1) + a2
//     order, cross / normalized by Eigen's coefficient formulas;
//   * Quaterniond (w,x,y,z constructor, product by Eigen's generic coefficient formula,
//     conjugate, from AngleAxis, to / from rotation matrix), AngleAxisd, Matrix3d::eulerAngles.
// Eigen's own vectorised reduction order is not observable here (no Eigen in the image), so a
// genuine Eigen build may differ from this in the last bit of a norm or dot product.
#ifndef ORACLE_MINI_EIGEN_H_
#define ORACLE_MINI_EIGEN_H_

#include <cassert>
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, Dynamic = -1 };

template <typename T, int R, int C, int Opt = ColMajor>
class Matrix;

namespace mini {

// rectangular view into another matrix' storage (row(i), col(j), tail(n))
struct View {
  double* p;
  int rows, cols;
  std::ptrdiff_t rstride, cstride;
  double& at(int i, int j) const { return p[i * rstride + j * cstride]; }
  int size() const { return rows * cols; }
  double& lin(int k) const { return rows == 1 ? at(0, k) : (cols == 1 ? at(k, 0) : at(k % rows, k / rows)); }
};

}  // namespace mini

template <typename Derived>
struct CommaInit;

template <typename T, int R, int C, int Opt>
class Matrix {
 public:
  typedef Matrix Self;
  // ---- construction ----------------------------------------------------------
  Matrix() { resize_(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C); }
  explicit Matrix(int n) {  // VectorXd v(n)
    if (C == 1) resize_(n, 1);
    else if (R == 1) resize_(1, n);
    else resize_(n, n);
  }
  Matrix(int r, int c) { resize_(r, c); }
  Matrix(double a, double b) { resize_(2, 1); d_[0] = a; d_[1] = b; fixup_vec_(); }
  Matrix(double a, double b, double c) { resize_(3, 1); d_[0] = a; d_[1] = b; d_[2] = c; fixup_vec_(); }
  Matrix(double a, double b, double c, double d) { resize_(4, 1); d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = d; fixup_vec_(); }
  explicit Matrix(const double* src) {  // Matrix3d C_(C.data())
    resize_(R, C);
    for (int k = 0; k < r_ * c_; k++) d_[k] = src[k];
  }
  template <int R2, int C2, int O2>
  Matrix(const Matrix<T, R2, C2, O2>& o) { assign_(o.view_()); }
  Matrix(const mini::View& v) { assign_(v); }
  template <int R2, int C2, int O2>
  Matrix& operator=(const Matrix<T, R2, C2, O2>& o) { assign_(o.view_()); return *this; }
  Matrix& operator=(const mini::View& v) { assign_(v); return *this; }

  static Matrix Zero() { Matrix m; m.fill_(0.0); return m; }
  static Matrix Zero(int r, int c) { Matrix m(r, c); m.fill_(0.0); return m; }
  static Matrix Zero(int n) { Matrix m(n); m.fill_(0.0); return m; }
  static Matrix UnitX() { Matrix m = Zero(); m[0] = 1.0; return m; }
  static Matrix UnitY() { Matrix m = Zero(); m[1] = 1.0; return m; }
  static Matrix UnitZ() { Matrix m = Zero(); m[2] = 1.0; return m; }

  // ---- shape / access --------------------------------------------------------
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double& operator()(int i, int j) { return d_[idx_(i, j)]; }
  double operator()(int i, int j) const { return d_[idx_(i, j)]; }
  double& operator()(int k) { return d_[k]; }
  double operator()(int k) const { return d_[k]; }
  double& operator[](int k) { return d_[k]; }
  double operator[](int k) const { return d_[k]; }
  double& x() { return d_[0]; }
  double& y() { return d_[1]; }
  double& z() { return d_[2]; }

  // blocks: assignable views (acc.row(i) = v) that also convert to matrices
  struct Block : mini::View {
    Block(const mini::View& v) : mini::View(v) {}
    template <int R2, int C2, int O2>
    Block& operator=(const Matrix<T, R2, C2, O2>& o) {
      mini::View s = o.view_();
      assert(s.size() == this->size());
      for (int k = 0; k < this->size(); k++) this->lin(k) = s.lin(k);
      return *this;
    }
  };
  Block row(int i) { return Block(mini::View{&d_[idx_(i, 0)], 1, c_, 0, cs_()}); }
  Block col(int j) { return Block(mini::View{&d_[idx_(0, j)], r_, 1, rs_(), 0}); }
  Block tail(int n) { assert(c_ == 1 || r_ == 1); return Block(mini::View{&d_[size() - n], n, 1, 1, 0}); }
  mini::View view_() const {
    return mini::View{const_cast<double*>(d_.data()), r_, c_, rs_(), cs_()};
  }

  // ---- arithmetic --------------------------------------------------------------
  Matrix operator-() const { Matrix m(*this); for (auto& v : m.d_) v = -v; return m; }
  template <int R2, int C2, int O2>
  Matrix operator+(const Matrix<T, R2, C2, O2>& o) const { return zip_(o.view_(), +1.0); }
  template <int R2, int C2, int O2>
  Matrix operator-(const Matrix<T, R2, C2, O2>& o) const { return zip_(o.view_(), -1.0); }
  Matrix operator*(double s) const { Matrix m(*this); for (auto& v : m.d_) v = v * s; return m; }
  Matrix operator/(double s) const { Matrix m(*this); for (auto& v : m.d_) v = v / s; return m; }
  friend Matrix operator*(double s, const Matrix& a) { Matrix m(a); for (auto& v : m.d_) v = s * v; return m; }

  // reductions: coefficients in storage order, ((a0*b0 + a1*b1) + a2*b2) + ...
  template <int R2, int C2, int O2>
  double dot(const Matrix<T, R2, C2, O2>& o) const {
    mini::View a = view_(), b = o.view_();
    assert(a.size() == b.size());
    double s = a.lin(0) * b.lin(0);
    for (int k = 1; k < a.size(); k++) s = s + a.lin(k) * b.lin(k);
    return s;
  }
  double squaredNorm() const { return dot(*this); }
  double norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const {  // Eigen: n = squaredNorm(); n > 0 ? *this / sqrt(n) : *this
    double n2 = squaredNorm();
    if (n2 > 0.0) return *this / std::sqrt(n2);
    return *this;
  }
  template <int R2, int C2, int O2>
  Matrix<T, 3, 1> cross(const Matrix<T, R2, C2, O2>& o) const {
    mini::View a = view_(), b = o.view_();
    return Matrix<T, 3, 1>(a.lin(1) * b.lin(2) - a.lin(2) * b.lin(1), a.lin(2) * b.lin(0) - a.lin(0) * b.lin(2),
                           a.lin(0) * b.lin(1) - a.lin(1) * b.lin(0));
  }
  Matrix<T, C, R, Opt> transpose() const {
    Matrix<T, C, R, Opt> m(c_, r_);
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++) m(j, i) = (*this)(i, j);
    return m;
  }
  // product of two 3x3 matrices is not needed by the reference; matrix * vector neither.

  // Eigen/src/Geometry/EulerAngles.h, the (2, 1, 0) call the reference makes is the
  // generic algorithm with i = a0, j = (i+1+odd)%3, k = (i+2-odd)%3, odd = ((a0+1)%3 != a1)
  Matrix<T, 3, 1> eulerAngles(int a0, int a1, int a2) const {
    const Matrix& m = *this;
    Matrix<T, 3, 1> res;
    const int odd = ((a0 + 1) % 3 == a1) ? 0 : 1;
    const int i = a0, j = (a0 + 1 + odd) % 3, k = (a0 + 2 - odd) % 3;
    if (a0 == a2) {
      res[0] = std::atan2(m(j, i), m(k, i));
      if ((odd && res[0] < 0.0) || ((!odd) && res[0] > 0.0)) {
        if (res[0] > 0.0) res[0] -= M_PI; else res[0] += M_PI;
        double s2 = std::sqrt(m(j, i) * m(j, i) + m(k, i) * m(k, i));
        res[1] = -std::atan2(s2, m(i, i));
      } else {
        double s2 = std::sqrt(m(j, i) * m(j, i) + m(k, i) * m(k, i));
        res[1] = std::atan2(s2, m(i, i));
      }
      double s1 = std::sin(res[0]), c1 = std::cos(res[0]);
      res[2] = std::atan2(c1 * m(j, k) - s1 * m(k, k), c1 * m(j, j) - s1 * m(k, j));
    } else {
      res[0] = std::atan2(m(j, k), m(k, k));
      double c2 = std::sqrt(m(i, i) * m(i, i) + m(i, j) * m(i, j));
      if ((odd && res[0] < 0.0) || ((!odd) && res[0] > 0.0)) {
        if (res[0] > 0.0) res[0] -= M_PI; else res[0] += M_PI;
        res[1] = std::atan2(-m(i, k), -c2);
      } else {
        res[1] = std::atan2(-m(i, k), c2);
      }
      double s1 = std::sin(res[0]), c1 = std::cos(res[0]);
      res[2] = std::atan2(s1 * m(k, i) - c1 * m(j, i), c1 * m(j, j) - s1 * m(k, j));
    }
    if (!odd) res = -res;
    return res;
  }

  // comma initialisation: out << a, b, c;  scalars fill in storage-independent row-major
  // order; column vectors fill successive columns (the one use: Coordinate.cpp:188)
  CommaInit<Matrix> operator<<(double v);
  template <int R2, int C2, int O2>
  CommaInit<Matrix> operator<<(const Matrix<T, R2, C2, O2>& v);

 private:
  template <typename, int, int, int>
  friend class Matrix;
  std::vector<double> d_;
  int r_ = 0, c_ = 0;
  void resize_(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
  void fixup_vec_() { if (R == 1 && C != 1) { r_ = 1; c_ = (int)d_.size(); } }
  void fill_(double v) { for (auto& e : d_) e = v; }
  std::ptrdiff_t rs_() const { return (Opt & RowMajor) ? c_ : 1; }
  std::ptrdiff_t cs_() const { return (Opt & RowMajor) ? 1 : r_; }
  size_t idx_(int i, int j) const { return (size_t)(i * rs_() + j * cs_()); }
  void assign_(const mini::View& v) {
    int r = v.rows, c = v.cols;
    // vectors convert between row and column shape (Eigen allows it for vectors)
    if ((R == Dynamic || R == r) && (C == Dynamic || C == c)) {
      resize_(r, c);
      for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) d_[idx_(i, j)] = v.at(i, j);
    } else {
      assert((r == 1 || c == 1) && (R == 1 || C == 1 || R == Dynamic || C == Dynamic));
      int n = r * c;
      if (C == 1) resize_(n, 1); else resize_(1, n);
      for (int k = 0; k < n; k++) d_[k] = v.lin(k);
    }
  }
  Matrix zip_(const mini::View& b, double sign) const {
    Matrix m(*this);
    mini::View a = m.view_();
    assert(a.size() == b.size());
    for (int k = 0; k < a.size(); k++) a.lin(k) = sign > 0 ? a.lin(k) + b.lin(k) : a.lin(k) - b.lin(k);
    return m;
  }
};

template <typename M>
struct CommaInit {
  M& m;
  int pos;  // scalars written so far (row-major order) or columns written so far
  bool by_col;
  CommaInit& operator,(double v) {
    m(pos / m.cols(), pos % m.cols()) = v;
    pos++;
    return *this;
  }
  template <typename T, int R2, int C2, int O2>
  CommaInit& operator,(const Matrix<T, R2, C2, O2>& v) {
    for (int i = 0; i < v.size(); i++) m(i, pos) = v[i];
    pos++;
    by_col = true;
    return *this;
  }
};
template <typename T, int R, int C, int Opt>
CommaInit<Matrix<T, R, C, Opt>> Matrix<T, R, C, Opt>::operator<<(double v) {
  CommaInit<Matrix> ci{*this, 0, false};
  ci, v;
  return ci;
}
template <typename T, int R, int C, int Opt>
template <int R2, int C2, int O2>
CommaInit<Matrix<T, R, C, Opt>> Matrix<T, R, C, Opt>::operator<<(const Matrix<T, R2, C2, O2>& v) {
  CommaInit<Matrix> ci{*this, 0, true};
  ci, v;
  return ci;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

// ---- geometry ---------------------------------------------------------------------
class AngleAxisd;

class Quaterniond {
 public:
  Quaterniond() : w_(1), x_(0), y_(0), z_(0) {}
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  Quaterniond(const AngleAxisd& aa);
  // Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl<Other,3,3>
  explicit Quaterniond(const Matrix3d& mat) {
    double t = mat(0, 0) + mat(1, 1) + mat(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      w_ = 0.5 * t;
      t = 0.5 / t;
      x_ = (mat(2, 1) - mat(1, 2)) * t;
      y_ = (mat(0, 2) - mat(2, 0)) * t;
      z_ = (mat(1, 0) - mat(0, 1)) * t;
    } else {
      int i = 0;
      if (mat(1, 1) > mat(0, 0)) i = 1;
      if (mat(2, 2) > mat(i, i)) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(mat(i, i) - mat(j, j) - mat(k, k) + 1.0);
      double v[3];
      v[i] = 0.5 * t;
      t = 0.5 / t;
      w_ = (mat(k, j) - mat(j, k)) * t;
      v[j] = (mat(j, i) + mat(i, j)) * t;
      v[k] = (mat(k, i) + mat(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  Quaterniond conjugate() const { return Quaterniond(w_, -x_, -y_, -z_); }
  // Eigen/src/Geometry/Quaternion.h quat_product (generic)
  Quaterniond operator*(const Quaterniond& b) const {
    const Quaterniond& a = *this;
    return Quaterniond(a.w_ * b.w_ - a.x_ * b.x_ - a.y_ * b.y_ - a.z_ * b.z_,
                       a.w_ * b.x_ + a.x_ * b.w_ + a.y_ * b.z_ - a.z_ * b.y_,
                       a.w_ * b.y_ + a.y_ * b.w_ + a.z_ * b.x_ - a.x_ * b.z_,
                       a.w_ * b.z_ + a.z_ * b.w_ + a.x_ * b.y_ - a.y_ * b.x_);
  }
  Matrix3d toRotationMatrix() const {
    Matrix3d res;
    const double tx = 2.0 * x_, ty = 2.0 * y_, tz = 2.0 * z_;
    const double twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const double txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const double tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = 1.0 - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = 1.0 - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = 1.0 - (txx + tyy);
    return res;
  }

 private:
  double w_, x_, y_, z_;
};

class AngleAxisd {
 public:
  AngleAxisd(double angle, const Vector3d& axis) : angle_(angle), axis_(axis) {}
  double angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }
  Quaterniond operator*(const AngleAxisd& o) const { return Quaterniond(*this) * Quaterniond(o); }
  friend Quaterniond operator*(const Quaterniond& a, const AngleAxisd& b) { return a * Quaterniond(b); }

 private:
  double angle_;
  Vector3d axis_;
};
inline Quaterniond::Quaterniond(const AngleAxisd& aa) {
  const double ha = 0.5 * aa.angle();
  w_ = std::cos(ha);
  const double s = std::sin(ha);
  x_ = s * aa.axis()[0];
  y_ = s * aa.axis()[1];
  z_ = s * aa.axis()[2];
}

}  // namespace Eigen
#endif  // ORACLE_MINI_EIGEN_H_
