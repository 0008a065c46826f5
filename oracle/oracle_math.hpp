// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (parity oracle).  Never linked, imported or called by the
// product path (syropod_highlevel_controller_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
//
// The reference ships no tests, golden vectors or fixtures; the restatement is pinned to the reference's
// own sources compiled against stand-in headers (oracle/_ref, tests/test_reference_pin.py — see shc_oracle.hpp).
//
// Dependency-free IEEE-double restatement of the small part of Eigen 3.3 (unpinned by the reference's
// CMakeLists.txt:33; 3.3.4 / 3.3.7 ship with the supported ROS distros) that the hot path uses, plus the scalar helpers of
// /root/reference/include/syropod_highlevel_controller/standard_includes.h.  Each function cites what it restates.
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstring>

namespace om {

constexpr double UNASSIGNED_VALUE = double(INT_MAX);  // standard_includes.h:52

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() = default;
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Vec3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
  Vec3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
  Vec3 operator-() const { return {-x, -y, -z}; }
  Vec3 operator*(double s) const { return {x * s, y * s, z * s}; }
  Vec3 operator/(double s) const { return {x / s, y / s, z / s}; }
  Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  Vec3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  Vec3& operator*=(double s) { x *= s; y *= s; z *= s; return *this; }
  double dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
  Vec3 cross(const Vec3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
  double squaredNorm() const { return x * x + y * y + z * z; }
  double norm() const { return std::sqrt(squaredNorm()); }
  // Eigen MatrixBase::normalized(): returns the input unchanged when the squared norm is not > 0.
  Vec3 normalized() const {
    double z2 = squaredNorm();
    return z2 > 0.0 ? (*this) / std::sqrt(z2) : *this;
  }
  bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Vec3& o) const { return !(*this == o); }
};
inline Vec3 operator*(double s, const Vec3& v) { return v * s; }
inline Vec3 UnitX() { return {1, 0, 0}; }
inline Vec3 UnitY() { return {0, 1, 0}; }
inline Vec3 UnitZ() { return {0, 0, 1}; }

// Eigen DenseBase::isApprox (Fuzzy.h): |a-b|^2 <= prec^2 * min(|a|^2,|b|^2), prec = 1e-12 for double.
inline bool isApproxVec(const Vec3& a, const Vec3& b) {
  const double prec = 1e-12;
  return (a - b).squaredNorm() <= prec * prec * std::min(a.squaredNorm(), b.squaredNorm());
}

struct Mat3 {
  double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  Vec3 operator*(const Vec3& v) const {
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
  }
};

struct Quat {  // Eigen::Quaterniond(w, x, y, z)
  double w = 1, x = 0, y = 0, z = 0;
  Quat() = default;
  Quat(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
  static Quat Identity() { return {1, 0, 0, 0}; }
  Vec3 vec() const { return {x, y, z}; }
  double squaredNorm() const { return x * x + y * y + z * z + w * w; }
  double dot(const Quat& o) const { return x * o.x + y * o.y + z * o.z + w * o.w; }
  Quat conjugate() const { return {w, -x, -y, -z}; }
  // Eigen QuaternionBase::inverse(): conjugate / squaredNorm, or the all-zero quaternion when the norm is 0.
  Quat inverse() const {
    double n2 = squaredNorm();
    if (n2 > 0.0) return {w / n2, -x / n2, -y / n2, -z / n2};
    return {0, 0, 0, 0};
  }
  Quat normalized() const {
    double n2 = squaredNorm();
    if (n2 > 0.0) {
      double n = std::sqrt(n2);
      return {w / n, x / n, y / n, z / n};
    }
    return *this;
  }
  // Hamilton product (Eigen quat_product).
  Quat operator*(const Quat& b) const {
    return {w * b.w - x * b.x - y * b.y - z * b.z, w * b.x + x * b.w + y * b.z - z * b.y,
            w * b.y + y * b.w + z * b.x - x * b.z, w * b.z + z * b.w + x * b.y - y * b.x};
  }
  // Eigen QuaternionBase::_transformVector: v + w*uv + vec x uv, uv = 2 * vec x v.
  Vec3 transformVector(const Vec3& v) const {
    Vec3 uv = vec().cross(v);
    uv += uv;
    return v + w * uv + vec().cross(uv);
  }
  // Eigen QuaternionBase::toRotationMatrix.
  Mat3 toRotationMatrix() const {
    Mat3 r;
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    r.m[0][0] = 1.0 - (tyy + tzz); r.m[0][1] = txy - twz;         r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;         r.m[1][1] = 1.0 - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;         r.m[2][1] = tyz + twx;         r.m[2][2] = 1.0 - (txx + tyy);
    return r;
  }
  bool isZero() const { return w == 0.0 && x == 0.0 && y == 0.0 && z == 0.0; }
};

// isApprox against another quaternion (coeff-wise Fuzzy test as for vectors).
inline bool isApproxQuat(const Quat& a, const Quat& b) {
  const double prec = 1e-12;
  double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
  return dx * dx + dy * dy + dz * dz + dw * dw <= prec * prec * std::min(a.squaredNorm(), b.squaredNorm());
}
inline Quat UndefinedRotation() { return {0, 0, 0, 0}; }  // standard_includes.h:56
inline Vec3 UndefinedPosition() { return {UNASSIGNED_VALUE, UNASSIGNED_VALUE, UNASSIGNED_VALUE}; }  // :57

// Eigen quaternion-from-rotation-matrix (Shoemake; Quaternion.h quaternionbase_assign_impl<Other,3,3>).
inline Quat quatFromMatrix(const Mat3& mat) {
  Quat q;
  double t = mat.m[0][0] + mat.m[1][1] + mat.m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (mat.m[2][1] - mat.m[1][2]) * t;
    q.y = (mat.m[0][2] - mat.m[2][0]) * t;
    q.z = (mat.m[1][0] - mat.m[0][1]) * t;
  } else {
    int i = 0;
    if (mat.m[1][1] > mat.m[0][0]) i = 1;
    if (mat.m[2][2] > mat.m[i][i]) i = 2;
    int j = (i + 1) % 3;
    int k = (j + 1) % 3;
    t = std::sqrt(mat.m[i][i] - mat.m[j][j] - mat.m[k][k] + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (mat.m[k][j] - mat.m[j][k]) * t;
    c[j] = (mat.m[j][i] + mat.m[i][j]) * t;
    c[k] = (mat.m[k][i] + mat.m[i][k]) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}

// Eigen AngleAxis -> quaternion: w = cos(a/2), vec = sin(a/2) * axis.
inline Quat quatFromAngleAxis(double angle, const Vec3& axis) {
  double ha = 0.5 * angle;
  double s = std::sin(ha);
  return {std::cos(ha), s * axis.x, s * axis.y, s * axis.z};
}
// Eigen AngleAxis::toRotationMatrix() * v  (used where the reference multiplies an AngleAxisd with a vector).
inline Vec3 angleAxisRotate(double angle, const Vec3& axis, const Vec3& v) {
  Mat3 res;
  double s = std::sin(angle), c = std::cos(angle);
  Vec3 sin_axis = s * axis;
  Vec3 cos1_axis = (1.0 - c) * axis;
  double tmp;
  tmp = cos1_axis.x * axis.y; res.m[0][1] = tmp - sin_axis.z; res.m[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z; res.m[0][2] = tmp + sin_axis.y; res.m[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z; res.m[1][2] = tmp - sin_axis.x; res.m[2][1] = tmp + sin_axis.x;
  res.m[0][0] = cos1_axis.x * axis.x + c;
  res.m[1][1] = cos1_axis.y * axis.y + c;
  res.m[2][2] = cos1_axis.z * axis.z + c;
  return res * v;
}
// Eigen 3.3 AngleAxis(const QuaternionBase&): angle = 2*atan2(|vec|, |w|), axis = vec / (+-|vec|).
inline void angleAxisFromQuat(const Quat& q, double* angle, Vec3* axis) {
  double n = q.vec().norm();
  if (n != 0.0) {
    *angle = 2.0 * std::atan2(n, std::fabs(q.w));
    if (q.w < 0.0) n = -n;
    *axis = q.vec() / n;
  } else {
    *angle = 0.0;
    *axis = Vec3(1, 0, 0);
  }
}

// Eigen Quaternion::FromTwoVectors / setFromTwoVectors.  The antiparallel fallback in Eigen takes the last right
// singular vector of [v0;v1] (JacobiSVD); that vector is not unique, so here any unit vector orthogonal to v0 is
// used instead.  No hot-path call site reaches the fallback (it needs exactly opposed vectors).
inline Quat fromTwoVectors(const Vec3& a, const Vec3& b) {
  Vec3 v0 = a.normalized();
  Vec3 v1 = b.normalized();
  double c = v1.dot(v0);
  Quat q;
  if (c < -1.0 + 1e-12) {
    c = std::max(c, -1.0);
    Vec3 helper = std::fabs(v0.x) < 0.9 ? Vec3(1, 0, 0) : Vec3(0, 1, 0);
    Vec3 axis = v0.cross(helper).normalized();
    double w2 = (1.0 + c) * 0.5;
    q.w = std::sqrt(w2);
    Vec3 v = axis * std::sqrt(1.0 - w2);
    q.x = v.x; q.y = v.y; q.z = v.z;
    return q;
  }
  Vec3 axis = v0.cross(v1);
  double s = std::sqrt((1.0 + c) * 2.0);
  double invs = 1.0 / s;
  q.x = axis.x * invs; q.y = axis.y * invs; q.z = axis.z * invs;
  q.w = s * 0.5;
  return q;
}

// Eigen QuaternionBase::slerp.
inline Quat slerp(const Quat& a, double t, const Quat& b) {
  const double one = 1.0 - 2.220446049250313e-16;
  double d = a.dot(b);
  double absD = std::fabs(d);
  double scale0, scale1;
  if (absD >= one) {
    scale0 = 1.0 - t;
    scale1 = t;
  } else {
    double theta = std::acos(absD);
    double sinTheta = std::sin(theta);
    scale0 = std::sin((1.0 - t) * theta) / sinTheta;
    scale1 = std::sin((t * theta)) / sinTheta;
  }
  if (d < 0.0) scale1 = -scale1;
  return {scale0 * a.w + scale1 * b.w, scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y,
          scale0 * a.z + scale1 * b.z};
}

// Eigen 3.3 MatrixBase::eulerAngles(a0,a1,a2) for a0 != a2 (Graphics Gems IV), first angle in [0,pi].
inline Vec3 eulerAngles(const Mat3& M, int a0, int a1, int a2) {
  (void)a2;
  Vec3 res;
  const int odd = ((a0 + 1) % 3 == a1) ? 0 : 1;
  const int i = a0;
  const int j = (a0 + 1 + odd) % 3;
  const int k = (a0 + 2 - odd) % 3;
  res[0] = std::atan2(M.m[j][k], M.m[k][k]);
  double c2 = std::sqrt(M.m[i][i] * M.m[i][i] + M.m[i][j] * M.m[i][j]);
  if ((odd && res[0] < 0.0) || ((!odd) && res[0] > 0.0)) {
    if (res[0] > 0.0) res[0] -= M_PI;
    else res[0] += M_PI;
    res[1] = std::atan2(-M.m[i][k], -c2);
  } else {
    res[1] = std::atan2(-M.m[i][k], c2);
  }
  double s1 = std::sin(res[0]);
  double c1 = std::cos(res[0]);
  res[2] = std::atan2(s1 * M.m[k][i] - c1 * M.m[j][i], c1 * M.m[j][j] - s1 * M.m[k][j]);
  if (!odd) res = -res;
  return res;
}

// ---- standard_includes.h scalar helpers ------------------------------------------------------------------------
inline double degreesToRadians(double d) { return d / 360.0 * 2.0 * M_PI; }       // :64
inline double radiansToDegrees(double r) { return (r / (2.0 * M_PI)) * 360.0; }   // :69
inline int mod(int a, int b) { return (a % b + b) % b; }                          // :76
inline double sqr(double v) { return v * v; }                                     // :82
inline double sign(double v) { return v > 0 ? 1 : -1; }                           // :88
inline int roundToInt(double x) { return x >= 0 ? int(x + 0.5) : -int(0.5 - x); } // :93
inline int roundToEvenInt(double x) { return int(x) % 2 == 0 ? int(x) : int(x) + 1; }  // :98
inline double clamped(double v, double lo, double hi) { return std::max(lo, std::min(v, hi)); }  // :106
inline int clampedInt(int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); }
inline double setPrecision(double v, int p) { return roundToInt(v * std::pow(10, p)) / std::pow(10, p); }  // :143
inline Vec3 setPrecision(const Vec3& v, int p) { return {setPrecision(v.x, p), setPrecision(v.y, p), setPrecision(v.z, p)}; }
inline double smoothStep(double c) {                                               // :163
  return 6.0 * std::pow(c, 5) - 15.0 * std::pow(c, 4) + 10.0 * std::pow(c, 3);
}
inline Vec3 getProjection(const Vec3& a, const Vec3& b) {                          // :173
  if (a.norm() == 0.0 || b.norm() == 0.0) return Vec3(0, 0, 0);
  return (a.dot(b) / b.dot(b)) * b;
}
inline Vec3 getRejection(const Vec3& a, const Vec3& b) { return a - getProjection(a, b); }  // :190
inline double interpolate(double o, double t, double c) { return (1.0 - c) * o + c * t; }   // :201
inline Vec3 interpolate(const Vec3& o, const Vec3& t, double c) { return (1.0 - c) * o + c * t; }
inline Quat correctRotation(const Quat& test, const Quat& ref) {                   // :211
  if (test.dot(ref) < 0.0) return {-test.w, -test.x, -test.y, -test.z};
  return test;
}
// clamped(vector, limit-vector) keeps the reference's upper bound limit[1] for every axis (:129-137, trap a22).
inline Vec3 clampedVecBuggy(const Vec3& v, const Vec3& limit) {
  Vec3 r;
  for (int i = 0; i < 3; ++i) r[i] = clamped(v[i], -limit[i], limit[1]);
  return r;
}
inline Quat eulerAnglesToQuaternion(const Vec3& e, bool intrinsic = false) {       // :227
  if (intrinsic)
    return quatFromAngleAxis(e[0], UnitX()) * quatFromAngleAxis(e[1], UnitY()) * quatFromAngleAxis(e[2], UnitZ());
  return quatFromAngleAxis(e[2], UnitZ()) * quatFromAngleAxis(e[1], UnitY()) * quatFromAngleAxis(e[0], UnitX());
}
inline Vec3 quaternionToEulerAngles(const Quat& q, bool intrinsic = false) {       // :248
  Vec3 result(0, 0, 0);
  if (intrinsic) result = eulerAngles(q.toRotationMatrix(), 0, 1, 2);
  else result = eulerAngles(q.toRotationMatrix(), 2, 1, 0);
  if (std::fabs(result[1]) > M_PI / 2 || std::fabs(result[2]) > M_PI / 2) {
    result[0] -= M_PI;
    if (result[1] > M_PI / 2.0) result[1] = -result[1] + M_PI;
    else if (result[1] < M_PI / 2.0) result[1] = -result[1] - M_PI;
    if (result[2] > M_PI / 2.0) result[2] -= M_PI;
    else if (result[2] < M_PI / 2.0) result[2] += M_PI;
  }
  return intrinsic ? result : Vec3(result[2], result[1], result[0]);
}
template <class T> inline T cubicBezier(const T* p, double t) {                    // :347
  double s = 1.0 - t;
  return p[0] * (s * s * s) + p[1] * (3.0 * t * s * s) + p[2] * (3.0 * t * t * s) + p[3] * (t * t * t);
}
template <class T> inline T quarticBezier(const T* p, double t) {                  // :402
  double s = 1.0 - t;
  return p[0] * (s * s * s * s) + p[1] * (4.0 * t * s * s * s) + p[2] * (6.0 * t * t * s * s) +
         p[3] * (4.0 * t * t * t * s) + p[4] * (t * t * t * t);
}
template <class T> inline T quarticBezierDot(const T* p, double t) {               // :415
  double s = 1.0 - t;
  return (4.0 * s * s * s * (p[1] - p[0]) + 12.0 * s * s * t * (p[2] - p[1]) + 12.0 * s * t * t * (p[3] - p[2]) +
          4.0 * t * t * t * (p[4] - p[3]));
}

// ---- small dense matrices (row-major, n <= 6) ------------------------------------------------------------------
struct MatX {
  int r = 0, c = 0;
  double a[36];
  MatX() { std::memset(a, 0, sizeof(a)); }
  MatX(int r_, int c_) : r(r_), c(c_) { std::memset(a, 0, sizeof(a)); }
  double& operator()(int i, int j) { return a[i * c + j]; }
  double operator()(int i, int j) const { return a[i * c + j]; }
  static MatX Identity(int n) {
    MatX m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
  MatX transpose() const {
    MatX t(c, r);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j);
    return t;
  }
  MatX operator*(const MatX& o) const {
    assert(c == o.r);
    MatX m(r, o.c);
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < o.c; ++j) {
        double s = 0.0;
        for (int k = 0; k < c; ++k) s += (*this)(i, k) * o(k, j);
        m(i, j) = s;
      }
    return m;
  }
  MatX operator+(const MatX& o) const {
    MatX m(r, c);
    for (int i = 0; i < r * c; ++i) m.a[i] = a[i] + o.a[i];
    return m;
  }
  MatX operator-(const MatX& o) const {
    MatX m(r, c);
    for (int i = 0; i < r * c; ++i) m.a[i] = a[i] - o.a[i];
    return m;
  }
  MatX operator*(double s) const {
    MatX m(r, c);
    for (int i = 0; i < r * c; ++i) m.a[i] = a[i] * s;
    return m;
  }
  // Dense inverse by Gauss-Jordan with partial pivoting.  Eigen uses PartialPivLU for dynamic sizes and cofactors
  // for fixed 4x4; every matrix inverted on this path is well conditioned (J J^T + 4e-4 I: cond <~ 1e2), so the
  // results agree to a few ulp.
  MatX inverse() const {
    assert(r == c);
    int n = r;
    MatX A = *this, I = Identity(n);
    for (int col = 0; col < n; ++col) {
      int piv = col;
      for (int i = col + 1; i < n; ++i)
        if (std::fabs(A(i, col)) > std::fabs(A(piv, col))) piv = i;
      if (piv != col)
        for (int j = 0; j < n; ++j) {
          std::swap(A(col, j), A(piv, j));
          std::swap(I(col, j), I(piv, j));
        }
      double d = A(col, col);
      for (int j = 0; j < n; ++j) {
        A(col, j) /= d;
        I(col, j) /= d;
      }
      for (int i = 0; i < n; ++i) {
        if (i == col) continue;
        double f = A(i, col);
        if (f == 0.0) continue;
        for (int j = 0; j < n; ++j) {
          A(i, j) -= f * A(col, j);
          I(i, j) -= f * I(col, j);
        }
      }
    }
    return I;
  }
};

struct Mat4 {  // homogeneous transform, row-major
  double m[4][4];
  static Mat4 Identity() {
    Mat4 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.m[i][j] = (i == j) ? 1.0 : 0.0;
    return r;
  }
  Mat4 operator*(const Mat4& o) const {
    Mat4 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += m[i][k] * o.m[k][j];
        r.m[i][j] = s;
      }
    return r;
  }
  Vec3 col3(int j) const { return {m[0][j], m[1][j], m[2][j]}; }
  Mat3 block3() const {
    Mat3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][j];
    return r;
  }
  Mat4 inverse() const {
    MatX a(4, 4);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) a(i, j) = m[i][j];
    MatX inv = a.inverse();
    Mat4 r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.m[i][j] = inv(i, j);
    return r;
  }
};

// Classical DH matrix (standard_includes.h:466).
inline Mat4 createDHMatrix(double d, double theta, double r, double alpha) {
  Mat4 m;
  m.m[0][0] = std::cos(theta); m.m[0][1] = -std::sin(theta) * std::cos(alpha); m.m[0][2] = std::sin(theta) * std::sin(alpha);  m.m[0][3] = r * std::cos(theta);
  m.m[1][0] = std::sin(theta); m.m[1][1] = std::cos(theta) * std::cos(alpha);  m.m[1][2] = -std::cos(theta) * std::sin(alpha); m.m[1][3] = r * std::sin(theta);
  m.m[2][0] = 0;               m.m[2][1] = std::sin(alpha);                    m.m[2][2] = std::cos(alpha);                    m.m[2][3] = d;
  m.m[3][0] = 0;               m.m[3][1] = 0;                                  m.m[3][2] = 0;                                  m.m[3][3] = 1;
  return m;
}

// ---- Pose (pose.h:17-216) --------------------------------------------------------------------------------------
struct Pose {
  Vec3 position_;
  Quat rotation_;
  Pose() = default;
  Pose(const Vec3& p, const Quat& q) : position_(p), rotation_(q) {}
  static Pose Identity() { return {Vec3(0, 0, 0), Quat::Identity()}; }                 // :199
  static Pose Undefined() { return {UndefinedPosition(), UndefinedRotation()}; }       // :206
  bool isValid() const {                                                               // :42
    return std::fabs(position_.x) < UNASSIGNED_VALUE && std::fabs(position_.y) < UNASSIGNED_VALUE &&
           std::fabs(position_.z) < UNASSIGNED_VALUE && std::fabs(rotation_.w) < UNASSIGNED_VALUE &&
           std::fabs(rotation_.x) < UNASSIGNED_VALUE && std::fabs(rotation_.y) < UNASSIGNED_VALUE &&
           std::fabs(rotation_.z) < UNASSIGNED_VALUE;
  }
  bool operator==(const Pose& o) const {                                               // :97
    return isApproxVec(position_, o.position_) && isApproxQuat(rotation_, o.rotation_);
  }
  bool operator!=(const Pose& o) const {                                               // :105
    return !isApproxVec(position_, o.position_) || !isApproxQuat(rotation_, o.rotation_);
  }
  Pose operator~() const {                                                             // :112
    return {rotation_.conjugate().transformVector(-position_), rotation_.conjugate()};
  }
  Pose transform(const Mat4& T) const {                                                // :135
    Pose r;
    double v[4] = {position_.x, position_.y, position_.z, 1.0};
    double o[4];
    for (int i = 0; i < 4; ++i) o[i] = T.m[i][0] * v[0] + T.m[i][1] * v[1] + T.m[i][2] * v[2] + T.m[i][3] * v[3];
    r.position_ = Vec3(o[0], o[1], o[2]);
    r.rotation_ = (quatFromMatrix(T.block3()) * rotation_).normalized();
    return r;
  }
  Vec3 transformVector(const Vec3& v) const { return position_ + rotation_.transformVector(v); }   // :151
  Vec3 inverseTransformVector(const Vec3& v) const { return (~*this).transformVector(v); }         // :159
  Pose addPose(const Pose& p) const {                                                  // :167
    Pose r = *this;
    r.position_ = transformVector(p.position_);
    r.rotation_ = r.rotation_ * p.rotation_;
    return r;
  }
  Pose removePose(const Pose& p) const {                                               // :178
    Pose r = *this;
    r.position_ = transformVector(-p.position_);
    r.rotation_ = r.rotation_ * p.rotation_.inverse();
    return r;
  }
  Pose interpolate(double c, const Pose& target) const {                               // :190
    Vec3 pos = c * target.position_ + (1.0 - c) * position_;
    Quat rot = slerp(rotation_, c, target.rotation_);
    return {pos, rot};
  }
};

}  // namespace om
