// shc_math.cuh — scalar/vector/quaternion primitives of the batched SHC engine, templated on the real type so the
// same source serves the fp64 ("reference-exact") and the mixed fp32 instantiations of the control-cycle kernel,
// and the host-side start-up code (compiled for the CPU by nvcc).
//
// Semantics follow the reference's helpers (include/syropod_highlevel_controller/standard_includes.h:64-291,
// pose.h:112-195) and the Eigen 3.3 routines they call; the implementation is new.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SHC_HD __host__ __device__ __forceinline__
#else
#define SHC_HD inline
#endif

namespace shc {

constexpr double kPi = 3.14159265358979323846;

// ---- overload set so templated code picks the right-precision routine ------------------------------------------
SHC_HD float rsqrt_(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
// Device 1/sqrt for POSITIVE NORMAL arguments (sums of squares plus lambda^2, squared norms guarded by "> 0"): the
// MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) refined by two Newton steps with the residual formed by FMA, < 1 ulp;
// the library routine's zero / infinity / denormal handling is not needed on these call sites.
SHC_HD double rsqrt_(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-h * y, y, 0.5);
  y = fma(y, e, y);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}
SHC_HD float sqrt_(float x) { return sqrtf(x); }
SHC_HD double sqrt_(double x) { return sqrt(x); }
SHC_HD float abs_(float x) { return fabsf(x); }
SHC_HD double abs_(double x) { return fabs(x); }
SHC_HD float atan2_(float y, float x) { return atan2f(y, x); }
SHC_HD double atan2_(double y, double x) { return atan2(y, x); }
SHC_HD float acos_(float x) { return acosf(x); }
SHC_HD double acos_(double x) { return acos(x); }
SHC_HD float sin_(float x) { return sinf(x); }
SHC_HD double sin_(double x) { return sin(x); }
SHC_HD float cos_(float x) { return cosf(x); }
SHC_HD double cos_(double x) { return cos(x); }
SHC_HD float tan_(float x) { return tanf(x); }
SHC_HD double tan_(double x) { return tan(x); }
SHC_HD void sincos_(float a, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(a, s, c);
#else
  *s = sinf(a); *c = cosf(a);
#endif
}
// Device sine/cosine pair for BOUNDED arguments (joint angles, Euler angles, half-angles: |a| well below 1e5 rad, so
// the three-term Cody-Waite reduction by pi/2 is exact enough and no Payne-Hanek slow path is needed), with the
// fdlibm kernel polynomials on [-pi/4, pi/4]: <= 1.6 ulp (measured against long double over +-1000 rad), like the CUDA
// library routine, at about half its instruction count and without its stack frame.
SHC_HD void sincos_(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  double j = fma(a, 0.6366197723675814, 6755399441055744.0);  // nearest integer to a * 2/pi in the low mantissa bits
  const int q = __double2loint(j);
  j -= 6755399441055744.0;
  double r = fma(-j, 1.5707963267948966, a);
  r = fma(-j, 6.123233995736766e-17, r);
  r = fma(-j, -1.4973849048591698e-33, r);
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double ss = (q & 1) ? cr : sr;
  const double cc = (q & 1) ? sr : cr;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
#else
  *s = sin(a); *c = cos(a);
#endif
}
template <class R> SHC_HD R min_(R a, R b) { return a < b ? a : b; }
template <class R> SHC_HD R max_(R a, R b) { return a > b ? a : b; }
// standard_includes.h:106 — std::max(lo, std::min(v, hi))
template <class R> SHC_HD R clamp_(R v, R lo, R hi) { return max_(lo, min_(v, hi)); }
// standard_includes.h:88 — +1 for > 0, otherwise -1
template <class R> SHC_HD R sign_(R v) { return v > R(0) ? R(1) : R(-1); }
SHC_HD int imod(int a, int b) { return (a % b + b) % b; }                              // :76
SHC_HD int round_to_int(double x) { return x >= 0 ? int(x + 0.5) : -int(0.5 - x); }    // :93
SHC_HD int round_to_even_int(double x) { return int(x) % 2 == 0 ? int(x) : int(x) + 1; }  // :98
// standard_includes.h:163 (6c^5 - 15c^4 + 10c^3), Horner form
template <class R> SHC_HD R smooth_step(R c) { return c * c * c * (c * (R(6) * c - R(15)) + R(10)); }

// ---- 3-vectors ---------------------------------------------------------------------------------------------------
template <class R> struct V3 {
  R x, y, z;
};
template <class R> SHC_HD V3<R> v3(R x, R y, R z) { return V3<R>{x, y, z}; }
template <class R> SHC_HD V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> SHC_HD V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> SHC_HD V3<R> operator-(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R> SHC_HD V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> SHC_HD V3<R> operator*(R s, V3<R> a) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> SHC_HD R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> SHC_HD V3<R> cross(V3<R> a, V3<R> b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class R> SHC_HD R norm2(V3<R> a) { return dot(a, a); }
template <class R> SHC_HD R norm(V3<R> a) { return sqrt_(dot(a, a)); }
// Eigen normalized(): unchanged when the squared norm is not > 0
template <class R> SHC_HD V3<R> normalized(V3<R> a) {
  R n2 = dot(a, a);
  return n2 > R(0) ? a * rsqrt_(n2) : a;
}
template <class A, class B> SHC_HD V3<A> cvt(V3<B> v) { return {A(v.x), A(v.y), A(v.z)}; }
// standard_includes.h:173 / :190
template <class R> SHC_HD V3<R> projection(V3<R> a, V3<R> b) {
  R bb = dot(b, b);
  if (dot(a, a) == R(0) || bb == R(0)) return {R(0), R(0), R(0)};
  return b * (dot(a, b) / bb);
}
template <class R> SHC_HD V3<R> rejection(V3<R> a, V3<R> b) { return a - projection(a, b); }

// ---- quaternions (w,x,y,z) --------------------------------------------------------------------------------------
template <class R> struct Q4 {
  R w, x, y, z;
};
template <class R> SHC_HD Q4<R> qidentity() { return {R(1), R(0), R(0), R(0)}; }
template <class A, class B> SHC_HD Q4<A> cvt(Q4<B> q) { return {A(q.w), A(q.x), A(q.y), A(q.z)}; }
template <class R> SHC_HD Q4<R> qmul(Q4<R> a, Q4<R> b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
template <class R> SHC_HD Q4<R> qconj(Q4<R> q) { return {q.w, -q.x, -q.y, -q.z}; }
template <class R> SHC_HD R qdot(Q4<R> a, Q4<R> b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
// Eigen inverse(): conj / |q|^2, all-zero when |q| = 0
template <class R> SHC_HD Q4<R> qinverse(Q4<R> q) {
  R n2 = qdot(q, q);
  if (n2 > R(0)) {
    R i = R(1) / n2;
    return {q.w * i, -q.x * i, -q.y * i, -q.z * i};
  }
  return {R(0), R(0), R(0), R(0)};
}
template <class R> SHC_HD Q4<R> qnormalized(Q4<R> q) {
  R n2 = qdot(q, q);
  if (n2 > R(0)) {
    R i = R(1) / sqrt_(n2);
    return {q.w * i, q.x * i, q.y * i, q.z * i};
  }
  return q;
}
// Eigen _transformVector: v + w*(2 u x v) + u x (2 u x v)
template <class R> SHC_HD V3<R> qrot(Q4<R> q, V3<R> v) {
  V3<R> u{q.x, q.y, q.z};
  V3<R> uv = cross(u, v);
  uv = uv + uv;
  return v + uv * q.w + cross(u, uv);
}
// standard_includes.h:211
template <class R> SHC_HD Q4<R> correct_rotation(Q4<R> t, Q4<R> ref) {
  if (qdot(t, ref) < R(0)) return {-t.w, -t.x, -t.y, -t.z};
  return t;
}
template <class R> SHC_HD Q4<R> q_axis_x(R a) { R s, c; sincos_(R(0.5) * a, &s, &c); return {c, s, R(0), R(0)}; }
template <class R> SHC_HD Q4<R> q_axis_y(R a) { R s, c; sincos_(R(0.5) * a, &s, &c); return {c, R(0), s, R(0)}; }
template <class R> SHC_HD Q4<R> q_axis_z(R a) { R s, c; sincos_(R(0.5) * a, &s, &c); return {c, R(0), R(0), s}; }
// standard_includes.h:227
template <class R> SHC_HD Q4<R> euler_to_quat(V3<R> e, bool intrinsic) {
  if (intrinsic) return qmul(qmul(q_axis_x(e.x), q_axis_y(e.y)), q_axis_z(e.z));
  return qmul(qmul(q_axis_z(e.z), q_axis_y(e.y)), q_axis_x(e.x));
}
// Rotation matrix entries of a quaternion (Eigen toRotationMatrix), row-major m[3][3]
template <class R> SHC_HD void quat_to_matrix(Q4<R> q, R m[3][3]) {
  const R tx = R(2) * q.x, ty = R(2) * q.y, tz = R(2) * q.z;
  const R twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const R txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const R tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0][0] = R(1) - (tyy + tzz); m[0][1] = txy - twz;          m[0][2] = txz + twy;
  m[1][0] = txy + twz;          m[1][1] = R(1) - (txx + tzz); m[1][2] = tyz - twx;
  m[2][0] = txz - twy;          m[2][1] = tyz + twx;          m[2][2] = R(1) - (txx + tyy);
}
// Eigen Quaternion(Matrix3) (QuaternionBase::operator=(MatrixBase), "Shoemake" branch on the trace), row-major m[3][3]
template <class R> SHC_HD Q4<R> matrix_to_quat(const R m[3][3]) {
  R t = m[0][0] + m[1][1] + m[2][2];
  R q[4];  // x y z w
  if (t > R(0)) {
    t = sqrt_(t + R(1));
    q[3] = R(0.5) * t;
    t = R(0.5) / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt_(m[i][i] - m[j][j] - m[k][k] + R(1));
    q[i] = R(0.5) * t;
    t = R(0.5) / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  return {q[3], q[0], q[1], q[2]};
}
// Eigen 3.3 eulerAngles(a0,a1,a2), a0 != a2 (first angle in [0,pi]) followed by the reference's range fix-up
// (standard_includes.h:248-291).  Returns (roll, pitch, yaw).
template <class R> SHC_HD V3<R> quat_to_euler(Q4<R> q, bool intrinsic) {
  R M[3][3];
  quat_to_matrix(q, M);
  const int a0 = intrinsic ? 0 : 2, a1 = 1;
  const int odd = ((a0 + 1) % 3 == a1) ? 0 : 1;
  const int i = a0, j = (a0 + 1 + odd) % 3, k = (a0 + 2 - odd) % 3;
  const R pi = R(kPi);
  R r0 = atan2_(M[j][k], M[k][k]);
  R c2 = sqrt_(M[i][i] * M[i][i] + M[i][j] * M[i][j]);
  R r1;
  if ((odd && r0 < R(0)) || ((!odd) && r0 > R(0))) {
    if (r0 > R(0)) r0 -= pi;
    else r0 += pi;
    r1 = atan2_(-M[i][k], -c2);
  } else {
    r1 = atan2_(-M[i][k], c2);
  }
  R s1, c1;
  sincos_(r0, &s1, &c1);
  R r2 = atan2_(s1 * M[k][i] - c1 * M[j][i], c1 * M[j][j] - s1 * M[k][j]);
  if (!odd) { r0 = -r0; r1 = -r1; r2 = -r2; }
  const R hp = pi / R(2);
  if (abs_(r1) > hp || abs_(r2) > hp) {
    r0 -= pi;
    if (r1 > hp) r1 = -r1 + pi;
    else if (r1 < hp) r1 = -r1 - pi;
    if (r2 > hp) r2 -= pi;
    else if (r2 < hp) r2 += pi;
  }
  return intrinsic ? V3<R>{r0, r1, r2} : V3<R>{r2, r1, r0};
}
// Eigen FromTwoVectors(a, b) for non-opposed vectors (every hot-path call site has a = unit axis, b within 90 deg)
template <class R> SHC_HD Q4<R> from_two_vectors(V3<R> a, V3<R> b) {
  V3<R> v0 = normalized(a), v1 = normalized(b);
  R c = dot(v1, v0);
  V3<R> axis = cross(v0, v1);
  R s = sqrt_((R(1) + c) * R(2));
  R invs = R(1) / s;
  return {s * R(0.5), axis.x * invs, axis.y * invs, axis.z * invs};
}
// Eigen AngleAxis(Quaternion): angle = 2 atan2(|vec|, |w|), axis = vec / (+-|vec|); (0, x-axis) for a null vector part.
// Returns axis * angle, the rotation vector the reference feeds to solveIK (model.cpp:893-894).
template <class R> SHC_HD V3<R> rotation_vector(Q4<R> q) {
  V3<R> v{q.x, q.y, q.z};
  R n = norm(v);
  if (n == R(0)) return {R(0), R(0), R(0)};
  const R angle = R(2) * atan2_(n, abs_(q.w));
  if (q.w < R(0)) n = -n;
  return V3<R>{v.x / n, v.y / n, v.z / n} * angle;
}
template <class R> struct Eps;
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; };
template <> struct Eps<float> { static constexpr float v = 1.1920929e-07f; };
// Eigen slerp
template <class R> SHC_HD Q4<R> slerp(Q4<R> a, R t, Q4<R> b) {
  const R one = R(1) - Eps<R>::v;
  R d = qdot(a, b);
  R ad = abs_(d);
  R s0, s1;
  if (ad >= one) {
    s0 = R(1) - t;
    s1 = t;
  } else {
    R th = acos_(ad);
    R st = sin_(th);
    s0 = sin_((R(1) - t) * th) / st;
    s1 = sin_(t * th) / st;
  }
  if (d < R(0)) s1 = -s1;
  return {s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}

// ---- poses (pose.h) ---------------------------------------------------------------------------------------------
template <class R> struct PoseT {
  V3<R> p;
  Q4<R> q;
};
template <class R> SHC_HD PoseT<R> pose_identity() { return {{R(0), R(0), R(0)}, qidentity<R>()}; }
template <class A, class B> SHC_HD PoseT<A> cvt(PoseT<B> a) { return {cvt<A>(a.p), cvt<A>(a.q)}; }
template <class R> SHC_HD V3<R> pose_transform(PoseT<R> a, V3<R> v) { return a.p + qrot(a.q, v); }   // :151
template <class R> SHC_HD PoseT<R> pose_inverse(PoseT<R> a) {                                         // :112
  Q4<R> c = qconj(a.q);
  return {qrot(c, -a.p), c};
}
template <class R> SHC_HD V3<R> pose_inverse_transform(PoseT<R> a, V3<R> v) {                         // :159
  return pose_transform(pose_inverse(a), v);
}
template <class R> SHC_HD PoseT<R> pose_add(PoseT<R> a, PoseT<R> b) {                                 // :167
  return {pose_transform(a, b.p), qmul(a.q, b.q)};
}
template <class R> SHC_HD PoseT<R> pose_remove(PoseT<R> a, PoseT<R> b) {                              // :178
  return {pose_transform(a, -b.p), qmul(a.q, qinverse(b.q))};
}
template <class R> SHC_HD PoseT<R> pose_interpolate(PoseT<R> a, R c, PoseT<R> t) {                    // :190
  return {t.p * c + a.p * (R(1) - c), slerp(a.q, c, t.q)};
}

// ---- quartic Bezier (standard_includes.h:402-420) -----------------------------------------------------------------
template <class R> SHC_HD V3<R> quartic_bezier(const V3<R> p[5], R t) {
  R s = R(1) - t;
  return p[0] * (s * s * s * s) + p[1] * (R(4) * t * s * s * s) + p[2] * (R(6) * t * t * s * s) +
         p[3] * (R(4) * t * t * t * s) + p[4] * (t * t * t * t);
}
template <class R> SHC_HD V3<R> quartic_bezier_dot(V3<R> p0, V3<R> p1, V3<R> p2, V3<R> p3, V3<R> p4, R t) {
  R s = R(1) - t;
  return (p1 - p0) * (R(4) * s * s * s) + (p2 - p1) * (R(12) * s * s * t) + (p3 - p2) * (R(12) * s * t * t) +
         (p4 - p3) * (R(4) * t * t * t);
}

}  // namespace shc
