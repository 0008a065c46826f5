// shim_eigen_check.cpp — TEST INFRASTRUCTURE ONLY: extern "C" probes of the stand-in Eigen (oracle/shim/Eigen) so that
// tests/test_shim_eigen.py can check it against numpy / scipy.  The stand-in is what the reference's own sources are compiled
// against for the parity pin (oracle/Makefile.ref); the reference's arithmetic that runs through it deserves its own check.
#include <vector>

#include "Eigen/Geometry"
#include "boost/numeric/odeint.hpp"

using namespace Eigen;
static Quaterniond Q(const double* q) { return Quaterniond(q[0], q[1], q[2], q[3]); }  // w x y z
static void put(double* o, const Quaterniond& q) { o[0] = q.w(); o[1] = q.x(); o[2] = q.y(); o[3] = q.z(); }

extern "C" {
void shim_quat_mul(const double* a, const double* b, double* out) { put(out, Q(a) * Q(b)); }
void shim_quat_rotate(const double* q, const double* v, double* out) {
  Vector3d r = Q(q)._transformVector(Vector3d(v[0], v[1], v[2]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void shim_quat_to_matrix(const double* q, double* out) {  // row-major 3x3
  Matrix3d m = Q(q).toRotationMatrix();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[i * 3 + j] = m(i, j);
}
void shim_matrix_to_quat(const double* m9, double* out) {
  Matrix3d m;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = m9[i * 3 + j];
  put(out, Quaterniond(m));
}
void shim_quat_inverse(const double* q, double* out) { put(out, Q(q).inverse()); }
void shim_slerp(const double* a, double t, const double* b, double* out) { put(out, Q(a).slerp(t, Q(b))); }
void shim_from_two_vectors(const double* a, const double* b, double* out) {
  put(out, Quaterniond::FromTwoVectors(Vector3d(a[0], a[1], a[2]), Vector3d(b[0], b[1], b[2])));
}
void shim_euler_angles(const double* q, int a0, int a1, int a2, double* out) {
  Vector3d e = Q(q).toRotationMatrix().eulerAngles(a0, a1, a2);
  out[0] = e[0]; out[1] = e[1]; out[2] = e[2];
}
void shim_angle_axis(const double* q, double* out) {  // angle, axis xyz
  AngleAxisd aa(Q(q));
  out[0] = aa.angle(); out[1] = aa.axis()[0]; out[2] = aa.axis()[1]; out[3] = aa.axis()[2];
}
void shim_angle_axis_rotate(double angle, const double* axis, const double* v, double* out) {
  Vector3d r = AngleAxisd(angle, Vector3d(axis[0], axis[1], axis[2]))._transformVector(Vector3d(v[0], v[1], v[2]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void shim_angle_axis_to_quat(double angle, const double* axis, double* out) {
  put(out, Quaterniond(AngleAxisd(angle, Vector3d(axis[0], axis[1], axis[2]))));
}
void shim_inverse(const double* a, int n, double* out) {  // row-major n x n, dynamic matrix
  MatrixXd m(n, n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) m(i, j) = a[i * n + j];
  MatrixXd r = m.inverse();
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[i * n + j] = r(i, j);
}
void shim_matmul(const double* a, const double* b, int r, int k, int c, double* out) {
  MatrixXd A(r, k), B(k, c);
  for (int i = 0; i < r; ++i) for (int j = 0; j < k; ++j) A(i, j) = a[i * k + j];
  for (int i = 0; i < k; ++i) for (int j = 0; j < c; ++j) B(i, j) = b[i * c + j];
  MatrixXd C = A * B.transpose().transpose();
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) out[i * c + j] = C(i, j);
}
int shim_is_approx(const double* a, const double* b) { return Q(a).isApprox(Q(b)) ? 1 : 0; }
void shim_block_write(double* m16, int col, const double* v3) {  // Matrix4d: block<3,1>(0, col) = v, row-major in/out
  Matrix4d m;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = m16[i * 4 + j];
  m.block<3, 1>(0, col) = Vector3d(v3[0], v3[1], v3[2]);
  Vector3d back = m.block<3, 1>(0, col);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m16[i * 4 + j] = m(i, j);
  m16[15] = back[0] + back[1] + back[2];
}
// Boost.Odeint stand-in: the admittance controller's call (admittance_controller.cpp:42-52) on x'' = -f/m - c/m x' - k/m x;
// returns the number of steps taken.
int shim_admittance_integrate(double f, double m, double c, double k, double step_time, double* x2) {
  std::vector<double> x = {x2[0], x2[1]};
  boost::numeric::odeint::runge_kutta4<std::vector<double>> stepper;
  size_t steps = integrate_const(stepper,
                                 [&](const std::vector<double>& s, std::vector<double>& d, double) {
                                   d[0] = s[1];
                                   d[1] = -f / m - c / m * s[1] - k / m * s[0];
                                 },
                                 x, 0.0, step_time, step_time / 30);
  x2[0] = x[0]; x2[1] = x[1];
  return int(steps);
}
}
