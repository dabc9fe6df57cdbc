// =====================================================================================
// oracle/te_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (see te_oracle.hpp header).
// PARITY UNPINNED by reference golden vectors (none exist); see te_oracle.hpp.
// Build: g++ -O2 -std=c++17 -ffp-contract=off  (no FMA contraction: the reference's
// default x86-64 build has none).
// =====================================================================================
#include "te_oracle.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

namespace oracle {

// -------------------------------------------------------------------------------------
// Dense helpers.  (Eigen) products are materialised left-to-right; inner products
// accumulate in increasing k.
// -------------------------------------------------------------------------------------
Mat mul(const Mat& A, const Mat& B) {
  assert(A.c == B.r);
  Mat C(A.r, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int i = 0; i < A.r; ++i) {
      double s = A(i, 0) * B(0, j);
      for (int k = 1; k < A.c; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < A.c; ++j) T(j, i) = A(i, j);
  return T;
}
Mat add(const Mat& A, const Mat& B) {
  assert(A.r == B.r && A.c == B.c);
  Mat C(A.r, A.c);
  for (size_t i = 0; i < C.d.size(); ++i) C.d[i] = A.d[i] + B.d[i];
  return C;
}
Mat sub(const Mat& A, const Mat& B) {
  assert(A.r == B.r && A.c == B.c);
  Mat C(A.r, A.c);
  for (size_t i = 0; i < C.d.size(); ++i) C.d[i] = A.d[i] - B.d[i];
  return C;
}
Vec mulv(const Mat& A, const Vec& x) {
  assert(A.c == (int)x.size());
  Vec y(A.r);
  for (int i = 0; i < A.r; ++i) {
    double s = A(i, 0) * x[0];
    for (int k = 1; k < A.c; ++k) s += A(i, k) * x[k];
    y[i] = s;
  }
  return y;
}

// (Eigen) MatrixXd::inverse() on a dynamic matrix = PartialPivLU(M).inverse():
// unblocked Doolittle LU with row partial pivoting (first largest |entry|), then
// solve L U X = P I by unit-lower forward and upper backward substitution (the upper
// solve multiplies by the reciprocal of the diagonal).  SURVEY.md Appendix B.
Mat inversePartialPivLU(const Mat& M) {
  const int n = M.r;
  assert(M.r == M.c);
  Mat lu = M;
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  std::vector<int> transpositions(n);
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(lu(k, k));
    for (int i = k + 1; i < n; ++i) {
      double v = std::fabs(lu(i, k));
      if (v > best) { best = v; piv = i; }
    }
    transpositions[k] = piv;
    if (best != 0.0) {
      if (piv != k)
        for (int j = 0; j < n; ++j) std::swap(lu(k, j), lu(piv, j));
      for (int i = k + 1; i < n; ++i) lu(i, k) /= lu(k, k);
    }
    for (int j = k + 1; j < n; ++j)
      for (int i = k + 1; i < n; ++i) lu(i, j) -= lu(i, k) * lu(k, j);
  }
  // permutation P = product of transpositions, applied to identity rows
  Mat X = Mat::Identity(n);
  for (int k = 0; k < n; ++k)
    if (transpositions[k] != k)
      for (int j = 0; j < n; ++j) std::swap(X(k, j), X(transpositions[k], j));
  // unit-lower forward substitution
  for (int j = 0; j < n; ++j)
    for (int k = 0; k < n; ++k) {
      double xk = X(k, j);
      for (int i = k + 1; i < n; ++i) X(i, j) -= xk * lu(i, k);
    }
  // upper backward substitution
  for (int j = 0; j < n; ++j)
    for (int k = n - 1; k >= 0; --k) {
      double a = 1.0 / lu(k, k);
      X(k, j) *= a;
      double xk = X(k, j);
      for (int i = 0; i < k; ++i) X(i, j) -= xk * lu(i, k);
    }
  return X;
}

// -------------------------------------------------------------------------------------
// geometry.hpp
// -------------------------------------------------------------------------------------
double constrainAngle(double x) {   // geometry.hpp:31-36
  x = std::fmod(x + M_PI, 2 * M_PI);
  if (x < 0) x += 2 * M_PI;
  return x - M_PI;
}
double angleConv(double angle) {   // geometry.hpp:43-45
  return std::fmod(constrainAngle(angle), 2 * M_PI);
}
double angleDiff(double a, double b) {   // geometry.hpp:53-58
  double dif = std::fmod(b - a + M_PI, 2 * M_PI);
  if (dif < 0) dif += 2 * M_PI;
  return dif - M_PI;
}
void unwrap3(const double prev[3], const double nw[3], double out[3]) {   // geometry.hpp:70-76
  for (unsigned i = 0; i < 3; i++) out[i] = prev[i] - angleDiff(nw[i], angleConv(prev[i]));
}
double wrapMax(double x, double max) { return std::fmod(max + std::fmod(x, max), max); }       // :79-83
double wrapMinMax(double x, double min, double max) { return min + wrapMax(x - min, max - min); }  // :85-88

void quatNormalize(Quat& q) {   // (Eigen) coeffs /= norm
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

void quatToRpy(const Quat& q, double rpy[3]) {   // geometry.hpp:154-176
  if (-2 * (q.x * q.z - q.w * q.y) > 0.9999) {
    rpy[0] = 0;
    rpy[1] = M_PI / 2;
    rpy[2] = 2 * std::atan2(q.z, q.w);
  } else if (-2 * (q.x * q.z - q.w * q.y) < -0.9999) {
    rpy[0] = 0;
    rpy[1] = -M_PI / 2;
    rpy[2] = 2 * std::atan2(q.z, q.w);
  } else {
    rpy[0] = std::atan2(2 * (q.y * q.z + q.w * q.x), (q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z));
    rpy[1] = std::asin(-2 * (q.x * q.z - q.w * q.y));
    rpy[2] = std::atan2(2 * (q.x * q.y + q.w * q.z), (q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z));
  }
}

void rpyToQuat(const double rpy[3], Quat& q) {   // geometry.hpp:178-189
  double phi = rpy[0] / 2, the = rpy[1] / 2, psi = rpy[2] / 2;
  q.w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
  q.x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
  q.y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
  q.z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
  quatNormalize(q);
}

void rotToRpy(const Mat3& R, double rpy[3]) {   // geometry.hpp:191-196
  rpy[0] = std::atan2(R.m[2][1], R.m[2][2]);
  rpy[1] = std::atan2(-R.m[2][0], std::sqrt(R.m[2][1] * R.m[2][1] + R.m[2][2] * R.m[2][2]));
  rpy[2] = std::atan2(R.m[1][0], R.m[0][0]);
}

Mat3 quatToRotationMatrix(const Quat& q) {   // (Eigen) QuaternionBase::toRotationMatrix
  Mat3 R;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz;       R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;       R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;       R.m[2][1] = tyz + twx;       R.m[2][2] = 1 - (txx + tyy);
  return R;
}

Quat rotationMatrixToQuat(const Mat3& R) {   // (Eigen) quaternionbase_assign_impl<Matrix3>
  double qv[4];   // x y z w
  double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    qv[3] = 0.5 * t;
    t = 0.5 / t;
    qv[0] = (R.m[2][1] - R.m[1][2]) * t;
    qv[1] = (R.m[0][2] - R.m[2][0]) * t;
    qv[2] = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (R.m[1][1] > R.m[0][0]) i = 1;
    if (R.m[2][2] > R.m[i][i]) i = 2;
    int j = (i + 1) % 3;
    int k = (j + 1) % 3;
    t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    qv[3] = (R.m[k][j] - R.m[j][k]) * t;
    qv[j] = (R.m[j][i] + R.m[i][j]) * t;
    qv[k] = (R.m[k][i] + R.m[i][k]) * t;
  }
  Quat q; q.x = qv[0]; q.y = qv[1]; q.z = qv[2]; q.w = qv[3];
  return q;
}

void rpyToEarBase(const double rpy[3], Mat3& E) {   // geometry.hpp:333-351
  double c_r = std::cos(rpy[0]), s_r = std::sin(rpy[0]);
  double c_p = std::cos(rpy[1]), s_p = std::sin(rpy[1]);
  E.m[0][0] = 1; E.m[0][1] = 0;    E.m[0][2] = -s_p;
  E.m[1][0] = 0; E.m[1][1] = c_r;  E.m[1][2] = c_p * s_r;
  E.m[2][0] = 0; E.m[2][1] = -s_r; E.m[2][2] = c_p * c_r;
}

void rpyToEarBaseInv(const double rpy[3], Mat3& E) {   // geometry.hpp:359-374
  double c_r = std::cos(rpy[0]), s_r = std::sin(rpy[0]);
  double c_p = std::cos(rpy[1]), s_p = std::sin(rpy[1]);
  E.m[0][0] = 1; E.m[0][1] = (s_p * s_r) / c_p; E.m[0][2] = (c_r * s_p) / c_p;
  E.m[1][0] = 0; E.m[1][1] = c_r;               E.m[1][2] = -s_r;
  E.m[2][0] = 0; E.m[2][1] = s_r / c_p;         E.m[2][2] = c_r / c_p;
}

Mat3 EarBaseInvJacobianRpy(const double rpy[3], const double omega[3], double dt) {   // geometry.hpp:394-410
  Mat3 o;
  double wy = omega[1], wz = omega[2];
  double c_r = std::cos(rpy[0]), c_p = std::cos(rpy[1]);
  double s_r = std::sin(rpy[0]), s_p = std::sin(rpy[1]);
  o.m[0][0] = (dt * (wy * c_r * s_p - wz * s_p * s_r)) / c_p + 1;
  o.m[0][1] = (dt * (wz * c_r + wy * s_r)) / (c_p * c_p);
  o.m[0][2] = 0;
  o.m[1][0] = -dt * (wz * c_r + wy * s_r);
  o.m[1][1] = 1;
  o.m[1][2] = 0;
  o.m[2][0] = (dt * (wy * c_r - wz * s_r)) / c_p;
  o.m[2][1] = (dt * s_p * (wz * c_r + wy * s_r)) / (c_p * c_p);
  o.m[2][2] = 1;
  return o;
}

Mat3 EarBaseInvJacobianOmega(const double rpy[3], double dt) {   // geometry.hpp:412-426
  Mat3 o;
  double c_r = std::cos(rpy[0]), c_p = std::cos(rpy[1]);
  double s_r = std::sin(rpy[0]), s_p = std::sin(rpy[1]);
  o.m[0][0] = dt; o.m[0][1] = (dt * s_p * s_r) / c_p; o.m[0][2] = (dt * c_r * s_p) / c_p;
  o.m[1][0] = 0;  o.m[1][1] = dt * c_r;               o.m[1][2] = -dt * s_r;
  o.m[2][0] = 0;  o.m[2][1] = (dt * s_r) / c_p;       o.m[2][2] = (dt * c_r) / c_p;
  return o;
}

void Qtran(double dt, const double omega[3], double Q[4][4]) {   // geometry.hpp:448-465,493-504
  double omega_norm = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
  double tmp = omega_norm * dt / 2.0;
  double S[4][4] = {{0, -omega[2], omega[1], omega[0]},
                    {omega[2], 0, -omega[0], omega[1]},
                    {-omega[1], omega[0], 0, omega[2]},
                    {-omega[0], -omega[1], -omega[2], 0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i][j] = 0.5 * S[i][j];
  if (omega_norm > 0.0) {
    double c = std::cos(tmp);
    double f = 2.0 / omega_norm * std::sin(tmp);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) Q[i][j] = c * (i == j ? 1.0 : 0.0) + f * S[i][j];
  } else {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) Q[i][j] = (i == j ? 1.0 : 0.0);
  }
}

void pose7dToPose6d(const double p7[7], double p6[6]) {   // geometry.hpp:619-628
  p6[0] = p7[0]; p6[1] = p7[1]; p6[2] = p7[2];
  Quat q; q.x = p7[3]; q.y = p7[4]; q.z = p7[5]; q.w = p7[6];
  quatNormalize(q);
  quatToRpy(q, p6 + 3);
}

Quat quatMul(const Quat& a, const Quat& b) {   // (Eigen) quat_product
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
Quat quatInverse(const Quat& q) {   // (Eigen) QuaternionBase::inverse
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  Quat r;
  if (n2 > 0.0) { r.w = q.w / n2; r.x = -q.x / n2; r.y = -q.y / n2; r.z = -q.z / n2; }
  else { r.x = r.y = r.z = r.w = 0.0; }
  return r;
}
double computeQuaternionErrorAngle(const Quat& q_des, const Quat& q) {   // geometry.hpp:630-657
  Quat q_e = quatMul(q_des, quatInverse(q));
  quatNormalize(q_e);
  return 2 * std::acos(q_e.w);
}
double toSec(uint32_t sec, uint32_t nsec) {   // utils.hpp:59-62
  return static_cast<double>(sec) + 1e-9 * static_cast<double>(nsec);
}

// -------------------------------------------------------------------------------------
// kalman.cpp
// -------------------------------------------------------------------------------------
void KalmanFilterInterface::init(const Vec& x0) {   // src/kalman.cpp:16-21
  x_hat_ = x0;
  P_ = P0_;
  initialized_ = true;
}
void KalmanFilterInterface::update(const Vec& y) {   // src/kalman.cpp:30-42
  if (!initialized_) throw std::runtime_error("Filter is not initialized!");
  predict();
  estimate(y);
  x_hat_ = x_hat_new_;
}
void KalmanFilterInterface::update() {   // src/kalman.cpp:44-54
  if (!initialized_) throw std::runtime_error("Filter is not initialized!");
  predict();
  x_hat_ = x_hat_new_;
}

LinearKalmanFilter::LinearKalmanFilter(const Mat& A, const Mat& C, const Mat& Q, const Mat& R, const Mat& P) {  // :62-82
  A_ = A; C_ = C; Q_ = Q; R_ = R; P0_ = P;
  m_ = C.rows();
  n_ = A.rows();
  initialized_ = false;
  I_ = Mat::Identity(n_);
  K_ = Mat(n_, m_);
  x_hat_.assign(n_, 0.0);
  x_hat_new_.assign(n_, 0.0);
}
void LinearKalmanFilter::predict() {   // src/kalman.cpp:84-88
  x_hat_new_ = mulv(A_, x_hat_);
  P_ = add(mul(mul(A_, P_), transpose(A_)), Q_);
}
void LinearKalmanFilter::estimate(const Vec& y) {   // src/kalman.cpp:90-95
  Mat Ct = transpose(C_);
  K_ = mul(mul(P_, Ct), inversePartialPivLU(add(mul(mul(C_, P_), Ct), R_)));
  Vec Cx = mulv(C_, x_hat_new_);
  Vec innov(y.size());
  for (size_t i = 0; i < y.size(); ++i) innov[i] = y[i] - Cx[i];
  Vec Kv = mulv(K_, innov);
  for (size_t i = 0; i < x_hat_new_.size(); ++i) x_hat_new_[i] += Kv[i];
  P_ = mul(sub(I_, mul(K_, C_)), P_);
}
void LinearKalmanFilter::updateA(const Mat& A) { A_ = A; KalmanFilterInterface::update(); }                   // :97-101
void LinearKalmanFilter::updateA(const Vec& y, const Mat& A) { A_ = A; KalmanFilterInterface::update(y); }    // :103-107

ExtendedKalmanFilter::ExtendedKalmanFilter(fn_t f, fn_t h, const Mat& A, const Mat& C, const Mat& Q, const Mat& R, const Mat& P)
    : LinearKalmanFilter(A, C, Q, R, P) { f_ = f; h_ = h; }
void ExtendedKalmanFilter::predict() {   // src/kalman.cpp:129-133
  x_hat_new_ = f_(x_hat_);
  P_ = add(mul(mul(A_, P_), transpose(A_)), Q_);
}
void ExtendedKalmanFilter::estimate(const Vec& y) {   // src/kalman.cpp:135-140
  Mat Ct = transpose(C_);
  K_ = mul(mul(P_, Ct), inversePartialPivLU(add(mul(mul(C_, P_), Ct), R_)));
  Vec hx = h_(x_hat_new_);
  Vec innov(y.size());
  for (size_t i = 0; i < y.size(); ++i) innov[i] = y[i] - hx[i];
  Vec Kv = mulv(K_, innov);
  for (size_t i = 0; i < x_hat_new_.size(); ++i) x_hat_new_[i] += Kv[i];
  P_ = mul(sub(I_, mul(K_, C_)), P_);
}
void ExtendedKalmanFilter::updateF(fn_t f, const Mat& A) { f_ = f; LinearKalmanFilter::updateA(A); }                 // :142-146
void ExtendedKalmanFilter::updateF(const Vec& y, fn_t f, const Mat& A) { f_ = f; LinearKalmanFilter::updateA(y, A); }  // :148-152

// -------------------------------------------------------------------------------------
// target_interface.cpp
// -------------------------------------------------------------------------------------
TargetInterface::TargetInterface(unsigned id, const Mat& P0, double t0) {   // src/target_interface.cpp:18-41
  assert(t0 >= 0);
  id_ = static_cast<int>(id);
  n_meas_ = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rot_.m[i][j] = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i < 6; ++i) pose_internal_[i] = 0.0;
  for (int i = 0; i < 7; ++i) measured_pose_[i] = (i == 6) ? 1.0 : 0.0;
  for (int i = 0; i < 6; ++i) { twist_[i] = 0.0; acceleration_[i] = 0.0; }
  t_ = t0;
  P_ = P0;
}
double TargetInterface::getPeriodEstimate() {   // src/target_interface.cpp:80-87
  double n = std::sqrt(twist_[3] * twist_[3] + twist_[4] * twist_[4] + twist_[5] * twist_[5]);
  if (n > 0) return 2 * M_PI / n;
  return -1.0;
}
void TargetInterface::getEstimatedPose(double out[7]) {   // :100-104 + geometry.hpp:590-594
  out[0] = trans_[0]; out[1] = trans_[1]; out[2] = trans_[2];
  Quat q = rotationMatrixToQuat(rot_);
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}
void TargetInterface::getEstimatedPoseAt(double, double out[7]) { getEstimatedPose(out); }              // :123-128
void TargetInterface::getEstimatedTwistAt(double, double out[6]) { getEstimatedTwist(out); }            // :130-134
void TargetInterface::getEstimatedAccelerationAt(double, double out[6]) { getEstimatedAcceleration(out); }  // :136-140
void TargetInterface::updateMeasurement(const double meas[7]) {   // :142-146
  std::memcpy(measured_pose_, meas, 7 * sizeof(double));
  n_meas_ += 1;
}
void TargetInterface::updateTime(double dt) {   // :148-152
  assert(dt >= 0.0);
  t_ = t_ + dt;
}

static const double kZero6[6] = {0, 0, 0, 0, 0, 0};

static void isometryToPose6d(const double trans[3], const Mat3& R, double p[6]) {   // geometry.hpp:602-608
  p[0] = trans[0]; p[1] = trans[1]; p[2] = trans[2];
  rotToRpy(R, p + 3);
}

// ---- uniform velocity (src/types/uniform_velocity.cpp) ----
TargetUniformVelocity::TargetUniformVelocity(unsigned id, double dt0, double t0, const Mat& Q, const Mat& R, const Mat& P0,
                                             const double p0[7], const double v0[6], const double*)
    : TargetInterface(id, P0, t0) {   // :16-61
  n_ = (unsigned)Q.rows();
  m_ = (unsigned)R.rows();
  assert(n_ == 6);
  assert(m_ <= n_);
  assert(dt0 >= 0.0);
  A_ = Mat(n_, n_);
  updateA(dt0);
  C_ = Mat(m_, n_);
  for (unsigned i = 0; i < m_; ++i) C_(i, i) = 1.0;
  estimator_.reset(new LinearKalmanFilter(A_, C_, Q, R, P0));
  x_.assign(n_, 0.0);
  for (int i = 0; i < 3; ++i) { x_[i] = p0[i]; x_[3 + i] = v0[i]; }
  estimator_->init(x_);
  updateTargetState();
}
void TargetUniformVelocity::addMeasurement(double dt, const double meas[7]) {   // :63-76
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  updateMeasurement(meas);
  Vec y = {measured_pose_[0], measured_pose_[1], measured_pose_[2]};
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(y, A_);
  updateTargetState();
  updateTime(dt);
}
void TargetUniformVelocity::update(double dt) {   // :78-88
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(A_);
  updateTargetState();
  updateTime(dt);
}
void TargetUniformVelocity::updateA(double dt) {   // :90-96
  for (unsigned i = 0; i < n_; ++i) A_(i, i) = 1.0;
  for (unsigned i = 0; i < n_ / 2; ++i) A_(i, i + n_ / 2) = 1.0 * dt;
}
void TargetUniformVelocity::updateTargetState() {   // :98-115
  x_ = estimator_->getState();
  P_ = estimator_->getP();
  for (int i = 0; i < 3; ++i) trans_[i] = x_[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rot_.m[i][j] = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i < 3; ++i) { twist_[i] = x_[3 + i]; twist_[3 + i] = 0.0; }
  for (int i = 0; i < 6; ++i) acceleration_[i] = 0.0;
  isometryToPose6d(trans_, rot_, pose_internal_);
}
void TargetUniformVelocity::getEstimatedPoseAt(double t1, double out[7]) {   // :117-127
  for (int i = 0; i < 3; ++i) out[i] = trans_[i] + twist_[i] * (t1 - t_);
  out[3] = 0; out[4] = 0; out[5] = 0; out[6] = 1;
}
void TargetUniformVelocity::getEstimatedTwistAt(double, double out[6]) { getEstimatedTwist(out); }   // :129-133

// ---- uniform acceleration (src/types/uniform_acceleration.cpp) ----
TargetUniformAcceleration::TargetUniformAcceleration(unsigned id, double dt0, double t0, const Mat& Q, const Mat& R, const Mat& P0,
                                                     const double p0[7], const double v0[6], const double a0[6])
    : TargetInterface(id, P0, t0) {   // :17-62
  n_ = (unsigned)Q.rows();
  m_ = (unsigned)R.rows();
  assert(n_ == 9);
  assert(m_ <= n_);
  assert(dt0 >= 0.0);
  A_ = Mat(n_, n_);
  updateA(dt0);
  C_ = Mat(m_, n_);
  for (unsigned i = 0; i < m_; ++i) C_(i, i) = 1.0;
  estimator_.reset(new LinearKalmanFilter(A_, C_, Q, R, P0));
  x_.assign(n_, 0.0);
  for (int i = 0; i < 3; ++i) { x_[i] = p0[i]; x_[3 + i] = v0[i]; x_[6 + i] = a0[i]; }
  estimator_->init(x_);
  updateTargetState();
}
void TargetUniformAcceleration::addMeasurement(double dt, const double meas[7]) {   // :64-77
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  updateMeasurement(meas);
  Vec y = {measured_pose_[0], measured_pose_[1], measured_pose_[2]};
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(y, A_);
  updateTargetState();
  updateTime(dt);
}
void TargetUniformAcceleration::update(double dt) {   // :79-89
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(A_);
  updateTargetState();
  updateTime(dt);
}
void TargetUniformAcceleration::updateA(double dt) {   // :91-99
  std::fill(A_.d.begin(), A_.d.end(), 0.0);
  for (unsigned i = 0; i < n_; ++i) A_(i, i) = 1.0;
  for (unsigned i = 0; i < (n_ * 2) / 3; ++i) A_(i, i + n_ / 3) = 1.0 * dt;
  for (unsigned i = 0; i < n_ / 3; ++i) A_(i, i + (n_ * 2) / 3) = 1.0 * 0.5 * dt * dt;
}
void TargetUniformAcceleration::updateTargetState() {   // :101-118
  x_ = estimator_->getState();
  P_ = estimator_->getP();
  for (int i = 0; i < 3; ++i) trans_[i] = x_[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rot_.m[i][j] = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i < 3; ++i) { twist_[i] = x_[3 + i]; twist_[3 + i] = 0.0; }
  for (int i = 0; i < 3; ++i) { acceleration_[i] = x_[6 + i]; acceleration_[3 + i] = 0.0; }
  isometryToPose6d(trans_, rot_, pose_internal_);
}
void TargetUniformAcceleration::getEstimatedPoseAt(double t1, double out[7]) {   // :120-130
  for (int i = 0; i < 3; ++i)
    out[i] = trans_[i] + twist_[i] * (t1 - t_) + 0.5 * acceleration_[i] * (t1 - t_) * (t1 - t_);
  out[3] = 0; out[4] = 0; out[5] = 0; out[6] = 1;
}
void TargetUniformAcceleration::getEstimatedTwistAt(double t1, double out[6]) {   // :132-136
  for (int i = 0; i < 6; ++i) out[i] = twist_[i] + acceleration_[i] * (t1 - t_);
}

// ---- angular rates (src/types/angular_rates.cpp) ----
TargetAngularRates::TargetAngularRates(unsigned id, double dt0, double t0, const Mat& Q, const Mat& R, const Mat& P0,
                                       const double p0[7], const double v0[6], const double a0[6])
    : TargetInterface(id, P0, t0) {   // :21-70
  n_ = (unsigned)Q.rows();
  m_ = (unsigned)R.rows();
  assert(n_ == 18);
  assert(m_ <= n_);
  assert(dt0 >= 0.0);
  A_ = Mat(n_, n_);
  updateA(dt0);
  C_ = Mat(m_, n_);
  for (unsigned i = 0; i < m_; ++i) C_(i, i) = 1.0;
  estimator_.reset(new LinearKalmanFilter(A_, C_, Q, R, P0));
  x_.assign(n_, 0.0);
  pose7dToPose6d(p0, pose_internal_);
  for (int i = 0; i < 6; ++i) { x_[i] = pose_internal_[i]; x_[6 + i] = v0[i]; x_[12 + i] = a0[i]; }
  estimator_->init(x_);
  updateTargetState();
}
void TargetAngularRates::addMeasurement(double dt, const double meas[7]) {   // :72-94
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  updateMeasurement(meas);
  Vec y(6);
  y[0] = measured_pose_[0]; y[1] = measured_pose_[1]; y[2] = measured_pose_[2];
  Quat q; q.x = measured_pose_[3]; q.y = measured_pose_[4]; q.z = measured_pose_[5]; q.w = measured_pose_[6];
  quatNormalize(q);
  double rpy[3], un[3];
  quatToRpy(q, rpy);
  unwrap3(meas_rpy_internal_, rpy, un);
  y[3] = un[0]; y[4] = un[1]; y[5] = un[2];
  meas_rpy_internal_[0] = un[0]; meas_rpy_internal_[1] = un[1]; meas_rpy_internal_[2] = un[2];
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(y, A_);
  updateTargetState();
  updateTime(dt);
}
void TargetAngularRates::update(double dt) {   // :96-106
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt);
  static_cast<LinearKalmanFilter*>(estimator_.get())->updateA(A_);
  updateTargetState();
  updateTime(dt);
}
void TargetAngularRates::updateA(double dt) {   // :108-115
  for (unsigned i = 0; i < n_; ++i) A_(i, i) = 1.0;
  for (unsigned i = 0; i < (n_ * 2) / 3; ++i) A_(i, i + n_ / 3) = 1.0 * dt;
  for (unsigned i = 0; i < n_ / 3; ++i) A_(i, i + (n_ * 2) / 3) = 1.0 * 0.5 * dt * dt;
}
void TargetAngularRates::updateTargetState() {   // :117-138
  x_ = estimator_->getState();
  P_ = estimator_->getP();
  for (int i = 0; i < 3; ++i) trans_[i] = x_[i];
  double rpy[3] = {x_[3], x_[4], x_[5]};
  Quat q;
  rpyToQuat(rpy, q);
  rot_ = quatToRotationMatrix(q);
  for (int i = 0; i < 3; ++i) twist_[i] = x_[6 + i];
  rotToRpy(rot_, rpy);
  Mat3 Ear;
  rpyToEarBase(rpy, Ear);
  for (int i = 0; i < 3; ++i)
    twist_[3 + i] = Ear.m[i][0] * x_[9] + Ear.m[i][1] * x_[10] + Ear.m[i][2] * x_[11];
  for (int i = 0; i < 6; ++i) acceleration_[i] = x_[12 + i];
  isometryToPose6d(trans_, rot_, pose_internal_);
}
void TargetAngularRates::getEstimatedPoseAt(double t1, double out[7]) {   // :140-151
  double v6[6];
  for (int i = 0; i < 6; ++i)
    v6[i] = pose_internal_[i] + twist_[i] * (t1 - t_) + 0.5 * acceleration_[i] * (t1 - t_) * (t1 - t_);
  Quat q;
  rpyToQuat(v6 + 3, q);
  quatNormalize(q);
  out[0] = v6[0]; out[1] = v6[1]; out[2] = v6[2];
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}
void TargetAngularRates::getEstimatedTwistAt(double t1, double out[6]) {   // :153-157
  for (int i = 0; i < 6; ++i) out[i] = twist_[i] + acceleration_[i] * (t1 - t_);
}

// ---- angular velocities, EKF (src/types/angular_velocities.cpp) ----
TargetAngularVelocities::TargetAngularVelocities(unsigned id, double dt0, double t0, const Mat& Q, const Mat& R, const Mat& P0,
                                                 const double p0[7], const double v0[6], const double*)
    : TargetInterface(id, P0, t0) {   // :21-78
  n_ = (unsigned)Q.rows();
  m_ = (unsigned)R.rows();
  assert(n_ == 12);
  assert(m_ <= n_);
  assert(dt0 >= 0.0);
  x_.assign(n_, 0.0);
  pose7dToPose6d(p0, pose_internal_);
  for (int i = 0; i < 6; ++i) { x_[i] = pose_internal_[i]; x_[6 + i] = v0[i]; }
  A_ = Mat(n_, n_);
  updateA(dt0, &x_[3], &x_[9]);
  C_ = Mat(m_, n_);
  for (unsigned i = 0; i < m_; ++i) C_(i, i) = 1.0;
  estimator_.reset(new ExtendedKalmanFilter(
      [this, dt0](const Vec& x) { return this->f(x, dt0); }, [this](const Vec& x) { return this->h(x); }, A_, C_, Q, R, P0));
  estimator_->init(x_);
  updateTargetState();
}
void TargetAngularVelocities::addMeasurement(double dt, const double meas[7]) {   // :80-102
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt, &x_[3], &x_[9]);
  updateMeasurement(meas);
  Vec y(6);
  y[0] = measured_pose_[0]; y[1] = measured_pose_[1]; y[2] = measured_pose_[2];
  Quat q; q.x = measured_pose_[3]; q.y = measured_pose_[4]; q.z = measured_pose_[5]; q.w = measured_pose_[6];
  quatNormalize(q);
  double rpy[3], un[3];
  quatToRpy(q, rpy);
  unwrap3(meas_rpy_internal_, rpy, un);
  y[3] = un[0]; y[4] = un[1]; y[5] = un[2];
  meas_rpy_internal_[0] = un[0]; meas_rpy_internal_[1] = un[1]; meas_rpy_internal_[2] = un[2];
  static_cast<ExtendedKalmanFilter*>(estimator_.get())->updateF(y, [this, dt](const Vec& x) { return this->f(x, dt); }, A_);
  updateTargetState();
  updateTime(dt);
}
void TargetAngularVelocities::update(double dt) {   // :104-114
  std::lock_guard<std::mutex> lg(data_lock_);
  updateA(dt, &x_[3], &x_[9]);
  static_cast<ExtendedKalmanFilter*>(estimator_.get())->updateF([this, dt](const Vec& x) { return this->f(x, dt); }, A_);
  updateTargetState();
  updateTime(dt);
}
void TargetAngularVelocities::updateA(double dt, const double rpy[3], const double omega[3]) {   // :116-124
  Mat3 Jr = EarBaseInvJacobianRpy(rpy, omega, dt);
  Mat3 Jo = EarBaseInvJacobianOmega(rpy, dt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double I = (i == j) ? 1.0 : 0.0;
      A_(i, j) = I;
      A_(i, 6 + j) = I * dt;
      A_(3 + i, 3 + j) = Jr.m[i][j];
      A_(3 + i, 9 + j) = Jo.m[i][j];
      A_(6 + i, 6 + j) = I;
      A_(9 + i, 9 + j) = I;
    }
}
Vec TargetAngularVelocities::f(const Vec& x, double dt) {   // :126-140
  assert(x.size() == n_);
  Vec o(n_, 0.0);
  Mat3 E;
  rpyToEarBaseInv(&x[3], E);
  for (int i = 0; i < 3; ++i) {
    o[i] = x[i] + dt * x[6 + i];
    o[6 + i] = x[6 + i];
    o[9 + i] = x[9 + i];
  }
  for (int i = 0; i < 3; ++i) {
    double s = (dt * E.m[i][0]) * x[9] + (dt * E.m[i][1]) * x[10] + (dt * E.m[i][2]) * x[11];
    o[3 + i] = x[3 + i] + s;
  }
  return o;
}
Vec TargetAngularVelocities::h(const Vec& x) {   // :142-151
  assert(x.size() == n_);
  Vec y(6);
  for (int i = 0; i < 6; ++i) y[i] = x[i];
  return y;
}
void TargetAngularVelocities::updateTargetState() {   // :153-169
  x_ = estimator_->getState();
  P_ = estimator_->getP();
  for (int i = 0; i < 3; ++i) trans_[i] = x_[i];
  double rpy[3] = {x_[3], x_[4], x_[5]};
  Quat q;
  rpyToQuat(rpy, q);
  rot_ = quatToRotationMatrix(q);
  for (int i = 0; i < 3; ++i) { twist_[i] = x_[6 + i]; twist_[3 + i] = x_[9 + i]; }
  isometryToPose6d(trans_, rot_, pose_internal_);
}
void TargetAngularVelocities::getEstimatedPoseAt(double t1, double out[7]) {   // :171-184
  for (int i = 0; i < 3; ++i) out[i] = trans_[i] + twist_[i] * (t1 - t_);
  Quat q;
  rpyToQuat(pose_internal_ + 3, q);
  double Qm[4][4];
  Qtran(t1 - t_, twist_ + 3, Qm);
  double c[4] = {q.x, q.y, q.z, q.w}, r[4];
  for (int i = 0; i < 4; ++i) r[i] = Qm[i][0] * c[0] + Qm[i][1] * c[1] + Qm[i][2] * c[2] + Qm[i][3] * c[3];
  q.x = r[0]; q.y = r[1]; q.z = r[2]; q.w = r[3];
  quatNormalize(q);
  out[3] = q.x; out[4] = q.y; out[5] = q.z; out[6] = q.w;
}

// -------------------------------------------------------------------------------------
// target_manager.cpp
// -------------------------------------------------------------------------------------
bool selectTargetType(const std::string& s, target_t& type) {   // src/target_manager.cpp:52-65
  if (s == "angular_rates") type = ANGULAR_RATES;
  else if (s == "angular_velocities") type = ANGULAR_VELOCITIES;
  else if (s == "uniform_acceleration") type = UNIFORM_ACCELERATION;
  else if (s == "uniform_velocity") type = UNIFORM_VELOCITY;
  else return false;
  return true;
}

// yaml-cpp stand-in for the flat "key: value" / "key: [a, b, ...]" files of models/*.yaml.
// parseSquareMatrix (src/target_manager.cpp:18-33): size = sqrt(len), column-major Map.
static bool parseYamlLists(const std::string& file, std::map<std::string, std::string>& kv) {
  std::ifstream in(file.c_str());
  if (!in.is_open()) return false;
  std::string line;
  while (std::getline(in, line)) {
    size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string key = line.substr(0, c);
    std::string val = line.substr(c + 1);
    auto trim = [](std::string& s) {
      size_t a = s.find_first_not_of(" \t\r\n");
      size_t b = s.find_last_not_of(" \t\r\n");
      s = (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
    };
    trim(key); trim(val);
    kv[key] = val;
  }
  return true;
}
static bool parseSquareMatrix(const std::map<std::string, std::string>& kv, const std::string& name, Mat& M) {
  auto it = kv.find(name);
  if (it == kv.end()) return false;
  std::string v = it->second;
  size_t a = v.find('['), b = v.rfind(']');
  if (a == std::string::npos || b == std::string::npos) return false;
  v = v.substr(a + 1, b - a - 1);
  std::vector<double> Mv;
  const char* p = v.c_str();
  while (*p) {
    while (*p == ' ' || *p == ',' || *p == '\t') ++p;
    if (!*p) break;
    char* e = nullptr;
    double x = std::strtod(p, &e);
    if (e == p) return false;
    Mv.push_back(x);
    p = e;
  }
  unsigned size = static_cast<unsigned>(std::sqrt((double)Mv.size()));
  if ((size_t)size * size > Mv.size()) return false;
  M = Mat::MapColMajor(Mv.data(), (int)size);
  return true;
}
bool loadYamlFile(const std::string& file, Mat& Q, Mat& R, Mat& P, target_t& type, double* frequency) {   // :67-104
  std::map<std::string, std::string> kv;
  if (!parseYamlLists(file, kv)) return false;
  bool ok = true;
  if (!parseSquareMatrix(kv, "Q", Q)) ok = false;
  if (!parseSquareMatrix(kv, "R", R)) ok = false;
  if (!parseSquareMatrix(kv, "P", P)) ok = false;
  auto it = kv.find("type");
  if (it == kv.end()) ok = false;
  else selectTargetType(it->second, type);   // the reference only warns on unknown type (:41-42)
  if (frequency) {
    auto f = kv.find("frequency");
    *frequency = (f == kv.end()) ? 0.0 : std::strtod(f->second.c_str(), nullptr);
  }
  return ok;
}

TargetManager::TargetManager(const std::string& file) {   // :112-118
  if (!loadYamlFile(file, default_Q_, default_R_, default_P_, default_type_))
    throw "TargetManager default constructor failed!";
  default_values_loaded_ = true;
}
std::vector<unsigned> TargetManager::getAvailableTargets() {   // :126-133
  std::vector<unsigned> ids;
  std::lock_guard<std::mutex> lg(target_lock_);
  for (auto const& kv : targets_) ids.push_back(kv.first);
  return ids;
}
void TargetManager::init(unsigned id, double dt0, double t0, const double p0[7], const double v0[6], const double a0[6]) {  // :135-142
  if (default_values_loaded_) init(default_type_, id, dt0, t0, default_Q_, default_R_, default_P_, p0, v0, a0);
  else throw "TargetManager::init failed, can not find default values to load!";
}
void TargetManager::init(target_t type, unsigned id, double dt0, double t0, const Mat& Q, const Mat& R, const Mat& P0,
                         const double p0[7], const double v0in[6], const double a0in[6]) {   // :144-179
  const double* v0 = v0in ? v0in : kZero6;
  const double* a0 = a0in ? a0in : kZero6;
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.find(id) == targets_.end()) {
    switch (type) {
      case ANGULAR_RATES: targets_[id].reset(new TargetAngularRates(id, dt0, t0, Q, R, P0, p0, v0, a0)); break;
      case ANGULAR_VELOCITIES: targets_[id].reset(new TargetAngularVelocities(id, dt0, t0, Q, R, P0, p0, v0, a0)); break;
      case UNIFORM_ACCELERATION: targets_[id].reset(new TargetUniformAcceleration(id, dt0, t0, Q, R, P0, p0, v0, a0)); break;
      case UNIFORM_VELOCITY: targets_[id].reset(new TargetUniformVelocity(id, dt0, t0, Q, R, P0, p0, v0, a0)); break;
    }
  } else if (!quiet) {
    std::cout << "Target(" << id << ") already exists!" << std::endl;
  }
}
void TargetManager::init(const std::string& file, unsigned id, double dt0, double t0, const double p0[7],
                         const double v0[6], const double a0[6]) {   // :181-188
  Mat Q, P, R;
  target_t type = UNIFORM_VELOCITY;
  loadYamlFile(file, Q, R, P, type);
  init(type, id, dt0, t0, Q, R, P, p0, v0, a0);
}
bool TargetManager::update(unsigned id, double dt, const double meas[7]) {   // :190-202
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.find(id) == targets_.end()) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  targets_[id]->addMeasurement(dt, meas);
  return true;
}
bool TargetManager::update(unsigned id, double dt) {   // :204-218
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.find(id) == targets_.end()) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  targets_[id]->update(dt);
  return true;
}
void TargetManager::update(double dt) {   // :220-225
  std::lock_guard<std::mutex> lg(target_lock_);
  for (const auto& kv : targets_) kv.second->update(dt);
}
bool TargetManager::erase(unsigned id) {   // :227-241
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.find(id) == targets_.end()) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  targets_.erase(id);
  return true;
}
TargetInterface::Ptr TargetManager::getTarget(unsigned id) {   // :243-250
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.count(id) != 0) return targets_[id];
  return nullptr;
}
bool TargetManager::getTargetPose(unsigned id, double pose[7]) {   // :252-261
  if (getTarget(id) != nullptr) { getTarget(id)->getEstimatedPose(pose); return true; }
  return false;
}
bool TargetManager::getTargetTwist(unsigned id, double twist[6]) {   // :263-272
  if (getTarget(id) != nullptr) { getTarget(id)->getEstimatedTwist(twist); return true; }
  return false;
}
bool TargetManager::getTargetAcceleration(unsigned id, double acc[6]) {   // :274-283
  if (getTarget(id) != nullptr) { getTarget(id)->getEstimatedAcceleration(acc); return true; }
  return false;
}
long long TargetManager::getNumberMeasurements(unsigned id) {   // :285-295
  std::lock_guard<std::mutex> lg(target_lock_);
  if (targets_.count(id) != 0) return targets_[id]->getNumberMeasurements();
  if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
  return 0;
}

// -------------------------------------------------------------------------------------
// utils.hpp
// -------------------------------------------------------------------------------------
double MovingAvgFilter::update(double value) {   // utils.hpp:222-251
  unsigned n = (unsigned)window_.size();
  double res = 0.0;
  sum_ -= window_[window_idx_];
  sum_ += value;
  window_[window_idx_] = value;
  if (!filter_complete_ && window_idx_ == n - 1) filter_complete_ = true;
  unsigned num = n;
  if (!filter_complete_) num = window_idx_ + 1;
  res = sum_ / num;
  window_idx_ = (window_idx_ + 1) % n;
  double variance_sum = 0.0;
  for (double val : window_) variance_sum += std::pow((val - res), 2);
  variance_ = variance_sum / num;
  return res;
}

std::vector<std::string> splitString(const std::string& s, const std::string& delimiter) {   // utils.hpp:273-294
  size_t pos_start = 0, pos_end, delim_len = delimiter.length();
  std::vector<std::string> res;
  pos_end = s.find(delimiter, pos_start);
  while (pos_end != std::string::npos) {
    res.push_back(s.substr(pos_start, pos_end - pos_start));
    pos_start = pos_end + delim_len;
    pos_end = s.find(delimiter, pos_start);
  }
  res.push_back(s.substr(pos_start));
  return res;
}
bool getId(const std::string& s, unsigned& id) {   // utils.hpp:302-313
  auto strings = splitString(s);
  if (strings.size() == 2) {
    id = std::stoi(strings[1]);
    return true;
  }
  return false;
}

// -------------------------------------------------------------------------------------
// (Eigen) unsupported/Polynomials: PolynomialSolver<double,Dynamic>::compute restated.
// companion matrix -> balance() -> EigenSolver (Hessenberg is a no-op on a companion
// matrix; real Schur by Francis double-shift QR as in Eigen's RealSchur.h) -> roots;
// Eigen 3.4 imaginary-noise cleanup.  SURVEY.md Appendix B.
// -------------------------------------------------------------------------------------
namespace {

bool companionBalanced(double colNorm, double rowNorm, bool& isBalanced, double& colB, double& rowB) {
  if (0.0 == colNorm || 0.0 == rowNorm || !std::isfinite(colNorm) || !std::isfinite(rowNorm)) return true;
  const double radix = 2.0, radix2 = 4.0;
  rowB = rowNorm / radix;
  colB = 1.0;
  const double s = colNorm + rowNorm;
  double scout = colNorm;
  while (scout < rowB) { colB *= radix; scout *= radix2; }
  scout = colNorm * (colB / radix) * colB;
  while (scout >= rowNorm) { colB /= radix; scout /= radix2; }
  if ((rowNorm + radix * scout) < 0.95 * s * colB) {
    isBalanced = false;
    rowB = 1.0 / colB;
    return false;
  }
  return true;
}

struct RealSchurT {
  int n;
  Mat T;
  explicit RealSchurT(const Mat& H) : n(H.r), T(H) {}

  static void makeHouseholder(const double* v, int size, double* ess, double& tau, double& beta) {
    double tailSqNorm = 0.0;
    for (int i = 1; i < size; ++i) tailSqNorm += v[i] * v[i];
    double c0 = v[0];
    const double tol = std::numeric_limits<double>::min();
    if (tailSqNorm <= tol) {
      tau = 0.0;
      beta = c0;
      for (int i = 1; i < size; ++i) ess[i - 1] = 0.0;
    } else {
      beta = std::sqrt(c0 * c0 + tailSqNorm);
      if (c0 >= 0.0) beta = -beta;
      for (int i = 1; i < size; ++i) ess[i - 1] = v[i] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
  }
  // block(r0,c0,nr,nc).applyHouseholderOnTheLeft(ess,tau)
  void houseLeft(int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nr == 1) { for (int j = 0; j < nc; ++j) T(r0, c0 + j) *= (1.0 - tau); return; }
    if (tau == 0.0) return;
    for (int j = 0; j < nc; ++j) {
      double tmp = 0.0;
      for (int i = 1; i < nr; ++i) tmp += ess[i - 1] * T(r0 + i, c0 + j);
      tmp += T(r0, c0 + j);
      T(r0, c0 + j) -= tau * tmp;
      for (int i = 1; i < nr; ++i) T(r0 + i, c0 + j) -= (tau * ess[i - 1]) * tmp;
    }
  }
  void houseRight(int r0, int c0, int nr, int nc, const double* ess, double tau) {
    if (nc == 1) { for (int i = 0; i < nr; ++i) T(r0 + i, c0) *= (1.0 - tau); return; }
    if (tau == 0.0) return;
    for (int i = 0; i < nr; ++i) {
      double tmp = 0.0;
      for (int j = 1; j < nc; ++j) tmp += T(r0 + i, c0 + j) * ess[j - 1];
      tmp += T(r0 + i, c0);
      T(r0 + i, c0) -= tau * tmp;
      for (int j = 1; j < nc; ++j) T(r0 + i, c0 + j) -= (tau * tmp) * ess[j - 1];
    }
  }
  double computeNormOfT() const {
    double norm = 0.0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < std::min(n, j + 2); ++i) norm += std::fabs(T(i, j));
    return norm;
  }
  int findSmallSubdiagEntry(int iu, double considerAsZero) const {
    int res = iu;
    while (res > 0) {
      double s = std::fabs(T(res - 1, res - 1)) + std::fabs(T(res, res));
      s = std::max(s * std::numeric_limits<double>::epsilon(), considerAsZero);
      if (std::fabs(T(res, res - 1)) <= s) break;
      res--;
    }
    return res;
  }
  void splitOffTwoRows(int iu, double exshift) {
    double p = 0.5 * (T(iu - 1, iu - 1) - T(iu, iu));
    double q = p * p + T(iu, iu - 1) * T(iu - 1, iu);
    T(iu, iu) += exshift;
    T(iu - 1, iu - 1) += exshift;
    if (q >= 0.0) {
      double z = std::sqrt(std::fabs(q));
      double gp = (p >= 0.0) ? (p + z) : (p - z);
      double gq = T(iu, iu - 1);
      double c, s;   // JacobiRotation::makeGivens(gp, gq)
      if (gq == 0.0) { c = gp < 0.0 ? -1.0 : 1.0; s = 0.0; }
      else if (gp == 0.0) { c = 0.0; s = gq < 0.0 ? 1.0 : -1.0; }
      else if (std::fabs(gp) > std::fabs(gq)) {
        double t = gq / gp;
        double u = std::sqrt(1.0 + t * t);
        if (gp < 0.0) u = -u;
        c = 1.0 / u;
        s = -t * c;
      } else {
        double t = gp / gq;
        double u = std::sqrt(1.0 + t * t);
        if (gq < 0.0) u = -u;
        s = -1.0 / u;
        c = -t * s;
      }
      // rightCols(size-iu+1).applyOnTheLeft(iu-1, iu, rot.adjoint())
      for (int j = iu - 1; j < n; ++j) {
        double xi = T(iu - 1, j), yi = T(iu, j);
        T(iu - 1, j) = c * xi - s * yi;
        T(iu, j) = s * xi + c * yi;
      }
      // topRows(iu+1).applyOnTheRight(iu-1, iu, rot)
      for (int i = 0; i <= iu; ++i) {
        double xi = T(i, iu - 1), yi = T(i, iu);
        T(i, iu - 1) = c * xi - s * yi;
        T(i, iu) = s * xi + c * yi;
      }
      T(iu, iu - 1) = 0.0;
    }
    if (iu > 1) T(iu - 1, iu - 2) = 0.0;
  }
  void computeShift(int iu, int iter, double& exshift, double shiftInfo[3]) {
    shiftInfo[0] = T(iu, iu);
    shiftInfo[1] = T(iu - 1, iu - 1);
    shiftInfo[2] = T(iu, iu - 1) * T(iu - 1, iu);
    if (iter == 10) {
      exshift += shiftInfo[0];
      for (int i = 0; i <= iu; ++i) T(i, i) -= shiftInfo[0];
      double s = std::fabs(T(iu, iu - 1)) + std::fabs(T(iu - 1, iu - 2));
      shiftInfo[0] = 0.75 * s;
      shiftInfo[1] = 0.75 * s;
      shiftInfo[2] = -0.4375 * s * s;
    }
    if (iter == 30) {
      double s = (shiftInfo[1] - shiftInfo[0]) / 2.0;
      s = s * s + shiftInfo[2];
      if (s > 0.0) {
        s = std::sqrt(s);
        if (shiftInfo[1] < shiftInfo[0]) s = -s;
        s = s + (shiftInfo[1] - shiftInfo[0]) / 2.0;
        s = shiftInfo[0] - shiftInfo[2] / s;
        exshift += s;
        for (int i = 0; i <= iu; ++i) T(i, i) -= s;
        shiftInfo[0] = shiftInfo[1] = shiftInfo[2] = 0.964;
      }
    }
  }
  void initFrancisQRStep(int il, int iu, const double shiftInfo[3], int& im, double v[3]) const {
    for (im = iu - 2; im >= il; --im) {
      const double Tmm = T(im, im);
      const double r = shiftInfo[0] - Tmm;
      const double s = shiftInfo[1] - Tmm;
      v[0] = (r * s - shiftInfo[2]) / T(im + 1, im) + T(im, im + 1);
      v[1] = T(im + 1, im + 1) - Tmm - r - s;
      v[2] = T(im + 2, im + 1);
      if (im == il) break;
      const double lhs = T(im, im - 1) * (std::fabs(v[1]) + std::fabs(v[2]));
      const double rhs = v[0] * (std::fabs(T(im - 1, im - 1)) + std::fabs(Tmm) + std::fabs(T(im + 1, im + 1)));
      if (std::fabs(lhs) < std::numeric_limits<double>::epsilon() * rhs) break;
    }
  }
  void performFrancisQRStep(int il, int im, int iu, const double firstV[3]) {
    for (int k = im; k <= iu - 2; ++k) {
      bool first = (k == im);
      double v[3];
      if (first) { v[0] = firstV[0]; v[1] = firstV[1]; v[2] = firstV[2]; }
      else { v[0] = T(k, k - 1); v[1] = T(k + 1, k - 1); v[2] = T(k + 2, k - 1); }
      double tau, beta, ess[2];
      makeHouseholder(v, 3, ess, tau, beta);
      if (beta != 0.0) {
        if (first && k > il) T(k, k - 1) = -T(k, k - 1);
        else if (!first) T(k, k - 1) = beta;
        houseLeft(k, k, 3, n - k, ess, tau);
        houseRight(0, k, std::min(iu, k + 3) + 1, 3, ess, tau);
      }
    }
    double v2[2] = {T(iu - 1, iu - 2), T(iu, iu - 2)};
    double tau, beta, ess[1];
    makeHouseholder(v2, 2, ess, tau, beta);
    if (beta != 0.0) {
      T(iu - 1, iu - 2) = beta;
      houseLeft(iu - 1, iu - 1, 2, n - iu + 1, ess, tau);
      houseRight(0, iu - 1, iu + 1, 2, ess, tau);
    }
    for (int i = im + 2; i <= iu; ++i) {
      T(i, i - 2) = 0.0;
      if (i > im + 2) T(i, i - 3) = 0.0;
    }
  }
  bool compute() {
    int maxIters = 40 * n;
    int iu = n - 1, iter = 0, totalIter = 0;
    double exshift = 0.0;
    double norm = computeNormOfT();
    double eps = std::numeric_limits<double>::epsilon();
    double considerAsZero = std::max(norm * eps * eps, std::numeric_limits<double>::min());
    if (norm != 0.0) {
      while (iu >= 0) {
        int il = findSmallSubdiagEntry(iu, considerAsZero);
        if (il == iu) {
          T(iu, iu) = T(iu, iu) + exshift;
          if (iu > 0) T(iu, iu - 1) = 0.0;
          iu--;
          iter = 0;
        } else if (il == iu - 1) {
          splitOffTwoRows(iu, exshift);
          iu -= 2;
          iter = 0;
        } else {
          double firstV[3] = {0, 0, 0}, shiftInfo[3];
          computeShift(iu, iter, exshift, shiftInfo);
          iter = iter + 1;
          totalIter = totalIter + 1;
          if (totalIter > maxIters) break;
          int im;
          initFrancisQRStep(il, iu, shiftInfo, im, firstV);
          performFrancisQRStep(il, im, iu, firstV);
        }
      }
    }
    return totalIter <= maxIters;
  }
  std::vector<std::complex<double>> eigenvalues() const {   // (Eigen) EigenSolver::compute, values only
    std::vector<std::complex<double>> ev(n);
    int i = 0;
    while (i < n) {
      if (i == n - 1 || T(i + 1, i) == 0.0) {
        ev[i] = T(i, i);
        ++i;
      } else {
        double p = 0.5 * (T(i, i) - T(i + 1, i + 1));
        double z;
        {
          double t0 = T(i + 1, i), t1 = T(i, i + 1);
          double maxval = std::max(std::fabs(p), std::max(std::fabs(t0), std::fabs(t1)));
          t0 /= maxval;
          t1 /= maxval;
          double p0 = p / maxval;
          z = maxval * std::sqrt(std::fabs(p0 * p0 + t0 * t1));
        }
        ev[i] = std::complex<double>(T(i + 1, i + 1) + p, z);
        ev[i + 1] = std::complex<double>(T(i + 1, i + 1) + p, -z);
        i += 2;
      }
    }
    return ev;
  }
};

std::complex<double> polyEval(const std::vector<double>& poly, const std::complex<double>& x) {   // (Eigen) poly_eval
  if (std::norm(x) <= 1.0) {
    std::complex<double> val = poly.back();
    for (int i = (int)poly.size() - 2; i >= 0; --i) val = val * x + poly[i];
    return val;
  }
  std::complex<double> val = poly[0];
  std::complex<double> inv_x = std::complex<double>(1.0) / x;
  for (size_t i = 1; i < poly.size(); ++i) val = val * inv_x + poly[i];
  return std::pow(x, (double)(poly.size() - 1)) * val;
}

}  // namespace

std::vector<std::complex<double>> polynomialRoots(const std::vector<double>& poly) {
  const int deg = (int)poly.size() - 1;
  assert(deg >= 1 && poly[deg] != 0.0);
  std::vector<std::complex<double>> roots;
  if (deg == 1) {
    roots.push_back(-poly[0] / poly[1]);
    return roots;
  }
  // companion<>::setPolynomial
  std::vector<double> monic(deg), bl_diag(deg - 1, 1.0);
  for (int i = 0; i < deg; ++i) monic[i] = -poly[i] / poly[deg];
  // companion<>::balance()
  {
    const int deg_1 = deg - 1;
    bool hasConverged = false;
    while (!hasConverged) {
      hasConverged = true;
      double colNorm, rowNorm, colB, rowB;
      colNorm = std::fabs(bl_diag[0]);
      rowNorm = std::fabs(monic[0]);
      if (!companionBalanced(colNorm, rowNorm, hasConverged, colB, rowB)) {
        bl_diag[0] *= colB;
        monic[0] *= rowB;
      }
      for (int i = 1; i < deg_1; ++i) {
        colNorm = std::fabs(bl_diag[i]);
        rowNorm = std::fabs(bl_diag[i - 1]) + std::fabs(monic[i]);
        if (!companionBalanced(colNorm, rowNorm, hasConverged, colB, rowB)) {
          bl_diag[i] *= colB;
          bl_diag[i - 1] *= rowB;
          monic[i] *= rowB;
        }
      }
      const int ebl = (int)bl_diag.size() - 1;
      colNorm = 0.0;
      for (int i = 0; i < deg_1; ++i) colNorm += std::fabs(monic[i]);
      rowNorm = std::fabs(bl_diag[ebl]);
      if (!companionBalanced(colNorm, rowNorm, hasConverged, colB, rowB)) {
        for (int i = 0; i < deg_1; ++i) monic[i] *= colB;
        bl_diag[ebl] *= rowB;
      }
    }
  }
  // companion<>::denseMatrix(): sub-diagonal = bl_diag, last column = monic
  Mat C(deg, deg);
  for (int i = 0; i < deg - 1; ++i) C(i + 1, i) = bl_diag[i];
  for (int i = 0; i < deg; ++i) C(i, deg - 1) = monic[i];
  RealSchurT rs(C);
  rs.compute();
  roots = rs.eigenvalues();
  // Eigen 3.4: cleanup noise in imaginary part of real roots
  const double coarse_prec = std::pow(4.0, (double)(poly.size() + 1)) * std::numeric_limits<double>::epsilon();
  for (size_t i = 0; i < roots.size(); ++i) {
    if (std::fabs(roots[i].imag()) <= std::fabs(roots[i].real()) * coarse_prec) {
      std::complex<double> as_real(roots[i].real(), 0.0);
      if (std::abs(polyEval(poly, as_real)) <= std::abs(polyEval(poly, roots[i]))) roots[i] = as_real;
    }
  }
  return roots;
}

double lowestRealRoot(const std::vector<double>& coeffs) {   // src/intersection_solver.cpp:4-17
  if (!(std::fabs(coeffs[coeffs.size() - 1]) > 0.0)) return -1;
  std::vector<std::complex<double>> roots = polynomialRoots(coeffs);
  bool reRootExists = false;
  const double imThreshold = 1e-10;
  double res = 0.0;
  // (Eigen) PolynomialSolverBase::smallestRealRoot: min real part among |imag| < thr
  for (size_t i = 0; i < roots.size(); ++i) {
    if (std::fabs(roots[i].imag()) < imThreshold) {
      if (!reRootExists) { reRootExists = true; res = roots[i].real(); }
      else if (roots[i].real() < res) res = roots[i].real();
    }
  }
  if (!reRootExists) return -1;
  return res;
}

IntersectionSolver::IntersectionSolver(TargetManager::Ptr tm, unsigned filters_length) {   // :19-40
  assert(tm);
  target_manager_ = tm;
  pos_error_filter_.reset(new MovingAvgFilter(filters_length));
  ang_error_filter_.reset(new MovingAvgFilter(filters_length));
  for (int i = 0; i < 7; ++i) intersection_pose_prev_[i] = (i == 6) ? 1.0 : 0.0;
}

double IntersectionSolver::getIntersectionTimeWithSphere(unsigned id, double t1, const double origin[3], double radius) {  // :42-89
  if (target_manager_->getTarget(id)) {
    double pose[7], tw[6], ac[6];
    target_manager_->getTarget(id)->getEstimatedPoseAt(t1, pose);
    target_manager_->getTarget(id)->getEstimatedTwistAt(t1, tw);
    target_manager_->getTarget(id)->getEstimatedAccelerationAt(t1, ac);
    double x = pose[0] - origin[0], y = pose[1] - origin[1], z = pose[2] - origin[2];
    double vx = tw[0], vy = tw[1], vz = tw[2];
    double ax = ac[0], ay = ac[1], az = ac[2];
    double R = radius;
    std::vector<double> coeff(5);
    coeff[4] = 0.25 * (ax * ax + ay * ay + az * az);
    coeff[3] = vx * ax + vy * ay + vz * az;
    coeff[2] = vx * vx + vy * vy + vz * vz + x * ax + y * ay + z * az;
    coeff[1] = 2 * (x * vx + y * vy + z * vz);
    coeff[0] = x * x + y * y + z * z - R * R;
    double delta = lowestRealRoot(coeff);
    if (delta < 0) return -1;
    return delta;
  }
  return -1;
}

bool IntersectionSolver::getIntersectionPoseWithSphere(unsigned id, double t1, double pos_th, double ang_th,
                                                       const double origin[3], double radius, double pose[7]) {  // :91-124
  assert(t1 >= 0.0);
  assert(pos_th >= 0.0);
  assert(ang_th >= 0.0);
  double delta = -1;
  bool converged = false;
  for (int i = 0; i < 7; ++i) pose[i] = (i == 6) ? 1.0 : 0.0;
  delta = getIntersectionTimeWithSphere(id, t1, origin, radius);
  if (delta > -1) {
    target_manager_->getTarget(id)->getEstimatedPoseAt(delta + t1, pose);
    double dx = pose[0] - intersection_pose_prev_[0], dy = pose[1] - intersection_pose_prev_[1],
           dz = pose[2] - intersection_pose_prev_[2];
    double pos_error = std::sqrt(dx * dx + dy * dy + dz * dz);
    Quat q1, q2;
    q1.x = pose[3]; q1.y = pose[4]; q1.z = pose[5]; q1.w = pose[6];
    q2.x = intersection_pose_prev_[3]; q2.y = intersection_pose_prev_[4];
    q2.z = intersection_pose_prev_[5]; q2.w = intersection_pose_prev_[6];
    quatNormalize(q1);
    quatNormalize(q2);
    double ang_error = std::fabs(wrapMinMax(computeQuaternionErrorAngle(q1, q2), -M_PI, M_PI));
    double pos_error_filt = pos_error_filter_->update(pos_error);
    double ang_error_filt = ang_error_filter_->update(ang_error);
    std::memcpy(intersection_pose_prev_, pose, sizeof(intersection_pose_prev_));
    if (pos_error_filt <= pos_th && ang_error_filt <= ang_th) converged = true;
  }
  return converged;
}

// -------------------------------------------------------------------------------------
// target_manager_ros.{hpp,cpp} tick semantics
// -------------------------------------------------------------------------------------
void Measurement::update(const StampedPose& tr) {   // target_manager_ros.hpp:96-115
  double current_time_stamp = toSec(tr.sec, tr.nsec);
  double prev_time_stamp = toSec(tr_.sec, tr_.nsec);
  if (current_time_stamp > prev_time_stamp) {
    new_meas_ = true;
    last_meas_time_ = current_time_stamp;
  } else {
    new_meas_ = false;
  }
  tr_ = tr;
}

TickTargetManager::TickTargetManager(target_t type, const Mat& Q, const Mat& R, const Mat& P)
    : type_(type), Q_(Q), P_(P), R_(R), token_name_("target"), t_(0.0), expiration_time_(1000.0) {}   // :6-24

void TickTargetManager::measurementCallBack(const std::vector<TfRecord>& msg) {   // :26-39
  for (size_t i = 0; i < msg.size(); i++) {
    const std::string& name = msg[i].child_frame_id;
    if (name.find(token_name_) != std::string::npos) {
      unsigned id;
      if (!getId(name, id)) break;
      measurements_[id].update(msg[i].tr);
    }
  }
}

void TickTargetManager::tick(double dt, uint32_t now_sec, uint32_t now_nsec, std::vector<unsigned>* erased) {   // :41-92
  const double now = toSec(now_sec, now_nsec);   // ros::Time::toSec()
  auto it = measurements_.begin();
  while (it != measurements_.end()) {
    unsigned id = it->first;
    double last_meas_time = it->second.getTime();
    StampedPose tmp;
    if (it->second.read(tmp)) {
      if (getTarget(id) == nullptr) init(type_, id, dt, t_, Q_, R_, P_, tmp.pose);
      TargetManager::update(id, dt, tmp.pose);
    } else {
      TargetManager::update(id, dt);
    }
    if (last_meas_time > 0.0 && (now - last_meas_time) >= expiration_time_) {
      it = measurements_.erase(it);
      erase(id);
      if (erased) erased->push_back(id);
    } else {
      ++it;
    }
  }
  t_ = t_ + dt;
}

}  // namespace oracle
