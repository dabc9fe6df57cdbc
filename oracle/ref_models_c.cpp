// oracle/ref_models_c.cpp -- TEST INFRASTRUCTURE: extern "C" handle on the REFERENCE's own target models
// (src/types/*.cpp + src/target_interface.cpp + src/kalman.cpp, compiled unmodified from /root/reference against
// oracle/eigen_standin; see the header of oracle/eigen_standin/Eigen/Dense for what that does and does not pin).
// Only tests load the resulting oracle/_ref/libref_models.so.
#include <cstring>
#include <iostream>
#include <sstream>

#include "target_estimation/types/angular_rates.hpp"
#include "target_estimation/types/angular_velocities.hpp"
#include "target_estimation/types/uniform_acceleration.hpp"
#include "target_estimation/types/uniform_velocity.hpp"

namespace {
Eigen::MatrixXd fromColMajor(const double* v, int r, int c) {
  Eigen::MatrixXd m(r, c);
  std::memcpy(m.data(), v, sizeof(double) * (size_t)r * c);
  return m;
}
template <class V> V fromVec(const double* v, int n) {
  V x;
  for (int i = 0; i < n; ++i) x(i) = v ? v[i] : 0.0;
  return x;
}
// the constructors print every matrix to stdout (src/target_interface.cpp:57-78): swallow that
struct Quiet {
  std::ostringstream sink;
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// type: the reference's enum target_t {ANGULAR_RATES = 0, ANGULAR_VELOCITIES, UNIFORM_ACCELERATION, UNIFORM_VELOCITY}
// (target_manager.hpp:38); matrices column-major
void* ref_target_new(int type, unsigned id, double dt0, double t0, const double* Q, int n, const double* R, int m, const double* P0,
                     const double* p0, const double* v0, const double* a0) {
  Quiet q;
  const Eigen::MatrixXd Qm = fromColMajor(Q, n, n), Rm = fromColMajor(R, m, m), Pm = fromColMajor(P0, n, n);
  const Eigen::Vector7d p = fromVec<Eigen::Vector7d>(p0, 7);
  const Eigen::Vector6d v = fromVec<Eigen::Vector6d>(v0, 6), a = fromVec<Eigen::Vector6d>(a0, 6);
  TargetInterface* t = nullptr;
  switch (type) {
    case 0: t = new TargetAngularRates(id, dt0, t0, Qm, Rm, Pm, p, v, a); break;
    case 1: t = new TargetAngularVelocities(id, dt0, t0, Qm, Rm, Pm, p, v, a); break;
    case 2: t = new TargetUniformAcceleration(id, dt0, t0, Qm, Rm, Pm, p, v, a); break;
    case 3: t = new TargetUniformVelocity(id, dt0, t0, Qm, Rm, Pm, p, v, a); break;
    default: break;
  }
  return t;
}
void ref_target_delete(void* h) { delete static_cast<TargetInterface*>(h); }
void ref_target_add_measurement(void* h, double dt, const double* meas7) {
  static_cast<TargetInterface*>(h)->addMeasurement(dt, fromVec<Eigen::Vector7d>(meas7, 7));
}
void ref_target_update(void* h, double dt) { static_cast<TargetInterface*>(h)->update(dt); }
// x [n], P [n*n] column-major, t, n_meas
int ref_target_state(void* h, double* x, double* P, double* t, long long* n_meas) {
  TargetInterface* tg = static_cast<TargetInterface*>(h);
  const Eigen::VectorXd& xs = tg->getEstimator()->getState();
  const Eigen::MatrixXd& Ps = tg->getEstimator()->getP();
  if (x) std::memcpy(x, xs.data(), sizeof(double) * (size_t)xs.size());
  if (P) std::memcpy(P, Ps.data(), sizeof(double) * (size_t)Ps.size());
  if (t) *t = tg->getTime();
  if (n_meas) *n_meas = tg->getNumberMeasurements();
  return (int)xs.size();
}
// current estimates (use_t1 == 0) or getEstimatedPose/Twist/Acceleration(t1)
void ref_target_estimates(void* h, int use_t1, double t1, double* pose7, double* twist6, double* acc6) {
  TargetInterface* tg = static_cast<TargetInterface*>(h);
  const Eigen::Vector7d p = use_t1 ? tg->getEstimatedPose(t1) : tg->getEstimatedPose();
  const Eigen::Vector6d tw = use_t1 ? tg->getEstimatedTwist(t1) : tg->getEstimatedTwist();
  const Eigen::Vector6d ac = use_t1 ? tg->getEstimatedAcceleration(t1) : tg->getEstimatedAcceleration();
  if (pose7) std::memcpy(pose7, p.data(), 7 * sizeof(double));
  if (twist6) std::memcpy(twist6, tw.data(), 6 * sizeof(double));
  if (acc6) std::memcpy(acc6, ac.data(), 6 * sizeof(double));
}

}  // extern "C"
