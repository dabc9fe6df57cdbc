// oracle/ref_manager_c.cpp -- TEST INFRASTRUCTURE: extern "C" handle on the REFERENCE's own TargetManager
// (src/target_manager.cpp), IntersectionSolver (src/intersection_solver.cpp), MovingAvgFilter / AvgFilter / getId (utils.hpp),
// compiled unmodified from /root/reference against oracle/eigen_standin (Eigen, yaml-cpp and the ROS adapter header are
// absent from the image; see the stand-in headers for what each of them does and does not pin).  The reference's own C-ABI
// (src/target_manager_c.cpp) is compiled into the same library under its own symbol names.  Only tests load
// oracle/_ref/libref_manager.so.
#include <cstring>
#include <iostream>
#include <sstream>

#include "target_estimation/intersection_solver.hpp"
#include "target_estimation/target_manager.hpp"
#include "target_estimation/utils.hpp"

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
struct Quiet {   // the reference prints matrices / "does not exist" lines to stdout
  std::ostringstream sink;
  std::streambuf *old, *olde;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())), olde(std::cerr.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); std::cerr.rdbuf(olde); }
};
struct Mgr { TargetManager::Ptr m; };
}  // namespace

extern "C" {

void* refm_new(const char* file) {
  Quiet q;
  try {
    Mgr* h = new Mgr;
    h->m.reset(file && file[0] ? new TargetManager(std::string(file)) : new TargetManager());
    return h;
  } catch (const char*) {   // `throw "TargetManager default constructor failed!"` (src/target_manager.cpp:115)
    return nullptr;
  } catch (...) {
    return nullptr;
  }
}
void refm_delete(void* h) { delete static_cast<Mgr*>(h); }
// init(type, id, dt0, t0, Q, R, P0, p0, v0, a0): matrices column-major
void refm_init_full(void* h, int type, unsigned id, double dt0, double t0, const double* Q, int n, const double* R, int m, const double* P0,
                    const double* p0, const double* v0, const double* a0) {
  Quiet q;
  static_cast<Mgr*>(h)->m->init((TargetManager::target_t)type, id, dt0, t0, fromColMajor(Q, n, n), fromColMajor(R, m, m), fromColMajor(P0, n, n),
                                fromVec<Eigen::Vector7d>(p0, 7), fromVec<Eigen::Vector6d>(v0, 6), fromVec<Eigen::Vector6d>(a0, 6));
}
// init(id, dt0, t0, p0) with the default model of the YAML file; returns -1 if the reference throws (no default model)
int refm_init_default(void* h, unsigned id, double dt0, double t0, const double* p0) {
  Quiet q;
  try {
    static_cast<Mgr*>(h)->m->init(id, dt0, t0, fromVec<Eigen::Vector7d>(p0, 7));
  } catch (...) {
    return -1;
  }
  return 0;
}
int refm_update_meas(void* h, unsigned id, double dt, const double* meas) { Quiet q; return static_cast<Mgr*>(h)->m->update(id, dt, fromVec<Eigen::Vector7d>(meas, 7)) ? 1 : 0; }
int refm_update(void* h, unsigned id, double dt) { Quiet q; return static_cast<Mgr*>(h)->m->update(id, dt) ? 1 : 0; }
void refm_update_all(void* h, double dt) { Quiet q; static_cast<Mgr*>(h)->m->update(dt); }
int refm_erase(void* h, unsigned id) { Quiet q; return static_cast<Mgr*>(h)->m->erase(id) ? 1 : 0; }
int refm_ids(void* h, unsigned* out, int cap) {
  const std::vector<unsigned int> ids = static_cast<Mgr*>(h)->m->getAvailableTargets();
  for (int i = 0; i < (int)ids.size() && i < cap; ++i) out[i] = ids[(size_t)i];
  return (int)ids.size();
}
int refm_state(void* h, unsigned id, double* x, double* P /* column-major */, double* t, long long* n_meas) {
  Quiet q;
  TargetInterface::Ptr tg = static_cast<Mgr*>(h)->m->getTarget(id);
  if (!tg) return 0;
  const Eigen::VectorXd& xs = tg->getEstimator()->getState();
  const Eigen::MatrixXd& Ps = tg->getEstimator()->getP();
  if (x) std::memcpy(x, xs.data(), sizeof(double) * (size_t)xs.size());
  if (P) std::memcpy(P, Ps.data(), sizeof(double) * (size_t)Ps.size());
  if (t) *t = tg->getTime();
  if (n_meas) *n_meas = tg->getNumberMeasurements();
  return (int)xs.size();
}
int refm_pose(void* h, unsigned id, double* out) { Quiet q; Eigen::Vector7d v = fromVec<Eigen::Vector7d>(out, 7); const bool ok = static_cast<Mgr*>(h)->m->getTargetPose(id, v); std::memcpy(out, v.data(), 56); return ok; }
int refm_twist(void* h, unsigned id, double* out) { Quiet q; Eigen::Vector6d v = fromVec<Eigen::Vector6d>(out, 6); const bool ok = static_cast<Mgr*>(h)->m->getTargetTwist(id, v); std::memcpy(out, v.data(), 48); return ok; }
int refm_acc(void* h, unsigned id, double* out) { Quiet q; Eigen::Vector6d v = fromVec<Eigen::Vector6d>(out, 6); const bool ok = static_cast<Mgr*>(h)->m->getTargetAcceleration(id, v); std::memcpy(out, v.data(), 48); return ok; }
long long refm_n_meas(void* h, unsigned id) { Quiet q; return static_cast<Mgr*>(h)->m->getNumberMeasurements(id); }

// ---- IntersectionSolver ----
void* refs_new(void* h, unsigned filters_length) { return new IntersectionSolver(static_cast<Mgr*>(h)->m, filters_length); }
void refs_delete(void* s) { delete static_cast<IntersectionSolver*>(s); }
double refs_time(void* s, unsigned id, double t1, const double* origin, double radius) {
  Quiet q;
  return static_cast<IntersectionSolver*>(s)->getIntersectionTimeWithSphere(id, t1, Eigen::Vector3d(origin[0], origin[1], origin[2]), radius);
}
int refs_pose(void* s, unsigned id, double t1, double pos_th, double ang_th, const double* origin, double radius, double* pose) {
  Quiet q;
  Eigen::Vector7d p;
  const bool ok = static_cast<IntersectionSolver*>(s)->getIntersectionPoseWithSphere(id, t1, pos_th, ang_th, Eigen::Vector3d(origin[0], origin[1], origin[2]),
                                                                                   radius, p);
  std::memcpy(pose, p.data(), 56);
  return ok ? 1 : 0;
}

// ---- utils.hpp ----
void* refu_mavg_new(unsigned n) { return new MovingAvgFilter(n); }
void refu_mavg_delete(void* f) { delete static_cast<MovingAvgFilter*>(f); }
double refu_mavg_update(void* f, double v) { return static_cast<MovingAvgFilter*>(f)->update(v); }
double refu_mavg_variance(void* f) { return static_cast<MovingAvgFilter*>(f)->getVariance(); }
void* refu_avg_new(unsigned n) { return new AvgFilter(n); }
void refu_avg_delete(void* f) { delete static_cast<AvgFilter*>(f); }
double refu_avg_update(void* f, double v) { return static_cast<AvgFilter*>(f)->update(v); }
int refu_get_id(const char* s, unsigned* id) {
  try { return getId(std::string(s), *id) ? 1 : 0; } catch (...) { return -1; }   // std::stoi may throw
}
double refu_to_sec(unsigned sec, unsigned nsec) { return toSec(sec, nsec); }

}  // extern "C"
