// oracle/ref_manager_c.cpp -- TEST INFRASTRUCTURE: extern "C" handle on the REFERENCE's own TargetManager
// (src/target_manager.cpp), IntersectionSolver (src/intersection_solver.cpp), MovingAvgFilter / AvgFilter / getId (utils.hpp),
// compiled unmodified from /root/reference against oracle/eigen_standin (Eigen, yaml-cpp and the ROS adapter header are
// absent from the image; see the stand-in headers for what each of them does and does not pin).  The reference's own C-ABI
// (src/target_manager_c.cpp) is compiled into the same library under its own symbol names.  Only tests load
// oracle/_ref/libref_manager.so.
#include <cstring>
#include <iostream>
#include <sstream>

#include <chrono>
#include <thread>
#include <memory>
#include <vector>

#include "target_estimation/intersection_solver.hpp"
#include "target_estimation/target_manager.hpp"
#include "target_estimation/target_manager_ros.hpp"
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
struct Mgr { TargetManager::Ptr m; std::shared_ptr<ros::NodeHandle::State> nh_state; };
std::shared_ptr<ros::NodeHandle::State>& g_last_nh_state(void* h) { return static_cast<Mgr*>(h)->nh_state; }
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

// ---- RosTargetManager (src/target_manager_ros.cpp) on the ROS stand-ins of oracle/eigen_standin ----
// The handle is a Mgr too: refm_ids / refm_state / refm_pose ... work on it.
void* refr_new(const char* type, const double* Q, int nq, const double* R, int nr, const double* P, int np) {
  Quiet q;
  try {
    ros::NodeHandle nh;
    nh.st_->lists["Q"].assign(Q, Q + nq);      // flat lists, as `rosparam load models/*.yaml` puts them on the server
    nh.st_->lists["R"].assign(R, R + nr);
    nh.st_->lists["P"].assign(P, P + np);
    nh.st_->strings["type"] = type;
    Mgr* h = new Mgr;
    h->m.reset(new RosTargetManager(nh));
    h->nh_state = nh.st_;
    return h;
  } catch (...) {
    return nullptr;
  }
}
static RosTargetManager* R_(void* h) { return static_cast<RosTargetManager*>(static_cast<Mgr*>(h)->m.get()); }
void refr_set_expiration(void* h, double t) { R_(h)->setExpirationTime(t); }
void refr_set_token(void* h, const char* s) { R_(h)->setTargetTokenName(s); }
// deliver one TFMessage to the subscribed (private) callback: frames separated by '\n', stamps [n][2], poses [n][7]
void refr_callback(void* h, int n, const char* frames, const unsigned* stamps, const double* poses, const char* frame_id) {
  Quiet q;
  std::shared_ptr<tf2_msgs::TFMessage> msg(new tf2_msgs::TFMessage);
  std::istringstream is(frames);
  std::string name;
  for (int i = 0; i < n; ++i) {
    std::getline(is, name);
    geometry_msgs::TransformStamped t;
    t.header.stamp.sec = stamps[2 * i];
    t.header.stamp.nsec = stamps[2 * i + 1];
    t.header.frame_id = frame_id ? frame_id : "";
    t.child_frame_id = name;
    t.transform.translation.x = poses[7 * i + 0]; t.transform.translation.y = poses[7 * i + 1]; t.transform.translation.z = poses[7 * i + 2];
    t.transform.rotation.x = poses[7 * i + 3]; t.transform.rotation.y = poses[7 * i + 4]; t.transform.rotation.z = poses[7 * i + 5];
    t.transform.rotation.w = poses[7 * i + 6];
    msg->transforms.push_back(t);
  }
  // the NodeHandle copy inside the manager shares its state with the one the constructor received: find the callback there
  RosTargetManager* r = R_(h);
  (void)r;
  g_last_nh_state(h)->callback(std::static_pointer_cast<const void>(std::shared_ptr<const tf2_msgs::TFMessage>(msg)));
}
// RosTargetManager::update(dt) with ros::Time::now() = (now_sec, now_nsec); returns the number of transforms broadcast
int refr_update(void* h, double dt, unsigned now_sec, unsigned now_nsec) {
  Quiet q;
  ros::Time::clock() = ros::Time(now_sec, now_nsec);
  tf::TransformBroadcaster::log().clear();
  R_(h)->update(dt);
  return (int)tf::TransformBroadcaster::log().size();
}
// k-th transform of the last update: child frame name (<= 63 chars), parent frame, origin xyz + rotation xyzw
void refr_broadcast(int k, char* child, char* parent, double* pose7) {
  const tf::StampedTransform& t = tf::TransformBroadcaster::log()[(size_t)k];
  std::strncpy(child, t.child_frame_id.c_str(), 63); child[63] = 0;
  std::strncpy(parent, t.frame_id.c_str(), 63); parent[63] = 0;
  pose7[0] = t.origin.x_; pose7[1] = t.origin.y_; pose7[2] = t.origin.z_;
  pose7[3] = t.rotation.x_; pose7[4] = t.rotation.y_; pose7[5] = t.rotation.z_; pose7[6] = t.rotation.w_;
}

// ---- CPU baseline timing on the reference's own TargetManager (bench.py cpu_baseline / --impl reference) ----
// Same workload, arguments and return value as orc_bench_steps (oracle_c.cpp): n_targets of `type` (ids 0..n-1, id % threads
// over one reference TargetManager per thread -- the reference serialises a manager behind one mutex, so more threads on ONE
// manager would not help it), then n_ticks ticks of TargetManager::update(id, dt, meas) / update(id, dt) are timed.
// NOTE: this is the reference's code on the stand-in Eigen of oracle/eigen_standin (heap-backed dynamic matrices, no
// vectorisation), not on real Eigen: bench.py reports it beside the oracle port and says so.
double refm_bench_steps(int type, const double* Q, int n, const double* R, int m, const double* P0, int n_targets, int n_ticks,
                        int threads, double dt, const double* meas, double miss_prob, double* checksum) {
  Quiet q;
  if (threads < 1) threads = 1;
  std::vector<std::unique_ptr<TargetManager>> mgrs;
  for (int t = 0; t < threads; ++t) mgrs.emplace_back(new TargetManager());
  const Eigen::MatrixXd Qm = fromColMajor(Q, n, n), Rm = fromColMajor(R, m, m), Pm = fromColMajor(P0, n, n);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n_targets; ++i)
    mgrs[(size_t)(i % threads)]->init((TargetManager::target_t)type, (unsigned)i, dt, 0.0, Qm, Rm, Pm, fromVec<Eigen::Vector7d>(meas + 7 * (size_t)i, 7),
                                      fromVec<Eigen::Vector6d>(zero6, 6), fromVec<Eigen::Vector6d>(zero6, 6));
  std::vector<double> sums((size_t)threads, 0.0);
  auto worker = [&](int t) {
    TargetManager* mg = mgrs[(size_t)t].get();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);
    Eigen::Vector7d mm;
    for (int k = 0; k < n_ticks; ++k) {
      for (int i = t; i < n_targets; i += threads) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        double u = (double)(rng >> 11) * (1.0 / 9007199254740992.0);
        if (u < miss_prob) { mg->update((unsigned)i, dt); continue; }
        const double* b = meas + 7 * (size_t)i;
        mm(0) = b[0] + 0.01 * k * dt; mm(1) = b[1] - 0.02 * k * dt; mm(2) = b[2] + (u - 0.5) * 0.01;
        mm(3) = b[3]; mm(4) = b[4]; mm(5) = b[5]; mm(6) = b[6];
        mg->update((unsigned)i, dt, mm);
      }
    }
    Eigen::Vector7d p;
    for (int i = t; i < n_targets; i += threads) { mg->getTargetPose((unsigned)i, p); sums[(size_t)t] += p(0) + p(1) + p(2); }
  };
  auto t0 = std::chrono::steady_clock::now();
  if (threads == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  double s = 0.0;
  for (double v : sums) s += v;
  if (checksum) *checksum = s;
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- utils.hpp ----
// writeTxtFile (utils.hpp:78-120), both overloads; values row-major
void refu_write_txt_vec(const char* fn, const double* v, int n) {
  Quiet q;
  Eigen::VectorXd x(n);
  for (int i = 0; i < n; ++i) x(i) = v[i];
  writeTxtFile(std::string(fn), x);
}
void refu_write_txt_mat(const char* fn, const double* v, int r, int c) {
  Quiet q;
  Eigen::MatrixXd m(r, c);
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) m(i, j) = v[i * c + j];
  writeTxtFile(std::string(fn), m);
}

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
