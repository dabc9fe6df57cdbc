// =====================================================================================
// oracle/oracle_c.cpp  --  TEST INFRASTRUCTURE ONLY (see te_oracle.hpp header).
// extern "C" surface over the oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// All symbols are prefixed orc_ so they can never be confused with the product's C-ABI.
// =====================================================================================
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>

#include "te_oracle.hpp"

using namespace oracle;

namespace {
struct OrcManager {
  std::shared_ptr<TargetManager> mgr;
  TickTargetManager* tick = nullptr;   // non-null when created through orc_tick_new
};
inline TargetManager* M(void* h) { return static_cast<OrcManager*>(h)->mgr.get(); }
}  // namespace

extern "C" {

// ---- manager lifecycle (mirrors target_manager_c.h:28-37) ---------------------------
void* orc_manager_new(const char* file) {
  try {
    OrcManager* h = new OrcManager();
    if (file && file[0]) h->mgr.reset(new TargetManager(std::string(file)));
    else h->mgr.reset(new TargetManager());
    return h;
  } catch (const char*) {
    return nullptr;
  }
}
void orc_manager_delete(void* h) { delete static_cast<OrcManager*>(h); }

int orc_load_yaml(const char* file, double* Q, double* R, double* P, int* n, int* m, int* type, double* freq) {
  Mat q, r, p;
  target_t t = UNIFORM_VELOCITY;
  if (!loadYamlFile(file, q, r, p, t, freq)) return 0;
  *n = q.r; *m = r.r; *type = (int)t;
  std::memcpy(Q, q.d.data(), sizeof(double) * q.d.size());
  std::memcpy(R, r.d.data(), sizeof(double) * r.d.size());
  std::memcpy(P, p.d.data(), sizeof(double) * p.d.size());
  return 1;
}

// default-model init (target_manager_c.cpp:20-24): v0 = a0 = 0
int orc_init_default(void* h, unsigned id, double dt0, const double* p0, double t0) {
  try { M(h)->init(id, dt0, t0, p0); return 1; } catch (const char*) { return 0; }
}
// full init (target_manager.hpp init(type,id,dt0,t0,Q,R,P0,p0,v0,a0)); Q/R/P0 column-major flat
void orc_init_full(void* h, int type, unsigned id, double dt0, double t0, const double* Q, int n, const double* R, int m,
                   const double* P0, const double* p0, const double* v0, const double* a0) {
  M(h)->init((target_t)type, id, dt0, t0, Mat::MapColMajor(Q, n), Mat::MapColMajor(R, m), Mat::MapColMajor(P0, n), p0, v0, a0);
}
int orc_update_meas(void* h, unsigned id, double dt, const double* meas) { return M(h)->update(id, dt, meas) ? 1 : 0; }
int orc_update(void* h, unsigned id, double dt) { return M(h)->update(id, dt) ? 1 : 0; }
void orc_update_all(void* h, double dt) { M(h)->update(dt); }
int orc_erase(void* h, unsigned id) { return M(h)->erase(id) ? 1 : 0; }
int orc_get_est_pose(void* h, unsigned id, double* pose) { return M(h)->getTargetPose(id, pose) ? 1 : 0; }
int orc_get_est_twist(void* h, unsigned id, double* tw) { return M(h)->getTargetTwist(id, tw) ? 1 : 0; }
int orc_get_est_acceleration(void* h, unsigned id, double* a) { return M(h)->getTargetAcceleration(id, a) ? 1 : 0; }
long long orc_get_n_measurements(void* h, unsigned id) { return M(h)->getNumberMeasurements(id); }
int orc_num_targets(void* h) { return (int)M(h)->getAvailableTargets().size(); }
int orc_get_ids(void* h, unsigned* out, int cap) {
  auto ids = M(h)->getAvailableTargets();
  int n = (int)std::min((size_t)cap, ids.size());
  for (int i = 0; i < n; ++i) out[i] = ids[i];
  return (int)ids.size();
}

// ---- per-target views (TargetInterface getters) -------------------------------------
int orc_get_state(void* h, unsigned id, double* x, double* P /*row-major n*n*/, double* t, long long* n_meas, double* prev_rpy) {
  auto tg = M(h)->getTarget(id);
  if (!tg) return 0;
  int n = (int)tg->getN();
  const Vec& xs = tg->getEstimator()->getState();
  const Mat& Pm = tg->getEstimator()->getP();
  if (x) for (int i = 0; i < n; ++i) x[i] = xs[i];
  if (P) for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) P[i * n + j] = Pm(i, j);
  if (t) *t = tg->getTime();
  if (n_meas) *n_meas = tg->getNumberMeasurements();
  if (prev_rpy) for (int i = 0; i < 3; ++i) prev_rpy[i] = tg->prevRpy()[i];
  return n;
}
int orc_get_pose_at(void* h, unsigned id, double t1, double* pose) {
  auto tg = M(h)->getTarget(id); if (!tg) return 0; tg->getEstimatedPoseAt(t1, pose); return 1;
}
int orc_get_twist_at(void* h, unsigned id, double t1, double* tw) {
  auto tg = M(h)->getTarget(id); if (!tg) return 0; tg->getEstimatedTwistAt(t1, tw); return 1;
}
int orc_get_acc_at(void* h, unsigned id, double t1, double* a) {
  auto tg = M(h)->getTarget(id); if (!tg) return 0; tg->getEstimatedAccelerationAt(t1, a); return 1;
}
int orc_get_measured_pose(void* h, unsigned id, double* p) {
  auto tg = M(h)->getTarget(id); if (!tg) return 0; tg->getMeasuredPose(p); return 1;
}
int orc_get_pose_internal(void* h, unsigned id, double* p6) {
  auto tg = M(h)->getTarget(id); if (!tg) return 0; tg->getPoseInternal(p6); return 1;
}
double orc_get_period_estimate(void* h, unsigned id) {
  auto tg = M(h)->getTarget(id); if (!tg) return -2.0; return tg->getPeriodEstimate();
}

// ---- batched drivers (amortise ctypes overhead; semantics = loops over the calls above) ----
// one tick over n ids: action 0 = nothing, 1 = update(id,dt) predict-only, 2 = update(id,dt,meas)
void orc_step_batch(void* h, int n, const unsigned* ids, double dt, const double* meas /*[n][7]*/, const unsigned char* action) {
  TargetManager* m = M(h);
  for (int i = 0; i < n; ++i) {
    if (action[i] == 2) m->update(ids[i], dt, meas + 7 * (size_t)i);
    else if (action[i] == 1) m->update(ids[i], dt);
  }
}
// The same over a SET of managers (one per host thread; id k lives in manager k % n_mgr): the per-target arithmetic is the
// single-threaded TargetManager::update of each manager, only the loop over independent targets is spread over threads, so
// that the parity tests at bench scale (tens of thousands of targets, SURVEY.md 8(d) 4096 x 2000) finish in seconds.
void orc_step_batch_mt(void** hs, int n_mgr, int n, const unsigned* ids, double dt, const double* meas /*[n][7]*/, const unsigned char* action) {
  auto worker = [&](int t) {
    TargetManager* m = M(hs[t]);
    for (int i = 0; i < n; ++i) {
      if ((int)(ids[i] % (unsigned)n_mgr) != t) continue;
      if (action[i] == 2) m->update(ids[i], dt, meas + 7 * (size_t)i);
      else if (action[i] == 1) m->update(ids[i], dt);
    }
  };
  if (n_mgr == 1) { worker(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_mgr; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
}
// n_ticks ticks in one call: meas [n_ticks][n][7], action [n_ticks][n]
void orc_step_ticks_mt(void** hs, int n_mgr, int n, const unsigned* ids, int n_ticks, double dt, const double* meas, const unsigned char* action) {
  auto worker = [&](int t) {
    TargetManager* m = M(hs[t]);
    for (int k = 0; k < n_ticks; ++k) {
      const double* mk = meas + (size_t)k * n * 7;
      const unsigned char* ak = action + (size_t)k * n;
      for (int i = 0; i < n; ++i) {
        if ((int)(ids[i] % (unsigned)n_mgr) != t) continue;
        if (ak[i] == 2) m->update(ids[i], dt, mk + 7 * (size_t)i);
        else if (ak[i] == 1) m->update(ids[i], dt);
      }
    }
  };
  if (n_mgr == 1) { worker(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_mgr; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
}
// bulk init with one (Q, R, P0) and a per-target P0 scale (scale may be null)
void orc_init_batch_mt(void** hs, int n_mgr, int type, int n, const unsigned* ids, double dt0, const double* t0 /*[n] or null*/, const double* Q, int nq,
                       const double* R, int mr, const double* P0, const double* scale, const double* p0 /*[n][7]*/) {
  Mat Qm = Mat::MapColMajor(Q, nq), Rm = Mat::MapColMajor(R, mr), Pm = Mat::MapColMajor(P0, nq);
  auto worker = [&](int t) {
    TargetManager* m = M(hs[t]);
    for (int i = 0; i < n; ++i) {
      if ((int)(ids[i] % (unsigned)n_mgr) != t) continue;
      Mat Ps = Pm;
      if (scale) for (auto& v : Ps.d) v = scale[i] * v;
      m->init((target_t)type, ids[i], dt0, t0 ? t0[i] : 0.0, Qm, Rm, Ps, p0 + 7 * (size_t)i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_mgr; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
}
// bulk state read-back (row-major P); found[i] = 0 for unknown ids
void orc_get_states_mt(void** hs, int n_mgr, int n, const unsigned* ids, int N, double* x, double* P, double* t, long long* n_meas, double* prev_rpy,
                       unsigned char* found) {
  for (int i = 0; i < n; ++i) {
    auto tg = M(hs[ids[i] % (unsigned)n_mgr])->getTarget(ids[i]);
    if (found) found[i] = tg ? 1 : 0;
    if (!tg) continue;
    const Vec& xs = tg->getEstimator()->getState();
    const Mat& Pm = tg->getEstimator()->getP();
    for (int a = 0; a < N; ++a) x[(size_t)i * N + a] = xs[a];
    for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) P[(size_t)i * N * N + a * N + b] = Pm(a, b);
    t[i] = tg->getTime();
    n_meas[i] = tg->getNumberMeasurements();
    for (int a = 0; a < 3; ++a) prev_rpy[(size_t)i * 3 + a] = tg->prevRpy()[a];
  }
}
// n_steps ticks of one target; records x,P (row-major) every `every` steps into out_x/out_P (may be null)
void orc_run_stream(void* h, unsigned id, int n_steps, double dt, const double* meas /*[n_steps][7]*/,
                    const unsigned char* action /*[n_steps] or null = all 2*/, int every, double* out_x, double* out_P,
                    double* out_pose /*[n_steps][7] or null*/, double* out_twist /*[n_steps][6] or null*/) {
  TargetManager* m = M(h);
  auto tg = m->getTarget(id);
  if (!tg) return;
  const int n = (int)tg->getN();
  int rec = 0;
  for (int k = 0; k < n_steps; ++k) {
    unsigned char a = action ? action[k] : 2;
    if (a == 2) m->update(id, dt, meas + 7 * (size_t)k);
    else if (a == 1) m->update(id, dt);
    if (out_pose) m->getTargetPose(id, out_pose + 7 * (size_t)k);
    if (out_twist) m->getTargetTwist(id, out_twist + 6 * (size_t)k);
    if (every > 0 && ((k + 1) % every == 0)) {
      const Vec& xs = tg->getEstimator()->getState();
      const Mat& Pm = tg->getEstimator()->getP();
      if (out_x) for (int i = 0; i < n; ++i) out_x[(size_t)rec * n + i] = xs[i];
      if (out_P) for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out_P[(size_t)rec * n * n + i * n + j] = Pm(i, j);
      ++rec;
    }
  }
}

// ---- the reference's own test scenario generator (test/target_manager_test.cpp:9-20,82-115) ----
// One global default-seeded std::default_random_engine shared by the four tests (draw order
// UV -> UA -> AR -> AV, 3 draws per sample); libstdc++'s engine/distribution are deterministic.
// out_meas/out_real: [n_tests][n_points][7]
void orc_reftest_streams(double dt, int n_points, int n_tests, double* out_meas, double* out_real) {
  std::default_random_engine gen;
  std::normal_distribution<double> normal_dist(0.0, 0.01);
  const double end_goal[3] = {0.2, 0.3, 0.4};
  const double omega[3] = {3.0, 0.01, 0.1};
  for (int tst = 0; tst < n_tests; ++tst) {
    double* meas = out_meas + (size_t)tst * n_points * 7;
    double* real = out_real + (size_t)tst * n_points * 7;
    Quat q;   // identity
    for (int i = 0; i < n_points; ++i) {
      // (Eigen) VectorXd::LinSpaced(n, 0, high): low + i*step, last element = high
      for (int a = 0; a < 3; ++a) {
        double step = (end_goal[a] - 0.0) / (double)(n_points - 1);
        real[7 * i + a] = (i == n_points - 1) ? end_goal[a] : (0.0 + i * step);
      }
      meas[7 * i + 0] = real[7 * i + 0] + normal_dist(gen);
      meas[7 * i + 1] = real[7 * i + 1] + normal_dist(gen);
      meas[7 * i + 2] = real[7 * i + 2] + normal_dist(gen);
      meas[7 * i + 3] = real[7 * i + 3] = q.x;
      meas[7 * i + 4] = real[7 * i + 4] = q.y;
      meas[7 * i + 5] = real[7 * i + 5] = q.z;
      meas[7 * i + 6] = real[7 * i + 6] = q.w;
      double Qm[4][4];
      Qtran(dt, omega, Qm);
      double c[4] = {q.x, q.y, q.z, q.w}, r[4];
      for (int a = 0; a < 4; ++a) r[a] = Qm[a][0] * c[0] + Qm[a][1] * c[1] + Qm[a][2] * c[2] + Qm[a][3] * c[3];
      q.x = r[0]; q.y = r[1]; q.z = r[2]; q.w = r[3];
      quatNormalize(q);
    }
  }
}
// first k draws of the N(mean,std) stream from a default-seeded engine (SURVEY.md section 4 anchor)
void orc_libstdcxx_normal(double mean, double stddev, int k, double* out) {
  std::default_random_engine gen;
  std::normal_distribution<double> d(mean, stddev);
  for (int i = 0; i < k; ++i) out[i] = d(gen);
}

// ---- geometry / filters / polynomial -------------------------------------------------
void orc_quat_to_rpy(const double* q4 /*x y z w*/, double* rpy) {
  Quat q; q.x = q4[0]; q.y = q4[1]; q.z = q4[2]; q.w = q4[3];
  quatToRpy(q, rpy);
}
void orc_rpy_to_quat(const double* rpy, double* q4) {
  Quat q; rpyToQuat(rpy, q); q4[0] = q.x; q4[1] = q.y; q4[2] = q.z; q4[3] = q.w;
}
void orc_quat_to_rot(const double* q4, double* R9) {
  Quat q; q.x = q4[0]; q.y = q4[1]; q.z = q4[2]; q.w = q4[3];
  Mat3 R = quatToRotationMatrix(q);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[3 * i + j] = R.m[i][j];
}
void orc_rot_to_quat(const double* R9, double* q4) {
  Mat3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = R9[3 * i + j];
  Quat q = rotationMatrixToQuat(R); q4[0] = q.x; q4[1] = q.y; q4[2] = q.z; q4[3] = q.w;
}
void orc_rot_to_rpy(const double* R9, double* rpy) {
  Mat3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = R9[3 * i + j];
  rotToRpy(R, rpy);
}
void orc_unwrap3(const double* prev, const double* nw, double* out) { unwrap3(prev, nw, out); }
double orc_constrain_angle(double x) { return constrainAngle(x); }
double orc_angle_diff(double a, double b) { return angleDiff(a, b); }
double orc_wrap_min_max(double x, double mn, double mx) { return wrapMinMax(x, mn, mx); }
void orc_qtran(double dt, const double* omega, double* Q16) {
  double Q[4][4]; Qtran(dt, omega, Q);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Q16[4 * i + j] = Q[i][j];
}
double orc_to_sec(unsigned sec, unsigned nsec) { return toSec(sec, nsec); }
void orc_inverse(const double* Mrow, int n, double* out_row) {
  Mat A(n, n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A(i, j) = Mrow[i * n + j];
  Mat X = inversePartialPivLU(A);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out_row[i * n + j] = X(i, j);
}

void* orc_mavg_new(unsigned n) { return new MovingAvgFilter(n); }
void orc_mavg_delete(void* f) { delete static_cast<MovingAvgFilter*>(f); }
double orc_mavg_update(void* f, double v) { return static_cast<MovingAvgFilter*>(f)->update(v); }
double orc_mavg_variance(void* f) { return static_cast<MovingAvgFilter*>(f)->getVariance(); }
void* orc_avg_new(unsigned n) { return new AvgFilter(n); }
void orc_avg_delete(void* f) { delete static_cast<AvgFilter*>(f); }
double orc_avg_update(void* f, double v) { return static_cast<AvgFilter*>(f)->update(v); }

int orc_get_id(const char* s, unsigned* id) {
  try { return getId(std::string(s), *id) ? 1 : 0; } catch (...) { return -1; }   // std::stoi may throw
}

int orc_poly_roots(const double* coeffs, int ncoef, double* re, double* im) {
  std::vector<double> c(coeffs, coeffs + ncoef);
  auto r = polynomialRoots(c);
  for (size_t i = 0; i < r.size(); ++i) { re[i] = r[i].real(); im[i] = r[i].imag(); }
  return (int)r.size();
}
double orc_lowest_real_root(const double* coeffs, int ncoef) {
  std::vector<double> c(coeffs, coeffs + ncoef);
  return lowestRealRoot(c);
}

// ---- IntersectionSolver ---------------------------------------------------------------
void* orc_isolver_new(void* h, unsigned filters_length) {
  return new IntersectionSolver(static_cast<OrcManager*>(h)->mgr, filters_length);
}
void orc_isolver_delete(void* s) { delete static_cast<IntersectionSolver*>(s); }
double orc_isolver_time(void* s, unsigned id, double t1, const double* origin, double radius) {
  return static_cast<IntersectionSolver*>(s)->getIntersectionTimeWithSphere(id, t1, origin, radius);
}
int orc_isolver_pose(void* s, unsigned id, double t1, double pos_th, double ang_th, const double* origin, double radius, double* pose) {
  return static_cast<IntersectionSolver*>(s)->getIntersectionPoseWithSphere(id, t1, pos_th, ang_th, origin, radius, pose) ? 1 : 0;
}

// ---- RosTargetManager tick semantics ---------------------------------------------------
void* orc_tick_new(int type, const double* Q, int n, const double* R, int m, const double* P) {
  OrcManager* h = new OrcManager();
  h->tick = new TickTargetManager((target_t)type, Mat::MapColMajor(Q, n), Mat::MapColMajor(R, m), Mat::MapColMajor(P, n));
  h->mgr.reset(h->tick);
  return h;
}
void orc_tick_set_expiration(void* h, double t) { static_cast<OrcManager*>(h)->tick->setExpirationTime(t); }
void orc_tick_set_token(void* h, const char* s) { static_cast<OrcManager*>(h)->tick->setTargetTokenName(s); }
// frames: '\n'-separated child_frame_id list; stamps [n][2] (sec,nsec); poses [n][7]
void orc_tick_callback(void* h, int n, const char* frames, const unsigned* stamps, const double* poses) {
  std::vector<TfRecord> msg(n);
  const char* p = frames;
  for (int i = 0; i < n; ++i) {
    const char* e = std::strchr(p, '\n');
    msg[i].child_frame_id = e ? std::string(p, e - p) : std::string(p);
    p = e ? e + 1 : p + std::strlen(p);
    msg[i].tr.sec = stamps[2 * i]; msg[i].tr.nsec = stamps[2 * i + 1];
    std::memcpy(msg[i].tr.pose, poses + 7 * (size_t)i, 7 * sizeof(double));
  }
  static_cast<OrcManager*>(h)->tick->measurementCallBack(msg);
}
// numeric-id fast path of the callback: frame name "<token>_<id>" for every record
void orc_tick_callback_ids(void* h, int n, const unsigned* ids, const unsigned* stamps, const double* poses) {
  std::vector<TfRecord> msg(n);
  for (int i = 0; i < n; ++i) {
    msg[i].child_frame_id = "target_" + std::to_string(ids[i]);
    msg[i].tr.sec = stamps[2 * i]; msg[i].tr.nsec = stamps[2 * i + 1];
    std::memcpy(msg[i].tr.pose, poses + 7 * (size_t)i, 7 * sizeof(double));
  }
  static_cast<OrcManager*>(h)->tick->measurementCallBack(msg);
}
int orc_tick_update(void* h, double dt, unsigned now_sec, unsigned now_nsec, unsigned* erased, int cap) {
  std::vector<unsigned> er;
  static_cast<OrcManager*>(h)->tick->tick(dt, now_sec, now_nsec, &er);
  for (int i = 0; i < (int)er.size() && i < cap; ++i) erased[i] = er[i];
  return (int)er.size();
}
double orc_tick_time(void* h) { return static_cast<OrcManager*>(h)->tick->time(); }
int orc_tick_mailboxes(void* h) { return (int)static_cast<OrcManager*>(h)->tick->mailboxCount(); }

// ---- CPU baseline timing (bench.py cpu_baseline / --impl reference) ----------------------
// n_targets of `type` are created (ids 0..n-1, split id % threads over one manager per thread --
// the reference serialises everything behind one manager mutex, so threads>1 is a best case),
// then n_ticks ticks of update(id,dt,meas) are timed.  meas: [n_targets][7] base poses; tick k
// perturbs xyz deterministically so the filter sees a moving target.  Returns seconds.
double orc_bench_steps(int type, const double* Q, int n, const double* R, int m, const double* P0, int n_targets, int n_ticks,
                       int threads, double dt, const double* meas, double miss_prob, double* checksum) {
  if (threads < 1) threads = 1;
  std::vector<std::unique_ptr<TargetManager>> mgrs;
  for (int t = 0; t < threads; ++t) mgrs.emplace_back(new TargetManager());
  Mat Qm = Mat::MapColMajor(Q, n), Rm = Mat::MapColMajor(R, m), Pm = Mat::MapColMajor(P0, n);
  for (int i = 0; i < n_targets; ++i) mgrs[i % threads]->init((target_t)type, (unsigned)i, dt, 0.0, Qm, Rm, Pm, meas + 7 * (size_t)i);
  std::vector<double> sums(threads, 0.0);
  auto worker = [&](int t) {
    TargetManager* mg = mgrs[t].get();
    uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);
    double mm[7];
    for (int k = 0; k < n_ticks; ++k) {
      for (int i = t; i < n_targets; i += threads) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        double u = (double)(rng >> 11) * (1.0 / 9007199254740992.0);
        if (u < miss_prob) { mg->update((unsigned)i, dt); continue; }
        const double* b = meas + 7 * (size_t)i;
        mm[0] = b[0] + 0.01 * k * dt; mm[1] = b[1] - 0.02 * k * dt; mm[2] = b[2] + (u - 0.5) * 0.01;
        mm[3] = b[3]; mm[4] = b[4]; mm[5] = b[5]; mm[6] = b[6];
        mg->update((unsigned)i, dt, mm);
      }
    }
    double p[7];
    for (int i = t; i < n_targets; i += threads) { mg->getTargetPose((unsigned)i, p); sums[t] += p[0] + p[1] + p[2]; }
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

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
