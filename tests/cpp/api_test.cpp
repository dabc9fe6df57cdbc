// tests/cpp/api_test.cpp -- the C++ host surface (include/target_estimation_b200/target_manager.hpp) used the way the reference's
// own test/target_manager_test.cpp uses TargetManager: load a model, init one target from the first measurement, then per
// step update(id, dt, meas) -> getTargetPose -> getTargetTwist -> getTarget(id)->getEstimator()->getP(), and the reference's
// tolerances (0.01 on the final position and the mean velocity, test/target_manager_test.cpp:179-189).  Plus what that test
// never touches: erase / unknown ids, the tick manager (frames, first-sight init, expiry, published poses), ids the manager
// knows under ANOTHER model type than the tick's (the by-id host path), the intersection solver.
// Built and run by tests/test_gpu_cpp_api.py (needs a GPU).  Usage: api_test <models dir>
#include <cmath>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "target_estimation_b200/target_manager.hpp"

using namespace target_estimation_b200;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    ++g_checks;                                                                  \
    if (!(cond)) { ++g_failed; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)
#define NEAR(a, b, tol) CHECK(std::fabs((a) - (b)) <= (tol))

static Vector7d pose_at(double x, double y, double z) { return Vector7d{x, y, z, 0.0, 0.0, 0.0, 1.0}; }

int main(int argc, char** argv) {
  const std::string models = argc > 1 ? argv[1] : "models";
  const char* names[4] = {"angular_rates", "angular_velocities", "uniform_acceleration", "uniform_velocity"};
  const double dt = 1.0 / 250.0;
  const unsigned n_points = 10000;
  const double goal[3] = {0.2, 0.3, 0.4};
  std::default_random_engine gen;
  std::normal_distribution<double> noise(0.0, 0.01);

  // ---- the reference's convergence scenario for every model, one manager holding all four targets -----------------
  TargetManager manager;
  manager.quiet = true;
  for (unsigned id = 0; id < 4; ++id) {
    MatrixXd Q, R, P;
    TargetManager::target_t type;
    CHECK(manager.loadYamlFile(models + "/model_" + names[id] + "_params.yaml", Q, R, P, type));
    CHECK((int)type == (int)id);
    std::vector<Vector7d> meas(n_points);
    for (unsigned i = 0; i < n_points; ++i) {
      const double s = (double)i / (n_points - 1);
      meas[i] = pose_at(goal[0] * s + noise(gen), goal[1] * s + noise(gen), goal[2] * s + noise(gen));
    }
    manager.init(type, id, dt, 0.0, Q, R, P, meas[0]);
    manager.init(type, id, dt, 5.0, Q, R, P, meas[10]);   // "already exists": a no-op
    Vector7d pose{};
    Vector6d twist{};
    double mean_v[3] = {0, 0, 0};
    for (unsigned i = 0; i < n_points; ++i) {
      CHECK(manager.update(id, dt, meas[i]));
      CHECK(manager.getTargetPose(id, pose));
      CHECK(manager.getTargetTwist(id, twist));
      for (int k = 0; k < 3; ++k) mean_v[k] += twist[k] / n_points;
      if (i % 2500 == 0) {
        const MatrixXd Pk = manager.getTarget(id)->getEstimator()->getP();
        CHECK(Pk.rows() == Q.rows() && Pk(0, 0) > 0.0 && std::fabs(Pk(0, 1) - Pk(1, 0)) <= 1e-12 * Pk(0, 0));
      }
    }
    for (int k = 0; k < 3; ++k) {
      NEAR(pose[k], goal[k], 0.01);
      NEAR(mean_v[k], goal[k] / (n_points * dt), 0.01);
    }
    CHECK(manager.getNumberMeasurements(id) == (long long)n_points);
    NEAR(manager.getTarget(id)->getTime(), n_points * dt, 1e-6);
  }
  {
    const std::vector<unsigned> ids = manager.getAvailableTargets();
    CHECK(ids.size() == 4 && ids[0] == 0 && ids[3] == 3);
    CHECK(!manager.update(77, dt));                    // unknown id
    Vector7d p{};
    CHECK(!manager.getTargetPose(77, p));
    CHECK(manager.getTarget(77) == nullptr);
    CHECK(manager.erase(1) && !manager.erase(1));
    CHECK(manager.getAvailableTargets().size() == 3);
    manager.update(dt);                                // predict-only for everybody
    NEAR(manager.getTarget(2)->getTime(), (n_points + 1) * dt, 1e-6);
  }

  // ---- intersection solver on the uniform-acceleration target: a falling target meets a sphere below it ----------
  {
    TargetManager::Ptr mp(new TargetManager());
    mp->quiet = true;
    MatrixXd Q, R, P;
    TargetManager::target_t type;
    CHECK(mp->loadYamlFile(models + "/model_uniform_acceleration_params.yaml", Q, R, P, type));
    mp->init(type, 5, dt, 0.0, Q, R, P, pose_at(0.0, 0.0, 2.0));
    for (unsigned i = 1; i <= 250; ++i) {
      const double t = i * dt;
      CHECK(mp->update(5, dt, pose_at(2.0 * t, 0.0, 2.0 - 0.5 * 9.81 * t * t)));
    }
    IntersectionSolver solver(mp, 10);
    const double t1 = 250 * dt;                        // z(t1) = -2.9, falling at 9 m/s, 2 m/s sideways
    const Vector3d origin{2.8, 0.0, -8.0};
    const double d = solver.getIntersectionTimeWithSphere(5, t1, origin, 1.0);
    NEAR(d, 0.4062777585451441, 1e-6);                 // the oracle's value for this stream (tests/orc.py, same scenario)
    // the reference returns the LOWEST real root: a sphere the backward-extended parabola also crosses gives a negative root -> -1
    CHECK(solver.getIntersectionTimeWithSphere(5, t1, Vector3d{-3.2, 0.0, -8.0}, 4.0) == -1.0);
    CHECK(solver.getIntersectionTimeWithSphere(99, t1, origin, 1.0) == -1.0);   // unknown id
  }

  // ---- tick manager: frames, first-sight init, sticky update, expiry, published poses ----------------------------
  {
    TickTargetManager tick(models + "/model_uniform_acceleration_params.yaml");
    tick.quiet = true;
    tick.setExpirationTime(0.05);
    // an id the manager knows under ANOTHER model type (created by hand): its /tf records feed THAT target, by id
    MatrixXd Q, R, P;
    TargetManager::target_t uv;
    CHECK(tick.loadYamlFile(models + "/model_uniform_velocity_params.yaml", Q, R, P, uv));
    tick.init(uv, 7, dt, 0.0, Q, R, P, pose_at(1.0, 1.0, 1.0));
    const char* frames[5] = {"target_3", "camera_link", "target_7", "target_filt_3", "target_9"};   // the 4th breaks the loop: 9 never arrives
    std::vector<unsigned> erased;
    for (unsigned k = 0; k < 40; ++k) {
      const uint32_t sec = 100, nsec = k * 4000000u;
      if (k < 10) {
        const uint32_t s[5] = {sec, sec, sec, sec, sec}, ns[5] = {nsec, nsec, nsec, nsec, nsec};
        double poses[5 * 7];
        for (int r = 0; r < 5; ++r) {
          const Vector7d p = pose_at(0.1 * r + 0.01 * k, 0.2, 0.3);
          for (int e = 0; e < 7; ++e) poses[7 * r + e] = p[e];
        }
        tick.measurementCallBack(5, frames, s, ns, poses);
      }
      tick.tick(dt, sec, nsec, &erased);
      const std::vector<unsigned> ids = tick.getAvailableTargets();
      if (k < 10) {
        CHECK(ids.size() == 2 && ids[0] == 3 && ids[1] == 7 && erased.empty());
        CHECK(tick.publishedIds() == ids && tick.publishedPoses().size() == 14);
        CHECK(tick.mailboxCount() == 2);
        CHECK(tick.getNumberMeasurements(7) == (long long)k + 1);   // the by-hand uniform-velocity target is the one being fed
        int type = -1;
        CHECK(tick.typeOf(7, type) && type == (int)uv);
        CHECK(tick.typeOf(3, type) && type == (int)TargetManager::UNIFORM_ACCELERATION);
      }
      if (!erased.empty()) {
        CHECK(erased.size() == 2 && erased[0] == 3 && erased[1] == 7 && ids.empty());   // both fall silent at k = 10, expire together
        CHECK(k >= 21 && k <= 23);
      }
    }
    CHECK(tick.getAvailableTargets().empty() && tick.mailboxCount() == 0);
    NEAR(tick.time(), 40 * dt, 1e-9);
  }

  std::printf("%s: %d checks, %d failed\n", g_failed ? "FAILED" : "ok", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
