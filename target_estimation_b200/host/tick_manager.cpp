// tick_manager.cpp -- RosTargetManager's tick / mailbox / expiry semantics without ROS
// (/root/reference/include/target_estimation/target_manager_ros.hpp:74-134, src/target_manager_ros.cpp:6-107),
// driving the device pool with one batch per tick instead of one call per id.
#include <cstring>
#include <stdexcept>

#include "target_estimation_b200/target_manager.hpp"

namespace target_estimation_b200 {

double toSec(uint32_t sec, uint32_t nsec) {   // utils.hpp:59-62: (double)sec + 1e-9*(double)nsec, two roundings
  volatile double ns = 1e-9 * (double)nsec;
  return (double)sec + ns;
}

void Measurement::update(const StampedPose& tr) {   // target_manager_ros.hpp:96-115
  const double current_time_stamp = toSec(tr.sec, tr.nsec);
  const double prev_time_stamp = toSec(tr_.sec, tr_.nsec);
  if (current_time_stamp > prev_time_stamp) {   // new measurement
    new_meas_ = true;
    last_meas_time_ = current_time_stamp;
    acc_sec_ = tr.sec;
    acc_nsec_ = tr.nsec;
    stamp_dirty_ = true;
  } else {
    new_meas_ = false;
  }
  tr_ = tr;
}

TickTargetManager::TickTargetManager(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P, int device)
    : TargetManager(device), type_(type), Q_(Q), P_(P), R_(R), token_name_("target"), t_(0.0), expiration_time_(1000.0) {}   // :6-24

TickTargetManager::TickTargetManager(const std::string& yaml_file, int device)
    : TargetManager(device), type_(UNIFORM_VELOCITY), token_name_("target"), t_(0.0), expiration_time_(1000.0) {
  if (!loadYamlFile(yaml_file, Q_, R_, P_, type_)) throw std::runtime_error("Can not load the Cov Matrices!");   // :19-23
}

void TickTargetManager::setExpirationTime(double t) {
  if (!(t >= 0.0)) throw std::invalid_argument("expiration time must be >= 0 (assert of src/target_manager_ros.cpp:101)");
  expiration_time_ = t;
}

void TickTargetManager::measurementCallBack(long long n, const char* const* frames, const uint32_t* sec, const uint32_t* nsec,
                                            const double* poses) {   // :26-39
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  for (long long i = 0; i < n; ++i) {
    const std::string name(frames[i]);
    if (name.find(token_name_) != std::string::npos) {
      unsigned id;
      if (!getId(name, id)) break;   // e.g. the node's own "<token>_filt_<id>" output frames
      StampedPose tr;
      tr.sec = sec[i];
      tr.nsec = nsec[i];
      std::memcpy(tr.pose.data(), poses + 7 * i, 7 * sizeof(double));
      measurements_[id].update(tr);
    }
  }
}

void TickTargetManager::measurementCallBackIds(long long n, const unsigned* ids, const uint32_t* sec, const uint32_t* nsec,
                                               const double* poses) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  for (long long i = 0; i < n; ++i) {
    StampedPose tr;
    tr.sec = sec[i];
    tr.nsec = nsec[i];
    std::memcpy(tr.pose.data(), poses + 7 * i, 7 * sizeof(double));
    measurements_[ids[i]].update(tr);
  }
}

void TickTargetManager::tick(const double& dt, uint32_t now_sec, uint32_t now_nsec, std::vector<unsigned>* erased) {   // :41-92
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
  flushLocked();
  const size_t nm = measurements_.size();
  // 1. walk the mailboxes in ascending id order (std::map) and turn the per-id decisions into batches
  std::vector<unsigned> new_ids, step_ids, stamp_ids;
  std::vector<double> new_p0, step_meas;
  std::vector<uint8_t> step_act;
  std::vector<uint32_t> stamp_sec, stamp_nsec;
  new_ids.reserve(16); step_ids.reserve(nm); step_meas.reserve(nm * 7); step_act.reserve(nm);
  stamp_ids.reserve(nm); stamp_sec.reserve(nm); stamp_nsec.reserve(nm);
  for (auto& kv : measurements_) {
    const unsigned id = kv.first;
    StampedPose tr;
    const bool known = targets_.count(id) != 0;
    if (kv.second.read(tr)) {
      if (!known) {   // target does not exist: create it from the measurement (:54-58), then update with the same one
        new_ids.push_back(id);
        new_p0.insert(new_p0.end(), tr.pose.begin(), tr.pose.end());
      }
      step_ids.push_back(id);
      step_act.push_back((uint8_t)TE_ACT_UPDATE);
      step_meas.insert(step_meas.end(), tr.pose.begin(), tr.pose.end());
    } else if (known) {
      step_ids.push_back(id);
      step_act.push_back((uint8_t)TE_ACT_PREDICT);
      step_meas.insert(step_meas.end(), 7, 0.0);
    } else if (!quiet) {
      std::printf("Target(%u) does not exist!\n", id);   // TargetManager::update(id,dt) on an unknown id (:209)
    }
    if (kv.second.stampDirty() && (known || !new_ids.empty() && new_ids.back() == id)) {
      // the device keeps the same last_meas_time_ for the expiry predicate: push accepted stamps that changed
      stamp_ids.push_back(id);
      stamp_sec.push_back(kv.second.acceptedSec());
      stamp_nsec.push_back(kv.second.acceptedNsec());
      kv.second.clearStampDirty();
    }
  }
  // 2. init on first sight: p0 = the measurement, t0 = t_, v0 = a0 = 0 (:57)
  if (!new_ids.empty()) {
    std::vector<double> t0(new_ids.size(), t_);
    const bool q = quiet;
    quiet = true;   // one message per tick instead of one per target
    initBatch(type_, Q_, R_, P_, (long long)new_ids.size(), new_ids.data(), dt, t0.data(), new_p0.data());
    quiet = q;
  }
  // 3. one launch: update where a (possibly stale) measurement is readable, predict elsewhere (:59,:64)
  if (!step_ids.empty()) updateBatch((long long)step_ids.size(), step_ids.data(), dt, step_meas.data(), step_act.data());
  // 4. expiry: last_meas_time > 0 && (now - last_meas_time) >= expiration_time_, evaluated on the device (:67-72)
  te_pool* pool = poolOf((int)type_, false);
  std::vector<uint32_t> gone;
  if (pool && te_pool_size(pool) > 0) {
    if (!stamp_ids.empty() &&
        te_pool_set_stamps(pool, (long long)stamp_ids.size(), stamp_ids.data(), stamp_sec.data(), stamp_nsec.data()) < 0)
      throw std::runtime_error(te_last_error());
    gone.resize((size_t)te_pool_size(pool));
    long long n_gone = te_pool_expire(pool, now_sec, now_nsec, expiration_time_, gone.data(), (long long)gone.size());
    if (n_gone < 0) throw std::runtime_error(te_last_error());
    gone.resize((size_t)n_gone);
    for (uint32_t id : gone) {
      if (!quiet) std::printf("Timeout for target %u\n", id);
      measurements_.erase(id);
      targets_.erase(id);
    }
  }
  if (erased) erased->assign(gone.begin(), gone.end());
  // 5. the filtered poses the node broadcasts (:76-87)
  pub_ids_.clear();
  pub_poses_.clear();
  if (publish) {
    pub_ids_ = getAvailableTargets();
    pub_poses_.resize(pub_ids_.size() * 7);
    if (!pub_ids_.empty()) getEstimatesBatch((long long)pub_ids_.size(), pub_ids_.data(), nullptr, pub_poses_.data(), nullptr, nullptr, nullptr);
  }
  t_ = t_ + dt;   // :89
}

}  // namespace target_estimation_b200
