// target_manager.cpp -- host surface (TargetManager / TargetInterface / IntersectionSolver) over the
// device pools of include/te_pool.h.  Mirrors /root/reference/src/target_manager.cpp:18-295,
// src/target_interface.cpp:80-152 and src/intersection_solver.cpp:19-124 call for call; the arithmetic
// itself runs in the CUDA layer (no computation of filter state happens on the host).
#include "target_estimation_b200/target_manager.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <unordered_set>

namespace target_estimation_b200 {

namespace {

struct PoolError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void ck(long long rc) {
  if (rc < 0) throw PoolError(te_last_error());
}

// --- minimal reader for the reference's model files (models/*.yaml: `key: scalar` and
// `key: [v, v, ...]` lines; yaml-cpp is not available in this image) -------------------------
bool yamlLite(const std::string& file, std::map<std::string, std::string>& scalars, std::map<std::string, std::vector<double>>& lists) {
  std::ifstream in(file);
  if (!in) return false;
  std::stringstream ss;
  ss << in.rdbuf();
  std::string text = ss.str();
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    std::string line = text.substr(pos, eol - pos);
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    size_t colon = line.find(':');
    if (colon == std::string::npos) { pos = eol + 1; continue; }
    std::string key = line.substr(0, colon);
    key.erase(0, key.find_first_not_of(" \t"));
    key.erase(key.find_last_not_of(" \t") + 1);
    std::string val = line.substr(colon + 1);
    size_t lb = val.find('[');
    if (lb != std::string::npos) {
      // flow sequence, possibly spanning lines
      size_t abs_lb = pos + colon + 1 + lb;
      size_t rb = text.find(']', abs_lb);
      if (rb == std::string::npos) return false;
      std::string body = text.substr(abs_lb + 1, rb - abs_lb - 1);
      for (char& ch : body) if (ch == ',' || ch == '\n' || ch == '\r') ch = ' ';
      std::stringstream ls(body);
      std::vector<double> v;
      std::string tok;
      while (ls >> tok) {
        try { v.push_back(std::stod(tok)); } catch (...) { return false; }
      }
      lists[key] = v;
      eol = text.find('\n', rb);
      if (eol == std::string::npos) eol = text.size();
    } else {
      val.erase(0, val.find_first_not_of(" \t\"'"));
      size_t e = val.find_last_not_of(" \t\r\"'");
      val = (e == std::string::npos) ? std::string() : val.substr(0, e + 1);
      scalars[key] = val;
    }
    pos = eol + 1;
  }
  return true;
}

// parseSquareMatrix (src/target_manager.cpp:18-33): size = sqrt(len); Eigen::Map<MatrixXd> is COLUMN-major,
// so M(i,j) = v[i + s*j] (the transpose of the file's visual layout).
bool squareFromList(const std::map<std::string, std::vector<double>>& lists, const std::string& key, MatrixXd& M) {
  auto it = lists.find(key);
  if (it == lists.end()) return false;
  const std::vector<double>& v = it->second;
  const unsigned s = static_cast<unsigned>(std::sqrt((double)v.size()));
  if (s == 0) return false;
  M = MatrixXd((int)s, (int)s);
  for (unsigned i = 0; i < s; ++i)
    for (unsigned j = 0; j < s; ++j) M((int)i, (int)j) = v[i + (size_t)s * j];
  return true;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// TargetManager
// ----------------------------------------------------------------------------------------------
TargetManager::TargetManager(int device) : device_(device) {}

TargetManager::TargetManager(const std::string& file, int device) : TargetManager(device) {
  if (!loadYamlFile(file, default_Q_, default_R_, default_P_, default_type_))
    throw "TargetManager default constructor failed!";
  default_values_loaded_ = true;
}

TargetManager::~TargetManager() {
  for (te_pool*& p : pools_) {
    if (p) te_pool_destroy(p);
    p = nullptr;
  }
}

bool TargetManager::selectTargetType(const std::string& s, target_t& type) {   // src/target_manager.cpp:52-65
  if (s == "angular_rates") type = ANGULAR_RATES;
  else if (s == "angular_velocities") type = ANGULAR_VELOCITIES;
  else if (s == "uniform_acceleration") type = UNIFORM_ACCELERATION;
  else if (s == "uniform_velocity") type = UNIFORM_VELOCITY;
  else return false;
  return true;
}

bool TargetManager::loadYamlFile(const std::string& file, MatrixXd& Q, MatrixXd& R, MatrixXd& P, target_t& type) {   // :67-104
  std::map<std::string, std::string> scalars;
  std::map<std::string, std::vector<double>> lists;
  bool success = true;
  if (!yamlLite(file, scalars, lists)) {
    std::cerr << "Can not parse file: " << file << std::endl;
    return false;
  }
  if (!squareFromList(lists, "Q", Q)) { std::cerr << "Can not load matrix Q from file: " << file << std::endl; success = false; }
  if (!squareFromList(lists, "R", R)) { std::cerr << "Can not load matrix R from file: " << file << std::endl; success = false; }
  if (!squareFromList(lists, "P", P)) { std::cerr << "Can not load matrix P from file: " << file << std::endl; success = false; }
  auto it = scalars.find("type");
  if (it == scalars.end()) {
    std::cerr << "Can not load type from file: " << file << std::endl;
    success = false;
  } else if (!selectTargetType(it->second, type)) {
    std::cerr << "Can not parse type: " << it->second << std::endl;   // the reference keeps going with an unset type (:40-41)
    success = false;
  }
  return success;
}

te_pool* TargetManager::poolOf(int type, bool create) {
  if (type < 0 || type > 3) return nullptr;
  if (!pools_[type] && create) {
    pools_[type] = te_pool_create(type, device_, nullptr);
    if (!pools_[type]) throw PoolError(te_last_error());
  }
  return pools_[type];
}

// ---- IdRegistry ------------------------------------------------------------------------------------------
IdRegistry::iterator IdRegistry::find(unsigned id) {
  auto it = std::lower_bound(v_.begin(), v_.end(), id, [](const value_type& a, unsigned key) { return a.first < key; });
  return (it != v_.end() && it->first == id) ? it : v_.end();
}
uint8_t& IdRegistry::operator[](unsigned id) {
  if (v_.empty() || v_.back().first < id) {   // ascending arrivals (track ids grow): append
    v_.emplace_back(id, (uint8_t)0);
    return v_.back().second;
  }
  auto it = std::lower_bound(v_.begin(), v_.end(), id, [](const value_type& a, unsigned key) { return a.first < key; });
  if (it == v_.end() || it->first != id) it = v_.insert(it, value_type(id, (uint8_t)0));
  return it->second;
}
size_t IdRegistry::erase(unsigned id) {
  auto it = find(id);
  if (it == v_.end()) return 0;
  v_.erase(it);
  return 1;
}
void IdRegistry::mergeSorted(const std::vector<value_type>& add) {
  if (add.empty()) return;
  if (v_.empty() || v_.back().first < add.front().first) {   // append-only batch
    v_.insert(v_.end(), add.begin(), add.end());
    return;
  }
  std::vector<value_type> out;
  out.reserve(v_.size() + add.size());
  size_t i = 0, j = 0;
  while (i < v_.size() && j < add.size()) {
    if (v_[i].first < add[j].first) out.push_back(v_[i++]);
    else if (add[j].first < v_[i].first) out.push_back(add[j++]);
    else { out.push_back(v_[i++]); ++j; }   // existing entry wins
  }
  out.insert(out.end(), v_.begin() + (long)i, v_.end());
  out.insert(out.end(), add.begin() + (long)j, add.end());
  v_.swap(out);
}
void IdRegistry::insertSorted(const uint32_t* ids, size_t n, uint8_t type) {
  std::vector<value_type> add;
  add.reserve(n);
  for (size_t k = 0; k < n; ++k) add.emplace_back(ids[k], type);
  mergeSorted(add);
}
void IdRegistry::eraseSorted(const uint32_t* ids, size_t n) {
  if (n == 0 || v_.empty()) return;
  size_t w = 0, j = 0;
  auto first = std::lower_bound(v_.begin(), v_.end(), ids[0], [](const value_type& a, unsigned key) { return a.first < key; });
  w = (size_t)(first - v_.begin());
  for (size_t i = w; i < v_.size(); ++i) {
    while (j < n && ids[j] < v_[i].first) ++j;
    if (j < n && ids[j] == v_[i].first) continue;   // dropped
    v_[w++] = v_[i];
  }
  v_.resize(w);
}

int TargetManager::registerClass(int type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0) {
  int n = 0, m = 0;
  te_model_dims(type, &n, &m);
  // asserts of the model constructors (e.g. src/types/uniform_acceleration.cpp:35-37) + the hard-wired
  // measurement sizes of addMeasurement (m must be 3 / 6)
  if (Q.rows() != n || Q.cols() != n || P0.rows() != n || P0.cols() != n || R.rows() != m || R.cols() != m)
    throw std::invalid_argument("model matrices do not match the target type (n=" + std::to_string(n) + ", m=" + std::to_string(m) + ")");
  te_pool* p = poolOf(type, true);
  int cls = te_pool_register_class(p, Q.data(), R.data(), P0.data());
  ck(cls);
  return cls;
}

bool TargetManager::typeOf(unsigned id, int& type) {
  auto it = targets_.find(id);
  if (it == targets_.end()) return false;
  type = it->second;
  return true;
}

void TargetManager::init(const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0, const Vector6d& v0,
                         const Vector6d& a0) {   // :135-142
  if (default_values_loaded_) init(default_type_, id, dt0, t0, default_Q_, default_R_, default_P_, p0, v0, a0);
  else throw "TargetManager::init failed, can not find default values to load!";
}

void TargetManager::init(const target_t& type, const unsigned int& id, const double& dt0, const double& t0, const MatrixXd& Q,
                         const MatrixXd& R, const MatrixXd& P0, const Vector7d& p0, const Vector6d& v0, const Vector6d& a0) {   // :144-179
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (targets_.find(id) != targets_.end()) {
    if (!quiet) std::cout << "Target(" << id << ") already exists!" << std::endl;
    return;
  }
  if (!(dt0 >= 0.0) || !(t0 >= 0.0)) throw std::invalid_argument("dt0 and t0 must be >= 0 (asserts of src/target_interface.cpp:20)");
  flushLocked();
  const int cls = registerClass((int)type, Q, R, P0);
  const uint32_t uid = id;
  const uint16_t c16 = (uint16_t)cls;
  ck(te_pool_add_batch(poolOf((int)type, true), 1, &uid, &c16, &t0, p0.data(), v0.data(), a0.data(), nullptr));
  targets_[id] = (uint8_t)type;
  if (!quiet) {
    static const char* msg[4] = {"Using angular rates for the orientation", "Using angular velocities for the orientation",
                                 "Uniformly accelerated motion", "Uniform rectilinear motion"};
    std::cout << msg[(int)type] << std::endl;
  }
}

void TargetManager::init(const std::string& file, const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0,
                         const Vector6d& v0, const Vector6d& a0) {   // :181-188
  MatrixXd Q, P, R;
  target_t type = UNIFORM_VELOCITY;
  if (!loadYamlFile(file, Q, R, P, type)) throw std::invalid_argument("cannot load model file " + file);
  init(type, id, dt0, t0, Q, R, P, p0, v0, a0);
}

long long TargetManager::initBatch(long long n, const unsigned* ids, double dt0, const double* t0, const double* p0, const double* v0,
                                   const double* a0) {
  if (!default_values_loaded_) throw "TargetManager::init failed, can not find default values to load!";
  return initBatch(default_type_, default_Q_, default_R_, default_P_, n, ids, dt0, t0, p0, v0, a0, nullptr);
}

long long TargetManager::initBatch(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0, long long n, const unsigned* ids,
                                   double dt0, const double* t0, const double* p0, const double* v0, const double* a0,
                                   const double* p0_scale) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (n <= 0) return 0;
  if (!(dt0 >= 0.0)) throw std::invalid_argument("dt0 must be >= 0");
  flushLocked();
  const int cls = registerClass((int)type, Q, R, P0);
  // ids already known to the manager under ANY type are skipped ("already exists"); first occurrence wins
  std::vector<long long> keep;
  keep.reserve((size_t)n);
  std::vector<std::pair<unsigned, uint8_t>> fresh;   // ascending (id, type) pairs for the registry
  bool ascending = true;
  for (long long k = 1; k < n && ascending; ++k) ascending = ids[k - 1] < ids[k];
  if (ascending) {
    // the usual batch (track ids grow): no id repeats inside it, and one walk along the sorted registry finds the known ones
    fresh.reserve((size_t)n);
    auto it = targets_.begin();
    for (long long k = 0; k < n; ++k) {
      while (it != targets_.end() && it->first < ids[k]) ++it;
      if (it != targets_.end() && it->first == ids[k]) {
        if (!quiet) std::cout << "Target(" << ids[k] << ") already exists!" << std::endl;
        continue;
      }
      fresh.emplace_back(ids[k], (uint8_t)type);
      keep.push_back(k);
    }
  } else {
    std::map<unsigned, uint8_t> seen;
    for (long long k = 0; k < n; ++k) {
      if (targets_.count(ids[k]) || seen.count(ids[k])) {
        if (!quiet) std::cout << "Target(" << ids[k] << ") already exists!" << std::endl;
        continue;
      }
      seen[ids[k]] = (uint8_t)type;
      keep.push_back(k);
    }
    fresh.assign(seen.begin(), seen.end());
  }
  if (keep.empty()) return 0;
  const long long na = (long long)keep.size();
  std::vector<uint32_t> g_ids((size_t)na);
  std::vector<uint16_t> g_cls((size_t)na, (uint16_t)cls);
  std::vector<double> g_t0, g_p0((size_t)na * 7), g_v0, g_a0, g_sc;
  auto gather = [&](const double* src, int w, std::vector<double>& dst) -> const double* {
    if (!src) return nullptr;
    dst.resize((size_t)na * w);
    for (long long k = 0; k < na; ++k) std::memcpy(&dst[(size_t)k * w], src + (size_t)keep[k] * w, sizeof(double) * w);
    return dst.data();
  };
  for (long long k = 0; k < na; ++k) g_ids[k] = ids[keep[k]];
  const bool all = na == n;
  const double* s_p0 = all ? p0 : gather(p0, 7, g_p0);
  const double* s_t0 = all ? t0 : gather(t0, 1, g_t0);
  const double* s_v0 = all ? v0 : gather(v0, 6, g_v0);
  const double* s_a0 = all ? a0 : gather(a0, 6, g_a0);
  const double* s_sc = all ? p0_scale : gather(p0_scale, 1, g_sc);
  long long added = te_pool_add_batch(poolOf((int)type, true), na, g_ids.data(), g_cls.data(), s_t0, s_p0, s_v0, s_a0, s_sc);
  ck(added);
  targets_.insert(fresh.begin(), fresh.end());
  return added;
}

void TargetManager::queue(int type, unsigned id, double dt, const double* meas, int action) {
  if (pending_ids_.count(id)) flushLocked();   // an id may appear once per launch: keep call order observable
  Pending& q = pending_[type];
  q.ids.push_back(id);
  q.dt.push_back(dt);
  q.action.push_back((uint8_t)action);
  const size_t o = q.meas.size();
  q.meas.resize(o + 7, 0.0);
  if (meas) std::memcpy(&q.meas[o], meas, 7 * sizeof(double));
  pending_ids_[id] = (uint8_t)type;
}

void TargetManager::flushLocked() {
  if (pending_ids_.empty()) return;
  for (int t = 0; t < 4; ++t) {
    Pending& q = pending_[t];
    if (q.ids.empty()) continue;
    long long rc = te_pool_step_ids(pools_[t], (long long)q.ids.size(), q.ids.data(), q.dt.data(), 0.0, q.meas.data(), q.action.data());
    q.ids.clear(); q.dt.clear(); q.meas.clear(); q.action.clear();
    if (rc < 0) {
      pending_ids_.clear();
      throw PoolError(te_last_error());
    }
  }
  pending_ids_.clear();
}

void TargetManager::flush() {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
}

bool TargetManager::update(const unsigned int& id, const double& dt, const Vector7d& meas) {   // :190-202
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0 (assert of src/target_interface.cpp:150)");
  queue(type, id, dt, meas.data(), TE_ACT_UPDATE);
  return true;
}

bool TargetManager::update(const unsigned int& id, const double& dt) {   // :204-218
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
  queue(type, id, dt, nullptr, TE_ACT_PREDICT);
  return true;
}

void TargetManager::update(const double& dt) {   // :220-225
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
  for (int t = 0; t < 4; ++t)
    if (pools_[t]) ck(te_pool_predict_all(pools_[t], dt));
}

long long TargetManager::updateBatch(long long n, const unsigned* ids, double dt, const double* meas, const unsigned char* action) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (n <= 0) return 0;
  flushLocked();
  // te_pool_step_ids applies one op per id per launch.  A batch that names an id again stands for the reference's sequential
  // update() calls, which apply every record in order: cut the batch in front of each repeat and launch the pieces one after
  // the other (strictly ascending ids -- the usual batch -- cannot repeat and skip the hash set).
  bool ascending = true;
  for (long long k = 1; k < n && ascending; ++k) ascending = ids[k - 1] < ids[k];
  if (ascending) return updateBatchUnique(n, ids, dt, meas, action);
  {
    // ids in arbitrary order (a /tf message lists its targets as they come): a homogeneous manager hands the batch over as it is --
    // the device finds a repeated id while it looks the ids up, before anything is stepped (te_pool_step_ids returns -2) -- and
    // only then the batch is cut on the host (a hash set of four million ids costs 50 x the tick it guards)
    int n_pools = 0, only = -1;
    for (int t = 0; t < 4; ++t)
      if (pools_[t] && te_pool_size(pools_[t]) > 0) { ++n_pools; only = t; }
    if (n_pools == 0) return 0;
    if (n_pools == 1) {
      const long long rc = te_pool_step_ids(pools_[only], n, ids, nullptr, dt, meas, action);
      if (rc != -2) {
        ck(rc);
        return rc;
      }
    }
  }
  long long applied = 0, start = 0;
  std::unordered_set<unsigned> seen;
  seen.reserve((size_t)std::min<long long>(n, 1 << 22));
  for (long long k = 0; k < n; ++k) {
    if (!seen.insert(ids[k]).second) {
      applied += updateBatchUnique(k - start, ids + start, dt, meas ? meas + 7 * (size_t)start : nullptr, action ? action + start : nullptr);
      start = k;
      seen.clear();
      seen.insert(ids[k]);
    }
  }
  return applied + updateBatchUnique(n - start, ids + start, dt, meas ? meas + 7 * (size_t)start : nullptr, action ? action + start : nullptr);
}

long long TargetManager::updateBatchUnique(long long n, const unsigned* ids, double dt, const double* meas, const unsigned char* action) {
  if (n <= 0) return 0;
  int n_pools = 0, only = -1;
  for (int t = 0; t < 4; ++t)
    if (pools_[t] && te_pool_size(pools_[t]) > 0) { ++n_pools; only = t; }
  if (n_pools == 0) return 0;
  if (n_pools == 1) {   // homogeneous manager: the batch goes to the device as it is
    long long rc = te_pool_step_ids(pools_[only], n, ids, nullptr, dt, meas, action);
    ck(rc);
    return rc;
  }
  // mixed manager: split the batch by model type on the host
  Pending split[4];
  for (long long k = 0; k < n; ++k) {
    auto it = targets_.find(ids[k]);
    if (it == targets_.end()) continue;
    Pending& q = split[it->second];
    q.ids.push_back(ids[k]);
    q.action.push_back(action ? action[k] : (uint8_t)TE_ACT_UPDATE);
    const size_t o = q.meas.size();
    q.meas.resize(o + 7, 0.0);
    if (meas) std::memcpy(&q.meas[o], meas + 7 * (size_t)k, 7 * sizeof(double));
  }
  long long applied = 0;
  for (int t = 0; t < 4; ++t) {
    if (split[t].ids.empty()) continue;
    long long rc = te_pool_step_ids(pools_[t], (long long)split[t].ids.size(), split[t].ids.data(), nullptr, dt, split[t].meas.data(),
                                    split[t].action.data());
    ck(rc);
    applied += rc;
  }
  return applied;
}

te_pool* TargetManager::densePool() {
  int n_pools = 0, only = -1;
  for (int t = 0; t < 4; ++t)
    if (pools_[t] && te_pool_size(pools_[t]) > 0) { ++n_pools; only = t; }
  if (n_pools != 1) throw std::invalid_argument("a dense tick needs a manager whose targets share one model type");
  return pools_[only];
}

long long TargetManager::updateDense(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
  if (targets_.empty()) return 0;
  te_pool* p = densePool();
  ck(te_pool_tick_host(p, dt, meas, meas_stride, action, TE_ACT_UPDATE, est_pos_out));
  return te_pool_size(p);
}

long long TargetManager::updateDenseAsync(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
  if (targets_.empty()) return 0;
  te_pool* p = densePool();
  ck(te_pool_tick_host_async(p, dt, meas, meas_stride, action, TE_ACT_UPDATE, est_pos_out));
  return te_pool_size(p);
}

void TargetManager::updateDenseWait(int lag) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  for (int t = 0; t < 4; ++t)
    if (pools_[t]) ck(te_pool_tick_host_wait(pools_[t], lag));
}

std::vector<unsigned int> TargetManager::denseIds() { return getAvailableTargets(); }

size_t TargetManager::size() {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  return targets_.size();
}

bool TargetManager::erase(const unsigned int& id) {   // :227-241
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  auto it = targets_.find(id);
  if (it == targets_.end()) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return false;
  }
  flushLocked();
  const uint32_t uid = id;
  ck(te_pool_erase_batch(pools_[it->second], 1, &uid));
  targets_.erase(it);
  return true;
}

long long TargetManager::eraseBatch(long long n, const unsigned* ids) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
  std::vector<uint32_t> per[4], found;
  for (long long k = 0; k < n; ++k) {
    auto it = targets_.find(ids[k]);
    if (it == targets_.end()) continue;
    per[it->second].push_back(ids[k]);
    found.push_back(ids[k]);
  }
  std::sort(found.begin(), found.end());
  targets_.eraseSorted(found.data(), found.size());
  long long erased = 0;
  for (int t = 0; t < 4; ++t) {
    if (per[t].empty()) continue;
    long long rc = te_pool_erase_batch(pools_[t], (long long)per[t].size(), per[t].data());
    ck(rc);
    erased += rc;
  }
  return erased;
}

TargetInterface::Ptr TargetManager::getTarget(const unsigned int& id) {   // :243-250
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (targets_.count(id) != 0) return std::make_shared<TargetInterface>(this, id);
  return nullptr;
}

void TargetManager::getEstimatesBatch(long long n, const unsigned* ids, const double* t1, double* pose7, double* twist6, double* acc6,
                                      unsigned char* found) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  if (n <= 0) return;
  flushLocked();
  if (found) std::memset(found, 0, (size_t)n);
  int n_pools = 0, only = -1;
  for (int t = 0; t < 4; ++t)
    if (pools_[t] && te_pool_size(pools_[t]) > 0) { ++n_pools; only = t; }
  if (n_pools == 0) return;
  if (n_pools == 1) {
    ck(te_pool_read_estimates(pools_[only], n, ids, t1, pose7, twist6, acc6, nullptr, found));
    return;
  }
  for (int t = 0; t < 4; ++t) {
    if (!pools_[t]) continue;
    std::vector<uint32_t> sub;
    std::vector<long long> where;
    std::vector<double> st1;
    for (long long k = 0; k < n; ++k) {
      auto it = targets_.find(ids[k]);
      if (it == targets_.end() || it->second != t) continue;
      sub.push_back(ids[k]);
      where.push_back(k);
      if (t1) st1.push_back(t1[k]);
    }
    if (sub.empty()) continue;
    const long long m = (long long)sub.size();
    std::vector<double> po((size_t)m * 7), tw((size_t)m * 6), ac((size_t)m * 6);
    std::vector<uint8_t> fo((size_t)m);
    ck(te_pool_read_estimates(pools_[t], m, sub.data(), t1 ? st1.data() : nullptr, po.data(), tw.data(), ac.data(), nullptr, fo.data()));
    for (long long j = 0; j < m; ++j) {
      const long long k = where[j];
      if (pose7) std::memcpy(pose7 + 7 * k, &po[7 * j], 56);
      if (twist6) std::memcpy(twist6 + 6 * k, &tw[6 * j], 48);
      if (acc6) std::memcpy(acc6 + 6 * k, &ac[6 * j], 48);
      if (found) found[k] = fo[j];
    }
  }
}

bool TargetManager::getTargetPose(const unsigned int& id, Vector7d& pose) {   // :252-261
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) return false;
  flushLocked();
  const uint32_t uid = id;
  ck(te_pool_read_estimates(pools_[type], 1, &uid, nullptr, pose.data(), nullptr, nullptr, nullptr, nullptr));
  return true;
}

bool TargetManager::getTargetTwist(const unsigned int& id, Vector6d& twist) {   // :263-272
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) return false;
  flushLocked();
  const uint32_t uid = id;
  ck(te_pool_read_estimates(pools_[type], 1, &uid, nullptr, nullptr, twist.data(), nullptr, nullptr, nullptr));
  return true;
}

bool TargetManager::getTargetAcceleration(const unsigned int& id, Vector6d& acc) {   // :274-283
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) return false;
  flushLocked();
  const uint32_t uid = id;
  ck(te_pool_read_estimates(pools_[type], 1, &uid, nullptr, nullptr, nullptr, acc.data(), nullptr, nullptr));
  return true;
}

long long TargetManager::getNumberMeasurements(const unsigned int& id) {   // :285-295
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  int type;
  if (!typeOf(id, type)) {
    if (!quiet) std::cout << "Target(" << id << ") does not exist!" << std::endl;
    return 0;
  }
  flushLocked();
  const uint32_t uid = id;
  long long nm = 0;
  ck(te_pool_read_state(pools_[type], 1, &uid, nullptr, nullptr, nullptr, &nm, nullptr, nullptr));
  return nm;
}

// ---- sampled logging (the reference's rt_logger publishers + writeTxtFile dumps) ----------------------------
void TargetManager::watch(long long n, const unsigned* ids, size_t max_samples) {
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  log_ids_.assign(ids, ids + (n > 0 ? n : 0));
  log_cap_ = max_samples;
  log_t_.clear(); log_rows_.clear(); log_P_.clear(); log_n_.clear();
}

void TargetManager::log() {   // :119-123: every target publishes measurement / pose / twist / acceleration / covariance
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  flushLocked();
  const size_t m = log_ids_.size();
  if (m == 0 || log_t_.size() >= log_cap_) return;
  const size_t base = log_t_.size();
  log_rows_.resize((base + 1) * m * 26, 0.0);
  log_P_.resize((base + 1) * m * 324, 0.0);
  log_n_.resize((base + 1) * m, 0);
  double t_sample = std::nan("");
  // one batched read-back per pool that holds watched ids
  for (int type = 0; type < 4; ++type) {
    if (!pools_[type]) continue;
    std::vector<uint32_t> sub;
    std::vector<size_t> where;
    for (size_t j = 0; j < m; ++j) {
      auto it = targets_.find(log_ids_[j]);
      if (it != targets_.end() && it->second == type) { sub.push_back(log_ids_[j]); where.push_back(j); }
    }
    if (sub.empty()) continue;
    int N = 0, M = 0;
    te_model_dims(type, &N, &M);
    const long long k = (long long)sub.size();
    std::vector<double> t((size_t)k), P((size_t)k * N * N), mp((size_t)k * 7), tw((size_t)k * 6), ac((size_t)k * 6), pi((size_t)k * 6);
    ck(te_pool_read_state(pools_[type], k, sub.data(), nullptr, P.data(), t.data(), nullptr, nullptr, mp.data()));
    ck(te_pool_read_estimates(pools_[type], k, sub.data(), nullptr, nullptr, tw.data(), ac.data(), pi.data(), nullptr));
    for (long long q = 0; q < k; ++q) {
      const size_t j = where[(size_t)q];
      double* row = &log_rows_[(base * m + j) * 26];
      row[0] = t[(size_t)q];
      std::memcpy(row + 1, &mp[(size_t)q * 7], 7 * sizeof(double));
      std::memcpy(row + 8, &pi[(size_t)q * 6], 6 * sizeof(double));
      std::memcpy(row + 14, &tw[(size_t)q * 6], 6 * sizeof(double));
      std::memcpy(row + 20, &ac[(size_t)q * 6], 6 * sizeof(double));
      std::memcpy(&log_P_[(base * m + j) * 324], &P[(size_t)q * N * N], sizeof(double) * (size_t)N * N);
      log_n_[base * m + j] = N;
      if (std::isnan(t_sample)) t_sample = t[(size_t)q];
    }
  }
  log_t_.push_back(t_sample);
}

bool TargetManager::logSample(size_t k, size_t j, double* row26, double* P, int* n_state) const {
  const size_t m = log_ids_.size();
  if (k >= log_t_.size() || j >= m) return false;
  const int N = log_n_[k * m + j];
  if (n_state) *n_state = N;
  if (row26) std::memcpy(row26, &log_rows_[(k * m + j) * 26], 26 * sizeof(double));
  if (P && N > 0) std::memcpy(P, &log_P_[(k * m + j) * 324], sizeof(double) * (size_t)N * N);
  return N > 0;
}

bool writeTxtFile(const std::string& filename, const double* values, size_t rows, size_t cols) {   // utils.hpp:78-120
  std::ofstream f(filename.c_str());
  if (!f.is_open()) {
    std::cerr << "Unable to open file : [" << filename << "]" << std::endl;
    return false;
  }
  for (size_t r = 0; r < rows; ++r) {
    if (cols == 1) {   // the VectorXd overload: "value\n"
      f << values[r] << "\n";
    } else {           // the MatrixXd overload: "value " per column, then "\n"
      for (size_t c = 0; c < cols; ++c) f << values[r * cols + c] << " ";
      f << "\n";
    }
  }
  return true;
}

int TargetManager::writeLog(const std::string& folder) const {
  const size_t m = log_ids_.size(), S = log_t_.size();
  int files = 0;
  for (size_t j = 0; j < m; ++j) {
    std::vector<double> time, meas, pose, twist, acc, diag;
    size_t n_diag = 0;
    for (size_t k = 0; k < S; ++k) {
      const int N = log_n_[k * m + j];
      if (N == 0) continue;   // the id did not exist at that sample
      const double* row = &log_rows_[(k * m + j) * 26];
      time.push_back(row[0]);
      meas.insert(meas.end(), row + 1, row + 8);
      pose.insert(pose.end(), row + 8, row + 14);
      twist.insert(twist.end(), row + 14, row + 20);
      acc.insert(acc.end(), row + 20, row + 26);
      n_diag = (size_t)N;
      for (int d = 0; d < N; ++d) diag.push_back(log_P_[(k * m + j) * 324 + (size_t)d * N + d]);
    }
    if (time.empty()) continue;
    const std::string id = std::to_string(log_ids_[j]);
    files += writeTxtFile(folder + "time_" + id, time.data(), time.size(), 1);
    files += writeTxtFile(folder + "meas_pose_" + id, meas.data(), time.size(), 7);
    files += writeTxtFile(folder + "est_pose_" + id, pose.data(), time.size(), 6);
    files += writeTxtFile(folder + "est_twist_" + id, twist.data(), time.size(), 6);
    files += writeTxtFile(folder + "est_acc_" + id, acc.data(), time.size(), 6);
    files += writeTxtFile(folder + "cov_diag_" + id, diag.data(), time.size(), n_diag);
  }
  return files;
}

std::vector<unsigned int> TargetManager::getAvailableTargets() {   // :125-133
  std::lock_guard<std::recursive_mutex> lg(target_lock_);
  std::vector<unsigned int> ids;
  ids.reserve(targets_.size());
  for (auto const& kv : targets_) ids.push_back(kv.first);
  return ids;
}

// ----------------------------------------------------------------------------------------------
// TargetInterface proxy / EstimatorView
// ----------------------------------------------------------------------------------------------
namespace {
struct SlotRead {
  te_pool* pool = nullptr;
  int type = -1;
  int n = 0, m = 0;
};
bool locate(TargetManager* mgr, unsigned id, SlotRead& s) {
  if (!mgr->typeOf(id, s.type)) return false;
  mgr->flush();
  s.pool = mgr->poolOf(s.type, false);
  te_model_dims(s.type, &s.n, &s.m);
  return s.pool != nullptr;
}
}  // namespace

VectorXd EstimatorView::getState() const {
  std::lock_guard<std::recursive_mutex> lg(mgr_->lock());
  SlotRead s;
  if (!locate(mgr_, id_, s)) return VectorXd();
  VectorXd x((size_t)s.n, 0.0);
  const uint32_t uid = id_;
  ck(te_pool_read_state(s.pool, 1, &uid, x.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
  return x;
}
MatrixXd EstimatorView::getP() const {
  std::lock_guard<std::recursive_mutex> lg(mgr_->lock());
  SlotRead s;
  if (!locate(mgr_, id_, s)) return MatrixXd();
  MatrixXd P(s.n, s.n);
  const uint32_t uid = id_;
  ck(te_pool_read_state(s.pool, 1, &uid, nullptr, P.d.data(), nullptr, nullptr, nullptr, nullptr));
  return P;
}
namespace {
MatrixXd classMatrix(TargetManager* mgr, unsigned id, int which) {
  std::lock_guard<std::recursive_mutex> lg(mgr->lock());
  SlotRead s;
  if (!locate(mgr, id, s)) return MatrixXd();
  const uint32_t uid = id;
  int cls = te_pool_class_of(s.pool, uid);
  ck(cls);
  MatrixXd Q(s.n, s.n), R(s.m, s.m), P0(s.n, s.n);
  ck(te_pool_get_class(s.pool, cls, Q.d.data(), R.d.data(), P0.d.data()));
  return which == 0 ? Q : (which == 1 ? R : P0);
}
}  // namespace
MatrixXd EstimatorView::getQ() const { return classMatrix(mgr_, id_, 0); }
MatrixXd EstimatorView::getR() const { return classMatrix(mgr_, id_, 1); }
MatrixXd EstimatorView::getP0() const { return classMatrix(mgr_, id_, 2); }

void TargetInterface::addMeasurement(const double& dt, const Vector7d& meas) { mgr_->update(id_, dt, meas); }
void TargetInterface::update(const double& dt) { mgr_->update(id_, dt); }

namespace {
template <int W> std::array<double, W> readEst(TargetManager* mgr, unsigned id, const double* t1, int which) {
  std::array<double, W> out{};
  std::lock_guard<std::recursive_mutex> lg(mgr->lock());
  SlotRead s;
  if (!locate(mgr, id, s)) return out;
  const uint32_t uid = id;
  ck(te_pool_read_estimates(s.pool, 1, &uid, t1, which == 0 ? out.data() : nullptr, which == 1 ? out.data() : nullptr,
                            which == 2 ? out.data() : nullptr, nullptr, nullptr));
  return out;
}
}  // namespace
Vector7d TargetInterface::getEstimatedPose() const { return readEst<7>(mgr_, id_, nullptr, 0); }
Vector6d TargetInterface::getEstimatedTwist() const { return readEst<6>(mgr_, id_, nullptr, 1); }
Vector6d TargetInterface::getEstimatedAcceleration() const { return readEst<6>(mgr_, id_, nullptr, 2); }
Vector7d TargetInterface::getEstimatedPose(const double& t1) const { return readEst<7>(mgr_, id_, &t1, 0); }
Vector6d TargetInterface::getEstimatedTwist(const double& t1) const { return readEst<6>(mgr_, id_, &t1, 1); }
Vector6d TargetInterface::getEstimatedAcceleration(const double& t1) const { return readEst<6>(mgr_, id_, &t1, 2); }

Vector7d TargetInterface::getMeasuredPose() const {
  Vector7d out{};
  std::lock_guard<std::recursive_mutex> lg(mgr_->lock());
  SlotRead s;
  if (!locate(mgr_, id_, s)) return out;
  const uint32_t uid = id_;
  ck(te_pool_read_state(s.pool, 1, &uid, nullptr, nullptr, nullptr, nullptr, nullptr, out.data()));
  return out;
}
double TargetInterface::getTime() const {
  double t = 0.0;
  std::lock_guard<std::recursive_mutex> lg(mgr_->lock());
  SlotRead s;
  if (!locate(mgr_, id_, s)) return t;
  const uint32_t uid = id_;
  ck(te_pool_read_state(s.pool, 1, &uid, nullptr, nullptr, &t, nullptr, nullptr, nullptr));
  return t;
}
long long TargetInterface::getNumberMeasurements() const {
  long long nm = 0;
  std::lock_guard<std::recursive_mutex> lg(mgr_->lock());
  SlotRead s;
  if (!locate(mgr_, id_, s)) return nm;
  const uint32_t uid = id_;
  ck(te_pool_read_state(s.pool, 1, &uid, nullptr, nullptr, nullptr, &nm, nullptr, nullptr));
  return nm;
}
double TargetInterface::getPeriodEstimate() const {   // src/target_interface.cpp:80-87
  const Vector6d tw = getEstimatedTwist();
  const double n = std::sqrt(tw[3] * tw[3] + tw[4] * tw[4] + tw[5] * tw[5]);
  if (n > 0) return 2 * M_PI / n;
  return -1.0;
}

// ----------------------------------------------------------------------------------------------
// MovingAvgFilter (utils.hpp:222-251)
// ----------------------------------------------------------------------------------------------
double MovingAvgFilter::update(double value) {
  const unsigned n = (unsigned)window_.size();
  sum_ -= window_[idx_];
  sum_ += value;
  window_[idx_] = value;
  if (!complete_ && idx_ == n - 1) complete_ = true;
  const unsigned num = complete_ ? n : idx_ + 1;
  const double res = sum_ / num;
  idx_ = (idx_ + 1) % n;
  variance_ = 0.0;
  for (unsigned k = 0; k < n; ++k) variance_ += (window_[k] - res) * (window_[k] - res);
  variance_ /= num;
  return res;
}

// ----------------------------------------------------------------------------------------------
// IntersectionSolver (src/intersection_solver.cpp:19-124)
// ----------------------------------------------------------------------------------------------
IntersectionSolver::IntersectionSolver(TargetManager::Ptr tm, const unsigned int filters_length)
    : target_manager_(tm), filters_length_(filters_length) {
  if (!tm) throw std::invalid_argument("IntersectionSolver needs a target manager (assert of src/intersection_solver.cpp:21)");
}
IntersectionSolver::~IntersectionSolver() {
  for (te_isolver*& s : solvers_) {
    if (s) te_isolver_destroy(s);
    s = nullptr;
  }
}
te_isolver* IntersectionSolver::solverFor(int type) {
  if (!solvers_[type]) {
    solvers_[type] = te_isolver_create(target_manager_->poolOf(type, false), 1, filters_length_);
    if (!solvers_[type]) throw PoolError(te_last_error());
  }
  return solvers_[type];
}
double IntersectionSolver::getIntersectionTimeWithSphere(const unsigned int& id, const double& t1, const Vector3d& origin,
                                                         const double& radius) {
  std::lock_guard<std::recursive_mutex> lg(target_manager_->lock());
  int type;
  if (!target_manager_->typeOf(id, type)) return -1;   // :44,:87-88
  target_manager_->flush();
  const uint32_t uid = id;
  double delta = -1.0;
  ck(te_isolver_query(solverFor(type), 1, &uid, nullptr, &t1, origin.data(), &radius, nullptr, nullptr, &delta, nullptr, nullptr));
  return delta;
}
bool IntersectionSolver::getIntersectionPoseWithSphere(const unsigned int& id, const double& t1, const double& pos_th, const double& ang_th,
                                                       const Vector3d& origin, const double& radius, Vector7d& intersection_pose) {
  if (!(t1 >= 0.0) || !(pos_th >= 0.0) || !(ang_th >= 0.0)) throw std::invalid_argument("t1, pos_th, ang_th must be >= 0 (:94-96)");
  intersection_pose = Vector7d{0, 0, 0, 0, 0, 0, 1};   // initPose (:99)
  std::lock_guard<std::recursive_mutex> lg(target_manager_->lock());
  int type;
  if (!target_manager_->typeOf(id, type)) return false;
  target_manager_->flush();
  const uint32_t uid = id;
  double delta = -1.0;
  uint8_t conv = 0;
  ck(te_isolver_query(solverFor(type), 1, &uid, nullptr, &t1, origin.data(), &radius, &pos_th, &ang_th, &delta, intersection_pose.data(),
                      &conv));
  return conv != 0;
}

// ----------------------------------------------------------------------------------------------
// utils.hpp:273-313
// ----------------------------------------------------------------------------------------------
std::vector<std::string> splitString(const std::string& s, const std::string& delimiter) {
  std::vector<std::string> res;
  size_t start = 0, end;
  while ((end = s.find(delimiter, start)) != std::string::npos) {
    res.push_back(s.substr(start, end - start));
    start = end + delimiter.length();
  }
  res.push_back(s.substr(start));
  return res;
}
bool getId(const std::string& s, unsigned int& id) {
  auto parts = splitString(s);
  if (parts.size() == 2) {   // 'xxx_id'
    id = (unsigned)std::stoi(parts[1]);   // throws like the reference on a non-numeric suffix
    return true;
  }
  return false;
}

}  // namespace target_estimation_b200
