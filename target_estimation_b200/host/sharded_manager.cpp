// sharded_manager.cpp -- ShardedTargetManager: one TargetManager per GPU of the box behind the reference's TargetManager API
// (/root/reference/include/target_estimation/target_manager.hpp:66-203), owner(id) = id mod G, no collective on the hot path;
// the optional exchange of estimates goes over NCCL (te_group_*, csrc/te_group.cu).  See the class comment in
// include/target_estimation_b200/target_manager.hpp.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>

#include "target_estimation_b200/target_manager.hpp"

namespace target_estimation_b200 {

// one worker thread per shard: a batched call runs on all shards at once (every shard's host-to-device copies, launches and
// read-backs are issued by its own thread against its own device)
class ShardWorkers {
 public:
  explicit ShardWorkers(int n) : n_(n), err_((size_t)n) {
    for (int r = 0; r < n; ++r) th_.emplace_back([this, r] { loop(r); });
  }
  ~ShardWorkers() {
    {
      std::lock_guard<std::mutex> lg(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& f) {
    {
      std::lock_guard<std::mutex> lg(m_);
      job_ = &f;
      left_ = n_;
      ++gen_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return left_ == 0; });
    job_ = nullptr;
    for (auto& e : err_)
      if (e) {
        std::exception_ptr x = e;
        for (auto& c : err_) c = nullptr;
        std::rethrow_exception(x);
      }
  }

 private:
  void loop(int r) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      try {
        (*job)(r);
      } catch (...) {
        err_[(size_t)r] = std::current_exception();
      }
      {
        std::lock_guard<std::mutex> lg(m_);
        if (--left_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::vector<std::exception_ptr> err_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  unsigned long gen_ = 0;
  int left_ = 0;
  bool stop_ = false;
};

namespace {
// id mod G without a division per record (the routing passes look at every id of a batch; a hardware divide per id cost more than
// the tick it feeds): a mask for powers of two, else Lemire's multiply-shift remainder, exact for 32-bit operands
struct FastMod {
  explicit FastMod(unsigned d) : d_(d), pow2_((d & (d - 1)) == 0), m_(~(unsigned long long)0 / d + 1) {}
  unsigned operator()(unsigned a) const {
    if (pow2_) return a & (d_ - 1);
    const unsigned long long low = m_ * a;
    return (unsigned)(((unsigned __int128)low * d_) >> 64);
  }
  unsigned d_;
  bool pow2_;
  unsigned long long m_;
};
// page-locked staging of one shard (the shard's batched calls copy from it at PCIe speed)
void stageReserve(double*& meas, unsigned char*& action, size_t& cap, size_t n) {
  if (n <= cap) return;
  const size_t want = n + n / 4 + 1024;
  if (meas) { te_host_unregister(meas); std::free(meas); }
  if (action) { te_host_unregister(action); std::free(action); }
  meas = (double*)std::malloc(want * 7 * sizeof(double));
  action = (unsigned char*)std::malloc(want);
  if (!meas || !action) throw std::bad_alloc();
  te_host_register(meas, want * 7 * sizeof(double));   // (a failed registration leaves pageable memory: slower, still correct)
  te_host_register(action, want);
  cap = want;
}
}  // namespace

ShardedTargetManager::ShardedTargetManager(const std::string& file, int n_shards, const int* devices) : TargetManager(file, devices ? devices[0] : 0) {
  if (n_shards <= 0) throw std::invalid_argument("a sharded manager needs at least one shard");
  const int n_dev = te_device_count();
  if (n_dev <= 0) throw std::runtime_error("no CUDA device available");
  for (int r = 0; r < n_shards; ++r) {
    const int d = devices ? devices[r] : r;
    if (d < 0 || d >= n_dev) throw std::invalid_argument("no such CUDA device: " + std::to_string(d));
    devices_.push_back(d);
    shard_.emplace_back(new TargetManager(file, d));
    shard_.back()->quiet = true;
  }
  stage_.resize((size_t)n_shards);
  workers_.reset(new ShardWorkers(n_shards));
  const unsigned hw = std::thread::hardware_concurrency();
  const int n_route = (int)std::max<unsigned>((unsigned)n_shards, std::min<unsigned>(hw ? hw : 1u, 16u));
  routers_.reset(new ShardWorkers(n_route));
  quiet = true;
}

ShardedTargetManager::~ShardedTargetManager() {
  workers_.reset();
  routers_.reset();
  if (group_) te_group_destroy(group_);
  for (Stage& s : stage_) {
    if (s.meas) { te_host_unregister(s.meas); std::free(s.meas); }
    if (s.action) { te_host_unregister(s.action); std::free(s.action); }
  }
}

template <class F> void ShardedTargetManager::forEachShard(F&& f) {
  if (shard_.size() == 1) { f(0); return; }
  const std::function<void(int)> fn = std::forward<F>(f);
  workers_->run(fn);
}

// ---- per-id calls: the owner's ------------------------------------------------------------------------------------------
void ShardedTargetManager::init(const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0, const Vector6d& v0, const Vector6d& a0) {
  shard_[(size_t)owner(id)]->init(id, dt0, t0, p0, v0, a0);
}
void ShardedTargetManager::init(const target_t& type, const unsigned int& id, const double& dt0, const double& t0, const MatrixXd& Q, const MatrixXd& R,
                                const MatrixXd& P0, const Vector7d& p0, const Vector6d& v0, const Vector6d& a0) {
  shard_[(size_t)owner(id)]->init(type, id, dt0, t0, Q, R, P0, p0, v0, a0);
}
bool ShardedTargetManager::update(const unsigned int& id, const double& dt, const Vector7d& meas) { return shard_[(size_t)owner(id)]->update(id, dt, meas); }
bool ShardedTargetManager::update(const unsigned int& id, const double& dt) { return shard_[(size_t)owner(id)]->update(id, dt); }
void ShardedTargetManager::update(const double& dt) {
  for (auto& s : shard_) s->update(dt);   // asynchronous launches: the devices predict concurrently
}
bool ShardedTargetManager::erase(const unsigned int& id) { return shard_[(size_t)owner(id)]->erase(id); }
TargetInterface::Ptr ShardedTargetManager::getTarget(const unsigned int& id) { return shard_[(size_t)owner(id)]->getTarget(id); }
bool ShardedTargetManager::getTargetPose(const unsigned int& id, Vector7d& pose) { return shard_[(size_t)owner(id)]->getTargetPose(id, pose); }
bool ShardedTargetManager::getTargetTwist(const unsigned int& id, Vector6d& twist) { return shard_[(size_t)owner(id)]->getTargetTwist(id, twist); }
bool ShardedTargetManager::getTargetAcceleration(const unsigned int& id, Vector6d& acc) { return shard_[(size_t)owner(id)]->getTargetAcceleration(id, acc); }
long long ShardedTargetManager::getNumberMeasurements(const unsigned int& id) { return shard_[(size_t)owner(id)]->getNumberMeasurements(id); }
void ShardedTargetManager::log() {
  for (auto& s : shard_) s->log();
}
void ShardedTargetManager::flush() {
  for (auto& s : shard_) s->flush();
}
void ShardedTargetManager::watch(long long n, const unsigned* ids, size_t max_samples) {
  std::vector<std::vector<unsigned>> per(shard_.size());
  for (long long k = 0; k < n; ++k) per[(size_t)owner(ids[k])].push_back(ids[k]);
  for (size_t r = 0; r < shard_.size(); ++r) shard_[r]->watch((long long)per[r].size(), per[r].data(), max_samples);
}

std::vector<unsigned int> ShardedTargetManager::getAvailableTargets() {
  std::vector<unsigned int> all;
  for (auto& s : shard_) {   // every shard's list is ascending: merge pairwise
    std::vector<unsigned int> part = s->getAvailableTargets();
    const size_t mid = all.size();
    all.insert(all.end(), part.begin(), part.end());
    std::inplace_merge(all.begin(), all.begin() + (long)mid, all.end());
  }
  return all;
}

// ---- batched calls: routed by owner, all shards at once -------------------------------------------------------------------------
long long ShardedTargetManager::initBatch(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0, long long n, const unsigned* ids,
                                          double dt0, const double* t0, const double* p0, const double* v0, const double* a0, const double* p0_scale) {
  if (n <= 0) return 0;
  const unsigned G = (unsigned)shard_.size();
  const FastMod owner_of(G);
  std::atomic<long long> added{0};
  forEachShard([&](int r) {
    std::vector<unsigned> s_ids;
    std::vector<double> s_t0, s_p0, s_v0, s_a0, s_sc;
    for (long long k = 0; k < n; ++k) {
      if (owner_of(ids[k]) != (unsigned)r) continue;
      s_ids.push_back(ids[k]);
      if (t0) s_t0.push_back(t0[k]);
      if (p0) s_p0.insert(s_p0.end(), p0 + 7 * k, p0 + 7 * k + 7);
      if (v0) s_v0.insert(s_v0.end(), v0 + 6 * k, v0 + 6 * k + 6);
      if (a0) s_a0.insert(s_a0.end(), a0 + 6 * k, a0 + 6 * k + 6);
      if (p0_scale) s_sc.push_back(p0_scale[k]);
    }
    if (s_ids.empty()) return;
    added += shard_[(size_t)r]->initBatch(type, Q, R, P0, (long long)s_ids.size(), s_ids.data(), dt0, t0 ? s_t0.data() : nullptr, p0 ? s_p0.data() : nullptr,
                                          v0 ? s_v0.data() : nullptr, a0 ? s_a0.data() : nullptr, p0_scale ? s_sc.data() : nullptr);
  });
  return added.load();
}

long long ShardedTargetManager::updateBatch(long long n, const unsigned* ids, double dt, const double* meas, const unsigned char* action) {
  if (n <= 0) return 0;
  const unsigned G = (unsigned)shard_.size();
  const FastMod owner_of(G);
  if (G == 1) return shard_[0]->updateBatch(n, ids, dt, meas, action);
  // Routing = a parallel stable partition of the batch by owner, every access sequential.  Routing thread t (there are as many as
  // the host offers, at most 16, at least G) takes the t-th contiguous chunk of the records: (1) it counts the owners in its chunk; the prefix sums give chunk t a place in every shard's staging behind the
  // chunks before it; (2) it streams its chunk once more and appends every record to its owner's staging (G write streams per
  // worker).  Inside a shard the records keep the caller's order (chunks are ordered, and so is each chunk), so the records of one
  // id stay in sequence.  (3) every shard runs its batched call on its own staging.  (A pass per shard over the whole batch, each
  // picking its own records out, gathered 56-byte records at a stride of G records: 300 ms for 33 M records on 8 shards.)
  const unsigned T = (unsigned)routers_->size();   // routing threads (>= G)
  std::vector<size_t> count((size_t)T * G, 0);     // [chunk t][owner r]
  auto chunk_lo = [&](unsigned t) { return (long long)((__int128)n * t / T); };
  const std::function<void(int)> count_pass = [&](int t) {
    std::vector<size_t> c(G, 0);   // (local: neighbouring rows of `count` share cache lines)
    for (long long k = chunk_lo((unsigned)t), e = chunk_lo((unsigned)t + 1); k < e; ++k) ++c[owner_of(ids[k])];
    std::copy(c.begin(), c.end(), count.begin() + (long)t * G);
  };
  routers_->run(count_pass);
  std::vector<size_t> offset((size_t)T * G, 0), total(G, 0);
  for (unsigned r = 0; r < G; ++r)
    for (unsigned t = 0; t < T; ++t) {
      offset[(size_t)t * G + r] = total[r];
      total[r] += count[(size_t)t * G + r];
    }
  forEachShard([&](int r) {   // (each worker sizes its own shard's staging: page-locking is the slow part and runs in parallel)
    Stage& st = stage_[(size_t)r];
    st.ids.resize(total[(size_t)r]);
    stageReserve(st.meas, st.action, st.cap, total[(size_t)r]);
  });
  const std::function<void(int)> scatter_pass = [&](int t) {
    std::vector<size_t> at(offset.begin() + (long)t * G, offset.begin() + (long)(t + 1) * G);
    for (long long k = chunk_lo((unsigned)t), e = chunk_lo((unsigned)t + 1); k < e; ++k) {
      const unsigned r = owner_of(ids[k]);
      Stage& st = stage_[r];
      const size_t j = at[r]++;
      st.ids[j] = ids[k];
      if (meas) std::memcpy(st.meas + 7 * j, meas + 7 * (size_t)k, 56);
      st.action[j] = action ? action[k] : (unsigned char)TE_ACT_UPDATE;
    }
  };
  routers_->run(scatter_pass);
  std::atomic<long long> applied{0};
  forEachShard([&](int r) {
    Stage& st = stage_[(size_t)r];
    if (total[(size_t)r] == 0) return;
    applied += shard_[(size_t)r]->updateBatch((long long)total[(size_t)r], st.ids.data(), dt, meas ? st.meas : nullptr, st.action);
  });
  return applied.load();
}

long long ShardedTargetManager::eraseBatch(long long n, const unsigned* ids) {
  if (n <= 0) return 0;
  const unsigned G = (unsigned)shard_.size();
  const FastMod owner_of(G);
  std::atomic<long long> erased{0};
  forEachShard([&](int r) {
    std::vector<unsigned> s_ids;
    for (long long k = 0; k < n; ++k)
      if (owner_of(ids[k]) == (unsigned)r) s_ids.push_back(ids[k]);
    if (!s_ids.empty()) erased += shard_[(size_t)r]->eraseBatch((long long)s_ids.size(), s_ids.data());
  });
  return erased.load();
}

void ShardedTargetManager::getEstimatesBatch(long long n, const unsigned* ids, const double* t1, double* pose7, double* twist6, double* acc6,
                                             unsigned char* found) {
  if (n <= 0) return;
  const unsigned G = (unsigned)shard_.size();
  const FastMod owner_of(G);
  if (G == 1) return shard_[0]->getEstimatesBatch(n, ids, t1, pose7, twist6, acc6, found);
  forEachShard([&](int r) {
    Stage& st = stage_[(size_t)r];
    st.ids.clear();
    st.where.clear();
    std::vector<double> s_t1;
    for (long long k = 0; k < n; ++k) {
      if (owner_of(ids[k]) != (unsigned)r) continue;
      st.ids.push_back(ids[k]);
      st.where.push_back(k);
      if (t1) s_t1.push_back(t1[k]);
    }
    const long long m = (long long)st.ids.size();
    if (m == 0) return;
    std::vector<double> po(pose7 ? (size_t)m * 7 : 0), tw(twist6 ? (size_t)m * 6 : 0), ac(acc6 ? (size_t)m * 6 : 0);
    std::vector<unsigned char> fo((size_t)m, 0);
    shard_[(size_t)r]->getEstimatesBatch(m, st.ids.data(), t1 ? s_t1.data() : nullptr, pose7 ? po.data() : nullptr, twist6 ? tw.data() : nullptr,
                                         acc6 ? ac.data() : nullptr, fo.data());
    for (long long j = 0; j < m; ++j) {   // rows of different shards are disjoint: no two threads write the same output row
      const long long k = st.where[(size_t)j];
      if (pose7) std::memcpy(pose7 + 7 * k, &po[7 * (size_t)j], 56);
      if (twist6) std::memcpy(twist6 + 6 * k, &tw[6 * (size_t)j], 48);
      if (acc6) std::memcpy(acc6 + 6 * k, &ac[6 * (size_t)j], 48);
      if (found) found[k] = fo[(size_t)j];
    }
  });
}

// ---- dense ticks: shard-major records, every shard's slice on its own device at once ---------------------------------------------
std::vector<unsigned int> ShardedTargetManager::denseIds() {
  std::vector<unsigned int> all;
  for (auto& s : shard_) {
    std::vector<unsigned int> part = s->getAvailableTargets();
    all.insert(all.end(), part.begin(), part.end());
  }
  return all;
}

long long ShardedTargetManager::updateDense(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) {
  std::vector<long long> off(shard_.size() + 1, 0);
  for (size_t r = 0; r < shard_.size(); ++r) off[r + 1] = off[r] + (long long)shard_[r]->size();
  forEachShard([&](int r) {
    const long long o = off[(size_t)r];
    if (off[(size_t)r + 1] == o) return;
    shard_[(size_t)r]->updateDense(dt, meas ? meas + o * meas_stride : nullptr, meas_stride, action ? action + o : nullptr, est_pos_out ? est_pos_out + 3 * o : nullptr);
  });
  return off.back();
}

long long ShardedTargetManager::updateDenseAsync(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) {
  // enqueueing costs microseconds per shard: one thread issues all of them, the devices then run concurrently
  long long o = 0;
  for (auto& s : shard_) {
    const long long n = (long long)s->size();
    if (n > 0) s->updateDenseAsync(dt, meas ? meas + o * meas_stride : nullptr, meas_stride, action ? action + o : nullptr, est_pos_out ? est_pos_out + 3 * o : nullptr);
    o += n;
  }
  return o;
}

void ShardedTargetManager::updateDenseWait(int lag) {
  for (auto& s : shard_) s->updateDenseWait(lag);
}

// ---- the optional exchange of estimates ----------------------------------------------------------------------------------------
bool ShardedTargetManager::gatherUsesNccl() {
  if (!group_) {
    group_ = te_group_create((int)devices_.size(), devices_.data());
    if (!group_) throw std::runtime_error(te_last_error());
  }
  return te_group_uses_nccl(group_) == 1;
}

long long ShardedTargetManager::gatherEstimates(std::vector<unsigned>* ids, std::vector<double>* records13, int publisher) {
  if (publisher < 0 || publisher >= (int)shard_.size()) throw std::invalid_argument("no such shard");
  gatherUsesNccl();   // (creates the group)
  flush();
  if (ids) ids->clear();
  if (records13) records13->clear();
  long long total = 0;
  gather_ms_ = 0.0;
  std::vector<te_pool*> pools(shard_.size());
  for (int type = 0; type < 4; ++type) {
    bool any = false;
    for (size_t r = 0; r < shard_.size(); ++r) {
      pools[r] = shard_[r]->poolOf(type, false);
      if (pools[r] && te_pool_size(pools[r]) == 0) pools[r] = nullptr;
      any = any || pools[r];
    }
    if (!any) continue;
    const long long n = te_group_allgather_estimates(group_, pools.data(), nullptr);
    if (n < 0) throw std::runtime_error(te_last_error());
    if (ids || records13) {
      const size_t o = (size_t)total;
      if (ids) ids->resize(o + (size_t)n);
      if (records13) records13->resize((o + (size_t)n) * 13);
      if (te_group_fetch(group_, publisher, records13 ? records13->data() + o * 13 : nullptr, ids ? ids->data() + o : nullptr, n) < 0)
        throw std::runtime_error(te_last_error());
    } else if (te_group_sync(group_) < 0) {
      throw std::runtime_error(te_last_error());
    }
    gather_ms_ += te_group_last_gather_ms(group_);
    total += n;
  }
  return total;
}

}  // namespace target_estimation_b200
