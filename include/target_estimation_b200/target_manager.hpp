// include/target_estimation_b200/target_manager.hpp -- host surface of the B200 target pool.
//
// Mirrors the reference's TargetManager / TargetInterface / IntersectionSolver API
// (/root/reference/include/target_estimation/target_manager.hpp:30-270, target_interface.hpp:39-283,
// intersection_solver.hpp:55-125; tick/expiry semantics of target_manager_ros.hpp:74-183) with the
// same method names, argument order, return conventions and quirks.  Storage is not one heap object per
// target: every call lands in the device-resident pools of include/te_pool.h.  Eigen is not a
// dependency of this build (absent from the image), so Vector7d / Vector6d / MatrixXd are the
// plain fixed/dynamic arrays below; INTEGRATION.md shows the Eigen::Map overloads a maintainer adds.
//
// Per-id calls (update(id,dt,meas), update(id,dt)) are queued and coalesced: the reference's
// "for every id: update(id, dt, meas)" loop of one tick becomes ONE kernel launch at the next
// getter / flush / repeated id.  Observable results are identical to immediate execution.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../te_pool.h"

namespace target_estimation_b200 {

typedef std::array<double, 3> Vector3d;
typedef std::array<double, 6> Vector6d;
typedef std::array<double, 7> Vector7d;

struct MatrixXd {   // dense row-major dynamic matrix (only what the API needs)
  int r = 0, c = 0;
  std::vector<double> d;
  MatrixXd() {}
  MatrixXd(int rows, int cols) : r(rows), c(cols), d((size_t)rows * cols, 0.0) {}
  double& operator()(int i, int j) { return d[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
  int rows() const { return r; }
  int cols() const { return c; }
  const double* data() const { return d.data(); }
};
typedef std::vector<double> VectorXd;

class TargetManager;

// id -> model type, ascending ids: the host registry behind getAvailableTargets / "already exists" / "does not exist".  The
// reference's std::map<unsigned, TargetInterface::Ptr> (target_manager.hpp:36) node-per-target layout costs a cache miss per
// tree level: erasing the 10 k ids that expire in one tick of a million-target pool took 10 ms.  A sorted flat vector keeps the
// map's interface for the by-id calls (binary search; appending ids in ascending order is O(1)) and adds linear-time merges
// for the batches a tick produces.
class IdRegistry {
 public:
  typedef std::pair<unsigned, uint8_t> value_type;
  typedef std::vector<value_type>::iterator iterator;
  typedef std::vector<value_type>::const_iterator const_iterator;
  iterator begin() { return v_.begin(); }
  iterator end() { return v_.end(); }
  const_iterator begin() const { return v_.begin(); }
  const_iterator end() const { return v_.end(); }
  size_t size() const { return v_.size(); }
  bool empty() const { return v_.empty(); }
  iterator find(unsigned id);
  size_t count(unsigned id) { return find(id) != v_.end() ? 1 : 0; }
  uint8_t& operator[](unsigned id);                 // inserts (type 0) if missing, like std::map
  iterator erase(iterator it) { return v_.erase(it); }
  size_t erase(unsigned id);
  template <class It> void insert(It first, It last) {   // range of (id, type) pairs in ascending id order; existing ids are kept
    std::vector<value_type> add(first, last);
    mergeSorted(add);
  }
  void insertSorted(const uint32_t* ids, size_t n, uint8_t type);   // ascending ids
  void eraseSorted(const uint32_t* ids, size_t n);                  // ascending ids; unknown ones are ignored
 private:
  void mergeSorted(const std::vector<value_type>& add);
  std::vector<value_type> v_;
};

// KalmanFilterInterface view (kalman.hpp:49-89): what test/target_manager_test.cpp:144 reads
class EstimatorView {
 public:
  EstimatorView(TargetManager* m, unsigned id) : mgr_(m), id_(id) {}
  VectorXd getState() const;
  MatrixXd getP() const;
  MatrixXd getQ() const;
  MatrixXd getR() const;
  MatrixXd getP0() const;
 private:
  TargetManager* mgr_;
  unsigned id_;
};

// TargetInterface proxy (target_interface.hpp:58-160): every getter reads the target's slot back from
// the device (a sync point).  A handle may outlive erase(); its getters then return zeros.
class TargetInterface {
 public:
  typedef std::shared_ptr<TargetInterface> Ptr;
  TargetInterface(TargetManager* m, unsigned id) : mgr_(m), id_(id), est_(m, id) {}
  void addMeasurement(const double& dt, const Vector7d& meas);
  void update(const double& dt);
  Vector7d getEstimatedPose() const;
  Vector6d getEstimatedTwist() const;
  Vector6d getEstimatedAcceleration() const;
  Vector7d getEstimatedPose(const double& t1) const;
  Vector6d getEstimatedTwist(const double& t1) const;
  Vector6d getEstimatedAcceleration(const double& t1) const;
  Vector7d getMeasuredPose() const;
  double getTime() const;
  double getPeriodEstimate() const;   // src/target_interface.cpp:80-87
  long long getNumberMeasurements() const;
  unsigned getID() const { return id_; }
  const EstimatorView* getEstimator() const { return &est_; }
 private:
  TargetManager* mgr_;
  unsigned id_;
  EstimatorView est_;
};

class TargetManager {
 public:
  typedef std::shared_ptr<TargetManager> Ptr;
  enum target_t { ANGULAR_RATES = 0, ANGULAR_VELOCITIES, UNIFORM_ACCELERATION, UNIFORM_VELOCITY };   // target_manager.hpp:38

  explicit TargetManager(int device = 0);
  explicit TargetManager(const std::string& file, int device = 0);   // throws const char* like the reference (:110-118)
  virtual ~TargetManager();

  // reference API ---------------------------------------------------------------------------
  virtual void init(const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0, const Vector6d& v0 = Vector6d{},
                    const Vector6d& a0 = Vector6d{});
  virtual void init(const target_t& type, const unsigned int& id, const double& dt0, const double& t0, const MatrixXd& Q, const MatrixXd& R,
                    const MatrixXd& P0, const Vector7d& p0, const Vector6d& v0 = Vector6d{}, const Vector6d& a0 = Vector6d{});
  void init(const std::string& file, const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0,
            const Vector6d& v0 = Vector6d{}, const Vector6d& a0 = Vector6d{});
  virtual bool update(const unsigned int& id, const double& dt, const Vector7d& meas);
  virtual bool update(const unsigned int& id, const double& dt);
  virtual void update(const double& dt);
  virtual bool erase(const unsigned int& id);
  virtual TargetInterface::Ptr getTarget(const unsigned int& id);
  virtual bool getTargetPose(const unsigned int& id, Vector7d& pose);
  virtual bool getTargetTwist(const unsigned int& id, Vector6d& twist);
  virtual bool getTargetAcceleration(const unsigned int& id, Vector6d& acc);
  virtual long long getNumberMeasurements(const unsigned int& id);
  virtual void log();
  virtual std::vector<unsigned int> getAvailableTargets();
  bool selectTargetType(const std::string& type_str, target_t& type);
  bool loadYamlFile(const std::string& file, MatrixXd& Q, MatrixXd& R, MatrixXd& P, target_t& type);

  // batched extensions (one kernel launch per call) ---------------------------------------------
  long long initBatch(long long n, const unsigned* ids, double dt0, const double* t0, const double* p0 /*[n][7]*/,
                      const double* v0 = nullptr, const double* a0 = nullptr);
  virtual long long initBatch(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0, long long n, const unsigned* ids, double dt0,
                              const double* t0, const double* p0, const double* v0 = nullptr, const double* a0 = nullptr,
                              const double* p0_scale = nullptr);
  // (an id may appear more than once: the records are applied in order, as the reference's sequential update() calls would --
  //  the batch is cut in front of every repeat and the pieces are launched one after the other)
  virtual long long updateBatch(long long n, const unsigned* ids, double dt, const double* meas /*[n][7]*/, const unsigned char* action = nullptr);
  virtual long long eraseBatch(long long n, const unsigned* ids);
  virtual void getEstimatesBatch(long long n, const unsigned* ids, const double* t1, double* pose7, double* twist6, double* acc6,
                                 unsigned char* found);
  virtual void flush();
  // One dense tick from host arrays, record k = the k-th target in denseIds() order (a homogeneous manager: getAvailableTargets()
  // order): no ids travel, nothing is looked up -- te_pool_tick_host: the copies, the step and the read-back of every target's
  // estimated position (est_pos_out [n][3], may be NULL: what the node broadcasts per tick, src/target_manager_ros.cpp:78-87) are
  // pipelined in chunks.  meas_stride 7 = poses, 3 = x y z (UV / UA).  action NULL = update(id, dt, meas) for every target.
  // updateDenseAsync enqueues the tick and returns (two ticks may be in flight; buffers must stay valid until the tick is done);
  // updateDenseWait(0) waits for all of them, (1) for all but the newest.  Returns the number of targets.
  virtual long long updateDense(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out);
  virtual long long updateDenseAsync(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out);
  virtual void updateDenseWait(int lag);
  virtual std::vector<unsigned int> denseIds();
  size_t size();   // number of targets (of this manager / shard)
  bool quiet = false;   // suppress the reference's stdout messages ("does not exist", "already exists", ...)

  // sampled logging -----------------------------------------------------------------------------
  // The reference's log() (src/target_manager.cpp:119-123 -> src/target_interface.cpp:50-55, under LOGGER_ON) publishes, per
  // target, measured_pose_, pose_internal_, twist_, acceleration_ and P_ through rt_logger (:32-40); its tests dump time / pose /
  // twist series with writeTxtFile (utils.hpp:78-120) for matlab/plot_*.m.  Here log() appends ONE batched read-back of those five
  // quantities for the watched ids to an in-memory series, and writeLog() dumps them in writeTxtFile's format.
  virtual void watch(long long n, const unsigned* ids, size_t max_samples = 1 << 20);   // n = 0: stop logging and drop the series
  size_t logSamples() const { return log_t_.size(); }
  // one sample k of watched id j: [t | measured_pose 7 | pose_internal 6 | twist 6 | acceleration 6] = 26 doubles, then P (N*N)
  bool logSample(size_t k, size_t j, double* row26, double* P, int* n_state) const;
  // <folder>time_<id>, meas_pose_<id>, est_pose_<id> (x y z roll pitch yaw), est_twist_<id>, est_acc_<id>, cov_diag_<id>:
  // one row per sample, values separated by a blank, default ostream precision (what writeTxtFile writes).  Returns #files.
  int writeLog(const std::string& folder) const;

  // access for the proxies / solver
  struct Slot { int type; };
  bool typeOf(unsigned id, int& type);
  te_pool* poolOf(int type, bool create);
  int device() const { return device_; }
  std::recursive_mutex& lock() { return target_lock_; }

 protected:
  struct Pending {
    std::vector<uint32_t> ids;
    std::vector<double> dt, meas;
    std::vector<uint8_t> action;
  };
  void flushLocked();
  int registerClass(int type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0);
  void queue(int type, unsigned id, double dt, const double* meas, int action);
  long long updateBatchUnique(long long n, const unsigned* ids, double dt, const double* meas, const unsigned char* action);   // no id twice
  te_pool* densePool();   // the one non-empty pool of a homogeneous manager (throws otherwise)

  int device_;
  te_pool* pools_[4] = {nullptr, nullptr, nullptr, nullptr};
  Pending pending_[4];
  std::unordered_map<unsigned, uint8_t> pending_ids_;
  IdRegistry targets_;   // id -> model type, ascending like the reference's std::map
  std::recursive_mutex target_lock_;
  MatrixXd default_Q_, default_P_, default_R_;
  target_t default_type_ = UNIFORM_VELOCITY;
  bool default_values_loaded_ = false;
  // sampled logging: watched ids, per-sample tick time, per-sample-per-id rows
  std::vector<unsigned> log_ids_;
  size_t log_cap_ = 0;
  std::vector<double> log_t_;                  // [samples] time of the first watched target that exists (else NaN)
  std::vector<double> log_rows_;               // [samples][ids][26]
  std::vector<double> log_P_;                  // [samples][ids][18*18] (n_state^2 used)
  std::vector<int> log_n_;                     // [samples][ids] state size, 0 = id unknown at that sample
};

// utils.hpp:78-120: one value (vector) or one row (matrix) per line, "value " per column, default ostream formatting
bool writeTxtFile(const std::string& filename, const double* values, size_t rows, size_t cols);

// utils.hpp:206-265 (host copy used by nothing on the hot path; kept for API completeness)
class MovingAvgFilter {
 public:
  typedef std::shared_ptr<MovingAvgFilter> Ptr;
  explicit MovingAvgFilter(unsigned n) : sum_(0.0), variance_(0.0), window_(n, 0.0), idx_(0), complete_(false) {}
  double update(double value);
  double getVariance() const { return variance_; }
 private:
  double sum_, variance_;
  std::vector<double> window_;
  unsigned idx_;
  bool complete_;
};

// intersection_solver.hpp:55-125.  One object = one reference solver (one pair of moving-average
// filters + previous intersection pose shared by every id queried through it).
class IntersectionSolver {
 public:
  typedef std::shared_ptr<IntersectionSolver> Ptr;
  IntersectionSolver(TargetManager::Ptr target_manager, const unsigned int filters_length = 250);
  ~IntersectionSolver();
  double getIntersectionTimeWithSphere(const unsigned int& id, const double& t1, const Vector3d& origin, const double& radius);
  bool getIntersectionPoseWithSphere(const unsigned int& id, const double& t1, const double& pos_th, const double& ang_th,
                                     const Vector3d& origin, const double& radius, Vector7d& intersection_pose);
 private:
  te_isolver* solverFor(int type);
  TargetManager::Ptr target_manager_;
  unsigned filters_length_;
  // the filters live on the device with the pool they were last used with; a reference solver used
  // across model types would share them -- documented limitation: one solver state per model type
  te_isolver* solvers_[4] = {nullptr, nullptr, nullptr, nullptr};
};

// ---------------------------------------------------------------------------------------------------
// RosTargetManager without ROS (target_manager_ros.hpp:74-183, src/target_manager_ros.cpp:6-107): the /tf
// mailbox per id, the per-tick init-on-first-sight / update / predict loop, expiry erase and the list of
// filtered poses the node broadcasts.  ros::Time stamps are (sec, nsec) pairs; ros::Time::now() is passed in.
// The mailboxes are DEVICE resident (te_pool_mailbox_*): one /tf message = one ingest call (lookup, stable sort by slot,
// Measurement::update per slot), one tick = one rebuild + one step launch + one estimate gather; host work per tick is
// proportional to the ids that appear or expire, not to the number of targets.
// ---------------------------------------------------------------------------------------------------
struct StampedPose {
  uint32_t sec = 0, nsec = 0;
  Vector7d pose{};   // geometry_msgs default-constructs to zeros
};
double toSec(uint32_t sec, uint32_t nsec);   // utils.hpp:59-62 (never contracted to an FMA)

class Measurement {   // target_manager_ros.hpp:74-134
 public:
  Measurement() : last_meas_time_(0.0), new_meas_(true) {}
  bool read(StampedPose& tr) const {
    if (new_meas_) { tr = tr_; return true; }   // new_meas_ is NOT cleared by read()
    return false;
  }
  void update(const StampedPose& tr);
  double getTime() const { return last_meas_time_; }
  // the accepted stamp behind last_meas_time_, and whether the device copy of it is stale
  uint32_t acceptedSec() const { return acc_sec_; }
  uint32_t acceptedNsec() const { return acc_nsec_; }
  bool stampDirty() const { return stamp_dirty_; }
  void clearStampDirty() { stamp_dirty_ = false; }
 private:
  double last_meas_time_;
  bool new_meas_;
  StampedPose tr_;
  uint32_t acc_sec_ = 0, acc_nsec_ = 0;
  bool stamp_dirty_ = false;
};

class TickTargetManager : public TargetManager {
 public:
  TickTargetManager(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P, int device = 0);
  explicit TickTargetManager(const std::string& yaml_file, int device = 0);
  ~TickTargetManager() override;
  // /tf callback (src/target_manager_ros.cpp:26-39): frame "<token>_<id>"; a frame that contains the token but does
  // not parse BREAKS the loop (the rest of the message is dropped), like the reference
  void measurementCallBack(long long n, const char* const* child_frame_ids, const uint32_t* sec, const uint32_t* nsec, const double* poses);
  void measurementCallBackIds(long long n, const unsigned* ids, const uint32_t* sec, const uint32_t* nsec, const double* poses);
  // RosTargetManager::update(dt) (:41-92) with ros::Time::now() = (now_sec, now_nsec).  erased (optional) receives
  // the ids removed by the expiry rule in ascending order.
  void tick(const double& dt, uint32_t now_sec, uint32_t now_nsec, std::vector<unsigned>* erased = nullptr);
  // the poses the node would broadcast after the last tick: ascending ids + [n][7]
  const std::vector<unsigned>& publishedIds() const { return pub_ids_; }
  const std::vector<double>& publishedPoses() const { return pub_poses_; }
  void setTargetTokenName(const std::string& token_name) { token_name_ = token_name; }
  void setExpirationTime(double t);
  double time() const { return t_; }
  size_t mailboxCount();   // measurements_.size() of the reference: device mailboxes + target-less ones + host ones
  bool publish = true;   // gather the filtered poses every tick (the TF broadcast of :78-87)
 private:
  te_pool* tickPool();
  void tickForeign(const double& dt, uint32_t now_sec, uint32_t now_nsec, std::vector<unsigned>& gone_out);
  int cls_ = -1;   // model class of (Q_, R_, P_) in the pool of type_
  target_t type_;
  MatrixXd Q_, P_, R_;
  std::string token_name_;
  double t_;
  std::map<unsigned, Measurement> measurements_;   // host mailboxes: only ids that live in a pool of another model type
  double expiration_time_;
  std::vector<unsigned> pub_ids_;
  std::vector<double> pub_poses_;
  std::vector<uint32_t> gone_buf_, born_buf_;   // reusable output buffers of te_pool_mailbox_tick
  void* pub_pinned_ = nullptr;                  // storage of pub_poses_ while it is page-locked
};

// ---------------------------------------------------------------------------------------------------
// One TargetManager per GPU of the box behind the TargetManager interface (BASELINE.json north_star: "targets shard by ID across
// the 8 GPUs of one box with no collective on the hot path ... the TargetManager API stays as the host surface in C++").
// owner(id) = id mod G; shard r is an ordinary TargetManager on device devices[r] with its own pools and streams.  Per-id calls
// go to the owner (their launches coalesce per shard as in TargetManager).  The batched calls route their records to the owners
// in one host pass -- every shard's worker thread picks its own records out of the caller's arrays into page-locked staging, in
// the caller's order, and runs the shard's batched call -- so the G devices copy and step concurrently.  Results are the
// reference's: an id lives in exactly one shard, its records are applied in call order.
// The optional exchange of estimates is gatherEstimates(): [pose7 | twist6] records of every target, all shards,
// gathered on every device over NCCL (te_group_*), then read back from the publishing shard's device.
// ---------------------------------------------------------------------------------------------------
class ShardWorkers;
class ShardedTargetManager : public TargetManager {
 public:
  // devices NULL = devices 0 .. n_shards-1; a device may repeat (several shards on one GPU: a one-GPU box still runs the
  // sharded host logic -- the estimate exchange then uses device copies, NCCL needs distinct devices)
  ShardedTargetManager(const std::string& file, int n_shards, const int* devices = nullptr);
  ~ShardedTargetManager() override;
  int shards() const { return (int)shard_.size(); }
  int owner(unsigned id) const { return (int)(id % (unsigned)shard_.size()); }
  TargetManager& shard(int r) { return *shard_[(size_t)r]; }

  void init(const unsigned int& id, const double& dt0, const double& t0, const Vector7d& p0, const Vector6d& v0 = Vector6d{},
            const Vector6d& a0 = Vector6d{}) override;
  void init(const target_t& type, const unsigned int& id, const double& dt0, const double& t0, const MatrixXd& Q, const MatrixXd& R,
            const MatrixXd& P0, const Vector7d& p0, const Vector6d& v0 = Vector6d{}, const Vector6d& a0 = Vector6d{}) override;
  bool update(const unsigned int& id, const double& dt, const Vector7d& meas) override;
  bool update(const unsigned int& id, const double& dt) override;
  void update(const double& dt) override;
  bool erase(const unsigned int& id) override;
  TargetInterface::Ptr getTarget(const unsigned int& id) override;
  bool getTargetPose(const unsigned int& id, Vector7d& pose) override;
  bool getTargetTwist(const unsigned int& id, Vector6d& twist) override;
  bool getTargetAcceleration(const unsigned int& id, Vector6d& acc) override;
  long long getNumberMeasurements(const unsigned int& id) override;
  void log() override;
  std::vector<unsigned int> getAvailableTargets() override;   // ascending ids over all shards
  long long initBatch(target_t type, const MatrixXd& Q, const MatrixXd& R, const MatrixXd& P0, long long n, const unsigned* ids, double dt0,
                      const double* t0, const double* p0, const double* v0 = nullptr, const double* a0 = nullptr,
                      const double* p0_scale = nullptr) override;
  using TargetManager::initBatch;
  long long updateBatch(long long n, const unsigned* ids, double dt, const double* meas, const unsigned char* action = nullptr) override;
  long long eraseBatch(long long n, const unsigned* ids) override;
  void getEstimatesBatch(long long n, const unsigned* ids, const double* t1, double* pose7, double* twist6, double* acc6,
                         unsigned char* found) override;
  void flush() override;
  // dense ticks: the records are SHARD-MAJOR (denseIds(): shard 0's targets in ascending id, then shard 1's, ...); every shard's
  // slice is copied, stepped and read back by its own worker thread on its own device
  long long updateDense(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) override;
  long long updateDenseAsync(double dt, const double* meas, int meas_stride, const unsigned char* action, double* est_pos_out) override;
  void updateDenseWait(int lag) override;
  std::vector<unsigned int> denseIds() override;
  void watch(long long n, const unsigned* ids, size_t max_samples = 1 << 20) override;   // forwarded to the owners

  // The optional all-gather of estimates: ids (shard-major: shard 0's targets of model type 0 in ascending id, ...) and their
  // [pose7 | twist6] records, exchanged between the devices over NCCL and read back from shard `publisher`'s device.  Either output
  // may be NULL (exchange only).  Returns the number of records; lastGatherMs() = device time of the exchange(s).
  long long gatherEstimates(std::vector<unsigned>* ids, std::vector<double>* records13, int publisher = 0);
  double lastGatherMs() const { return gather_ms_; }
  bool gatherUsesNccl();

 private:
  template <class F> void forEachShard(F&& f);   // f(r) on shard r's worker thread, all shards at once; rethrows the first error
  std::vector<std::unique_ptr<TargetManager>> shard_;
  std::vector<int> devices_;
  std::unique_ptr<ShardWorkers> workers_;   // one thread per shard: the shards' batched calls
  std::unique_ptr<ShardWorkers> routers_;   // the routing passes of a batch: as many threads as the host offers (at most 16)
  te_group* group_ = nullptr;
  double gather_ms_ = -1.0;
  struct Stage {   // per shard, page-locked, grow-only
    std::vector<unsigned> ids;
    std::vector<long long> where;
    double* meas = nullptr;
    unsigned char* action = nullptr;
    size_t cap = 0;
  };
  std::vector<Stage> stage_;
};

// utils.hpp:273-313
std::vector<std::string> splitString(const std::string& s, const std::string& delimiter = "_");
bool getId(const std::string& s, unsigned int& id);

}  // namespace target_estimation_b200
