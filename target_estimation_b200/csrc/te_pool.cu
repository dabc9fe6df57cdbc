// te_pool.cu -- host side of the thin extern "C" CUDA layer declared in include/te_pool.h.
//
// One te_pool = the targets of one model type on one device, stored as tiles of 32 targets
// ([tile][field][lane], te_device.cuh Layout) in ascending-id slot order -- the iteration order of
// the reference's std::map<unsigned, TargetInterface::Ptr> (include/target_estimation/
// target_manager.hpp:36).  Add / erase rebuild the pool by a stable gather into the second buffer
// (stream compaction; ids stay sorted); the append fast path (all new ids larger than every
// existing id) writes in place.  There is no CPU fallback: every entry point needs a CUDA device.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/te_pool.h"
#include "te_kernels.cuh"
#include "te_split.cuh"
#include "te_direct.cuh"
#include "te_ar_pair.cuh"

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// grow-only device scratch, bump-allocated per call: one main block sized by the previous call's demand,
// overflow goes to one-off blocks that are folded into the main block at the next reset
struct Arena {
  char* main = nullptr;
  size_t main_cap = 0, off = 0, want = 0;
  std::vector<void*> extra;
  void reset() {
    for (void* c : extra) cudaFree(c);
    extra.clear();
    if (want > main_cap) {
      cudaFree(main);
      main = nullptr;
      main_cap = 0;
      const size_t cap = want + want / 2;
      CK(cudaMalloc((void**)&main, cap));
      main_cap = cap;
    }
    off = 0;
    want = 0;
  }
  void* get(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    want += bytes;
    if (off + bytes <= main_cap) {
      void* p = main + off;
      off += bytes;
      return p;
    }
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    extra.push_back(p);
    return p;
  }
  template <class T> T* get_n(size_t n) { return (T*)get(n * sizeof(T)); }
  void destroy() {
    for (void* c : extra) cudaFree(c);
    extra.clear();
    cudaFree(main);
    main = nullptr;
    main_cap = 0;
  }
};

struct Buf {   // one generation of the pool's per-slot storage
  double* tiles = nullptr;
  te::ColdArrays cold{nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;   // slots (multiple of 32)
};

struct MailBuf {   // one generation of the per-slot mailboxes (te_pool_mailbox_*)
  te::MailArrays a{nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;
};

// A mailbox whose id has no target yet (Measurement of target_manager_ros.hpp:74-134 on the host): the message carried a
// stamp that is not newer than the initial one, or a newer record was followed by an older one before the tick.  Rare, so
// these stay in a host map; the tick promotes the readable ones to targets and expires the others by the same predicate.
struct HostMail {
  uint32_t sec = 0, nsec = 0;
  double last = 0.0;
  bool fresh = true;   // Measurement(): new_meas_ = true
  double pose[7] = {0, 0, 0, 0, 0, 0, 0};
};
inline double host_to_sec(uint32_t sec, uint32_t nsec) {   // utils.hpp:59-62, never contracted
  volatile double ns = 1e-9 * (double)nsec;
  return (double)sec + ns;
}
// a /tf record whose id has no target: kept in arrival order until the next tick folds it into a mailbox (the common case --
// an id seen for the first time, promoted by that tick -- then never touches the std::map)
struct PendingRec {
  uint32_t id, sec, nsec;
  double pose[7];
};
inline void apply_record(HostMail& m, const PendingRec& r) {   // Measurement::update (target_manager_ros.hpp:96-115)
  const double cur = host_to_sec(r.sec, r.nsec), prev = host_to_sec(m.sec, m.nsec);
  if (cur > prev) { m.fresh = true; m.last = cur; }
  else m.fresh = false;
  m.sec = r.sec;
  m.nsec = r.nsec;
  std::memcpy(m.pose, r.pose, sizeof(m.pose));
}

}  // namespace

struct te_pool {
  int model = 0, device = 0;
  int N = 0, M = 0, NF = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n_sm = 148;
  int variant = 0;
  int grid_cap = 0;      // test hook (te_pool_set_grid_cap): upper bound on the CTAs of a step launch, 0 = none
  bool all_sym = true;   // every registered class has bitwise-symmetric Q, R, P0 (symmetric-covariance kernels are legal)
  // The direct symmetric kernels maintain the UPPER triangle of every covariance only ("packed": 36 of UA's 92 fields are
  // neither read nor written per step).  lower_stale = the lower triangles in HBM are out of date; whoever needs the full
  // matrix (a full-matrix kernel, a state read-back) mirrors it first / on the fly.
  bool lower_stale = false;
  long long n = 0;   // live targets
  Buf buf[2];
  int cur = 0;
  // per-slot work arrays (capacity wcap slots)
  size_t wcap = 0;
  uint8_t* action = nullptr;
  double* dt_slot = nullptr;
  uint8_t* tile_flag = nullptr;
  int* tile_list = nullptr;
  int* alive = nullptr;
  int* pos = nullptr;
  int* srcmap = nullptr;
  int* d_counters = nullptr;   // [0] = n_work, [1] = applied
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  // model classes
  std::vector<std::vector<double>> hQ, hR, hP0;
  double *dQ = nullptr, *dR = nullptr, *dP0 = nullptr;
  int cls_cap = 0;
  // host mirror of the sorted ids (lazy)
  std::vector<uint32_t> h_ids;
  bool h_ids_valid = true;
  // largest live id, kept across compactions: enough to recognise an append-only add batch (monotonically increasing
  // track ids, the common case) without downloading the whole id array again
  uint32_t h_last_id = 0;
  bool h_last_valid = false;
  Arena arena;
  // device-resident mailboxes (te_pool_mailbox_*): allocated on first use, then carried through every compaction
  bool mb_on = false;
  MailBuf mb[2];
  int mb_cur = 0;
  te::MailAdd mb_add{nullptr, nullptr, nullptr};   // set by the mailbox tick around its merge
  std::map<uint32_t, HostMail> orphans;            // mailboxes without a target that outlived a tick (unreadable ones)
  std::vector<PendingRec> pending;                 // records of unknown ids since the last tick, arrival order
  char* h_stage = nullptr;                         // pinned staging for the tick's add arrays / the ingest's read-backs (grow-only)
  size_t h_stage_cap = 0;
  // chunk pipeline of te_pool_tick_host
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  std::vector<cudaEvent_t> events;
};

struct te_isolver {
  te_pool* pool = nullptr;
  te::IsolverState st{};
};

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) CK(cudaSetDevice(dev));
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

size_t tile_doubles(const te_pool* p) { return (size_t)p->NF * te::TILE; }

void free_buf(Buf& b) {
  cudaFree(b.tiles);
  cudaFree(b.cold.ids);
  cudaFree(b.cold.cls);
  cudaFree(b.cold.last_meas);
  cudaFree(b.cold.meas);
  b = Buf();
}

void alloc_buf(te_pool* p, Buf& b, size_t slots) {
  slots = (slots + te::TILE - 1) / te::TILE * te::TILE;
  if (slots == 0) slots = te::TILE;
  CK(cudaMalloc(&b.tiles, slots / te::TILE * tile_doubles(p) * sizeof(double)));
  CK(cudaMalloc(&b.cold.ids, slots * sizeof(uint32_t)));
  CK(cudaMalloc(&b.cold.cls, slots * sizeof(uint16_t)));
  CK(cudaMalloc(&b.cold.last_meas, slots * sizeof(double)));
  CK(cudaMalloc(&b.cold.meas, slots * 7 * sizeof(double)));
  b.cap = slots;
  // pad lanes of the last tile are streamed by the step kernel: keep them finite
  CK(cudaMemsetAsync(b.tiles, 0, slots / te::TILE * tile_doubles(p) * sizeof(double), p->stream));
  CK(cudaMemsetAsync(b.cold.meas, 0, slots * 7 * sizeof(double), p->stream));
  CK(cudaMemsetAsync(b.cold.last_meas, 0, slots * sizeof(double), p->stream));
}

// make sure the per-slot work arrays cover `slots`
void ensure_work(te_pool* p, size_t slots) {
  slots = (slots + te::TILE - 1) / te::TILE * te::TILE;
  if (slots <= p->wcap) return;
  size_t cap = std::max(slots, p->wcap + p->wcap / 2);
  cap = (cap + te::TILE - 1) / te::TILE * te::TILE;
  CK(cudaStreamSynchronize(p->stream));
  cudaFree(p->action); cudaFree(p->dt_slot); cudaFree(p->tile_flag); cudaFree(p->tile_list);
  cudaFree(p->alive); cudaFree(p->pos); cudaFree(p->srcmap);
  CK(cudaMalloc(&p->action, cap));
  CK(cudaMalloc(&p->dt_slot, cap * sizeof(double)));
  CK(cudaMalloc(&p->tile_flag, cap / te::TILE + 4));
  CK(cudaMalloc(&p->tile_list, cap / te::TILE * sizeof(int)));
  CK(cudaMalloc(&p->alive, cap * sizeof(int)));
  CK(cudaMalloc(&p->pos, cap * sizeof(int)));
  CK(cudaMalloc(&p->srcmap, cap * sizeof(int)));
  CK(cudaMemsetAsync(p->action, 0, cap, p->stream));
  CK(cudaMemsetAsync(p->tile_flag, 0, cap / te::TILE + 4, p->stream));
  p->wcap = cap;
  size_t need = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, need, p->alive, p->pos, (int)cap, p->stream));
  if (need > p->cub_bytes) {
    cudaFree(p->cub_tmp);
    CK(cudaMalloc(&p->cub_tmp, need));
    p->cub_bytes = need;
  }
}

// current buffer must hold `slots` (append path / reserve): grow by copy
void ensure_cur_capacity(te_pool* p, size_t slots) {
  Buf& b = p->buf[p->cur];
  if (slots <= b.cap) return;
  size_t cap = std::max(slots, b.cap + b.cap / 2);
  Buf nb;
  alloc_buf(p, nb, cap);
  if (p->n > 0) {
    size_t tiles = (size_t)cdiv(p->n, te::TILE);
    CK(cudaMemcpyAsync(nb.tiles, b.tiles, tiles * tile_doubles(p) * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nb.cold.ids, b.cold.ids, p->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nb.cold.cls, b.cold.cls, p->n * sizeof(uint16_t), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nb.cold.last_meas, b.cold.last_meas, p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nb.cold.meas, b.cold.meas, p->n * 7 * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
  }
  CK(cudaStreamSynchronize(p->stream));
  free_buf(b);
  b = nb;
  ensure_work(p, b.cap);
}

void ensure_other_capacity(te_pool* p, size_t slots) {
  Buf& b = p->buf[1 - p->cur];
  if (slots <= b.cap && b.tiles) return;
  CK(cudaStreamSynchronize(p->stream));
  free_buf(b);
  // grow geometrically: a pool that gains a few targets per tick must not re-allocate gigabytes on every tick
  alloc_buf(p, b, std::max(slots + slots / 4, p->buf[p->cur].cap));
  // the work arrays hold the live alive[] / pos[] of the running compaction: callers size them up-front
  if (p->wcap < slots) throw std::logic_error("work arrays not sized before compaction");
}

// ---- mailboxes -------------------------------------------------------------------------------
char* pinned_stage(te_pool* p, size_t bytes) {
  if (bytes > p->h_stage_cap) {
    CK(cudaStreamSynchronize(p->stream));
    if (p->h_stage) cudaFreeHost(p->h_stage);
    p->h_stage = nullptr;
    p->h_stage_cap = 0;
    const size_t cap = bytes + bytes / 2 + 4096;
    CK(cudaHostAlloc((void**)&p->h_stage, cap, cudaHostAllocDefault));
    p->h_stage_cap = cap;
  }
  return p->h_stage;
}
void free_mail(MailBuf& m) {
  cudaFree(m.a.sec); cudaFree(m.a.nsec); cudaFree(m.a.act); cudaFree(m.a.pose);
  m = MailBuf();
}
void alloc_mail(MailBuf& m, size_t slots) {
  slots = (slots + te::TILE - 1) / te::TILE * te::TILE;
  if (slots == 0) slots = te::TILE;
  CK(cudaMalloc(&m.a.sec, slots * sizeof(uint32_t)));
  CK(cudaMalloc(&m.a.nsec, slots * sizeof(uint32_t)));
  CK(cudaMalloc(&m.a.act, slots));
  CK(cudaMalloc(&m.a.pose, slots * 7 * sizeof(double)));   // cudaMalloc is 256-byte aligned: TMA-loadable measurement block
  m.cap = slots;
}
// the current generation holds `slots` mailboxes (grow by copy, like ensure_cur_capacity)
void ensure_mail_cur(te_pool* p, size_t slots) {
  MailBuf& m = p->mb[p->mb_cur];
  if (slots <= m.cap && m.a.sec) return;
  MailBuf nm;
  alloc_mail(nm, std::max({slots, m.cap + m.cap / 2, p->buf[p->cur].cap}));   // sized like the slot buffers: no per-tick regrowth
  if (p->n > 0 && m.a.sec) {
    CK(cudaMemcpyAsync(nm.a.sec, m.a.sec, p->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nm.a.nsec, m.a.nsec, p->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nm.a.act, m.a.act, p->n, cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaMemcpyAsync(nm.a.pose, m.a.pose, p->n * 7 * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
  }
  CK(cudaStreamSynchronize(p->stream));
  free_mail(m);
  m = nm;
}
void ensure_mail_other(te_pool* p, size_t slots) {
  MailBuf& m = p->mb[1 - p->mb_cur];
  if (slots <= m.cap && m.a.sec) return;
  CK(cudaStreamSynchronize(p->stream));
  free_mail(m);
  alloc_mail(m, std::max({slots + slots / 8, p->mb[p->mb_cur].cap, p->buf[p->cur].cap, p->buf[1 - p->cur].cap}));
}
// first use: every existing target gets an empty mailbox
void enable_mail(te_pool* p) {
  if (p->mb_on) return;
  ensure_mail_cur(p, std::max<size_t>((size_t)p->n, te::TILE));
  if (p->n > 0) {
    te::mb_clear_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->mb[p->mb_cur].a, 0, (int)p->n);
    CK(cudaGetLastError());
  }
  p->mb_on = true;
}

void sync_host_ids(te_pool* p) {
  if (p->h_ids_valid) return;
  p->h_ids.resize((size_t)p->n);
  if (p->n) {
    CK(cudaMemcpyAsync(p->h_ids.data(), p->buf[p->cur].cold.ids, p->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
  }
  p->h_ids_valid = true;
  if (p->n) { p->h_last_id = p->h_ids.back(); p->h_last_valid = true; }
}

void upload_classes(te_pool* p) {
  const int nc = (int)p->hQ.size();
  if (nc > p->cls_cap) {
    CK(cudaStreamSynchronize(p->stream));
    cudaFree(p->dQ); cudaFree(p->dR); cudaFree(p->dP0);
    int cap = std::max(nc, std::max(4, p->cls_cap * 2));
    CK(cudaMalloc(&p->dQ, (size_t)cap * p->N * p->N * sizeof(double)));
    CK(cudaMalloc(&p->dR, (size_t)cap * p->M * p->M * sizeof(double)));
    CK(cudaMalloc(&p->dP0, (size_t)cap * p->N * p->N * sizeof(double)));
    p->cls_cap = cap;
    for (int c = 0; c < nc; ++c) {
      CK(cudaMemcpyAsync(p->dQ + (size_t)c * p->N * p->N, p->hQ[c].data(), p->N * p->N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
      CK(cudaMemcpyAsync(p->dR + (size_t)c * p->M * p->M, p->hR[c].data(), p->M * p->M * sizeof(double), cudaMemcpyHostToDevice, p->stream));
      CK(cudaMemcpyAsync(p->dP0 + (size_t)c * p->N * p->N, p->hP0[c].data(), p->N * p->N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
  } else {
    const int c = nc - 1;
    CK(cudaMemcpyAsync(p->dQ + (size_t)c * p->N * p->N, p->hQ[c].data(), p->N * p->N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->dR + (size_t)c * p->M * p->M, p->hR[c].data(), p->M * p->M * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->dP0 + (size_t)c * p->N * p->N, p->hP0[c].data(), p->N * p->N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  }
  CK(cudaStreamSynchronize(p->stream));
}

template <class T> T* to_dev(te_pool* p, const T* host, size_t n) {
  if (!host || !n) return nullptr;
  T* d = p->arena.get_n<T>(n);
  CK(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, p->stream));
  return d;
}

// ---- step kernel launch -----------------------------------------------------------------
// te_pool_set_grid_cap: a small pool under a capped grid walks the same persistent loops (grid-stride tiles, the STAGES ring of
// the split kernel with its mbarrier phase flips) that a bench-size pool walks on the full machine
inline int capped(const te_pool* p, int grid) { return p->grid_cap > 0 ? std::min(grid, p->grid_cap) : grid; }
template <int TYPE, int WARPS, int STAGES, bool MULTI = false, int IMPL = 0>
void launch_step_t(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  auto kern = te::kf_step_kernel<TYPE, WARPS, STAGES, MULTI, IMPL>;
  const size_t smem = te::step_smem_bytes<TYPE>(WARPS, STAGES);
  static thread_local int configured_dev = -1;
  static bool configured[64] = {false};
  (void)configured_dev;
  if (!configured[p->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[p->device & 63] = true;
  }
  int grid = capped(p, std::min(p->n_sm, std::max(1, cdiv(n_work_hint, WARPS))));
  kern<<<grid, WARPS * 32, smem, p->stream>>>(a);
  CK(cudaGetLastError());
}

// row/column-split kernel (te_split.cuh): one CTA of 6*CS warps per tile, STAGES stages per CTA, CTAS CTAs per SM
template <int TYPE, int CS, int STAGES, int CTAS, bool COMPACT>
void launch_split_k(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  auto kern = te::kf_step_split_kernel<TYPE, CS, STAGES, CTAS, COMPACT>;
  const size_t smem = te::split_smem_bytes<TYPE>(STAGES);
  static bool configured[64] = {false};
  if (!configured[p->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[p->device & 63] = true;
  }
  int grid = capped(p, std::min(p->n_sm * CTAS, std::max(1, n_work_hint)));
  kern<<<grid, (te::SPLIT_RS * CS + te::split_nt<TYPE>()) * 32, smem, p->stream>>>(a);
  CK(cudaGetLastError());
}
template <int TYPE, int CS, int STAGES, int CTAS>
void launch_split_t(te_pool* p, const te::StepArgs& a, int n_work_hint) { launch_split_k<TYPE, CS, STAGES, CTAS, false>(p, a, n_work_hint); }

bool uses_direct(const te_pool* p);
void ensure_full(te_pool* p);
template <int TYPE, int WARPS, int CTAS> void launch_kin_direct(te_pool* p, const te::StepArgs& a, int n_work_hint);
void launch_step_multi(te_pool* p, const te::StepArgs& a_in, int n_work_hint) {
  te::StepArgs a = a_in;
  if (uses_direct(p) && p->model != te::ANGULAR_VELOCITIES) {   // UV / UA: the direct kernel keeps the target in registers for all ticks
    a.packed = (p->all_sym && p->variant != 12) ? 1 : 0;
    if (a.packed) p->lower_stale = true;
    if (p->model == te::UNIFORM_VELOCITY) launch_kin_direct<te::UNIFORM_VELOCITY, 4, 3>(p, a, n_work_hint);
    else launch_kin_direct<te::UNIFORM_ACCELERATION, 8, 1>(p, a, n_work_hint);
    return;
  }
  ensure_full(p);
  switch (p->model) {
    case te::UNIFORM_VELOCITY: launch_step_t<te::UNIFORM_VELOCITY, 8, 2, true>(p, a, n_work_hint); break;
    case te::UNIFORM_ACCELERATION: launch_step_t<te::UNIFORM_ACCELERATION, 4, 2, true>(p, a, n_work_hint); break;
    case te::ANGULAR_VELOCITIES: launch_step_t<te::ANGULAR_VELOCITIES, 5, 1, true>(p, a, n_work_hint); break;
    default: launch_step_t<te::ANGULAR_RATES, 2, 1, true>(p, a, n_work_hint); break;
  }
}

template <int WARPS, int ZF = 2>
void launch_av_direct(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  auto kern = te::kf_step_av_direct_kernel<WARPS, ZF>;
  const size_t smem = te::av_direct_smem_bytes(WARPS);
  static bool configured[64] = {false};
  if (!configured[p->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[p->device & 63] = true;
  }
  int grid = capped(p, std::min(p->n_sm, std::max(1, cdiv(n_work_hint, WARPS))));
  kern<<<grid, WARPS * 32, smem, p->stream>>>(a);
  CK(cudaGetLastError());
}

template <int TYPE, int WARPS, int CTAS>
void launch_kin_direct_k(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  auto kern = te::kf_step_kin_direct_kernel<TYPE, WARPS, CTAS>;
  int grid = capped(p, std::min(p->n_sm * CTAS, std::max(1, cdiv(n_work_hint, WARPS))));
  // programmatic stream serialization: see the kernel's griddepcontrol.wait
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = p->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // Only for small pools, where a tick is a few microseconds and the launch latency matters.  For big pools it buys nothing,
  // and in a process with an NCCL communicator it was measured to cost 37 % (4 Mi UA targets: 0.64 -> 0.86 ms per tick under
  // torchrun, 0.63 either way in a plain process) -- the early-scheduled grid and the running one compete for the SMs.
  static const bool no_pdl = std::getenv("TE_NO_PDL") != nullptr;   // debugging switch
  cfg.attrs = attr;
  cfg.numAttrs = (no_pdl || n_work_hint > 4 * p->n_sm) ? 0 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, a));
}
// small pools (fewer tiles than the SMs have scheduler partitions) spread over more, smaller CTAs: one warp per partition has
// the FP64 pipe to itself, which is what bounds a tick of a few hundred tiles
template <int TYPE, int WARPS, int CTAS>
void launch_kin_direct(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  if (n_work_hint <= 2 * p->n_sm) launch_kin_direct_k<TYPE, 2, 1>(p, a, n_work_hint);
  else if (n_work_hint <= 4 * p->n_sm) launch_kin_direct_k<TYPE, 4, 1>(p, a, n_work_hint);
  else launch_kin_direct_k<TYPE, WARPS, CTAS>(p, a, n_work_hint);
}

template <int WARPS>
void launch_ar_pair(te_pool* p, const te::StepArgs& a, int n_work_hint) {
  auto kern = te::kf_step_ar_pair_kernel<WARPS>;
  const size_t smem = te::ar_pair_smem_bytes(WARPS);
  static bool configured[64] = {false};
  if (!configured[p->device & 63]) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[p->device & 63] = true;
  }
  int grid = capped(p, std::min(p->n_sm, std::max(1, cdiv(n_work_hint, WARPS))));
  kern<<<grid, WARPS * 32, smem, p->stream>>>(a);
  CK(cudaGetLastError());
}

// does the current variant run a direct symmetric-covariance kernel (te_direct.cuh)?
bool uses_direct(const te_pool* p) {
  const int v = p->variant;
  const bool dflt = v == 0 && p->all_sym;
  switch (p->model) {
    case te::UNIFORM_VELOCITY:
    case te::UNIFORM_ACCELERATION: return dflt || v == 5 || v == 6 || v == 7 || v == 12;
    case te::ANGULAR_VELOCITIES: return dflt || (v >= 6 && v <= 9) || v == 12;
    default: return v == 12 || v == 13;   // AR: the two-lanes-per-target kernel (te_ar_pair.cuh); 13 = packed
  }
}
// full-matrix kernels (and anything else that reads both halves) first get the lower triangles back
void ensure_full(te_pool* p) {
  if (!p->lower_stale || p->n == 0) { p->lower_stale = false; return; }
  double* tiles = p->buf[p->cur].tiles;
  const int n = (int)p->n;
  switch (p->model) {
    case te::UNIFORM_VELOCITY: te::mirror_lower_kernel<te::UNIFORM_VELOCITY><<<cdiv(n, 128), 128, 0, p->stream>>>(tiles, n); break;
    case te::UNIFORM_ACCELERATION: te::mirror_lower_kernel<te::UNIFORM_ACCELERATION><<<cdiv(n, 128), 128, 0, p->stream>>>(tiles, n); break;
    case te::ANGULAR_VELOCITIES: te::mirror_lower_kernel<te::ANGULAR_VELOCITIES><<<cdiv(n, 128), 128, 0, p->stream>>>(tiles, n); break;
    default: te::mirror_lower_kernel<te::ANGULAR_RATES><<<cdiv(n, 128), 128, 0, p->stream>>>(tiles, n); break;
  }
  CK(cudaGetLastError());
  p->lower_stale = false;
}

// variant -> kernel.  0 = default: the direct symmetric-covariance kernels for UV / UA / AV when every class is symmetric
// (packed: upper triangle only), else the full-matrix kernels; 10 = force the full-matrix kernel (TMA-staged / row-split);
// 12 = direct kernel writing both halves; the others are launch shapes kept for experiments (tests cover all of them).
// Stage bytes of the staged kernel: UV 13056, UA 25344, AV 43008, AR 90624.
void launch_step(te_pool* p, const te::StepArgs& a_in, int n_work_hint) {
  const int v = p->variant;
  te::StepArgs a = a_in;
  if (uses_direct(p)) {
    a.packed = (p->all_sym && v != 12) ? 1 : 0;
    if (a.packed) p->lower_stale = true;
    else if (p->model == te::ANGULAR_RATES) ensure_full(p);   // (the other direct kernels never read the lower triangle)
    switch (p->model) {
      case te::ANGULAR_RATES:
        launch_ar_pair<8>(p, a, n_work_hint);
        return;
      case te::UNIFORM_VELOCITY:
        // measured, packed, 4 Mi targets: <4,3> (12 warps per SM) 1.17e10, <4,4> 1.11e10, <8,1> 1.09e10 steps/s
        if (v == 5) launch_kin_direct<te::UNIFORM_VELOCITY, 4, 4>(p, a, n_work_hint);
        else if (v == 6) launch_kin_direct<te::UNIFORM_VELOCITY, 8, 1>(p, a, n_work_hint);
        else launch_kin_direct<te::UNIFORM_VELOCITY, 4, 3>(p, a, n_work_hint);
        return;
      case te::UNIFORM_ACCELERATION:
        if (v == 6) launch_kin_direct<te::UNIFORM_ACCELERATION, 4, 3>(p, a, n_work_hint);
        else if (v == 7) launch_kin_direct<te::UNIFORM_ACCELERATION, 5, 2>(p, a, n_work_hint);
        else launch_kin_direct<te::UNIFORM_ACCELERATION, 8, 1>(p, a, n_work_hint);
        return;
      default:
        if (v == 7) launch_av_direct<6>(p, a, n_work_hint);
        else if (v == 8) launch_av_direct<8, 4>(p, a, n_work_hint);
        else if (v == 9) launch_av_direct<8, 12>(p, a, n_work_hint);
        else launch_av_direct<8>(p, a, n_work_hint);
        return;
    }
  }
  // AR, variant 11: the row-split kernel in packed form (only the upper-triangle field ranges travel; te_split.cuh).  Not the
  // default: it moves 44 % fewer bytes but is not faster (1.07e9 vs 1.09e9 steps/s) -- with the traffic gone the six main
  // warps' work per tile is the bound, and they still compute full rows.
  if (p->model == te::ANGULAR_RATES && v == 11 && p->all_sym) {
    a.packed = 1;
    p->lower_stale = true;
  } else {
    ensure_full(p);
  }
  if (a.dst_tiles) {
    // compacting tick: separate instantiations of the default split configurations, so that the in-place kernels carry
    // none of its code (the AV kernel at 128 registers lost 6 % to a few extra runtime branches)
    if (p->model == te::ANGULAR_VELOCITIES) return launch_split_k<te::ANGULAR_VELOCITIES, 1, 2, 2, true>(p, a, n_work_hint);
    if (p->model == te::ANGULAR_RATES) return launch_split_k<te::ANGULAR_RATES, 1, 2, 1, true>(p, a, n_work_hint);
  }
  switch (p->model) {
    case te::UNIFORM_VELOCITY:
      if (v == 1) launch_step_t<te::UNIFORM_VELOCITY, 16, 1>(p, a, n_work_hint);
      else if (v == 2) launch_step_t<te::UNIFORM_VELOCITY, 4, 4>(p, a, n_work_hint);
      else launch_step_t<te::UNIFORM_VELOCITY, 8, 2>(p, a, n_work_hint);
      break;
    case te::UNIFORM_ACCELERATION:
      if (v == 1) launch_step_t<te::UNIFORM_ACCELERATION, 8, 1>(p, a, n_work_hint);
      else if (v == 2) launch_step_t<te::UNIFORM_ACCELERATION, 2, 4>(p, a, n_work_hint);
      else launch_step_t<te::UNIFORM_ACCELERATION, 4, 2>(p, a, n_work_hint);
      break;
    case te::ANGULAR_VELOCITIES:
      if (v == 1) launch_step_t<te::ANGULAR_VELOCITIES, 5, 1>(p, a, n_work_hint);
      else if (v == 5) launch_step_t<te::ANGULAR_VELOCITIES, 5, 1, false, 1>(p, a, n_work_hint);
      else if (v == 2) launch_split_t<te::ANGULAR_VELOCITIES, 1, 2, 1>(p, a, n_work_hint);
      else if (v == 3) launch_split_t<te::ANGULAR_VELOCITIES, 1, 1, 3>(p, a, n_work_hint);
      else if (v == 4) launch_split_t<te::ANGULAR_VELOCITIES, 1, 3, 1>(p, a, n_work_hint);
      else launch_split_t<te::ANGULAR_VELOCITIES, 1, 2, 2>(p, a, n_work_hint);
      break;
    default:
      if (v == 1) launch_step_t<te::ANGULAR_RATES, 2, 1>(p, a, n_work_hint);
      else if (v == 2) launch_split_t<te::ANGULAR_RATES, 1, 1, 1>(p, a, n_work_hint);
      else launch_split_t<te::ANGULAR_RATES, 1, 2, 1>(p, a, n_work_hint);
      break;
  }
}

te::StepArgs base_args(te_pool* p) {
  te::StepArgs a{};
  Buf& b = p->buf[p->cur];
  a.tiles = b.tiles;
  a.n_slots = (int)p->n;
  a.n_tiles = cdiv(p->n, te::TILE);
  a.cls = b.cold.cls;
  a.Qtab = p->dQ;
  a.Rtab = p->dR;
  a.n_ticks = 1;
  a.cls_c = -1;
  if (!p->hQ.empty()) {   // class 0 rides in the parameter constant bank
    a.cls_c = 0;
    std::memcpy(a.Qc, p->hQ[0].data(), sizeof(double) * p->N * p->N);
    std::memcpy(a.Rc, p->hR[0].data(), sizeof(double) * p->M * p->M);
  }
  return a;
}

void check_meas_stride(te_pool* p, int stride) {
  if (stride == 7) return;
  if (stride == 3 && p->M == 3) return;
  throw std::invalid_argument("meas_stride must be 7 (pose) or 3 (xyz, UV/UA pools only)");
}

// ---- rebuild (stable gather of survivors + init of new targets) ----------------------------
template <int TYPE>
void rebuild_t(te_pool* p, int n_new, const te::AddData& ad) {
  Buf& ob = p->buf[p->cur];
  Buf& nb = p->buf[1 - p->cur];
  const dim3 grid(cdiv(n_new, 128), (te::Layout<TYPE>::NF + te::REBUILD_FPT - 1) / te::REBUILD_FPT);
  te::rebuild_kernel<TYPE><<<grid, 128, 0, p->stream>>>(n_new, p->srcmap, ob.tiles, ob.cold, nb.tiles, nb.cold, ad, p->dP0);
  CK(cudaGetLastError());
}
void rebuild(te_pool* p, int n_new, const te::AddData& ad) {
  switch (p->model) {
    case te::UNIFORM_VELOCITY: rebuild_t<te::UNIFORM_VELOCITY>(p, n_new, ad); break;
    case te::UNIFORM_ACCELERATION: rebuild_t<te::UNIFORM_ACCELERATION>(p, n_new, ad); break;
    case te::ANGULAR_VELOCITIES: rebuild_t<te::ANGULAR_VELOCITIES>(p, n_new, ad); break;
    default: rebuild_t<te::ANGULAR_RATES>(p, n_new, ad); break;
  }
}
template <int TYPE>
void init_append_t(te_pool* p, int base, const te::AddData& ad, long long n) {
  Buf& b = p->buf[p->cur];
  te::init_append_kernel<TYPE><<<cdiv(n, 128), 128, 0, p->stream>>>(b.tiles, b.cold, base, ad, n, p->dP0);
  CK(cudaGetLastError());
}
void init_append(te_pool* p, int base, const te::AddData& ad, long long n) {
  switch (p->model) {
    case te::UNIFORM_VELOCITY: init_append_t<te::UNIFORM_VELOCITY>(p, base, ad, n); break;
    case te::UNIFORM_ACCELERATION: init_append_t<te::UNIFORM_ACCELERATION>(p, base, ad, n); break;
    case te::ANGULAR_VELOCITIES: init_append_t<te::ANGULAR_VELOCITIES>(p, base, ad, n); break;
    default: init_append_t<te::ANGULAR_RATES>(p, base, ad, n); break;
  }
}

template <int TYPE>
void init_promoted_t(te_pool* p, int n_add, const int* new_dst, const Buf& nb, const te::AddData& ad, const te::MailArrays& mb) {
  te::init_promoted_kernel<TYPE><<<cdiv(n_add, 128), 128, 0, p->stream>>>(n_add, new_dst, nb.tiles, nb.cold, ad, p->mb_add, mb, p->dP0, p->action,
                                                                          p->tile_flag, p->tile_list, p->d_counters);
  CK(cudaGetLastError());
}
void init_promoted(te_pool* p, int n_add, const int* new_dst, const Buf& nb, const te::AddData& ad, const te::MailArrays& mb) {
  switch (p->model) {
    case te::UNIFORM_VELOCITY: init_promoted_t<te::UNIFORM_VELOCITY>(p, n_add, new_dst, nb, ad, mb); break;
    case te::UNIFORM_ACCELERATION: init_promoted_t<te::UNIFORM_ACCELERATION>(p, n_add, new_dst, nb, ad, mb); break;
    case te::ANGULAR_VELOCITIES: init_promoted_t<te::ANGULAR_VELOCITIES>(p, n_add, new_dst, nb, ad, mb); break;
    default: init_promoted_t<te::ANGULAR_RATES>(p, n_add, new_dst, nb, ad, mb); break;
  }
}

// after a compaction: the largest surviving id (4 bytes; the callers synchronise the stream before they return)
void fetch_last_id(te_pool* p) {
  p->h_last_valid = false;
  if (p->n > 0) {
    CK(cudaMemcpyAsync(&p->h_last_id, p->buf[p->cur].cold.ids + p->n - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->h_last_valid = true;
  }
}

// alive[] (n_old entries) is set on the device; compacts survivors, merges `n_add` sorted new ids.
// Returns the number of survivors.
int compact_and_merge(te_pool* p, const te::AddData& ad, const uint32_t* d_add_ids, int n_add, uint32_t* d_erased /*or null*/) {
  const int n_old = (int)p->n;
  int total_alive = 0;
  if (n_old > 0) {
    size_t tmp = p->cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(p->cub_tmp, tmp, p->alive, p->pos, n_old, p->stream));
    int last_pos = 0, last_alive = 0;
    CK(cudaMemcpyAsync(&last_pos, p->pos + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&last_alive, p->alive + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    total_alive = last_pos + last_alive;
  }
  const int n_new = total_alive + n_add;
  if (d_erased && n_old > total_alive) {
    te::collect_erased_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(p->alive, p->pos, p->buf[p->cur].cold.ids, n_old, d_erased);
    CK(cudaGetLastError());
  }
  if (n_new == n_old && n_add == 0) return total_alive;   // nothing erased, nothing added
  ensure_other_capacity(p, (size_t)n_new);
  if (n_old > 0) {
    te::map_existing_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(n_old, p->alive, p->pos, p->buf[p->cur].cold.ids, d_add_ids, n_add, p->srcmap);
    CK(cudaGetLastError());
  }
  if (n_add > 0) {
    te::map_new_kernel<<<cdiv(n_add, 256), 256, 0, p->stream>>>(n_add, d_add_ids, p->buf[p->cur].cold.ids, n_old, p->pos, total_alive, p->srcmap);
    CK(cudaGetLastError());
  }
  if (n_new > 0) rebuild(p, n_new, ad);
  if (p->mb_on) {   // the mailboxes follow their slots; promoted host mailboxes are filled in (after rebuild: it zeroes last_meas of new slots)
    ensure_mail_other(p, (size_t)n_new);
    if (n_new > 0) {
      te::mb_move_kernel<<<cdiv(n_new, 256), 256, 0, p->stream>>>(n_new, p->srcmap, p->mb[p->mb_cur].a, p->mb[1 - p->mb_cur].a, p->mb_add, ad.p0,
                                                                  p->buf[1 - p->cur].cold.last_meas);
      CK(cudaGetLastError());
    }
    p->mb_cur = 1 - p->mb_cur;
  }
  p->cur = 1 - p->cur;
  p->n = n_new;
  p->h_ids_valid = false;
  fetch_last_id(p);
  return total_alive;
}

int* lookup_slots(te_pool* p, const uint32_t* d_ids, long long n) {
  int* slots = p->arena.get_n<int>((size_t)n);
  te::lookup_slots_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.ids, (int)p->n, d_ids, n, slots);
  CK(cudaGetLastError());
  return slots;
}

// queued records of unknown ids -> their (host) mailboxes, in arrival order
void fold_pending(te_pool* p) {
  for (const PendingRec& r : p->pending) apply_record(p->orphans[r.id], r);
  p->pending.clear();
}

// A target erased by hand keeps its mailbox in the reference (TargetManager::erase does not know the adapter's map,
// src/target_manager.cpp:227-241): the mailboxes of the listed slots move to the host's target-less map before the compaction.
void demote_mailboxes(te_pool* p, const uint32_t* ids, const int* d_slots, long long n) {
  uint32_t* d_sec = p->arena.get_n<uint32_t>((size_t)n);
  uint32_t* d_nsec = p->arena.get_n<uint32_t>((size_t)n);
  uint8_t* d_act = p->arena.get_n<uint8_t>((size_t)n);
  double* d_last = p->arena.get_n<double>((size_t)n);
  double* d_pose = p->arena.get_n<double>((size_t)n * 7);
  te::mb_gather_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>((int)n, d_slots, p->mb[p->mb_cur].a, p->buf[p->cur].cold.last_meas, d_sec, d_nsec, d_act,
                                                            d_last, d_pose);
  CK(cudaGetLastError());
  std::vector<uint32_t> sec((size_t)n), nsec((size_t)n);
  std::vector<uint8_t> act((size_t)n);
  std::vector<double> last((size_t)n), pose((size_t)n * 7);
  CK(cudaMemcpyAsync(sec.data(), d_sec, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(nsec.data(), d_nsec, (size_t)n * 4, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(act.data(), d_act, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(last.data(), d_last, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(pose.data(), d_pose, (size_t)n * 56, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  for (long long k = 0; k < n; ++k) {
    if (act[(size_t)k] == 0xFF || act[(size_t)k] == (uint8_t)TE_ACT_NONE) continue;   // unknown id / target without a mailbox
    HostMail& m = p->orphans[ids[k]];
    m.sec = sec[(size_t)k];
    m.nsec = nsec[(size_t)k];
    m.last = last[(size_t)k];
    m.fresh = act[(size_t)k] == (uint8_t)TE_ACT_UPDATE;
    std::memcpy(m.pose, &pose[(size_t)k * 7], sizeof(m.pose));
  }
}
// ... and a target created by hand for an id that already has a (target-less) mailbox is fed by it from the next tick on
void attach_mailboxes(te_pool* p, const uint32_t* ids, long long n) {
  fold_pending(p);
  std::vector<uint32_t> a_ids, sec, nsec;
  std::vector<uint8_t> act;
  std::vector<double> last, pose;
  for (long long k = 0; k < n; ++k) {
    auto it = p->orphans.find(ids[k]);
    if (it == p->orphans.end()) continue;
    const HostMail& m = it->second;
    a_ids.push_back(ids[k]);
    sec.push_back(m.sec);
    nsec.push_back(m.nsec);
    act.push_back((uint8_t)(m.fresh ? TE_ACT_UPDATE : TE_ACT_PREDICT));
    last.push_back(m.last);
    pose.insert(pose.end(), m.pose, m.pose + 7);
    p->orphans.erase(it);
  }
  const long long na = (long long)a_ids.size();
  if (na == 0) return;
  uint32_t* d_ids = to_dev(p, a_ids.data(), (size_t)na);
  int* slots = lookup_slots(p, d_ids, na);
  te::mb_scatter_kernel<<<cdiv(na, 256), 256, 0, p->stream>>>((int)na, slots, to_dev(p, sec.data(), (size_t)na), to_dev(p, nsec.data(), (size_t)na),
                                                              to_dev(p, act.data(), (size_t)na), to_dev(p, last.data(), (size_t)na),
                                                              to_dev(p, pose.data(), (size_t)na * 7), p->mb[p->mb_cur].a,
                                                              p->buf[p->cur].cold.last_meas);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(p->stream));   // the host vectors above go out of scope
}

template <class F> int guarded(te_pool* p, F&& f) {
  try {
    if (!p) throw std::invalid_argument("null pool");
    DeviceGuard g(p->device);
    p->arena.reset();
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
template <class F> long long guarded_ll(te_pool* p, F&& f) {
  try {
    if (!p) throw std::invalid_argument("null pool");
    DeviceGuard g(p->device);
    p->arena.reset();
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // namespace

extern "C" {

const char* te_last_error(void) { return g_err.c_str(); }

int te_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int te_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return 0;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess) { cudaGetLastError(); g_err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return -1; }
  return 0;
}
int te_host_unregister(void* ptr) {
  if (!ptr) return 0;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); g_err = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return -1; }
  return 0;
}

int te_model_dims(int model, int* n, int* m) {
  if (model < 0 || model > 3) return -1;
  if (n) *n = te::model_n(model);
  if (m) *m = te::model_m(model);
  return 0;
}

size_t te_model_bytes_per_step(int model) {
  // read x, P, measurement (+prev rpy), write x, P (+prev rpy), + 32 B of t / n_meas
  switch (model) {
    case TE_UNIFORM_VELOCITY: return 728;
    case TE_UNIFORM_ACCELERATION: return 1496;
    case TE_ANGULAR_VELOCITIES: return 2632;
    case TE_ANGULAR_RATES: return 5608;
    default: return 0;
  }
}

te_pool* te_pool_create(int model, int device, void* cuda_stream) {
  te_pool* p = nullptr;
  try {
    if (model < 0 || model > 3) throw std::invalid_argument("unknown model type");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      throw std::runtime_error("no CUDA device available: the target pool has no CPU fallback");
    }
    if (device < 0 || device >= ndev) throw std::invalid_argument("bad device index");
    DeviceGuard g(device);
    p = new te_pool();
    p->model = model;
    p->device = device;
    p->N = te::model_n(model);
    p->M = te::model_m(model);
    p->NF = te::model_nf(model);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw std::runtime_error("te_pool needs an sm_100a (Blackwell) device");
    p->n_sm = prop.multiProcessorCount;
    if (cuda_stream) {
      p->stream = (cudaStream_t)cuda_stream;
    } else {
      CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
      p->own_stream = true;
    }
    CK(cudaMalloc(&p->d_counters, 4 * sizeof(int)));
    CK(cudaMemsetAsync(p->d_counters, 0, 4 * sizeof(int), p->stream));
    if (const char* cap = std::getenv("TE_GRID_CAP")) p->grid_cap = std::max(0, std::atoi(cap));   // test hook, see te_pool_set_grid_cap
    return p;
  } catch (const std::exception& e) {
    g_err = e.what();
    delete p;
    return nullptr;
  }
}

void te_pool_destroy(te_pool* p) {
  if (!p) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);
  free_buf(p->buf[0]);
  free_buf(p->buf[1]);
  free_mail(p->mb[0]);
  free_mail(p->mb[1]);
  if (p->h_stage) cudaFreeHost(p->h_stage);
  cudaFree(p->action); cudaFree(p->dt_slot); cudaFree(p->tile_flag); cudaFree(p->tile_list);
  cudaFree(p->alive); cudaFree(p->pos); cudaFree(p->srcmap); cudaFree(p->d_counters); cudaFree(p->cub_tmp);
  cudaFree(p->dQ); cudaFree(p->dR); cudaFree(p->dP0);
  p->arena.destroy();
  for (cudaEvent_t e : p->events) cudaEventDestroy(e);
  if (p->h2d_stream) cudaStreamDestroy(p->h2d_stream);
  if (p->d2h_stream) cudaStreamDestroy(p->d2h_stream);
  if (p->own_stream) cudaStreamDestroy(p->stream);
  cudaSetDevice(prev);
  delete p;
}

int te_pool_set_stream(te_pool* p, void* cuda_stream) {
  return guarded(p, [&] {
    CK(cudaStreamSynchronize(p->stream));
    if (p->own_stream) cudaStreamDestroy(p->stream);
    p->own_stream = false;
    if (cuda_stream) {
      p->stream = (cudaStream_t)cuda_stream;
    } else {
      CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
      p->own_stream = true;
    }
    return 0;
  });
}

int te_pool_sync(te_pool* p) {
  return guarded(p, [&] {
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

int te_pool_set_variant(te_pool* p, int variant) {
  if (!p) return -1;
  p->variant = variant;
  return 0;
}

int te_pool_set_grid_cap(te_pool* p, int max_ctas) {
  if (!p || max_ctas < 0) return -1;
  p->grid_cap = max_ctas;
  return 0;
}

int te_pool_reserve(te_pool* p, size_t n_targets) {
  return guarded(p, [&] {
    if (!p->buf[p->cur].tiles) {
      alloc_buf(p, p->buf[p->cur], n_targets);
      ensure_work(p, p->buf[p->cur].cap);
    } else {
      ensure_cur_capacity(p, n_targets);
    }
    ensure_work(p, std::max(n_targets, p->buf[p->cur].cap));   // (a flipped buffer can be larger than the work arrays)
    // the second generation too (every compaction -- erase, merge-add, expiry -- gathers into it): no tick pays for a
    // multi-gigabyte allocation later
    ensure_other_capacity(p, n_targets);
    if (p->mb_on) {
      ensure_mail_cur(p, n_targets);
      ensure_mail_other(p, n_targets);
    }
    return 0;
  });
}

long long te_pool_size(te_pool* p) { return p ? p->n : -1; }

size_t te_pool_device_bytes(te_pool* p) {
  if (!p) return 0;
  size_t b = 0;
  for (int i = 0; i < 2; ++i)
    if (p->buf[i].tiles) b += p->buf[i].cap / te::TILE * tile_doubles(p) * 8 + p->buf[i].cap * (4 + 2 + 8 + 56);
  b += p->wcap * (1 + 8 + 4 + 4 + 4) + p->wcap / te::TILE * 5 + p->cub_bytes;
  return b;
}

int te_pool_register_class(te_pool* p, const double* Q, const double* R, const double* P0) {
  return guarded(p, [&] {
    if (!Q || !R || !P0) throw std::invalid_argument("null model matrix");
    const size_t nn = (size_t)p->N * p->N, mm = (size_t)p->M * p->M;
    for (size_t c = 0; c < p->hQ.size(); ++c)
      if (!std::memcmp(p->hQ[c].data(), Q, nn * 8) && !std::memcmp(p->hR[c].data(), R, mm * 8) && !std::memcmp(p->hP0[c].data(), P0, nn * 8))
        return (int)c;
    if (p->hQ.size() >= 65535) throw std::runtime_error("too many model classes (max 65535)");
    p->hQ.emplace_back(Q, Q + nn);
    p->hR.emplace_back(R, R + mm);
    p->hP0.emplace_back(P0, P0 + nn);
    auto sym = [](const double* A, int n) {
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
          if (std::memcmp(&A[i * n + j], &A[j * n + i], 8) != 0) return false;
      return true;
    };
    if (!sym(Q, p->N) || !sym(R, p->M) || !sym(P0, p->N)) p->all_sym = false;
    upload_classes(p);
    return (int)p->hQ.size() - 1;
  });
}

int te_pool_class_count(te_pool* p) { return p ? (int)p->hQ.size() : -1; }

int te_pool_get_class(te_pool* p, int cls, double* Q, double* R, double* P0) {
  if (!p || cls < 0 || cls >= (int)p->hQ.size()) return -1;
  if (Q) std::memcpy(Q, p->hQ[cls].data(), p->hQ[cls].size() * 8);
  if (R) std::memcpy(R, p->hR[cls].data(), p->hR[cls].size() * 8);
  if (P0) std::memcpy(P0, p->hP0[cls].data(), p->hP0[cls].size() * 8);
  return 0;
}

long long te_pool_add_batch(te_pool* p, long long n, const uint32_t* ids, const uint16_t* cls, const double* t0, const double* p0,
                            const double* v0, const double* a0, const double* p0_scale) {
  return guarded_ll(p, [&]() -> long long {
    if (n <= 0) return 0;
    if (!ids || !p0) throw std::invalid_argument("ids and p0 are required");
    if (p->hQ.empty()) throw std::runtime_error("no model class registered");
    // order the batch by id (first occurrence wins), drop ids that already exist
    std::vector<long long> ord((size_t)n);
    std::iota(ord.begin(), ord.end(), 0LL);
    bool sorted = true;
    for (long long k = 1; k < n && sorted; ++k) sorted = ids[k - 1] < ids[k];
    if (!sorted) std::stable_sort(ord.begin(), ord.end(), [&](long long a, long long b) { return ids[a] < ids[b]; });
    std::vector<long long> keep;
    keep.reserve((size_t)n);
    if (p->n == 0) { p->h_ids.clear(); p->h_ids_valid = true; }
    if (p->n > 0 && !p->h_ids_valid && !p->h_last_valid) sync_host_ids(p);
    const bool append_only = p->n == 0 || ids[ord[0]] > (p->h_ids_valid ? p->h_ids.back() : p->h_last_id);
    if (!append_only) sync_host_ids(p);   // membership test and merge need the whole sorted id array
    for (long long k = 0; k < n; ++k) {
      const long long s = ord[k];
      if (!keep.empty() && ids[keep.back()] == ids[s]) continue;
      if (!append_only && std::binary_search(p->h_ids.begin(), p->h_ids.end(), ids[s])) continue;   // "already exists"
      if (cls && cls[s] >= p->hQ.size()) throw std::invalid_argument("unknown model class in add batch");
      keep.push_back(s);
    }
    const long long na = (long long)keep.size();
    if (na == 0) return 0;
    const bool identity = sorted && na == n;
    // gather payload in id order
    std::vector<uint32_t> g_ids;
    std::vector<uint16_t> g_cls;
    std::vector<double> g_t0, g_p0, g_v0, g_a0, g_sc;
    const uint32_t* s_ids = ids;
    const uint16_t* s_cls = cls;
    const double *s_t0 = t0, *s_p0 = p0, *s_v0 = v0, *s_a0 = a0, *s_sc = p0_scale;
    if (!identity) {
      g_ids.resize(na);
      for (long long k = 0; k < na; ++k) g_ids[k] = ids[keep[k]];
      s_ids = g_ids.data();
      auto gather = [&](const double* src, int w, std::vector<double>& dst) -> const double* {
        if (!src) return nullptr;
        dst.resize((size_t)na * w);
        for (long long k = 0; k < na; ++k) std::memcpy(&dst[(size_t)k * w], src + (size_t)keep[k] * w, w * 8);
        return dst.data();
      };
      s_t0 = gather(t0, 1, g_t0);
      s_p0 = gather(p0, 7, g_p0);
      s_v0 = gather(v0, 6, g_v0);
      s_a0 = gather(a0, 6, g_a0);
      s_sc = gather(p0_scale, 1, g_sc);
      if (cls) {
        g_cls.resize(na);
        for (long long k = 0; k < na; ++k) g_cls[k] = cls[keep[k]];
        s_cls = g_cls.data();
      }
    }
    te::AddData ad{};
    ad.ids = to_dev(p, s_ids, na);
    ad.cls = to_dev(p, s_cls, na);
    ad.t0 = to_dev(p, s_t0, na);
    ad.p0 = to_dev(p, s_p0, na * 7);
    ad.v0 = to_dev(p, s_v0, na * 6);
    ad.a0 = to_dev(p, s_a0, na * 6);
    ad.scale = to_dev(p, s_sc, na);
    if (append_only) {
      if (!p->buf[p->cur].tiles) {
        alloc_buf(p, p->buf[p->cur], (size_t)na);
        ensure_work(p, p->buf[p->cur].cap);
      } else {
        ensure_cur_capacity(p, (size_t)(p->n + na));
      }
      init_append(p, (int)p->n, ad, na);
      if (p->mb_on) {
        ensure_mail_cur(p, (size_t)(p->n + na));
        te::mb_clear_kernel<<<cdiv(na, 256), 256, 0, p->stream>>>(p->mb[p->mb_cur].a, (int)p->n, (int)na);
        CK(cudaGetLastError());
      }
      p->n += na;
      if (p->h_ids_valid) p->h_ids.insert(p->h_ids.end(), s_ids, s_ids + na);
      p->h_last_id = s_ids[na - 1];
      p->h_last_valid = true;
    } else {
      ensure_work(p, (size_t)(p->n + na));
      te::fill_i32_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->alive, (int)p->n, 1);
      CK(cudaGetLastError());
      compact_and_merge(p, ad, ad.ids, (int)na, nullptr);
      std::vector<uint32_t> merged(p->h_ids.size() + (size_t)na);
      std::merge(p->h_ids.begin(), p->h_ids.end(), s_ids, s_ids + na, merged.begin());
      p->h_ids.swap(merged);
      p->h_ids_valid = true;
    }
    if (p->mb_on && (!p->orphans.empty() || !p->pending.empty())) attach_mailboxes(p, s_ids, na);
    CK(cudaStreamSynchronize(p->stream));   // host payload vectors go out of scope
    return na;
  });
}

long long te_pool_erase_batch(te_pool* p, long long n, const uint32_t* ids) {
  return guarded_ll(p, [&]() -> long long {
    if (n <= 0 || p->n == 0) return 0;
    if (!ids) throw std::invalid_argument("null ids");
    ensure_work(p, (size_t)p->n);
    uint32_t* d_ids = to_dev(p, ids, n);
    int* slots = lookup_slots(p, d_ids, n);
    if (p->mb_on) demote_mailboxes(p, ids, slots, n);
    te::fill_i32_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->alive, (int)p->n, 1);
    te::clear_listed_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(p->alive, slots, n);
    CK(cudaGetLastError());
    const long long n_old = p->n;
    te::AddData ad{};
    const int alive = compact_and_merge(p, ad, nullptr, 0, nullptr);
    return n_old - alive;
  });
}

long long te_pool_ids(te_pool* p, uint32_t* out, long long cap) {
  return guarded_ll(p, [&]() -> long long {
    sync_host_ids(p);
    if (out) std::memcpy(out, p->h_ids.data(), (size_t)std::min<long long>(cap, p->n) * sizeof(uint32_t));
    return p->n;
  });
}

int te_pool_contains(te_pool* p, uint32_t id) {
  return guarded(p, [&] {
    sync_host_ids(p);
    return std::binary_search(p->h_ids.begin(), p->h_ids.end(), id) ? 1 : 0;
  });
}

int te_pool_class_of(te_pool* p, uint32_t id) {
  return guarded(p, [&] {
    sync_host_ids(p);
    auto it = std::lower_bound(p->h_ids.begin(), p->h_ids.end(), id);
    if (it == p->h_ids.end() || *it != id) throw std::invalid_argument("unknown target id");
    uint16_t c = 0;
    CK(cudaMemcpyAsync(&c, p->buf[p->cur].cold.cls + (it - p->h_ids.begin()), sizeof(uint16_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return (int)c;
  });
}

int te_pool_step_dense(te_pool* p, double dt, const double* dev_meas, int meas_stride, const uint8_t* dev_action, int default_action) {
  return guarded(p, [&] {
    if (p->n == 0) return 0;
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0 (assert of src/target_interface.cpp:150)");
    te::StepArgs a = base_args(p);
    a.dt = dt;
    a.meas = dev_meas;
    a.meas_stride = meas_stride;
    a.action = const_cast<uint8_t*>(dev_action);
    a.default_action = default_action;
    if (dev_meas) {
      check_meas_stride(p, meas_stride);
      a.meas_tma = ((uintptr_t)dev_meas % 16 == 0) ? 1 : 0;
    } else if (dev_action || default_action == TE_ACT_UPDATE) {
      // (an action array may name ACT_UPDATE for any slot: without measurements the kernel would read a null pointer)
      throw std::invalid_argument("update tick without measurements");
    }
    launch_step(p, a, a.n_tiles);
    return 0;
  });
}

int te_pool_step_dense_ticks(te_pool* p, int n_ticks, double dt, const double* dev_meas, int meas_stride, const uint8_t* dev_action,
                             int default_action) {
  return guarded(p, [&] {
    if (p->n == 0 || n_ticks <= 0) return 0;
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
    if (dev_meas) check_meas_stride(p, meas_stride);
    else if (default_action == TE_ACT_UPDATE || dev_action) throw std::invalid_argument("update ticks without measurements");
    te::StepArgs a = base_args(p);
    a.dt = dt;
    a.meas = dev_meas;
    a.meas_stride = meas_stride;
    a.meas_tma = 0;
    a.action = const_cast<uint8_t*>(dev_action);
    a.default_action = default_action;
    a.n_ticks = n_ticks;
    a.meas_tick_stride = (long long)p->n * meas_stride;
    a.action_tick_stride = p->n;
    launch_step_multi(p, a, a.n_tiles);
    return 0;
  });
}

int te_pool_step_dense_host(te_pool* p, double dt, const double* meas, int meas_stride, const uint8_t* action, int default_action) {
  return guarded(p, [&] {
    if (p->n == 0) return 0;
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
    te::StepArgs a = base_args(p);
    a.dt = dt;
    a.default_action = default_action;
    if (meas) {
      check_meas_stride(p, meas_stride);
      if (meas_stride == 7) {
        // measured_pose_ = meas (src/target_interface.cpp:142-146): land the batch in the pool's own record
        // for UPDATE slots only -> stage, then the kernel-side copy would cost a pass; instead keep the
        // staged batch as this tick's measurement block and refresh measured_pose_ with one D2D copy
        // when every slot is updated.
        double* d = to_dev(p, meas, (size_t)p->n * 7);
        a.meas = d;
        if (!action && default_action == TE_ACT_UPDATE)
          CK(cudaMemcpyAsync(p->buf[p->cur].cold.meas, d, (size_t)p->n * 7 * 8, cudaMemcpyDeviceToDevice, p->stream));
      } else {
        a.meas = to_dev(p, meas, (size_t)p->n * meas_stride);
      }
      a.meas_stride = meas_stride;
      a.meas_tma = 1;
    } else if (!action && default_action == TE_ACT_UPDATE) {
      throw std::invalid_argument("update tick without measurements");
    }
    if (action) a.action = to_dev(p, action, (size_t)p->n);
    launch_step(p, a, a.n_tiles);
    if (action && meas && meas_stride == 7) {
      // masked refresh of measured_pose_ for the UPDATE slots
      te::copy_meas_masked_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.meas, a.meas, a.action, (int)p->n);
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

int te_pool_tick_host(te_pool* p, double dt, const double* meas, int meas_stride, const uint8_t* action, int default_action,
                      double* est_pos_out) {
  return guarded(p, [&] {
    if (p->n == 0) return 0;
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0");
    if (meas) check_meas_stride(p, meas_stride);
    else if (!action && default_action == TE_ACT_UPDATE) throw std::invalid_argument("update tick without measurements");
    if (!p->h2d_stream) {
      CK(cudaStreamCreateWithFlags(&p->h2d_stream, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&p->d2h_stream, cudaStreamNonBlocking));
    }
    const long long n = p->n;
    const int n_tiles = cdiv(n, te::TILE);
    const int chunk_tiles = std::max(256, std::min(n_tiles, 8192));   // 262144 targets: 14.7 MB of pose measurements
    const int n_chunks = cdiv(n_tiles, chunk_tiles);
    while ((int)p->events.size() < 2 * n_chunks + 1) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      p->events.push_back(e);
    }
    double* d_meas = meas ? p->arena.get_n<double>((size_t)n * meas_stride) : nullptr;
    uint8_t* d_act = action ? p->arena.get_n<uint8_t>((size_t)n) : nullptr;
    double* d_pos = est_pos_out ? p->arena.get_n<double>((size_t)n * 3) : nullptr;
    // staging buffers may have been carved by earlier work on the pool stream
    CK(cudaEventRecord(p->events[2 * n_chunks], p->stream));
    CK(cudaStreamWaitEvent(p->h2d_stream, p->events[2 * n_chunks], 0));
    te::StepArgs a = base_args(p);
    a.dt = dt;
    a.meas = d_meas;
    a.meas_stride = meas_stride;
    a.meas_tma = d_meas ? 1 : 0;
    a.action = d_act;
    a.default_action = default_action;
    a.pos_out = d_pos;
    for (int c = 0; c < n_chunks; ++c) {
      const long long s0 = (long long)c * chunk_tiles * te::TILE;
      const long long s1 = std::min<long long>(n, s0 + (long long)chunk_tiles * te::TILE);
      if (d_meas) CK(cudaMemcpyAsync(d_meas + s0 * meas_stride, meas + s0 * meas_stride, (size_t)(s1 - s0) * meas_stride * 8, cudaMemcpyHostToDevice, p->h2d_stream));
      if (d_act) CK(cudaMemcpyAsync(d_act + s0, action + s0, (size_t)(s1 - s0), cudaMemcpyHostToDevice, p->h2d_stream));
      CK(cudaEventRecord(p->events[2 * c], p->h2d_stream));
      CK(cudaStreamWaitEvent(p->stream, p->events[2 * c], 0));
      a.tile_begin = c * chunk_tiles;
      a.n_tiles = std::min(chunk_tiles, n_tiles - c * chunk_tiles);
      launch_step(p, a, a.n_tiles);
      if (d_pos) {
        CK(cudaEventRecord(p->events[2 * c + 1], p->stream));
        CK(cudaStreamWaitEvent(p->d2h_stream, p->events[2 * c + 1], 0));
        CK(cudaMemcpyAsync(est_pos_out + s0 * 3, d_pos + s0 * 3, (size_t)(s1 - s0) * 24, cudaMemcpyDeviceToHost, p->d2h_stream));
      }
    }
    if (d_meas && meas_stride == 7) {   // measured_pose_ = meas for the updated slots (src/target_interface.cpp:142-146)
      if (d_act) te::copy_meas_masked_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.meas, d_meas, d_act, (int)n);
      else if (default_action == TE_ACT_UPDATE)
        CK(cudaMemcpyAsync(p->buf[p->cur].cold.meas, d_meas, (size_t)n * 56, cudaMemcpyDeviceToDevice, p->stream));
      CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(p->d2h_stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

long long te_pool_step_ids(te_pool* p, long long n, const uint32_t* ids, const double* dt, double dt_scalar, const double* meas,
                           const uint8_t* action) {
  return guarded_ll(p, [&]() -> long long {
    if (n <= 0 || p->n == 0) return 0;
    if (!ids) throw std::invalid_argument("null ids");
    if (!dt && !(dt_scalar >= 0.0)) throw std::invalid_argument("dt must be >= 0");
    ensure_work(p, (size_t)p->n);
    Buf& b = p->buf[p->cur];
    uint32_t* d_ids = to_dev(p, ids, n);
    double* d_dt = to_dev(p, dt, n);
    double* d_meas = to_dev(p, meas, n * 7);
    uint8_t* d_act = to_dev(p, action, n);
    if (!meas) {
      bool needs = !action;
      if (action) for (long long k = 0; k < n && !needs; ++k) needs = action[k] == TE_ACT_UPDATE;
      if (needs) throw std::invalid_argument("update ops without measurements");
    }
    CK(cudaMemsetAsync(p->d_counters, 0, 2 * sizeof(int), p->stream));
    te::scatter_ops_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(b.cold.ids, (int)p->n, n, d_ids, d_dt, dt_scalar, d_meas, d_act, p->action,
                                                                 p->dt_slot, b.cold.meas, p->tile_flag, p->tile_list, p->d_counters);
    CK(cudaGetLastError());
    te::StepArgs a = base_args(p);
    a.tile_list = p->tile_list;
    a.d_nwork = p->d_counters;
    a.dt = dt_scalar;
    a.dt_slot = p->dt_slot;
    a.meas = b.cold.meas;
    a.meas_stride = 7;
    a.meas_tma = 1;
    a.action = p->action;
    a.default_action = TE_ACT_NONE;
    a.clear_action = 1;
    a.tile_flag = p->tile_flag;
    launch_step(p, a, (int)std::min<long long>(n, a.n_tiles));
    int applied = 0;
    CK(cudaMemcpyAsync(&applied, p->d_counters + 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return applied;
  });
}

int te_pool_predict_all(te_pool* p, double dt) { return te_pool_step_dense(p, dt, nullptr, 7, nullptr, TE_ACT_PREDICT); }

int te_pool_read_state(te_pool* p, long long n, const uint32_t* ids, double* x, double* P, double* t, long long* n_meas, double* prev_rpy,
                       double* measured_pose) {
  return guarded(p, [&] {
    if (n <= 0) return 0;
    if (!ids && n != p->n) throw std::invalid_argument("ids == NULL requires n == pool size");
    const int N = p->N;
    int* slots = nullptr;
    if (ids) slots = lookup_slots(p, to_dev(p, ids, n), n);
    double* dx = x ? p->arena.get_n<double>((size_t)n * N) : nullptr;
    double* dP = P ? p->arena.get_n<double>((size_t)n * N * N) : nullptr;
    double* dt_ = t ? p->arena.get_n<double>((size_t)n) : nullptr;
    long long* dn = n_meas ? p->arena.get_n<long long>((size_t)n) : nullptr;
    double* dprev = prev_rpy ? p->arena.get_n<double>((size_t)n * 3) : nullptr;
    double* dmp = measured_pose ? p->arena.get_n<double>((size_t)n * 7) : nullptr;
    if (dx) CK(cudaMemsetAsync(dx, 0, (size_t)n * N * 8, p->stream));
    if (dP) CK(cudaMemsetAsync(dP, 0, (size_t)n * N * N * 8, p->stream));
    if (dt_) CK(cudaMemsetAsync(dt_, 0, (size_t)n * 8, p->stream));
    if (dn) CK(cudaMemsetAsync(dn, 0, (size_t)n * 8, p->stream));
    if (dprev) CK(cudaMemsetAsync(dprev, 0, (size_t)n * 24, p->stream));
    if (dmp) CK(cudaMemsetAsync(dmp, 0, (size_t)n * 56, p->stream));
    Buf& b = p->buf[p->cur];
    const int g = cdiv(n, 128);
    switch (p->model) {
      case te::UNIFORM_VELOCITY: te::gather_state_kernel<te::UNIFORM_VELOCITY><<<g, 128, 0, p->stream>>>(b.tiles, b.cold, slots, n, dx, dP, dt_, dn, dprev, dmp, p->lower_stale ? 1 : 0); break;
      case te::UNIFORM_ACCELERATION: te::gather_state_kernel<te::UNIFORM_ACCELERATION><<<g, 128, 0, p->stream>>>(b.tiles, b.cold, slots, n, dx, dP, dt_, dn, dprev, dmp, p->lower_stale ? 1 : 0); break;
      case te::ANGULAR_VELOCITIES: te::gather_state_kernel<te::ANGULAR_VELOCITIES><<<g, 128, 0, p->stream>>>(b.tiles, b.cold, slots, n, dx, dP, dt_, dn, dprev, dmp, p->lower_stale ? 1 : 0); break;
      default: te::gather_state_kernel<te::ANGULAR_RATES><<<g, 128, 0, p->stream>>>(b.tiles, b.cold, slots, n, dx, dP, dt_, dn, dprev, dmp, p->lower_stale ? 1 : 0); break;
    }
    CK(cudaGetLastError());
    if (x) CK(cudaMemcpyAsync(x, dx, (size_t)n * N * 8, cudaMemcpyDeviceToHost, p->stream));
    if (P) CK(cudaMemcpyAsync(P, dP, (size_t)n * N * N * 8, cudaMemcpyDeviceToHost, p->stream));
    if (t) CK(cudaMemcpyAsync(t, dt_, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (n_meas) CK(cudaMemcpyAsync(n_meas, dn, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (prev_rpy) CK(cudaMemcpyAsync(prev_rpy, dprev, (size_t)n * 24, cudaMemcpyDeviceToHost, p->stream));
    if (measured_pose) CK(cudaMemcpyAsync(measured_pose, dmp, (size_t)n * 56, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

static void launch_gather_estimates(te_pool* p, const int* slots, long long n, const double* d_t1, double* pose, double* twist, double* acc,
                                    double* pose6, uint8_t* found, int rec13) {
  Buf& b = p->buf[p->cur];
  const int g = cdiv(n, 128);
  switch (p->model) {
    case te::UNIFORM_VELOCITY: te::gather_estimates_kernel<te::UNIFORM_VELOCITY><<<g, 128, 0, p->stream>>>(b.tiles, slots, n, d_t1, pose, twist, acc, pose6, found, rec13); break;
    case te::UNIFORM_ACCELERATION: te::gather_estimates_kernel<te::UNIFORM_ACCELERATION><<<g, 128, 0, p->stream>>>(b.tiles, slots, n, d_t1, pose, twist, acc, pose6, found, rec13); break;
    case te::ANGULAR_VELOCITIES: te::gather_estimates_kernel<te::ANGULAR_VELOCITIES><<<g, 128, 0, p->stream>>>(b.tiles, slots, n, d_t1, pose, twist, acc, pose6, found, rec13); break;
    default: te::gather_estimates_kernel<te::ANGULAR_RATES><<<g, 128, 0, p->stream>>>(b.tiles, slots, n, d_t1, pose, twist, acc, pose6, found, rec13); break;
  }
  CK(cudaGetLastError());
}

int te_pool_read_estimates(te_pool* p, long long n, const uint32_t* ids, const double* t1, double* pose7, double* twist6, double* acc6,
                           double* pose6_internal, uint8_t* found) {
  return guarded(p, [&] {
    if (n <= 0) return 0;
    if (!ids && n != p->n) throw std::invalid_argument("ids == NULL requires n == pool size");
    int* slots = nullptr;
    if (ids) slots = lookup_slots(p, to_dev(p, ids, n), n);
    const double* d_t1 = to_dev(p, t1, n);
    double* dpose = pose7 ? p->arena.get_n<double>((size_t)n * 7) : nullptr;
    double* dtw = twist6 ? p->arena.get_n<double>((size_t)n * 6) : nullptr;
    double* dac = acc6 ? p->arena.get_n<double>((size_t)n * 6) : nullptr;
    double* dp6 = pose6_internal ? p->arena.get_n<double>((size_t)n * 6) : nullptr;
    uint8_t* dfound = found ? p->arena.get_n<uint8_t>((size_t)n) : nullptr;
    if (dpose) CK(cudaMemsetAsync(dpose, 0, (size_t)n * 56, p->stream));
    if (dtw) CK(cudaMemsetAsync(dtw, 0, (size_t)n * 48, p->stream));
    if (dac) CK(cudaMemsetAsync(dac, 0, (size_t)n * 48, p->stream));
    if (dp6) CK(cudaMemsetAsync(dp6, 0, (size_t)n * 48, p->stream));
    launch_gather_estimates(p, slots, n, d_t1, dpose, dtw, dac, dp6, dfound, 0);
    if (pose7) CK(cudaMemcpyAsync(pose7, dpose, (size_t)n * 56, cudaMemcpyDeviceToHost, p->stream));
    if (twist6) CK(cudaMemcpyAsync(twist6, dtw, (size_t)n * 48, cudaMemcpyDeviceToHost, p->stream));
    if (acc6) CK(cudaMemcpyAsync(acc6, dac, (size_t)n * 48, cudaMemcpyDeviceToHost, p->stream));
    if (pose6_internal) CK(cudaMemcpyAsync(pose6_internal, dp6, (size_t)n * 48, cudaMemcpyDeviceToHost, p->stream));
    if (found) CK(cudaMemcpyAsync(found, dfound, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

int te_pool_estimates_dev(te_pool* p, double* dev_out) {
  return guarded(p, [&] {
    if (p->n == 0) return 0;
    if (!dev_out) throw std::invalid_argument("null output");
    launch_gather_estimates(p, nullptr, p->n, nullptr, dev_out, nullptr, nullptr, nullptr, nullptr, 1);
    return 0;
  });
}

const uint32_t* te_pool_dev_ids(te_pool* p) { return p ? p->buf[p->cur].cold.ids : nullptr; }

int te_pool_set_stamps(te_pool* p, long long n, const uint32_t* ids, const uint32_t* sec, const uint32_t* nsec) {
  return guarded(p, [&] {
    if (n <= 0 || p->n == 0) return 0;
    if (!ids || !sec || !nsec) throw std::invalid_argument("null stamp arrays");
    uint32_t* d_ids = to_dev(p, ids, n);
    uint32_t* d_sec = to_dev(p, sec, n);
    uint32_t* d_nsec = to_dev(p, nsec, n);
    Buf& b = p->buf[p->cur];
    te::set_stamps_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(b.cold.ids, (int)p->n, n, d_ids, d_sec, d_nsec, b.cold.last_meas);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

int te_pool_stamp_dense(te_pool* p, const uint8_t* dev_action, int default_action, uint32_t sec, uint32_t nsec) {
  return guarded(p, [&] {
    if (p->n == 0) return 0;
    volatile double ns = 1e-9 * (double)nsec;   // toSec (utils.hpp:59-62), never contracted
    const double stamp = (double)sec + ns;
    te::stamp_dense_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(dev_action, default_action, (int)p->n, stamp, p->buf[p->cur].cold.last_meas);
    CK(cudaGetLastError());
    return 0;
  });
}

long long te_pool_expire(te_pool* p, uint32_t now_sec, uint32_t now_nsec, double timeout, uint32_t* erased_out, long long cap) {
  return guarded_ll(p, [&]() -> long long {
    if (p->n == 0) return 0;
    ensure_work(p, (size_t)p->n);
    const long long n_old = p->n;
    // toSec on the host in the same non-contracted arithmetic (utils.hpp:59-62)
    volatile double ns = 1e-9 * (double)now_nsec;
    const double now = (double)now_sec + ns;
    te::expire_flags_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.last_meas, (int)n_old, now, timeout, p->alive);
    CK(cudaGetLastError());
    uint32_t* d_erased = p->arena.get_n<uint32_t>((size_t)n_old);
    te::AddData ad{};
    const int alive = compact_and_merge(p, ad, nullptr, 0, d_erased);
    const long long n_er = n_old - alive;
    if (n_er > 0 && erased_out && cap > 0)
      CK(cudaMemcpyAsync(erased_out, d_erased, (size_t)std::min(cap, n_er) * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return n_er;
  });
}

long long te_pool_step_dense_expire(te_pool* p, double dt, const double* dev_meas, int meas_stride, const uint8_t* dev_action,
                                    int default_action, uint32_t stamp_sec, uint32_t stamp_nsec, uint32_t now_sec, uint32_t now_nsec,
                                    double timeout, uint32_t* erased_out, long long cap) {
  return guarded_ll(p, [&]() -> long long {
    if (p->n == 0) return 0;
    if (p->mb_on) throw std::logic_error("this pool keeps device mailboxes: its ticks go through te_pool_mailbox_tick");
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0 (assert of src/target_interface.cpp:150)");
    if (dev_meas) check_meas_stride(p, meas_stride);
    else if (dev_action || default_action == TE_ACT_UPDATE) throw std::invalid_argument("update tick without measurements");
    const int n_old = (int)p->n;
    ensure_work(p, (size_t)n_old);
    Buf& ob = p->buf[p->cur];
    // 1. this tick's stamps, then the expiry predicate (both as in te_pool_stamp_dense / te_pool_expire)
    volatile double sns = 1e-9 * (double)stamp_nsec;
    const double stamp = (double)stamp_sec + sns;
    volatile double nns = 1e-9 * (double)now_nsec;
    const double now = (double)now_sec + nns;
    te::stamp_dense_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(dev_action, default_action, n_old, stamp, ob.cold.last_meas);
    CK(cudaGetLastError());
    te::expire_flags_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(ob.cold.last_meas, n_old, now, timeout, p->alive);
    CK(cudaGetLastError());
    size_t tmp = p->cub_bytes;
    CK(cub::DeviceScan::ExclusiveSum(p->cub_tmp, tmp, p->alive, p->pos, n_old, p->stream));
    int last_pos = 0, last_alive = 0;
    CK(cudaMemcpyAsync(&last_pos, p->pos + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&last_alive, p->alive + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const int n_alive = last_pos + last_alive;
    const long long n_er = n_old - n_alive;
    // 2. the step: in place when nobody expired, else into the compacted slots of the other buffer
    te::StepArgs a = base_args(p);
    a.dt = dt;
    a.meas = dev_meas;
    a.meas_stride = meas_stride;
    a.meas_tma = (dev_meas && (uintptr_t)dev_meas % 16 == 0) ? 1 : 0;
    a.action = const_cast<uint8_t*>(dev_action);
    a.default_action = default_action;
    if (n_er == 0) {
      launch_step(p, a, a.n_tiles);
      return 0;
    }
    uint32_t* d_erased = p->arena.get_n<uint32_t>((size_t)n_er);
    te::collect_erased_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(p->alive, p->pos, ob.cold.ids, n_old, d_erased);
    CK(cudaGetLastError());
    if (n_alive > 0) {
      ensure_other_capacity(p, (size_t)n_alive);
      Buf& nb = p->buf[1 - p->cur];
      te::compact_cold_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(n_old, p->alive, p->pos, p->buf[p->cur].cold, nb.cold);
      CK(cudaGetLastError());
      a.dst_tiles = nb.tiles;
      a.dst_alive = p->alive;
      a.dst_pos = p->pos;
      launch_step(p, a, a.n_tiles);
    }
    p->cur = 1 - p->cur;
    p->n = n_alive;
    p->h_ids_valid = false;
    fetch_last_id(p);
    if (erased_out && cap > 0)
      CK(cudaMemcpyAsync(erased_out, d_erased, (size_t)std::min(cap, n_er) * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return n_er;
  });
}

// ---- device-resident mailboxes: measurementCallBack + update(dt) of RosTargetManager ----------------
namespace {
// one /tf message into the mailboxes.  Host source (ids .. poses non-null): the arrays are staged here; device source (d_* given,
// host pointers null): the records are used in place and the few records of unknown ids are read back for the host's queue.
int mailbox_ingest_impl(te_pool* p, long long n, const uint32_t* ids, const uint32_t* sec, const uint32_t* nsec, const double* poses,
                        const uint32_t* d_ids, const uint32_t* d_sec, const uint32_t* d_nsec, const double* d_pose) {
  if (n <= 0) return 0;
  if (n > 0x7FFFFFFF) throw std::invalid_argument("too many records in one message");
  const bool host_src = ids != nullptr;
  enable_mail(p);
  const int nr = (int)n;
  auto queue = [&](uint32_t id, uint32_t s, uint32_t ns, const double* pose) {   // unknown id: queued in arrival order for the next tick
    PendingRec r;
    r.id = id;
    r.sec = s;
    r.nsec = ns;
    std::memcpy(r.pose, pose, sizeof(r.pose));
    p->pending.push_back(r);
  };
  if (p->n == 0 && host_src) {   // no targets yet: every record belongs to a target-less mailbox
    for (long long k = 0; k < n; ++k) queue(ids[k], sec[k], nsec[k], poses + 7 * k);
    return 0;
  }
  if (host_src) {
    d_ids = to_dev(p, ids, (size_t)n);
    d_sec = to_dev(p, sec, (size_t)n);
    d_nsec = to_dev(p, nsec, (size_t)n);
    d_pose = to_dev(p, poses, (size_t)n * 7);
  }
  uint32_t* key_in = p->arena.get_n<uint32_t>((size_t)n);
  uint32_t* key_out = p->arena.get_n<uint32_t>((size_t)n);
  int* rec_in = p->arena.get_n<int>((size_t)n);
  int* rec_out = p->arena.get_n<int>((size_t)n);
  int* unknown = p->arena.get_n<int>((size_t)n);
  int* counter = p->arena.get_n<int>(1);
  CK(cudaMemsetAsync(counter, 0, sizeof(int), p->stream));
  Buf& b = p->buf[p->cur];
  te::mb_lookup_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(b.cold.ids, (int)p->n, d_ids, nr, key_in, rec_in, unknown, counter);
  CK(cudaGetLastError());
  if (p->n > 0) {
    // stable sort by slot: the records of one id stay in arrival order (a message may name an id more than once, and several
    // messages may be ingested between two ticks)
    int bits = 1;
    while (bits < 32 && (1ll << bits) <= p->n) ++bits;   // keys are slots < n, or n for unknown ids: 2^bits > n
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_in, key_out, rec_in, rec_out, nr, 0, bits, p->stream));
    void* tmp = p->arena.get(tmp_bytes);
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, rec_in, rec_out, nr, 0, bits, p->stream));
    te::mb_apply_kernel<<<cdiv(n, 256), 256, 0, p->stream>>>(nr, (int)p->n, key_out, rec_out, d_sec, d_nsec, d_pose, p->mb[p->mb_cur].a, b.cold.last_meas);
    CK(cudaGetLastError());
  }
  int n_unknown = 0;
  CK(cudaMemcpyAsync(&n_unknown, counter, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));   // also: a host caller's record arrays are free again
  if (n_unknown == 0) return 0;
  if (host_src) {
    std::vector<int> list((size_t)n_unknown);
    CK(cudaMemcpy(list.data(), unknown, (size_t)n_unknown * sizeof(int), cudaMemcpyDeviceToHost));
    std::sort(list.begin(), list.end());   // arrival order
    for (int k : list) queue(ids[k], sec[k], nsec[k], poses + 7 * (size_t)k);
    return 0;
  }
  // device source: pack the unknown records (device order) into one block [pose 7 | index | id | sec | nsec] x n_unknown, read it
  // back in one copy to pinned memory, queue the records in arrival order
  const size_t nu = (size_t)n_unknown;
  char* d_blk = (char*)p->arena.get(nu * 72);
  double* o_pose = (double*)d_blk;
  int* o_idx = (int*)(o_pose + 7 * nu);
  uint32_t* o_ids = (uint32_t*)(o_idx + nu);
  uint32_t* o_sec = o_ids + nu;
  uint32_t* o_nsec = o_sec + nu;
  te::mb_pack_unknown_kernel<<<cdiv(n_unknown, 256), 256, 0, p->stream>>>(n_unknown, unknown, d_ids, d_sec, d_nsec, d_pose, o_ids, o_sec, o_nsec, o_pose);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(o_idx, unknown, nu * sizeof(int), cudaMemcpyDeviceToDevice, p->stream));
  char* h_blk = pinned_stage(p, nu * 72);
  CK(cudaMemcpyAsync(h_blk, d_blk, nu * 72, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  const double* h_pose = (const double*)h_blk;
  const int* h_idx = (const int*)(h_pose + 7 * nu);
  const uint32_t* h_ids = (const uint32_t*)(h_idx + nu);
  const uint32_t* h_sec = h_ids + nu;
  const uint32_t* h_nsec = h_sec + nu;
  bool in_order = true;
  for (size_t k = 1; k < nu && in_order; ++k) in_order = h_idx[k - 1] < h_idx[k];
  if (in_order) {
    for (size_t k = 0; k < nu; ++k) queue(h_ids[k], h_sec[k], h_nsec[k], h_pose + 7 * k);
  } else {
    std::vector<int> order(nu);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int a, int c) { return h_idx[a] < h_idx[c]; });
    for (int k : order) queue(h_ids[(size_t)k], h_sec[(size_t)k], h_nsec[(size_t)k], h_pose + 7 * (size_t)k);
  }
  return 0;
}
}  // namespace

int te_pool_mailbox_ingest(te_pool* p, long long n, const uint32_t* ids, const uint32_t* sec, const uint32_t* nsec, const double* poses) {
  return guarded(p, [&] {
    if (n <= 0) return 0;
    if (!ids || !sec || !nsec || !poses) throw std::invalid_argument("null record arrays");
    return mailbox_ingest_impl(p, n, ids, sec, nsec, poses, nullptr, nullptr, nullptr, nullptr);
  });
}

int te_pool_mailbox_ingest_dev(te_pool* p, long long n, const uint32_t* dev_ids, const uint32_t* dev_sec, const uint32_t* dev_nsec,
                               const double* dev_poses) {
  return guarded(p, [&] {
    if (n <= 0) return 0;
    if (!dev_ids || !dev_sec || !dev_nsec || !dev_poses) throw std::invalid_argument("null record arrays");
    return mailbox_ingest_impl(p, n, nullptr, nullptr, nullptr, nullptr, dev_ids, dev_sec, dev_nsec, dev_poses);
  });
}

long long te_pool_mailbox_tick(te_pool* p, double dt, double t0_new, int cls_new, uint32_t now_sec, uint32_t now_nsec, double timeout,
                               uint32_t* erased_out, long long cap, uint32_t* added_out, long long added_cap, long long* n_added_out) {
  return guarded_ll(p, [&]() -> long long {
    if (!(dt >= 0.0)) throw std::invalid_argument("dt must be >= 0 (assert of src/target_interface.cpp:150)");
    if (p->hQ.empty()) throw std::runtime_error("no model class registered");
    if (cls_new < 0 || cls_new >= (int)p->hQ.size()) throw std::invalid_argument("unknown model class for the new targets");
    enable_mail(p);
    static const bool dbg = std::getenv("TE_MB_DEBUG") != nullptr;
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tw[6] = {0, 0, 0, 0, 0, 0};
    auto mark = [&](int i) { if (dbg) { cudaStreamSynchronize(p->stream); tw[i] = wall(); } };
    mark(0);
    const double now = host_to_sec(now_sec, now_nsec);
    // 1. target-less mailboxes, ascending id: readable -> init on first sight (src/target_manager_ros.cpp:54-58) unless the
    //    same tick would erase it again (:67-72; init + update + erase is unobservable); unreadable -> stays, or expires
    // the add arrays are written straight into ONE pinned block [pose 7 | t0 | last | id | sec | nsec] x max_add and go to the
    // device in one copy
    std::vector<uint32_t> host_erased;
    const size_t max_add = p->pending.size() + p->orphans.size();
    char* stage = max_add ? pinned_stage(p, max_add * 84) : nullptr;
    double* add_pose = (double*)stage;
    double* add_t0 = add_pose + 7 * max_add;
    double* add_last = add_t0 + max_add;
    uint32_t* add_ids = (uint32_t*)(add_last + max_add);
    uint32_t* add_sec = add_ids + max_add;
    uint32_t* add_nsec = add_sec + max_add;
    size_t n_promoted = 0;
    auto promote = [&](uint32_t id, const HostMail& m) {
      const size_t k = n_promoted++;
      add_ids[k] = id;
      add_sec[k] = m.sec;
      add_nsec[k] = m.nsec;
      add_last[k] = m.last;
      add_t0[k] = t0_new;
      std::memcpy(add_pose + 7 * k, m.pose, 56);
    };
    bool fast = p->orphans.empty();   // common case: every queued record is the first sight of a new id, ids ascending
    for (size_t k = 1; fast && k < p->pending.size(); ++k) fast = p->pending[k - 1].id < p->pending[k].id;
    if (fast) {
      for (const PendingRec& r : p->pending) {
        HostMail m;   // Measurement(): readable, stamp 0
        apply_record(m, r);
        if (m.last > 0.0 && (now - m.last) >= timeout) host_erased.push_back(r.id);
        else if (m.fresh) promote(r.id, m);
        else p->orphans.emplace_hint(p->orphans.end(), r.id, m);   // "Target(id) does not exist!" (src/target_manager.cpp:209)
      }
      p->pending.clear();
    } else {
      fold_pending(p);
      for (auto it = p->orphans.begin(); it != p->orphans.end();) {
        const HostMail& m = it->second;
        const bool expired = m.last > 0.0 && (now - m.last) >= timeout;
        if (expired) {
          host_erased.push_back(it->first);
          it = p->orphans.erase(it);
        } else if (m.fresh) {
          promote(it->first, m);
          it = p->orphans.erase(it);
        } else {
          ++it;
        }
      }
    }
    const int n_add = (int)n_promoted;
    const int n_old = (int)p->n;
    if (n_added_out) *n_added_out = n_add;
    if (added_out && added_cap > 0 && n_add > 0) std::memcpy(added_out, add_ids, (size_t)std::min<long long>(added_cap, n_add) * sizeof(uint32_t));
    mark(1);
    // 2. expiry flags of the existing targets, then ONE stable rebuild: survivors compacted, promoted mailboxes merged in by id
    ensure_work(p, (size_t)n_old + (size_t)n_add);
    uint32_t* d_erased = nullptr;
    long long n_dev_erased = 0;
    if (n_old > 0) {
      te::expire_flags_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.last_meas, n_old, now, timeout, p->alive);
      CK(cudaGetLastError());
      d_erased = p->arena.get_n<uint32_t>((size_t)n_old);
    }
    te::AddData ad{};
    std::vector<uint16_t> add_cls;
    if (n_add > 0) {
      char* d_stage = (char*)p->arena.get(max_add * 84);
      CK(cudaMemcpyAsync(d_stage, stage, max_add * 84, cudaMemcpyHostToDevice, p->stream));   // (pinned: the final sync of the tick covers it)
      const double* d_pose0 = (const double*)d_stage;
      const double* d_t0 = d_pose0 + 7 * max_add;
      const double* d_last = d_t0 + max_add;
      const uint32_t* d_aid = (const uint32_t*)(d_last + max_add);
      ad.ids = d_aid;
      if (cls_new != 0) {
        add_cls.assign((size_t)n_add, (uint16_t)cls_new);
        ad.cls = to_dev(p, add_cls.data(), (size_t)n_add);
      }
      ad.t0 = d_t0;
      ad.p0 = d_pose0;
      p->mb_add.sec = d_aid + max_add;
      p->mb_add.nsec = d_aid + 2 * max_add;
      p->mb_add.last = d_last;
    }
    const bool unfused_env = std::getenv("TE_MB_UNFUSED") != nullptr;   // debugging / test switch: the rebuild-then-step form
    const bool fused = !unfused_env && n_old > 0;
    if (!fused) {
      // reference form: stable rebuild (survivors gathered, new ids merged in and initialised), then the step in place
      if (n_old > 0 || n_add > 0) {
        int alive = 0;
        try {
          alive = compact_and_merge(p, ad, ad.ids, n_add, d_erased);
        } catch (...) {
          p->mb_add = te::MailAdd{nullptr, nullptr, nullptr};
          throw;
        }
        n_dev_erased = n_old - alive;
      }
      p->mb_add = te::MailAdd{nullptr, nullptr, nullptr};
      mark(2);
      // 3. the step: update where the mailbox is readable (the flag is sticky: a silent target re-applies its last pose),
      //    predict elsewhere (:59,:64).  The mailbox arrays are the kernel's measurement block and action array.
      if (p->n > 0) {
        te::StepArgs a = base_args(p);
        const te::MailArrays& mb = p->mb[p->mb_cur].a;
        a.dt = dt;
        a.meas = mb.pose;
        a.meas_stride = 7;
        a.meas_tma = 1;
        a.action = mb.act;
        a.default_action = TE_ACT_PREDICT;
        launch_step(p, a, a.n_tiles);
        te::copy_meas_masked_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->buf[p->cur].cold.meas, mb.pose, mb.act, (int)p->n);
        CK(cudaGetLastError());
      }
    } else {
      // fused form (same results bit for bit): the step kernel reads every tile in place and writes the survivors' columns
      // straight to their slots in the merged order (StepArgs::dst_*), so the state crosses HBM once per tick; the promoted
      // mailboxes are initialised in their slots afterwards and get their first update from a sparse follow-up launch
      Buf& ob = p->buf[p->cur];
      const te::MailArrays omb = p->mb[p->mb_cur].a;
      size_t tmp = p->cub_bytes;
      CK(cub::DeviceScan::ExclusiveSum(p->cub_tmp, tmp, p->alive, p->pos, n_old, p->stream));
      int last_pos = 0, last_alive = 0;
      CK(cudaMemcpyAsync(&last_pos, p->pos + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
      CK(cudaMemcpyAsync(&last_alive, p->alive + n_old - 1, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
      CK(cudaStreamSynchronize(p->stream));
      const int n_alive = last_pos + last_alive;
      n_dev_erased = n_old - n_alive;
      te::StepArgs a = base_args(p);
      a.dt = dt;
      a.meas = omb.pose;
      a.meas_stride = 7;
      a.meas_tma = 1;
      a.action = omb.act;
      a.default_action = TE_ACT_PREDICT;
      te::copy_meas_masked_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(ob.cold.meas, omb.pose, omb.act, n_old);   // measured_pose_
      CK(cudaGetLastError());
      if (n_dev_erased == 0 && n_add == 0) {
        launch_step(p, a, a.n_tiles);   // nothing moves: the ordinary in-place step
        p->mb_add = te::MailAdd{nullptr, nullptr, nullptr};
        mark(2);
      } else {
        const int n_new = n_alive + n_add;
        if (n_dev_erased > 0) {
          te::collect_erased_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(p->alive, p->pos, ob.cold.ids, n_old, d_erased);
          CK(cudaGetLastError());
        }
        ensure_other_capacity(p, (size_t)n_new);
        ensure_mail_other(p, (size_t)n_new);
        Buf& nb = p->buf[1 - p->cur];
        const te::MailArrays nmb = p->mb[1 - p->mb_cur].a;
        int* new_dst = nullptr;
        if (n_add > 0) {
          new_dst = p->arena.get_n<int>((size_t)n_add);
          te::merge_new_dst_kernel<<<cdiv(n_add, 256), 256, 0, p->stream>>>(n_add, ad.ids, ob.cold.ids, n_old, p->pos, n_alive, new_dst);
          te::merge_old_dst_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(n_old, p->alive, p->pos, ob.cold.ids, ad.ids, n_add);
          CK(cudaGetLastError());
        }
        te::compact_cold_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(n_old, p->alive, p->pos, ob.cold, nb.cold);
        te::mb_compact_kernel<<<cdiv(n_old, 256), 256, 0, p->stream>>>(n_old, p->alive, p->pos, omb, nmb);
        CK(cudaGetLastError());
        if (n_alive > 0) {
          a.dst_tiles = nb.tiles;
          a.dst_alive = p->alive;
          a.dst_pos = p->pos;
          launch_step(p, a, a.n_tiles);
        }
        if (n_add > 0) {
          CK(cudaMemsetAsync(p->d_counters, 0, 2 * sizeof(int), p->stream));
          init_promoted(p, n_add, new_dst, nb, ad, nmb);
        }
        p->mb_add = te::MailAdd{nullptr, nullptr, nullptr};
        p->cur = 1 - p->cur;
        p->mb_cur = 1 - p->mb_cur;
        p->n = n_new;
        p->h_ids_valid = false;
        mark(2);
        if (n_add > 0) {   // first update of the new targets with the pose that created them: only their tiles, only their lanes
          te::StepArgs b = base_args(p);
          b.tile_list = p->tile_list;
          b.d_nwork = p->d_counters;
          b.dt = dt;
          b.meas = nmb.pose;
          b.meas_stride = 7;
          b.meas_tma = 1;
          b.action = p->action;
          b.default_action = TE_ACT_NONE;
          b.clear_action = 1;
          b.tile_flag = p->tile_flag;
          launch_step(p, b, std::min(n_add, b.n_tiles));
        }
        fetch_last_id(p);
      }
    }
    mark(3);
    // 4. erased ids of this tick, ascending: targets the device expired + target-less mailboxes the host expired
    std::vector<uint32_t> dev_erased((size_t)n_dev_erased);
    if (n_dev_erased > 0)
      CK(cudaMemcpyAsync(dev_erased.data(), d_erased, (size_t)n_dev_erased * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const long long n_er = n_dev_erased + (long long)host_erased.size();
    if (erased_out && cap > 0 && n_er > 0) {
      std::vector<uint32_t> all((size_t)n_er);
      std::merge(dev_erased.begin(), dev_erased.end(), host_erased.begin(), host_erased.end(), all.begin());
      std::memcpy(erased_out, all.data(), (size_t)std::min(cap, n_er) * sizeof(uint32_t));
    }
    mark(4);
    if (dbg) std::fprintf(stderr, "[te mailbox tick] host mailboxes %.3f ms, flags + merge (fused: + step) %.3f ms, step (fused: first update of the new targets) %.3f ms, erase list %.3f ms (n %lld, +%d, -%lld)\n",
                          tw[1] - tw[0], tw[2] - tw[1], tw[3] - tw[2], tw[4] - tw[3], p->n, n_add, n_er);
    return n_er;
  });
}

long long te_pool_mailbox_count(te_pool* p) {
  return guarded_ll(p, [&]() -> long long {
    fold_pending(p);
    long long n = (long long)p->orphans.size();
    if (!p->mb_on || p->n == 0) return n;
    int* counter = p->arena.get_n<int>(1);
    CK(cudaMemsetAsync(counter, 0, sizeof(int), p->stream));
    te::mb_count_kernel<<<cdiv(p->n, 256), 256, 0, p->stream>>>(p->mb[p->mb_cur].a.act, (int)p->n, counter);
    CK(cudaGetLastError());
    int c = 0;
    CK(cudaMemcpyAsync(&c, counter, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return n + c;
  });
}

long long te_pool_mailbox_bound(te_pool* p) { return p ? p->n + (long long)p->orphans.size() + (long long)p->pending.size() : -1; }

const double* te_pool_mailbox_dev_pose(te_pool* p) { return (p && p->mb_on) ? p->mb[p->mb_cur].a.pose : nullptr; }
const uint8_t* te_pool_mailbox_dev_action(te_pool* p) { return (p && p->mb_on) ? p->mb[p->mb_cur].a.act : nullptr; }


// ---- batched IntersectionSolver -------------------------------------------------------------
te_isolver* te_isolver_create(te_pool* p, long long n_streams, unsigned filters_length) {
  te_isolver* s = nullptr;
  try {
    if (!p || n_streams <= 0 || filters_length == 0) throw std::invalid_argument("bad isolver arguments");
    DeviceGuard g(p->device);
    s = new te_isolver();
    s->pool = p;
    s->st.n_streams = n_streams;
    s->st.L = filters_length;
    CK(cudaMalloc(&s->st.prev_pose, (size_t)n_streams * 7 * 8));
    CK(cudaMalloc(&s->st.pos_win, (size_t)n_streams * filters_length * 8));
    CK(cudaMalloc(&s->st.ang_win, (size_t)n_streams * filters_length * 8));
    CK(cudaMalloc(&s->st.pos_sum, (size_t)n_streams * 8));
    CK(cudaMalloc(&s->st.ang_sum, (size_t)n_streams * 8));
    CK(cudaMalloc(&s->st.idx, (size_t)n_streams * 4));
    CK(cudaMalloc(&s->st.complete, (size_t)n_streams));
    CK(cudaMemsetAsync(s->st.pos_win, 0, (size_t)n_streams * filters_length * 8, p->stream));
    CK(cudaMemsetAsync(s->st.ang_win, 0, (size_t)n_streams * filters_length * 8, p->stream));
    CK(cudaMemsetAsync(s->st.pos_sum, 0, (size_t)n_streams * 8, p->stream));
    CK(cudaMemsetAsync(s->st.ang_sum, 0, (size_t)n_streams * 8, p->stream));
    CK(cudaMemsetAsync(s->st.idx, 0, (size_t)n_streams * 4, p->stream));
    CK(cudaMemsetAsync(s->st.complete, 0, (size_t)n_streams, p->stream));
    te::init_pose_kernel<<<cdiv(n_streams, 256), 256, 0, p->stream>>>(s->st.prev_pose, n_streams);   // initPose (:39)
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(p->stream));
    return s;
  } catch (const std::exception& e) {
    g_err = e.what();
    if (s) te_isolver_destroy(s);
    return nullptr;
  }
}

void te_isolver_destroy(te_isolver* s) {
  if (!s) return;
  cudaFree(s->st.prev_pose); cudaFree(s->st.pos_win); cudaFree(s->st.ang_win); cudaFree(s->st.pos_sum);
  cudaFree(s->st.ang_sum); cudaFree(s->st.idx); cudaFree(s->st.complete);
  delete s;
}

int te_isolver_query(te_isolver* s, long long n, const uint32_t* ids, const int32_t* stream, const double* t1, const double* origin,
                     const double* radius, const double* pos_th, const double* ang_th, double* delta_t, double* pose7, uint8_t* converged) {
  if (!s) { g_err = "null isolver"; return -1; }
  te_pool* p = s->pool;
  return guarded(p, [&] {
    if (n <= 0) return 0;
    if (!ids || !t1 || !origin || !radius) throw std::invalid_argument("ids, t1, origin and radius are required");
    if (pose7 && (!pos_th || !ang_th)) throw std::invalid_argument("thresholds are required with pose7");
    if (pose7 && !stream && n > s->st.n_streams) throw std::invalid_argument("more queries than solver streams");
    if (stream) for (long long k = 0; k < n; ++k) if (stream[k] < 0 || stream[k] >= s->st.n_streams) throw std::invalid_argument("bad stream index");
    int* slots = p->n ? lookup_slots(p, to_dev(p, ids, n), n) : nullptr;
    if (!slots) {
      slots = p->arena.get_n<int>((size_t)n);
      CK(cudaMemsetAsync(slots, 0xff, (size_t)n * sizeof(int), p->stream));
    }
    const int* d_stream = to_dev(p, stream, n);
    const double* d_t1 = to_dev(p, t1, n);
    const double* d_origin = to_dev(p, origin, n * 3);
    const double* d_radius = to_dev(p, radius, n);
    const double* d_pth = to_dev(p, pos_th, n);
    const double* d_ath = to_dev(p, ang_th, n);
    double* d_delta = delta_t ? p->arena.get_n<double>((size_t)n) : nullptr;
    double* d_pose = pose7 ? p->arena.get_n<double>((size_t)n * 7) : nullptr;
    uint8_t* d_conv = converged ? p->arena.get_n<uint8_t>((size_t)n) : nullptr;
    Buf& b = p->buf[p->cur];
    const int g = cdiv(n, 128);
    switch (p->model) {
      case te::UNIFORM_VELOCITY: te::isolver_kernel<te::UNIFORM_VELOCITY><<<g, 128, 0, p->stream>>>(b.tiles, slots, d_stream, n, d_t1, d_origin, d_radius, d_pth, d_ath, s->st, d_delta, d_pose, d_conv); break;
      case te::UNIFORM_ACCELERATION: te::isolver_kernel<te::UNIFORM_ACCELERATION><<<g, 128, 0, p->stream>>>(b.tiles, slots, d_stream, n, d_t1, d_origin, d_radius, d_pth, d_ath, s->st, d_delta, d_pose, d_conv); break;
      case te::ANGULAR_VELOCITIES: te::isolver_kernel<te::ANGULAR_VELOCITIES><<<g, 128, 0, p->stream>>>(b.tiles, slots, d_stream, n, d_t1, d_origin, d_radius, d_pth, d_ath, s->st, d_delta, d_pose, d_conv); break;
      default: te::isolver_kernel<te::ANGULAR_RATES><<<g, 128, 0, p->stream>>>(b.tiles, slots, d_stream, n, d_t1, d_origin, d_radius, d_pth, d_ath, s->st, d_delta, d_pose, d_conv); break;
    }
    CK(cudaGetLastError());
    if (delta_t) CK(cudaMemcpyAsync(delta_t, d_delta, (size_t)n * 8, cudaMemcpyDeviceToHost, p->stream));
    if (pose7) CK(cudaMemcpyAsync(pose7, d_pose, (size_t)n * 56, cudaMemcpyDeviceToHost, p->stream));
    if (converged) CK(cudaMemcpyAsync(converged, d_conv, (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  });
}

int te_isolver_query_dense(te_isolver* s, const double* dev_t1, const double* dev_origin, const double* dev_radius, double pos_th,
                           double ang_th, double* dev_delta, double* dev_pose7, uint8_t* dev_converged) {
  if (!s) { g_err = "null isolver"; return -1; }
  te_pool* p = s->pool;
  return guarded(p, [&] {
    const long long n = p->n;
    if (n == 0) return 0;
    if (!dev_origin || !dev_radius) throw std::invalid_argument("origin and radius are required");
    if (n > s->st.n_streams) throw std::invalid_argument("more targets than solver streams");
    te::IsolverState st = s->st;
    st.pos_th_all = pos_th;
    st.ang_th_all = ang_th;
    Buf& b = p->buf[p->cur];
    const int g = cdiv(n, 128);
    switch (p->model) {
      case te::UNIFORM_VELOCITY: te::isolver_kernel<te::UNIFORM_VELOCITY><<<g, 128, 0, p->stream>>>(b.tiles, nullptr, nullptr, n, dev_t1, dev_origin, dev_radius, nullptr, nullptr, st, dev_delta, dev_pose7, dev_converged); break;
      case te::UNIFORM_ACCELERATION: te::isolver_kernel<te::UNIFORM_ACCELERATION><<<g, 128, 0, p->stream>>>(b.tiles, nullptr, nullptr, n, dev_t1, dev_origin, dev_radius, nullptr, nullptr, st, dev_delta, dev_pose7, dev_converged); break;
      case te::ANGULAR_VELOCITIES: te::isolver_kernel<te::ANGULAR_VELOCITIES><<<g, 128, 0, p->stream>>>(b.tiles, nullptr, nullptr, n, dev_t1, dev_origin, dev_radius, nullptr, nullptr, st, dev_delta, dev_pose7, dev_converged); break;
      default: te::isolver_kernel<te::ANGULAR_RATES><<<g, 128, 0, p->stream>>>(b.tiles, nullptr, nullptr, n, dev_t1, dev_origin, dev_radius, nullptr, nullptr, st, dev_delta, dev_pose7, dev_converged); break;
    }
    CK(cudaGetLastError());
    return 0;
  });
}

}  // extern "C"

#ifdef TE_TIMELINE
// debug build only (tools/timeline.py): phase timestamps of CTA 0 recorded by kf_step_split_kernel
extern "C" int te_debug_timeline(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, te::g_timeline, sizeof(long long) * 2 * 64 * 12);
}
#endif
