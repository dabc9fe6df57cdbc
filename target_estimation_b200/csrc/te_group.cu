// te_group.cu -- several pools on several devices of one box: the optional NCCL all-gather of estimate records
// (BASELINE.json north_star: "NCCL over NVLink is used only for the optional all-gather of estimates / intercepts
// to the publishing rank").  The hot path has no collective: every shard steps its own pool on its own device and stream.
//
// One process drives all devices (the host surface is the C++ ShardedTargetManager of host/sharded_manager.cpp), so the
// communicators come from ncclCommInitAll and every collective is issued for all ranks inside one ncclGroupStart / ncclGroupEnd.
// NCCL is loaded with dlopen on first use: libte_pool.so carries no link-time dependency on it, and a process that already
// holds a libnccl.so.2 (PyTorch's) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include "te_pool_internal.cuh"

using namespace tehost;

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.h) return api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw std::runtime_error(std::string("NCCL is required for a multi-device group: ") + dlerror());
  auto sym = [&](const char* name) {
    void* s = dlsym(h, name);
    if (!s) throw std::runtime_error(std::string("libnccl.so.2 lacks ") + name);
    return s;
  };
  api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.h = h;
  return api;
}

#define NCK(call)                                                                                                        \
  do {                                                                                                                   \
    ncclResult_t r__ = (call);                                                                                           \
    if (r__ != ncclSuccess) throw std::runtime_error(std::string(#call) + ": " + nccl().GetErrorString(r__));            \
  } while (0)

struct Shard {
  int device = 0;
  cudaStream_t own_stream = nullptr;   // used when the shard has no pool in a gather
  cudaStream_t stream = nullptr;       // the stream of the running gather on this device
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double* rec = nullptr;               // this shard's records [count][13]
  size_t rec_cap = 0;
  double* out = nullptr;               // all shards' records [total][13]
  uint32_t* out_ids = nullptr;
  size_t out_cap = 0;
};

}  // namespace

struct te_group {
  std::vector<Shard> sh;
  std::vector<ncclComm_t> comms;
  bool use_nccl = false;
  std::vector<long long> counts;
  long long total = 0;
  bool timed = false;
};

namespace {

void grow(te_group* g, int r, size_t n_rec, size_t n_total) {
  Shard& s = g->sh[r];
  if (n_rec > s.rec_cap) {
    CK(cudaStreamSynchronize(s.stream));
    cudaFree(s.rec);
    s.rec = nullptr;
    const size_t cap = n_rec + n_rec / 8 + 32;
    CK(cudaMalloc(&s.rec, cap * 13 * sizeof(double)));
    s.rec_cap = cap;
  }
  if (n_total > s.out_cap) {
    CK(cudaStreamSynchronize(s.stream));
    cudaFree(s.out);
    cudaFree(s.out_ids);
    s.out = nullptr;
    s.out_ids = nullptr;
    const size_t cap = n_total + n_total / 8 + 32;
    CK(cudaMalloc(&s.out, cap * 13 * sizeof(double)));
    CK(cudaMalloc(&s.out_ids, cap * sizeof(uint32_t)));
    s.out_cap = cap;
  }
}

template <class F> long long group_guard(te_group* g, F&& f) {
  int prev = 0;
  cudaGetDevice(&prev);
  long long rc = -1;
  try {
    if (!g) throw std::invalid_argument("null group");
    rc = f();
  } catch (const std::exception& e) {
    last_error() = e.what();
    rc = -1;
  }
  cudaSetDevice(prev);
  return rc;
}

}  // namespace

extern "C" {

te_group* te_group_create(int n, const int* devices) {
  te_group* g = nullptr;
  int prev = 0;
  cudaGetDevice(&prev);
  try {
    if (n <= 0 || !devices) throw std::invalid_argument("a group needs at least one device");
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    g = new te_group();
    g->sh.resize((size_t)n);
    bool distinct = true;
    for (int r = 0; r < n; ++r) {
      if (devices[r] < 0 || devices[r] >= n_dev) throw std::invalid_argument("no such CUDA device: " + std::to_string(devices[r]));
      for (int q = 0; q < r; ++q) distinct = distinct && devices[q] != devices[r];
      Shard& s = g->sh[r];
      s.device = devices[r];
      CK(cudaSetDevice(s.device));
      CK(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
      s.stream = s.own_stream;
      CK(cudaEventCreate(&s.ev0));
      CK(cudaEventCreate(&s.ev1));
    }
    g->counts.assign((size_t)n, 0);
    g->use_nccl = distinct && n > 1;
    if (g->use_nccl) {
      g->comms.resize((size_t)n);
      NCK(nccl().CommInitAll(g->comms.data(), n, devices));
    }
  } catch (const std::exception& e) {
    last_error() = e.what();
    if (g) te_group_destroy(g);
    g = nullptr;
  }
  cudaSetDevice(prev);
  return g;
}

void te_group_destroy(te_group* g) {
  if (!g) return;
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto c : g->comms)
    if (c) nccl().CommDestroy(c);
  for (Shard& s : g->sh) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.rec);
    cudaFree(s.out);
    cudaFree(s.out_ids);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
  }
  cudaSetDevice(prev);
  delete g;
}

int te_group_size(te_group* g) { return g ? (int)g->sh.size() : -1; }
int te_group_uses_nccl(te_group* g) { return g ? (g->use_nccl ? 1 : 0) : -1; }

long long te_group_allgather_estimates(te_group* g, te_pool* const* pools, long long* counts_out) {
  return group_guard(g, [&]() -> long long {
    const int G = (int)g->sh.size();
    if (!pools) throw std::invalid_argument("null pool list");
    long long total = 0;
    bool equal = true;
    for (int r = 0; r < G; ++r) {
      te_pool* p = pools[r];
      if (p && p->device != g->sh[r].device) throw std::invalid_argument("pool of shard " + std::to_string(r) + " lives on another device");
      g->counts[r] = p ? p->n : 0;
      g->sh[r].stream = p ? p->stream : g->sh[r].own_stream;
      total += g->counts[r];
      equal = equal && g->counts[r] == g->counts[0];
      if (counts_out) counts_out[r] = g->counts[r];
    }
    g->total = total;
    g->timed = false;
    if (total == 0) return 0;
    // 1. every shard's records on its own device and stream
    for (int r = 0; r < G; ++r) {
      Shard& s = g->sh[r];
      CK(cudaSetDevice(s.device));
      grow(g, r, (size_t)g->counts[r], (size_t)total);
      if (g->counts[r] > 0 && te_pool_estimates_dev(pools[r], s.rec) < 0) throw std::runtime_error(last_error());
      CK(cudaEventRecord(s.ev0, s.stream));
    }
    // 2. the exchange
    if (g->use_nccl) {
      NCK(nccl().GroupStart());
      if (equal) {
        for (int d = 0; d < G; ++d) {
          Shard& s = g->sh[d];
          NCK(nccl().AllGather(s.rec, s.out, (size_t)g->counts[d] * 13, ncclDouble, g->comms[d], s.stream));
          NCK(nccl().AllGather(pools[d]->buf[pools[d]->cur].cold.ids, s.out_ids, (size_t)g->counts[d], ncclUint32, g->comms[d], s.stream));
        }
      } else {
        long long off = 0;
        for (int r = 0; r < G; ++r) {   // root r sends its block to everyone (ragged all-gather)
          if (g->counts[r] > 0) {
            for (int d = 0; d < G; ++d) {
              Shard& s = g->sh[d];
              const void* src = d == r ? (const void*)s.rec : (const void*)(s.out + off * 13);
              const void* src_ids = d == r ? (const void*)pools[r]->buf[pools[r]->cur].cold.ids : (const void*)(s.out_ids + off);
              NCK(nccl().Broadcast(src, s.out + off * 13, (size_t)g->counts[r] * 13, ncclDouble, r, g->comms[d], s.stream));
              NCK(nccl().Broadcast(src_ids, s.out_ids + off, (size_t)g->counts[r], ncclUint32, r, g->comms[d], s.stream));
            }
          }
          off += g->counts[r];
        }
      }
      NCK(nccl().GroupEnd());
    } else {
      // shards that share a device (one-GPU test of the sharded host logic) or a single shard: plain copies, ordered after the
      // producers through events
      for (int d = 0; d < G; ++d) {
        Shard& s = g->sh[d];
        CK(cudaSetDevice(s.device));
        long long off = 0;
        for (int r = 0; r < G; ++r) {
          if (g->counts[r] > 0) {
            if (r != d) CK(cudaStreamWaitEvent(s.stream, g->sh[r].ev0, 0));
            CK(cudaMemcpyPeerAsync(s.out + off * 13, s.device, g->sh[r].rec, g->sh[r].device, (size_t)g->counts[r] * 13 * sizeof(double), s.stream));
            CK(cudaMemcpyPeerAsync(s.out_ids + off, s.device, pools[r]->buf[pools[r]->cur].cold.ids, g->sh[r].device,
                                   (size_t)g->counts[r] * sizeof(uint32_t), s.stream));
          }
          off += g->counts[r];
        }
      }
    }
    for (int d = 0; d < G; ++d) {
      Shard& s = g->sh[d];
      CK(cudaSetDevice(s.device));
      CK(cudaEventRecord(s.ev1, s.stream));
    }
    g->timed = true;
    return total;
  });
}

int te_group_sync(te_group* g) {
  return (int)group_guard(g, [&]() -> long long {
    for (Shard& s : g->sh) {
      CK(cudaSetDevice(s.device));
      CK(cudaStreamSynchronize(s.stream));
    }
    return 0;
  });
}

const double* te_group_dev_records(te_group* g, int r) { return (g && r >= 0 && r < (int)g->sh.size()) ? g->sh[r].out : nullptr; }
const uint32_t* te_group_dev_ids(te_group* g, int r) { return (g && r >= 0 && r < (int)g->sh.size()) ? g->sh[r].out_ids : nullptr; }

long long te_group_fetch(te_group* g, int root, double* records_out, uint32_t* ids_out, long long cap) {
  return group_guard(g, [&]() -> long long {
    if (root < 0 || root >= (int)g->sh.size()) throw std::invalid_argument("no such shard");
    Shard& s = g->sh[root];
    const long long k = std::min(cap, g->total);
    CK(cudaSetDevice(s.device));
    if (k > 0 && records_out) CK(cudaMemcpyAsync(records_out, s.out, (size_t)k * 13 * sizeof(double), cudaMemcpyDeviceToHost, s.stream));
    if (k > 0 && ids_out) CK(cudaMemcpyAsync(ids_out, s.out_ids, (size_t)k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return g->total;
  });
}

double te_group_last_gather_ms(te_group* g) {
  double worst = -1.0;
  group_guard(g, [&]() -> long long {
    if (!g->timed) return 0;
    for (Shard& s : g->sh) {
      CK(cudaSetDevice(s.device));
      CK(cudaEventSynchronize(s.ev1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
      worst = std::max(worst, (double)ms);
    }
    return 0;
  });
  return worst;
}

}  // extern "C"
