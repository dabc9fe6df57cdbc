// te_isolver.cu -- batched IntersectionSolver (/root/reference/src/intersection_solver.cpp:42-124): te_isolver_* of
// include/te_pool.h over isolver_kernel (te_kernels.cuh) and the quartic root finder (te_quartic.h).
#include "te_pool_internal.cuh"

using namespace tehost;
#define g_err (tehost::last_error())

extern "C" {

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
