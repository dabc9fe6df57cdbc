// te_diag.cu -- device micro-benchmarks behind the roofline statements of DESIGN.md: the FP64 FMA rate of the machine (BASELINE.md
// section 2 asks for it before any "HBM-bound" claim) and a plain streaming copy on the same clocks.  Not on any product path.
#include "te_pool_internal.cuh"

using namespace tehost;

namespace {

// ILP independent dependent-FMA chains per thread: the FP64 pipe saturates once enough warps x chains cover its latency
template <int ILP>
__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = (double)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  if (s == 123.456) out[0] = s;   // keeps the chains alive
}

__global__ void __launch_bounds__(256) stream_copy_kernel(const double4* __restrict__ src, double4* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

extern "C" {

/* FP64 FMA throughput of `device` in TFLOP/s (2 flops per FMA), best of 5 launches of a dependent-chain kernel with 8 chains per
 * thread, 8 CTAs of 256 threads per SM; also the streaming copy rate in GB/s (read + write bytes) over a 2 GiB buffer pair with
 * the same timing.  Either output may be NULL.  Returns 0, -1 on error (te_last_error()). */
int te_diag_device_peaks(int device, double* fp64_tflops_out, double* copy_gbs_out) {
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = 0;
  double* buf = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  try {
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const size_t bytes = (size_t)1 << 30;
    CK(cudaMalloc(&buf, 2 * bytes));
    CK(cudaMemset(buf, 0, 2 * bytes));
    if (fp64_tflops_out) {
      constexpr int ILP = 8;
      const int iters = 20000, grid = prop.multiProcessorCount * 8;
      double best = 0.0;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        fp64_fma_kernel<ILP><<<grid, 256>>>(buf, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * ILP * (double)iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0) best = std::max(best, tf);
      }
      *fp64_tflops_out = best;
    }
    if (copy_gbs_out) {
      double best = 0.0;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        stream_copy_kernel<<<prop.multiProcessorCount * 8, 256>>>((const double4*)buf, (double4*)((char*)buf + bytes), bytes / 32);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::max(best, 2.0 * (double)bytes / (ms * 1e-3) / 1e9);
      }
      *copy_gbs_out = best;
    }
  } catch (const std::exception& e) {
    last_error() = e.what();
    rc = -1;
  }
  cudaFree(buf);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
