// oracle/ref_kalman_c.cpp -- TEST INFRASTRUCTURE: extern "C" handle on the REFERENCE's own LinearKalmanFilter /
// ExtendedKalmanFilter (src/kalman.cpp, compiled unmodified from /root/reference against oracle/eigen_standin; see the header
// of oracle/eigen_standin/Eigen/Dense for what that does and does not pin).  Only tests load the resulting
// oracle/_ref/libref_kalman.so.
#include <cstring>
#include <functional>
#include <stdexcept>

#include "target_estimation/kalman.hpp"

namespace {
Eigen::MatrixXd fromColMajor(const double* v, int r, int c) {
  Eigen::MatrixXd m(r, c);
  std::memcpy(m.data(), v, sizeof(double) * (size_t)r * c);
  return m;
}
Eigen::VectorXd fromVec(const double* v, int n) {
  Eigen::VectorXd x(n);
  std::memcpy(x.data(), v, sizeof(double) * (size_t)n);
  return x;
}
typedef void (*vecfn_t)(const double* x, int n, double* out, int n_out, void* ctx);
struct Handle {
  LinearKalmanFilter* kf = nullptr;
  ExtendedKalmanFilter* ekf = nullptr;   // == kf when the filter is an EKF
  int n = 0, m = 0;
};
}  // namespace

extern "C" {

// all matrices column-major (Eigen::MatrixXd storage)
void* ref_lkf_new(const double* A, const double* C, const double* Q, const double* R, const double* P, int n, int m) {
  Handle* h = new Handle;
  h->n = n; h->m = m;
  h->kf = new LinearKalmanFilter(fromColMajor(A, n, n), fromColMajor(C, m, n), fromColMajor(Q, n, n), fromColMajor(R, m, m), fromColMajor(P, n, n));
  return h;
}
// EKF with f: R^n -> R^n and h: R^n -> R^m given as C callbacks (the callers pass the oracle's restatement of the model's f / h)
void* ref_ekf_new(vecfn_t f, vecfn_t hfun, void* ctx, const double* A, const double* C, const double* Q, const double* R, const double* P, int n,
                  int m) {
  Handle* h = new Handle;
  h->n = n; h->m = m;
  auto wrap = [ctx](vecfn_t fn, int n_out) {
    return [fn, ctx, n_out](const Eigen::VectorXd& x) {
      Eigen::VectorXd out(n_out);
      fn(x.data(), (int)x.size(), out.data(), n_out, ctx);
      return out;
    };
  };
  h->ekf = new ExtendedKalmanFilter(wrap(f, n), wrap(hfun, m), fromColMajor(A, n, n), fromColMajor(C, m, n), fromColMajor(Q, n, n),
                                    fromColMajor(R, m, m), fromColMajor(P, n, n));
  h->kf = h->ekf;
  return h;
}
void ref_kf_delete(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  delete h->kf;
  delete h;
}
void ref_kf_init(void* hv, const double* x0) { Handle* h = static_cast<Handle*>(hv); h->kf->init(fromVec(x0, h->n)); }
// LinearKalmanFilter::update(y, A) / update(A) (src/kalman.cpp:97-107); returns -1 if the reference throws
int ref_kf_update_meas(void* hv, const double* y, const double* A) {
  Handle* h = static_cast<Handle*>(hv);
  try { h->kf->update(fromVec(y, h->m), fromColMajor(A, h->n, h->n)); } catch (const std::exception&) { return -1; }
  return 0;
}
int ref_kf_update(void* hv, const double* A) {
  Handle* h = static_cast<Handle*>(hv);
  try { h->kf->update(fromColMajor(A, h->n, h->n)); } catch (const std::exception&) { return -1; }
  return 0;
}
void ref_kf_get(void* hv, double* x, double* P /* column-major */) {
  Handle* h = static_cast<Handle*>(hv);
  if (x) std::memcpy(x, h->kf->getState().data(), sizeof(double) * (size_t)h->n);
  if (P) std::memcpy(P, h->kf->getP().data(), sizeof(double) * (size_t)h->n * h->n);
}

}  // extern "C"
