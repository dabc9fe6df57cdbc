// te_device.cuh -- device-side math of the batched Kalman-filter hot path (sm_100a, FP64).
//
// What the reference does per target per tick (one TargetManager::update(id,dt,meas),
// /root/reference/src/target_manager.cpp:190-202 -> src/types/*.cpp addMeasurement ->
// src/kalman.cpp:30-42,84-95) is done here for 32 targets of one pool tile by one warp,
// one target per lane.  The dense n x n products of the reference are replaced by the
// structure the models fix at compile time: A(dt) = I + dt*E_B + 0.5dt^2*E_2B (UV/UA/AR,
// src/types/uniform_acceleration.cpp:91-99) or the EKF block form (AV,
// src/types/angular_velocities.cpp:116-124) and C = [I_m 0] (e.g. uniform_velocity.cpp:43-45).
// Multiplying by the structural 1s/0s is exact, so the arithmetic below is the reference's
// arithmetic with the zero terms dropped and evaluated in the same k-order; Q, R, P0 stay
// dense and arbitrary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "te_fmod.h"

namespace te {

enum : int { ANGULAR_RATES = 0, ANGULAR_VELOCITIES = 1, UNIFORM_ACCELERATION = 2, UNIFORM_VELOCITY = 3 };
enum : int { ACT_NONE = 0, ACT_PREDICT = 1, ACT_UPDATE = 2 };

constexpr int TILE = 32;   // slots per pool tile = lanes per warp

// Per-model compile-time shape.  B = kinematic block size, NB = number of blocks,
// NPREV = extra persistent doubles (previous unwrapped rpy of the angular models,
// types/angular_rates.hpp:110).
template <int TYPE> struct Model;
template <> struct Model<UNIFORM_VELOCITY>     { static constexpr int N = 6,  M = 3, B = 3, NB = 2, NPREV = 0; static constexpr bool REGS = true;  };
template <> struct Model<UNIFORM_ACCELERATION> { static constexpr int N = 9,  M = 3, B = 3, NB = 3, NPREV = 0; static constexpr bool REGS = true;  };
template <> struct Model<ANGULAR_RATES>        { static constexpr int N = 18, M = 6, B = 6, NB = 3, NPREV = 3; static constexpr bool REGS = false; };
template <> struct Model<ANGULAR_VELOCITIES>   { static constexpr int N = 12, M = 6, B = 3, NB = 4, NPREV = 3; static constexpr bool REGS = false; };

// Tile record: NF fields x 32 lanes x 8 B, contiguous in HBM ("[tile][field][lane]").
// field order: x[N] | P[N*N] row-major | t | n_meas (int64 bits) | prev_rpy[NPREV]
template <int TYPE> struct Layout {
  using MT = Model<TYPE>;
  static constexpr int F_X = 0;
  static constexpr int F_P = MT::N;
  static constexpr int F_T = MT::N + MT::N * MT::N;
  static constexpr int F_NMEAS = F_T + 1;
  static constexpr int F_PREV = F_NMEAS + 1;
  static constexpr int NF = F_PREV + MT::NPREV;
  static constexpr int TILE_DOUBLES = NF * TILE;
  static constexpr int TILE_BYTES = TILE_DOUBLES * 8;
};

__host__ __device__ inline int model_nf(int type) {
  switch (type) {
    case UNIFORM_VELOCITY: return Layout<UNIFORM_VELOCITY>::NF;
    case UNIFORM_ACCELERATION: return Layout<UNIFORM_ACCELERATION>::NF;
    case ANGULAR_RATES: return Layout<ANGULAR_RATES>::NF;
    default: return Layout<ANGULAR_VELOCITIES>::NF;
  }
}
__host__ __device__ inline int model_n(int type) {
  switch (type) { case UNIFORM_VELOCITY: return 6; case UNIFORM_ACCELERATION: return 9; case ANGULAR_RATES: return 18; default: return 12; }
}
__host__ __device__ inline int model_m(int type) { return (type == UNIFORM_VELOCITY || type == UNIFORM_ACCELERATION) ? 3 : 6; }
__host__ __device__ inline int model_nprev(int type) { return (type == UNIFORM_VELOCITY || type == UNIFORM_ACCELERATION) ? 0 : 3; }

// -------------------------------------------------------------------------------------
// Covariance accessors: registers (UV/UA, fully unrolled constant indices) or the lane's
// column of the staged tile in shared memory (AR/AV: 324/144 doubles do not fit a thread).
// -------------------------------------------------------------------------------------
template <int N> struct RegP {
  double v[N * N];
  __device__ __forceinline__ double& operator()(int i, int j) { return v[i * N + j]; }
};
template <int N> struct SmemP {
  double* base;   // &stage[F_P * 32 + lane]
  __device__ __forceinline__ double& operator()(int i, int j) { return base[(i * N + j) * TILE]; }
};

// -------------------------------------------------------------------------------------
// geometry (restating include/target_estimation/geometry.hpp; (Eigen) where noted)
// -------------------------------------------------------------------------------------
#define TE_PI 3.14159265358979323846

__device__ __forceinline__ double constrain_angle(double x) {   // geometry.hpp:31-36
  x = fmod_exact(x + TE_PI, 2 * TE_PI);
  if (x < 0) x += 2 * TE_PI;
  return x - TE_PI;
}
__device__ __forceinline__ double angle_conv(double a) { return fmod_exact(constrain_angle(a), 2 * TE_PI); }   // :43-45
__device__ __forceinline__ double angle_diff(double a, double b) {   // geometry.hpp:53-58
  double dif = fmod_exact(b - a + TE_PI, 2 * TE_PI);
  if (dif < 0) dif += 2 * TE_PI;
  return dif - TE_PI;
}
__device__ __forceinline__ double unwrap1(double prev, double nw) {   // geometry.hpp:70-76
  return prev - angle_diff(nw, angle_conv(prev));
}
__device__ __forceinline__ double wrap_max(double x, double mx) { return fmod_exact(mx + fmod_exact(x, mx), mx); }          // :79-83
__device__ __forceinline__ double wrap_min_max(double x, double mn, double mx) { return mn + wrap_max(x - mn, mx - mn); }  // :85-88

struct Quat { double x, y, z, w; };

__device__ __forceinline__ void quat_normalize(Quat& q) {   // (Eigen) coeffs /= norm
  double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
__device__ __forceinline__ void quat_to_rpy(const Quat& q, double rpy[3]) {   // geometry.hpp:154-176
  double s = -2 * (q.x * q.z - q.w * q.y);
  if (s > 0.9999) {
    rpy[0] = 0; rpy[1] = TE_PI / 2; rpy[2] = 2 * atan2(q.z, q.w);
  } else if (s < -0.9999) {
    rpy[0] = 0; rpy[1] = -TE_PI / 2; rpy[2] = 2 * atan2(q.z, q.w);
  } else {
    rpy[0] = atan2(2 * (q.y * q.z + q.w * q.x), (q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z));
    rpy[1] = asin(s);
    rpy[2] = atan2(2 * (q.x * q.y + q.w * q.z), (q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z));
  }
}
__device__ __forceinline__ void rpy_to_quat(const double rpy[3], Quat& q) {   // geometry.hpp:178-189
  double sp, cp, st, ct, ss, cs;
  sincos(rpy[0] / 2, &sp, &cp);
  sincos(rpy[1] / 2, &st, &ct);
  sincos(rpy[2] / 2, &ss, &cs);
  q.w = cp * ct * cs + sp * st * ss;
  q.x = sp * ct * cs - cp * st * ss;
  q.y = cp * st * cs + sp * ct * ss;
  q.z = cp * ct * ss - sp * st * cs;
  quat_normalize(q);
}
__device__ __forceinline__ void quat_to_rot(const Quat& q, double R[3][3]) {   // (Eigen) toRotationMatrix
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
  R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
}
__device__ __forceinline__ void rot_to_quat(const double R[3][3], Quat& q) {   // (Eigen) Quaternion(Matrix3)
  double t = R[0][0] + R[1][1] + R[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[2][1] - R[1][2]) * t;
    q.y = (R[0][2] - R[2][0]) * t;
    q.z = (R[1][0] - R[0][1]) * t;
  } else {
    // i = argmax diag; written out per case to keep everything in registers
    int i = 0;
    if (R[1][1] > R[0][0]) i = 1;
    if (R[2][2] > (i == 0 ? R[0][0] : R[1][1])) i = 2;
    if (i == 0) {
      t = sqrt(R[0][0] - R[1][1] - R[2][2] + 1.0);
      q.x = 0.5 * t; t = 0.5 / t;
      q.w = (R[2][1] - R[1][2]) * t; q.y = (R[1][0] + R[0][1]) * t; q.z = (R[2][0] + R[0][2]) * t;
    } else if (i == 1) {
      t = sqrt(R[1][1] - R[2][2] - R[0][0] + 1.0);
      q.y = 0.5 * t; t = 0.5 / t;
      q.w = (R[0][2] - R[2][0]) * t; q.z = (R[2][1] + R[1][2]) * t; q.x = (R[0][1] + R[1][0]) * t;
    } else {
      t = sqrt(R[2][2] - R[0][0] - R[1][1] + 1.0);
      q.z = 0.5 * t; t = 0.5 / t;
      q.w = (R[1][0] - R[0][1]) * t; q.x = (R[0][2] + R[2][0]) * t; q.y = (R[1][2] + R[2][1]) * t;
    }
  }
}
__device__ __forceinline__ void rot_to_rpy(const double R[3][3], double rpy[3]) {   // geometry.hpp:191-196
  rpy[0] = atan2(R[2][1], R[2][2]);
  rpy[1] = atan2(-R[2][0], sqrt(R[2][1] * R[2][1] + R[2][2] * R[2][2]));
  rpy[2] = atan2(R[1][0], R[0][0]);
}
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {   // (Eigen) Hamilton product
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

// measurement conversion of the angular models (src/types/angular_rates.cpp:79-88):
// normalise quaternion -> rpy -> unwrap against the previous unwrapped rpy.
__device__ __forceinline__ void meas_to_unwrapped_rpy(const double q4[4], const double prev[3], double out[3]) {
  Quat q{q4[0], q4[1], q4[2], q4[3]};
  quat_normalize(q);
  double rpy[3];
  quat_to_rpy(q, rpy);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = unwrap1(prev[i], rpy[i]);
}

// -------------------------------------------------------------------------------------
// Kinematic predict (UV / UA / AR): x' = A x, P' = A P A^T + Q evaluated per position (r,c)
// of the B x B block grid on the NB x NB macro matrix, in the reference's order
// ((A P) first, then (.) A^T, then + Q; src/kalman.cpp:84-88).
// -------------------------------------------------------------------------------------
template <int N, int B, int NB, class PAcc>
__device__ __forceinline__ void predict_kinematic(PAcc& P, double* x, double dt, const double* __restrict__ Q) {
  const double h = 0.5 * dt * dt;   // Ones * 0.5 * dt * dt (uniform_acceleration.cpp:98)
#pragma unroll
  for (int i = 0; i < B; ++i) {
    if (NB == 3) {
      x[i] = x[i] + dt * x[i + B] + h * x[i + 2 * B];
      x[i + B] = x[i + B] + dt * x[i + 2 * B];
    } else {
      x[i] = x[i] + dt * x[i + B];
    }
  }
#pragma unroll
  for (int r = 0; r < B; ++r) {
#pragma unroll
    for (int c = 0; c < B; ++c) {
      double m[NB][NB];
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) m[a][b] = P(a * B + r, b * B + c);
      // A P : rows
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (NB == 3) {
          m[0][b] = m[0][b] + dt * m[1][b] + h * m[2][b];
          m[1][b] = m[1][b] + dt * m[2][b];
        } else {
          m[0][b] = m[0][b] + dt * m[1][b];
        }
      }
      // (A P) A^T : columns
#pragma unroll
      for (int a = 0; a < NB; ++a) {
        if (NB == 3) {
          m[a][0] = m[a][0] + dt * m[a][1] + h * m[a][2];
          m[a][1] = m[a][1] + dt * m[a][2];
        } else {
          m[a][0] = m[a][0] + dt * m[a][1];
        }
      }
#pragma unroll
      for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b)
          P(a * B + r, b * B + c) = m[a][b] + __ldg(&Q[(a * B + r) * N + (b * B + c)]);
    }
  }
}

// -------------------------------------------------------------------------------------
// EKF predict of the angular-velocities model (src/types/angular_velocities.cpp:116-140,
// src/kalman.cpp:129-133).  State [p rpy v w]; A linearised at the previous posterior.
// -------------------------------------------------------------------------------------
template <class PAcc>
__device__ __forceinline__ void predict_av(PAcc& P, double* x, double dt, const double* __restrict__ Q) {
  constexpr int N = 12;
  double s_r, c_r, s_p, c_p;
  sincos(x[3], &s_r, &c_r);
  sincos(x[4], &s_p, &c_p);
  const double wx = x[9], wy = x[10], wz = x[11];
  // EarBaseInvJacobianRpy (geometry.hpp:394-410)
  double J1[3][3];
  J1[0][0] = (dt * (wy * c_r * s_p - wz * s_p * s_r)) / c_p + 1;
  J1[0][1] = (dt * (wz * c_r + wy * s_r)) / (c_p * c_p);
  J1[0][2] = 0;
  J1[1][0] = -dt * (wz * c_r + wy * s_r);
  J1[1][1] = 1;
  J1[1][2] = 0;
  J1[2][0] = (dt * (wy * c_r - wz * s_r)) / c_p;
  J1[2][1] = (dt * s_p * (wz * c_r + wy * s_r)) / (c_p * c_p);
  J1[2][2] = 1;
  // EarBaseInvJacobianOmega (geometry.hpp:412-426)
  double J2[3][3];
  J2[0][0] = dt; J2[0][1] = (dt * s_p * s_r) / c_p; J2[0][2] = (dt * c_r * s_p) / c_p;
  J2[1][0] = 0;  J2[1][1] = dt * c_r;               J2[1][2] = -dt * s_r;
  J2[2][0] = 0;  J2[2][1] = (dt * s_r) / c_p;       J2[2][2] = (dt * c_r) / c_p;
  // f(x): rpyToEarBaseInv (geometry.hpp:359-374), angular_velocities.cpp:126-140
  double E[3][3];
  E[0][0] = 1; E[0][1] = (s_p * s_r) / c_p; E[0][2] = (c_r * s_p) / c_p;
  E[1][0] = 0; E[1][1] = c_r;               E[1][2] = -s_r;
  E[2][0] = 0; E[2][1] = s_r / c_p;         E[2][2] = c_r / c_p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x[i] = x[i] + dt * x[6 + i];
    double s = (dt * E[i][0]) * wx + (dt * E[i][1]) * wy + (dt * E[i][2]) * wz;
    x[3 + i] = x[3 + i] + s;
  }
  // A P : per column j.  rows 0..2 += dt*rows 6..8 ; rows 3..5 = J1*rows 3..5 + J2*rows 9..11
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double p0[3], p1[3], p2[3], p3[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { p0[i] = P(i, j); p1[i] = P(3 + i, j); p2[i] = P(6 + i, j); p3[i] = P(9 + i, j); }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      P(i, j) = p0[i] + dt * p2[i];
      P(3 + i, j) = J1[i][0] * p1[0] + J1[i][1] * p1[1] + J1[i][2] * p1[2] + J2[i][0] * p3[0] + J2[i][1] * p3[1] + J2[i][2] * p3[2];
    }
  }
  // (A P) A^T + Q : per row i.  cols 0..2 += dt*cols 6..8 ; cols 3..5 = r1*J1^T + r3*J2^T
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double r[N];
#pragma unroll
    for (int j = 0; j < N; ++j) r[j] = P(i, j);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      P(i, j) = (r[j] + r[6 + j] * dt) + __ldg(&Q[i * N + j]);
      double s = r[3] * J1[j][0] + r[4] * J1[j][1] + r[5] * J1[j][2] + r[9] * J2[j][0] + r[10] * J2[j][1] + r[11] * J2[j][2];
      P(i, 3 + j) = s + __ldg(&Q[i * N + 3 + j]);
      P(i, 6 + j) = r[6 + j] + __ldg(&Q[i * N + 6 + j]);
      P(i, 9 + j) = r[9 + j] + __ldg(&Q[i * N + 9 + j]);
    }
  }
}

// -------------------------------------------------------------------------------------
// Measurement update with selector C = [I_M 0] (src/kalman.cpp:90-95):
//   S = P[0:M,0:M] + R ; K = P[:,0:M] S^-1 ; x += K (y - x[0:M]) ; P = (I - K C) P
// evaluated as  v = S^-1 (y - x[0:M]),  x += P[:,0:M] v,  W = S^-1 P[0:M,:],
// P -= P[:,0:M] W, with S^-1 applied through an in-register Cholesky factor
// (S is symmetric positive definite: covariance block + R).  Differs from the reference's
// partial-pivot-LU inverse only in rounding (measured < 1e-12 relative, tests/).
// -------------------------------------------------------------------------------------
template <int M> struct Chol {
  // lower triangle in a full M x M register array (the upper half is never touched and costs nothing);
  // the diagonal holds 1/L_jj.  Every loop runs over a constant range with compile-time guards so that the
  // unroller resolves all indices (triangular trip counts left the factor in local memory).
  double L[M][M];
  __device__ __forceinline__ double& at(int i, int j) { return L[i][j]; }
  __device__ __forceinline__ void factor() {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double d = L[j][j];
#pragma unroll
      for (int k = 0; k < M; ++k)
        if (k < j) d -= L[j][k] * L[j][k];
      // reciprocal square root: one MUFU seed + Newton steps instead of sqrt followed by a division -- this scalar sits
      // on the longest dependent chain of the whole step (six of them in a row for M = 6); <= 2 ulp, far inside 1e-9
      const double inv = rsqrt(d);
#pragma unroll
      for (int i = 0; i < M; ++i) {
        if (i > j) {
          double s = L[i][j];
#pragma unroll
          for (int k = 0; k < M; ++k)
            if (k < j) s -= L[i][k] * L[j][k];
          L[i][j] = s * inv;
        }
      }
      L[j][j] = inv;
    }
  }
  // solve (L L^T) z = b in place
  __device__ __forceinline__ void solve(double* b) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      double s = b[i];
#pragma unroll
      for (int k = 0; k < M; ++k)
        if (k < i) s -= L[i][k] * b[k];
      b[i] = s * L[i][i];
    }
#pragma unroll
    for (int ii = 0; ii < M; ++ii) {
      const int i = M - 1 - ii;
      double s = b[i];
#pragma unroll
      for (int k = 0; k < M; ++k)
        if (k > i) s -= L[k][i] * b[k];
      b[i] = s * L[i][i];
    }
  }
};

template <int N, int M, class PAcc>
__device__ __forceinline__ void kf_update(PAcc& P, double* x, const double* y, const double* __restrict__ R) {
  static_assert(N % M == 0, "column groups of width M");
  Chol<M> ch;
#pragma unroll
  for (int i = 0; i < M; ++i)
#pragma unroll
    for (int j = 0; j < M; ++j)
      if (j <= i) ch.at(i, j) = P(i, j) + __ldg(&R[i * M + j]);
  ch.factor();
  double v[M];
#pragma unroll
  for (int k = 0; k < M; ++k) v[k] = y[k] - x[k];
  ch.solve(v);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) s += P(i, k) * v[k];
    x[i] += s;
  }
  // column groups from the last to the first so that P[:,0:M] is overwritten last
#pragma unroll
  for (int g = N / M - 1; g >= 0; --g) {
    double W[M][M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      double col[M];
#pragma unroll
      for (int k = 0; k < M; ++k) col[k] = P(k, g * M + j);
      ch.solve(col);
#pragma unroll
      for (int k = 0; k < M; ++k) W[k][j] = col[k];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double row[M];
#pragma unroll
      for (int k = 0; k < M; ++k) row[k] = P(i, k);
#pragma unroll
      for (int j = 0; j < M; ++j) {
        double s = P(i, g * M + j);
#pragma unroll
        for (int k = 0; k < M; ++k) s -= row[k] * W[k][j];
        P(i, g * M + j) = s;
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// One full target step for one lane (addMeasurement / update of src/types/*.cpp).
//   stage : this warp's staged tile in shared memory, lane-interleaved ([field][32])
//   meas  : the lane's 7-vector [x y z qx qy qz qw] (only read when action == ACT_UPDATE)
// -------------------------------------------------------------------------------------
template <int TYPE>
__device__ __forceinline__ void step_lane(double* stage, int lane, int action, double dt, const double* meas,
                                          const double* __restrict__ Q, const double* __restrict__ R) {
  using MT = Model<TYPE>;
  using LY = Layout<TYPE>;
  constexpr int N = MT::N, M = MT::M;
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = stage[(LY::F_X + i) * TILE + lane];

  double y[M];
  if (action == ACT_UPDATE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) y[k] = meas[k];
    if (M == 6) {
      double prev[3], un[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) prev[k] = stage[(LY::F_PREV + k) * TILE + lane];
      meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        y[3 + k] = un[k];
        stage[(LY::F_PREV + k) * TILE + lane] = un[k];   // meas_rpy_internal_ = unwrapped
      }
    }
  }

  if (MT::REGS) {
    RegP<N> P;
#pragma unroll
    for (int k = 0; k < N * N; ++k) P.v[k] = stage[(LY::F_P + k) * TILE + lane];
    predict_kinematic<N, MT::B, MT::NB>(P, x, dt, Q);
    if (action == ACT_UPDATE) kf_update<N, M>(P, x, y, R);
#pragma unroll
    for (int k = 0; k < N * N; ++k) stage[(LY::F_P + k) * TILE + lane] = P.v[k];
  } else {
    SmemP<N> P{stage + LY::F_P * TILE + lane};
    if (TYPE == ANGULAR_VELOCITIES) predict_av(P, x, dt, Q);
    else predict_kinematic<N, MT::B, MT::NB>(P, x, dt, Q);
    if (action == ACT_UPDATE) kf_update<N, M>(P, x, y, R);
  }

#pragma unroll
  for (int i = 0; i < N; ++i) stage[(LY::F_X + i) * TILE + lane] = x[i];
  // updateTime (src/target_interface.cpp:148-152) / updateMeasurement (:142-146)
  stage[LY::F_T * TILE + lane] = stage[LY::F_T * TILE + lane] + dt;
  if (action == ACT_UPDATE) {
    long long* nm = reinterpret_cast<long long*>(stage + LY::F_NMEAS * TILE + lane);
    *nm = *nm + 1;
  }
}

// -------------------------------------------------------------------------------------
// Derived outputs (updateTargetState + getters of src/types/*.cpp, src/target_interface.cpp).
// x: the lane's state, t: target time.  t1: query time (use_t1 == false -> current values).
// pose[7] twist[6] acc[6] pose6[6] (any may be null).
// -------------------------------------------------------------------------------------
template <int TYPE>
__device__ __forceinline__ void derive_outputs(const double* x, double t, bool use_t1, double t1,
                                               double* pose, double* twist, double* acc, double* pose6) {
  double tw[6], ac[6], p6[6], R[3][3];
  Quat q{0, 0, 0, 1};
  bool angular = (TYPE == ANGULAR_RATES || TYPE == ANGULAR_VELOCITIES);
  if (!angular) {
    // uniform_velocity.cpp:98-115, uniform_acceleration.cpp:101-118
#pragma unroll
    for (int i = 0; i < 3; ++i) { tw[i] = x[3 + i]; tw[3 + i] = 0.0; p6[i] = x[i]; p6[3 + i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { ac[i] = (TYPE == UNIFORM_ACCELERATION) ? x[6 + i] : 0.0; ac[3 + i] = 0.0; }
  } else {
    // angular_rates.cpp:117-138, angular_velocities.cpp:153-169
    double rpy[3] = {x[3], x[4], x[5]};
    rpy_to_quat(rpy, q);
    quat_to_rot(q, R);
    double rr[3];
    rot_to_rpy(R, rr);
#pragma unroll
    for (int i = 0; i < 3; ++i) { tw[i] = x[6 + i]; p6[i] = x[i]; p6[3 + i] = rr[i]; }
    if (TYPE == ANGULAR_RATES) {
      double s_r, c_r, s_p, c_p;   // rpyToEarBase (geometry.hpp:333-351)
      sincos(rr[0], &s_r, &c_r);
      sincos(rr[1], &s_p, &c_p);
      const double r0 = x[9], r1 = x[10], r2 = x[11];
      tw[3] = 1 * r0 + 0 * r1 + (-s_p) * r2;
      tw[4] = 0 * r0 + c_r * r1 + (c_p * s_r) * r2;
      tw[5] = 0 * r0 + (-s_r) * r1 + (c_p * c_r) * r2;
#pragma unroll
      for (int i = 0; i < 6; ++i) ac[i] = x[12 + i];
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) tw[3 + i] = x[9 + i];
#pragma unroll
      for (int i = 0; i < 6; ++i) ac[i] = 0.0;
    }
  }
  if (pose6) {
#pragma unroll
    for (int i = 0; i < 6; ++i) pose6[i] = p6[i];
  }
  if (!use_t1) {
    if (pose) {
      pose[0] = x[0]; pose[1] = x[1]; pose[2] = x[2];
      if (angular) { Quat qe; rot_to_quat(R, qe); pose[3] = qe.x; pose[4] = qe.y; pose[5] = qe.z; pose[6] = qe.w; }
      else { pose[3] = 0; pose[4] = 0; pose[5] = 0; pose[6] = 1; }
    }
    if (twist) {
#pragma unroll
      for (int i = 0; i < 6; ++i) twist[i] = tw[i];
    }
  } else {
    const double tau = t1 - t;
    if (pose) {
      if (TYPE == UNIFORM_VELOCITY) {   // uniform_velocity.cpp:117-127
#pragma unroll
        for (int i = 0; i < 3; ++i) pose[i] = x[i] + tw[i] * tau;
        pose[3] = 0; pose[4] = 0; pose[5] = 0; pose[6] = 1;
      } else if (TYPE == UNIFORM_ACCELERATION) {   // uniform_acceleration.cpp:120-130
#pragma unroll
        for (int i = 0; i < 3; ++i) pose[i] = x[i] + tw[i] * tau + 0.5 * ac[i] * tau * tau;
        pose[3] = 0; pose[4] = 0; pose[5] = 0; pose[6] = 1;
      } else if (TYPE == ANGULAR_RATES) {   // angular_rates.cpp:140-151
        double v6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) v6[i] = p6[i] + tw[i] * tau + 0.5 * ac[i] * tau * tau;
        Quat qq;
        rpy_to_quat(v6 + 3, qq);
        quat_normalize(qq);
        pose[0] = v6[0]; pose[1] = v6[1]; pose[2] = v6[2];
        pose[3] = qq.x; pose[4] = qq.y; pose[5] = qq.z; pose[6] = qq.w;
      } else {   // angular_velocities.cpp:171-184
#pragma unroll
        for (int i = 0; i < 3; ++i) pose[i] = x[i] + tw[i] * tau;
        Quat qq;
        rpy_to_quat(p6 + 3, qq);
        // Qtran(tau, omega) * q  (geometry.hpp:448-465,493-504)
        const double w0 = tw[3], w1 = tw[4], w2 = tw[5];
        const double wn = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
        double c[4] = {qq.x, qq.y, qq.z, qq.w}, r4[4];
        if (wn > 0.0) {
          double sn, cs;
          sincos(wn * tau / 2.0, &sn, &cs);
          const double f = 2.0 / wn * sn;
          const double S[4][4] = {{0.5 * 0, 0.5 * -w2, 0.5 * w1, 0.5 * w0},
                                  {0.5 * w2, 0.5 * 0, 0.5 * -w0, 0.5 * w1},
                                  {0.5 * -w1, 0.5 * w0, 0.5 * 0, 0.5 * w2},
                                  {0.5 * -w0, 0.5 * -w1, 0.5 * -w2, 0.5 * 0}};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) s += (cs * (i == j ? 1.0 : 0.0) + f * S[i][j]) * c[j];
            r4[i] = s;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) r4[i] = c[i];
        }
        Quat qo{r4[0], r4[1], r4[2], r4[3]};
        quat_normalize(qo);
        pose[3] = qo.x; pose[4] = qo.y; pose[5] = qo.z; pose[6] = qo.w;
      }
    }
    if (twist) {
      // getEstimatedTwist(t1): UA/AR extrapolate (uniform_acceleration.cpp:132-136,
      // angular_rates.cpp:153-157); UV/AV return the current twist.
#pragma unroll
      for (int i = 0; i < 6; ++i)
        twist[i] = (TYPE == UNIFORM_ACCELERATION || TYPE == ANGULAR_RATES) ? (tw[i] + ac[i] * tau) : tw[i];
    }
  }
  if (acc) {   // getEstimatedAcceleration([t]) (src/target_interface.cpp:114-121,136-140)
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i] = ac[i];
  }
}

// -------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copies (cp.async.bulk -> SASS UBLKCP).
// -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace te
