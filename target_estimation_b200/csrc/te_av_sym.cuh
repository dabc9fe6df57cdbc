// te_av_sym.cuh -- register-resident step of the angular-velocities EKF (n = 12, m = 6) for one lane.
//
// The full 12 x 12 covariance (144 doubles) does not fit a thread, its upper triangle (78) plus the state (12) does --
// exactly the 90 doubles of the uniform-acceleration model, whose thread-per-target kernel runs at 0.92 of the HBM
// peak.  The covariance is symmetric up to rounding in the reference too (P0, Q, R symmetric; P = (I - K C) P is
// symmetric in exact arithmetic; measured asymmetry <= 1e-15 relative), so this path carries the upper
// triangle only, reads the upper triangle of the stored matrix and writes both halves back.  The pool takes it only
// when every registered class has bitwise-symmetric Q, R and P0; otherwise the general row-split kernel runs.
//
//   predict (src/types/angular_velocities.cpp:116-140, src/kalman.cpp:129-133), block form over [p rpy v w]:
//       A = [I 0 dtI 0; 0 J1 0 J2; 0 0 I 0; 0 0 0 I],  P' = (A P) A^T + Q  block by block, in an order that lets every
//       block be overwritten in place
//   update  (src/kalman.cpp:135-140) with C = [I6 0]:  S = P'[0:6,0:6] + R = L L^T,  Z = L^-1 P'[0:6,:],
//       x += Z^T L^-1 (y - x'[0:6]),  P = P' - Z^T Z   ( = (I - K C) P' with K = P'[:,0:6] S^-1 )
//       Z (6 x 12) is parked in the lane's own column of the staged tile (the covariance fields are dead while P lives
//       in registers), one row at a time back into registers for the rank-1 downdates.
#pragma once
#include "te_device.cuh"

namespace te {

// upper triangle, row-major packed; both index orders address the same element
template <int N> struct SymP {
  double v[N * (N + 1) / 2];
  __device__ __forceinline__ double& operator()(int i, int j) {
    return i <= j ? v[i * N - (i * (i - 1)) / 2 + (j - i)] : v[j * N - (j * (j - 1)) / 2 + (i - j)];
  }
};

// What the step needs of the previous posterior besides the covariance: the entries of the two Jacobians and the attitude
// increment of f(x).  Depends on five state entries only (roll, pitch, body rates) -- the streaming kernel evaluates it
// from a prefetched copy of those while the tile itself is still in flight (te_av_stream.cuh).
struct AvFront {
  double j10, j11, j12, j13, j14;          // J1 = [j10 j11 0; j12 1 0; j13 j14 1]   (EarBaseInvJacobianRpy, geometry.hpp:394-410)
  double j20, j21, j22, j23, j24, j25;     // J2 = [dt j20 j21; 0 j22 j23; 0 j24 j25] (EarBaseInvJacobianOmega, :412-426)
  double d3, d4, d5;                       // f(x)[3..5] - x[3..5] = dt * EarBaseInv(rpy) * w (geometry.hpp:359-374, angular_velocities.cpp:126-140)
  double dt;
  __device__ __forceinline__ void eval(double roll, double pitch, double wx, double wy, double wz, double dt_) {
    dt = dt_;
    double s_r, c_r, s_p, c_p;
    sincos(roll, &s_r, &c_r);
    sincos(pitch, &s_p, &c_p);
    j10 = (dt * (wy * c_r * s_p - wz * s_p * s_r)) / c_p + 1;
    j11 = (dt * (wz * c_r + wy * s_r)) / (c_p * c_p);
    j12 = -dt * (wz * c_r + wy * s_r);
    j13 = (dt * (wy * c_r - wz * s_r)) / c_p;
    j14 = (dt * s_p * (wz * c_r + wy * s_r)) / (c_p * c_p);
    j20 = (dt * s_p * s_r) / c_p;
    j21 = (dt * c_r * s_p) / c_p;
    j22 = dt * c_r;
    j23 = -dt * s_r;
    j24 = (dt * s_r) / c_p;
    j25 = (dt * c_r) / c_p;
    const double E01 = (s_p * s_r) / c_p, E02 = (c_r * s_p) / c_p, E11 = c_r, E12 = -s_r, E21 = s_r / c_p, E22 = c_r / c_p;
    d3 = ((dt * 1.0) * wx + (dt * E01) * wy + (dt * E02) * wz);
    d4 = ((dt * 0.0) * wx + (dt * E11) * wy + (dt * E12) * wz);
    d5 = ((dt * 0.0) * wx + (dt * E21) * wy + (dt * E22) * wz);
  }
  // row i of J1 / J2 as compile-time-indexed values (structural 0 / 1 / dt entries fold away)
  __device__ __forceinline__ double J1(int i, int k) const {
    return i == 0 ? (k == 0 ? j10 : (k == 1 ? j11 : 0.0)) : (i == 1 ? (k == 0 ? j12 : (k == 1 ? 1.0 : 0.0)) : (k == 0 ? j13 : (k == 1 ? j14 : 1.0)));
  }
  __device__ __forceinline__ double J2(int i, int k) const {
    return i == 0 ? (k == 0 ? dt : (k == 1 ? j20 : j21)) : (i == 1 ? (k == 0 ? 0.0 : (k == 1 ? j22 : j23)) : (k == 0 ? 0.0 : (k == 1 ? j24 : j25)));
  }
};

// ---- covariance predict, blocks p = 0..2, r = 3..5, v = 6..8, w = 9..11.  T = A P, P' = T A^T + Q:
//   P'pp = Tpp + dt Tpv          Tpp = Ppp + dt Pvp, Tpv = Ppv + dt Pvv
//   P'pr = Tpr J1^T + Tpw J2^T   Tpr = Ppr + dt Pvr, Tpw = Ppw + dt Pvw
//   P'pv = Tpv,  P'pw = Tpw
//   P'rr = Trr J1^T + Trw J2^T   Trr = J1 Prr + J2 Pwr, Trw = J1 Prw + J2 Pww
//   P'rv = J1 Prv + J2 Pwv,  P'rw = Trw,  vv / vw / ww unchanged.
// In place, in an order in which every block reads only values that are still the old ones (or the T it needs), with at
// most nine temporaries alive.  QV(idx) = entry idx of the class's row-major Q.
template <class QV>
__device__ __forceinline__ void av_predict_cov(SymP<12>& P, const AvFront& F, QV Qv) {
  constexpr int N = 12;
  const double dt = F.dt;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (i <= j) {
        const double tpp = P(i, j) + dt * P(6 + i, j);
        const double tpv = P(i, 6 + j) + dt * P(6 + i, 6 + j);
        P(i, j) = (tpp + tpv * dt) + Qv(i * N + j);
      }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) P(i, 9 + j) = P(i, 9 + j) + dt * P(6 + i, 9 + j);   // Ppw <- Tpw (Q added below)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double tpr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) tpr[k] = P(i, 3 + k) + dt * P(6 + i, 3 + k);       // Pvr(i,k) lives at P(3+k, 6+i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double s = tpr[0] * F.J1(j, 0) + tpr[1] * F.J1(j, 1) + tpr[2] * F.J1(j, 2) + P(i, 9) * F.J2(j, 0) + P(i, 10) * F.J2(j, 1) + P(i, 11) * F.J2(j, 2);
      P(i, 3 + j) = s + Qv(i * N + 3 + j);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      P(i, 9 + j) = P(i, 9 + j) + Qv(i * N + 9 + j);
      P(i, 6 + j) = (P(i, 6 + j) + dt * P(6 + i, 6 + j)) + Qv(i * N + 6 + j);
    }
  {
    double trr[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += F.J1(i, k) * P(3 + k, 3 + j);
#pragma unroll
        for (int k = 0; k < 3; ++k) s += F.J2(i, k) * P(9 + k, 3 + j);              // Pwr(k,j) lives at P(3+j, 9+k)
        trr[i][j] = s;
      }
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // Prw <- Trw, one column at a time
      double c[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += F.J1(i, k) * P(3 + k, 9 + j);
#pragma unroll
        for (int k = 0; k < 3; ++k) s += F.J2(i, k) * P(9 + k, 9 + j);
        c[i] = s;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) P(3 + i, 9 + j) = c[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (i <= j) {
          const double s = trr[i][0] * F.J1(j, 0) + trr[i][1] * F.J1(j, 1) + trr[i][2] * F.J1(j, 2) + P(3 + i, 9) * F.J2(j, 0) + P(3 + i, 10) * F.J2(j, 1) +
                           P(3 + i, 11) * F.J2(j, 2);
          P(3 + i, 3 + j) = s + Qv((3 + i) * N + 3 + j);
        }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {     // Prv <- J1 Prv + J2 Pwv, one column at a time; Prw += Q
    double c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += F.J1(i, k) * P(3 + k, 6 + j);
#pragma unroll
      for (int k = 0; k < 3; ++k) s += F.J2(i, k) * P(9 + k, 6 + j);                // Pwv(k,j) lives at P(6+j, 9+k)
      c[i] = s;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      P(3 + i, 6 + j) = c[i] + Qv((3 + i) * N + 6 + j);
      P(3 + i, 9 + j) = P(3 + i, 9 + j) + Qv((3 + i) * N + 9 + j);
    }
  }
#pragma unroll
  for (int i = 6; i < N; ++i)
#pragma unroll
    for (int j = 6; j < N; ++j)
      if (i <= j) P(i, j) = P(i, j) + Qv(i * N + j);
}

// ---- update (src/kalman.cpp:135-140 with C = [I6 0]) ----
// xs  : the lane's column holding x' (entry j at xs[j * TILE]); updated in place
// zs  : the lane's column of a 72-field scratch: Z[k][j] at zs[(k * N + j) * TILE]
// ypos: measured position (3 values, registers); yang: the lane's column of the unwrapped measured angles (stride TILE)
// (the empty asm statements are scheduling fences for the front end: without them it hoists the shared-memory loads of
//  all six Z rows / interleaves all twelve column solves, and ptxas has to spill ~80 doubles)
template <int ZF, class RV>
__device__ __forceinline__ void av_update(SymP<12>& P, double* xs, double* zs, const double* ypos, const double* yang, RV Rv) {
  constexpr int N = 12, M = 6;
  {
    Chol<M> ch;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
      for (int j = 0; j < M; ++j)
        if (j <= i) ch.at(i, j) = P(j, i) + Rv(i * M + j);
    ch.factor();
    double u[M];   // L^-1 (y - x'[0:6])
#pragma unroll
    for (int k = 0; k < M; ++k) {
      const double yk = k < 3 ? ypos[k < 3 ? k : 0] : yang[(k < 3 ? 0 : k - 3) * TILE];
      double s = yk - xs[k * TILE];
#pragma unroll
      for (int m = 0; m < M; ++m)
        if (m < k) s -= ch.L[k][m] * u[m];
      u[k] = s * ch.L[k][k];
    }
    // Z = L^-1 P'[0:6,:] column by column (forward substitution; the diagonal of ch holds 1 / L_kk), and with column j
    // the state update x_j += Z[:,j] . u
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double z[M];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        double s = P(k, j);
#pragma unroll
        for (int m = 0; m < M; ++m)
          if (m < k) s -= ch.L[k][m] * z[m];
        z[k] = s * ch.L[k][k];
      }
      double xj = xs[j * TILE];
#pragma unroll
      for (int k = 0; k < M; ++k) {
        zs[(k * N + j) * TILE] = z[k];
        xj += z[k] * u[k];
      }
      xs[j * TILE] = xj;
      if (j % ZF == ZF - 1) asm volatile("" ::: "memory");
    }
  }
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double zr[N];
#pragma unroll
    for (int j = 0; j < N; ++j) zr[j] = zs[(k * N + j) * TILE];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (i <= j) P(i, j) -= zr[i] * zr[j];
    asm volatile("" ::: "memory");
  }
}

// ---- the same update as M scalar updates, everything in registers ----
// With R = L L^T and T = L^-1 (per class, from the host: StepArgs::Tc / Ttab) the whitened measurement T y = (T C) x + e has unit,
// uncorrelated noise, so its six components may be applied one after the other -- in exact arithmetic the posterior of the joint
// update (x + K (y - C x), (I - K C) P; src/kalman.cpp:135-140).  Row i of T C has the entries T[i][0..i] at the state indices
// 0..i, so scalar update i is   g = P[:, 0..i] T[i][0..i]^T,  s = T[i][0..i] g[0..i] + 1,  r = T[i][0..i] (y - x)[0..i],
// x += g r / s,  P -= g g^T / s.   No 6 x 12 intermediate (the joint form parks Z = L_S^-1 P'[0:6,:] in shared memory): the
// streaming kernel needs its shared memory for the NEXT tile.  Against the joint form on the SURVEY.md 8(d) streams (numpy, 24
// targets x 2000 ticks): <= 1.4e-11 relative on state and covariance.
// y: the measurement (position, unwrapped angles), entry k at y[k * YS] (registers: YS = 1; the lane's column of a parking area
// in shared memory: YS = TILE).  The empty asm statements keep the front end from interleaving two scalar updates (each needs the
// covariance the previous one left; hoisted loads only cost registers).
template <int YS, class TV>
__device__ __forceinline__ void av_update_seq(SymP<12>& P, double* x, const double* y, TV Tv) {
  constexpr int N = 12, M = 6;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double g[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = P(j, 0) * Tv(i * M + 0);
#pragma unroll
      for (int k = 1; k < M; ++k)
        if (k <= i) s += P(j, k) * Tv(i * M + k);
      g[j] = s;
    }
    double s = 1.0, r = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k)
      if (k <= i) {
        s += Tv(i * M + k) * g[k];
        r += Tv(i * M + k) * (y[k * YS] - x[k]);
      }
    const double inv = 1.0 / s;
    const double ri = r * inv;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      x[j] += g[j] * ri;
      const double gj = g[j] * inv;
#pragma unroll
      for (int l = 0; l < N; ++l)
        if (j <= l) P(j, l) -= gj * g[l];
    }
    if (i + 1 < M) asm volatile("" ::: "memory");   // (none after the last one: the caller's stores may start under its downdates)
  }
}

// in  : the lane's column of the source tile   (field f at in[f * TILE]); staged tile in shared memory or the tile in HBM
// out : the lane's column of the destination tile (may alias in)
// sc  : the lane's column of a shared-memory scratch with the tile's field numbering for x (F_X..) and for the first 72
//       covariance fields (Z), and the unwrapped measurement angles at field PREV_S
// The staged kernel passes in = out = sc = the stage (PREV_S = F_PREV, DIRECT = false); the direct kernel streams in / out
// from / to HBM (DIRECT = true: x is copied from the scratch to out at the end).
template <int PREV_S, bool DIRECT, int ZF = 2>
__device__ __forceinline__ void step_lane_av_sym(const double* in, double* out, double* sc, int action, double dt, const double* meas,
                                                 const double* __restrict__ Q, const double* __restrict__ R, bool packed = false) {
  using LY = Layout<ANGULAR_VELOCITIES>;
  constexpr int N = 12;
  // Register budget: the 78 covariance entries stay in registers from the load to the store; everything else is kept
  // short-lived (state, innovation and Z rows are parked in the scratch column between uses).

  // every load from `in` is issued up front: when `in` is HBM these ~100 independent loads per lane are the whole memory
  // latency of the step, and the measurement conversion below (a long dependent chain) runs underneath them
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = in[(LY::F_X + i) * TILE];
  const double t_in = in[LY::F_T * TILE];
  const long long nm_in = reinterpret_cast<const long long*>(in)[LY::F_NMEAS * TILE];
  double prev[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) prev[k] = in[(LY::F_PREV + k) * TILE];
  SymP<N> P;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (i <= j) P(i, j) = in[(LY::F_P + i * N + j) * TILE];

  // ---- measurement conversion (angular_velocities.cpp:87-96): unwrapped rpy = y[3..5] = new meas_rpy_internal_ ----
  if (action == ACT_UPDATE) {
    double un[3];
    meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sc[(PREV_S + k) * TILE] = un[k];
      prev[k] = un[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) out[(LY::F_PREV + k) * TILE] = prev[k];

  // ---- state predict x' = f(x) and the Jacobians at the previous posterior ----
  AvFront F;
  F.eval(x[3], x[4], x[9], x[10], x[11], dt);
  // f(x): p += dt v ; rpy += dt * EarBaseInv(rpy) * w
#pragma unroll
  for (int i = 0; i < 3; ++i) sc[(LY::F_X + i) * TILE] = x[i] + dt * x[6 + i];
  sc[(LY::F_X + 3) * TILE] = x[3] + F.d3;
  sc[(LY::F_X + 4) * TILE] = x[4] + F.d4;
  sc[(LY::F_X + 5) * TILE] = x[5] + F.d5;
#pragma unroll
  for (int i = 6; i < N; ++i) sc[(LY::F_X + i) * TILE] = x[i];

  av_predict_cov(P, F, [&](int idx) -> double { return __ldg(&Q[idx]); });
  if (action == ACT_UPDATE)
    av_update<ZF>(P, sc + LY::F_X * TILE, sc + LY::F_P * TILE, meas, sc + PREV_S * TILE, [&](int idx) -> double { return __ldg(&R[idx]); });

#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (i <= j || !packed) out[(LY::F_P + i * N + j) * TILE] = P(i, j);   // packed: the pool mirrors the upper triangle on demand
  if (DIRECT) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[(LY::F_X + i) * TILE] = sc[(LY::F_X + i) * TILE];
  }
  // updateTime (src/target_interface.cpp:148-152) / updateMeasurement (:142-146)
  out[LY::F_T * TILE] = t_in + dt;
  reinterpret_cast<long long*>(out)[LY::F_NMEAS * TILE] = nm_in + (action == ACT_UPDATE ? 1 : 0);
}

}  // namespace te
