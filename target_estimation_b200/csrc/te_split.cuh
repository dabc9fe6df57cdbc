// te_split.cuh -- row-split KF step kernel for the large-state models (AR n=18, AV n=12).
//
// One CTA of RS = 6 warps works on one tile of 32 targets: lane = target, warp r owns the rows
// {q*6 + r} of every target's covariance (AR: rows r, 6+r, 12+r = one row of each kinematic block;
// AV: rows r, 6+r = (p_r, v_r) for r < 3 and (rpy_i, w_i) for r = 3 + i), held in registers.  The tile itself is staged in
// shared memory by one TMA bulk copy and written back by one bulk store, as in kf_step_kernel; what the
// row owners must exchange -- the top M rows of the predicted covariance and W = S^-1 P'[0:M,:] -- goes
// through the staged tile in place, with CTA barriers between the phases:
//
//   A  predict own rows (x' = A x | f(x), P' = A P A^T + Q); warps 0..2 convert one Euler angle each
//      (quat -> rpy -> unwrap); publish row r of P', x'[r], y                                              | barrier
//   B  every warp factors S = P'[0:M,0:M] + R (in-register Cholesky), solves v = S^-1 (y - x'[0:M]) and
//      its N/6 columns of W into a separate W buffer                                                     | barrier
//   C  own rows: x += P'[rows,0:M] v ; P[rows,:] = P'[rows,:] - P'[rows,0:M] W
//   D  own rows back into the stage, t / n_meas bookkeeping                                              | barrier, bulk store
// (AV has one more barrier in A: warps 3..5 read the original rows 3..5 / 9..11 before they are republished.)
//
// Same arithmetic as te_device.cuh's step_lane (predict_kinematic / predict_av / kf_update); only the
// ownership of the rows differs.  Reference: src/kalman.cpp:84-95,129-140, src/types/angular_rates.cpp:72-115,
// src/types/angular_velocities.cpp:80-151.
#pragma once
#include "te_kernels.cuh"

namespace te {

template <int TYPE> struct Split { static constexpr int RS = 6; };   // AR: rows r, 6+r, 12+r ; AV: rows r, 6+r

// shared memory of one CTA: [mbarriers 1 KB][STAGES x (tile + measurement block)][W: M x N x 32][y: 6 x 32]
template <int TYPE> __host__ __device__ constexpr size_t split_smem_bytes(int stages, bool wsep) {
  return 1024 + ((size_t)stages * stage_doubles<TYPE>() + (wsep ? (size_t)Model<TYPE>::M * Model<TYPE>::N * TILE : 0) + 6 * TILE) * 8;
}

__device__ __forceinline__ double sel3(const double a[3], int r) { return r == 0 ? a[0] : (r == 1 ? a[1] : a[2]); }

// component k of quatToRpy (geometry.hpp:154-176), same expressions as quat_to_rpy()
__device__ __forceinline__ double quat_to_rpy_comp(const Quat& q, int k) {
  const double s = -2 * (q.x * q.z - q.w * q.y);
  if (s > 0.9999) return k == 0 ? 0.0 : (k == 1 ? TE_PI / 2 : 2 * atan2(q.z, q.w));
  if (s < -0.9999) return k == 0 ? 0.0 : (k == 1 ? -TE_PI / 2 : 2 * atan2(q.z, q.w));
  if (k == 0) return atan2(2 * (q.y * q.z + q.w * q.x), (q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z));
  if (k == 1) return asin(s);
  return atan2(2 * (q.x * q.y + q.w * q.z), (q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z));
}

// WSEP: W goes to its own buffer (3 barriers per tile) or overwrites the published top rows in place (2 more
// barriers, 27 KB less shared memory for AR -> two CTAs per SM)
template <int TYPE, int STAGES, int MIN_CTAS, bool WSEP>
__global__ void __launch_bounds__(Split<TYPE>::RS * 32, MIN_CTAS) kf_step_split_kernel(const StepArgs a) {
  using MT = Model<TYPE>;
  using LY = Layout<TYPE>;
  constexpr int N = MT::N, M = MT::M, RS = Split<TYPE>::RS;
  constexpr int RPT = N / RS;        // rows per thread: AR 3, AV 2
  constexpr int CW = N / RS;         // W columns per warp
  constexpr int STAGE_DOUBLES = stage_doubles<TYPE>();
  static_assert(N % RS == 0 && M == RS, "one measured row per thread");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  double* stage0 = reinterpret_cast<double*>(smem_raw + 1024);
  double* Wsep = stage0 + (size_t)STAGES * STAGE_DOUBLES;   // [M][N][32] (WSEP only)
  double* ybuf = Wsep + (WSEP ? (size_t)M * N * TILE : 0);  // [6][32]
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = threadIdx.x == 0;

  if (producer) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int n_work = a.d_nwork ? *a.d_nwork : a.n_tiles;
  const int n_my = (n_work > (int)blockIdx.x) ? (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int it) -> int {
    const int w = blockIdx.x + it * gridDim.x;
    return a.tile_list ? a.tile_list[w] : a.tile_begin + w;
  };
  auto use_meas_tma = [&](int tile) -> bool { return a.meas_tma && (tile * TILE + TILE <= a.n_slots); };
  auto issue = [&](int it) {
    const int tile = tile_of(it);
    const int s = it % STAGES;
    double* st = stage0 + (size_t)s * STAGE_DOUBLES;
    const bool mt = use_meas_tma(tile);
    const uint32_t mbytes = mt ? (uint32_t)a.meas_stride * TILE * 8u : 0u;
    mbar_expect_tx(&bars[s], (uint32_t)LY::TILE_BYTES + mbytes);
    bulk_g2s(st, a.tiles + (size_t)tile * LY::TILE_DOUBLES, LY::TILE_BYTES, &bars[s]);
    if (mt) bulk_g2s(st + LY::TILE_DOUBLES, a.meas + (size_t)tile * TILE * a.meas_stride, mbytes, &bars[s]);
  };
  if (producer) {
    for (int pre = 0; pre < STAGES - 1 && pre < n_my; ++pre) issue(pre);
  }

  // per-lane control words, fetched one tile ahead so that their global-load latency never sits in front of a tile
  int act_n = ACT_NONE, cls_n = 0;
  double dt_n = a.dt;
  auto load_ctrl = [&](int it) {
    act_n = ACT_NONE; cls_n = 0; dt_n = a.dt;
    if (it < n_my) {
      const int slot = tile_of(it) * TILE + lane;
      if (slot < a.n_slots) {
        act_n = a.action ? (int)a.action[slot] : a.default_action;
        if (a.dt_slot) dt_n = a.dt_slot[slot];
        cls_n = (int)a.cls[slot];
      }
    }
  };
  load_ctrl(0);

  for (int it = 0; it < n_my; ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    const int tile = tile_of(it);
    const int slot = tile * TILE + lane;
    const bool valid = slot < a.n_slots;
    const int act = act_n, cls = cls_n;
    const double dt = dt_n;
    load_ctrl(it + 1);
    const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
    const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;
    const bool mt = use_meas_tma(tile);

    if (producer && it + STAGES - 1 < n_my) {
      bulk_wait_read<0>();   // the previous iteration's store has drained its stage
      issue(it + STAGES - 1);
    }
    double* st = stage0 + (size_t)s * STAGE_DOUBLES;
    double* Wbuf = WSEP ? Wsep : st + LY::F_P * TILE;   // in place: W(k,c) takes the slot of P'(k,c), k < M
    mbar_wait(&bars[s], parity);

    // lane = target in every warp, so a warp ballot already is the tile-wide answer
    const unsigned any = __ballot_sync(0xffffffffu, act != ACT_NONE);
    if (any) {
      const bool active = act != ACT_NONE;
      const bool upd = act == ACT_UPDATE;
      double Pr[RPT][N];
      double xr[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int g = q * RS + r;
        xr[q] = st[(LY::F_X + g) * TILE + lane];
#pragma unroll
        for (int c = 0; c < N; ++c) Pr[q][c] = st[(LY::F_P + g * N + c) * TILE + lane];
      }

      // ---- phase A: measurement conversion (one Euler angle per warp 0..2) + predict of the own rows ----
      if (r < 3 && upd) {
        const double* mp = mt ? (st + LY::TILE_DOUBLES + lane * a.meas_stride) : (a.meas + (size_t)slot * a.meas_stride);
        Quat qm{mp[3], mp[4], mp[5], mp[6]};
        quat_normalize(qm);
        const double ang = quat_to_rpy_comp(qm, r);
        const double un = unwrap1(st[(LY::F_PREV + r) * TILE + lane], ang);
        st[(LY::F_PREV + r) * TILE + lane] = un;   // meas_rpy_internal_ = unwrapped (angular_rates.cpp:85-88)
        ybuf[(3 + r) * TILE + lane] = un;
        ybuf[r * TILE + lane] = mp[r];
      }

      if (active) {
        if (TYPE == ANGULAR_RATES) {
          constexpr int B = MT::B;
          const double h = 0.5 * dt * dt;
          xr[0] = xr[0] + dt * xr[1] + h * xr[2];
          xr[1] = xr[1] + dt * xr[2];
#pragma unroll
          for (int j = 0; j < N; ++j) {   // A P on the three rows of this thread (one per kinematic block)
            Pr[0][j] = Pr[0][j] + dt * Pr[1][j] + h * Pr[2][j];
            Pr[1][j] = Pr[1][j] + dt * Pr[2][j];
          }
#pragma unroll
          for (int q = 0; q < RPT; ++q) {   // (A P) A^T + Q within each row
#pragma unroll
            for (int c = 0; c < B; ++c) {
              Pr[q][c] = Pr[q][c] + dt * Pr[q][B + c] + h * Pr[q][2 * B + c];
              Pr[q][B + c] = Pr[q][B + c] + dt * Pr[q][2 * B + c];
            }
#pragma unroll
            for (int j = 0; j < N; ++j) Pr[q][j] = Pr[q][j] + __ldg(&Q[(q * RS + r) * N + j]);
          }
        } else {
          // EKF (angular_velocities.cpp:116-140): A linearised at the previous posterior, read from the stage.
          // Thread r < 3 owns rows (p_r, v_r); thread r >= 3 owns rows (rpy_i, w_i), i = r - 3.
          double s_r, c_r, s_p, c_p;
          sincos(st[(LY::F_X + 3) * TILE + lane], &s_r, &c_r);
          sincos(st[(LY::F_X + 4) * TILE + lane], &s_p, &c_p);
          const double wx = st[(LY::F_X + 9) * TILE + lane], wy = st[(LY::F_X + 10) * TILE + lane], wz = st[(LY::F_X + 11) * TILE + lane];
          double J1[3][3], J2[3][3];
          J1[0][0] = (dt * (wy * c_r * s_p - wz * s_p * s_r)) / c_p + 1;
          J1[0][1] = (dt * (wz * c_r + wy * s_r)) / (c_p * c_p);
          J1[0][2] = 0;
          J1[1][0] = -dt * (wz * c_r + wy * s_r);
          J1[1][1] = 1;
          J1[1][2] = 0;
          J1[2][0] = (dt * (wy * c_r - wz * s_r)) / c_p;
          J1[2][1] = (dt * s_p * (wz * c_r + wy * s_r)) / (c_p * c_p);
          J1[2][2] = 1;
          J2[0][0] = dt; J2[0][1] = (dt * s_p * s_r) / c_p; J2[0][2] = (dt * c_r * s_p) / c_p;
          J2[1][0] = 0;  J2[1][1] = dt * c_r;               J2[1][2] = -dt * s_r;
          J2[2][0] = 0;  J2[2][1] = (dt * s_r) / c_p;       J2[2][2] = (dt * c_r) / c_p;
          if (r < 3) {
            xr[0] = xr[0] + dt * xr[1];
#pragma unroll
            for (int j = 0; j < N; ++j) Pr[0][j] = Pr[0][j] + dt * Pr[1][j];
          } else {
            const int i = r - 3;
            double E[3][3];
            E[0][0] = 1; E[0][1] = (s_p * s_r) / c_p; E[0][2] = (c_r * s_p) / c_p;
            E[1][0] = 0; E[1][1] = c_r;               E[1][2] = -s_r;
            E[2][0] = 0; E[2][1] = s_r / c_p;         E[2][2] = c_r / c_p;
            double j1r[3], j2r[3], er[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double c0[3] = {J1[0][j], J1[1][j], J1[2][j]}, c1[3] = {J2[0][j], J2[1][j], J2[2][j]}, c2[3] = {E[0][j], E[1][j], E[2][j]};
              j1r[j] = sel3(c0, i); j2r[j] = sel3(c1, i); er[j] = sel3(c2, i);
            }
            xr[0] = xr[0] + ((dt * er[0]) * wx + (dt * er[1]) * wy + (dt * er[2]) * wz);
            // row 3+i of A P = J1[i,:] rows 3..5 + J2[i,:] rows 9..11 (original rows, from the stage)
#pragma unroll
            for (int j = 0; j < N; ++j) {
              double p1[3], p3[3];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                p1[k] = st[(LY::F_P + (3 + k) * N + j) * TILE + lane];
                p3[k] = st[(LY::F_P + (9 + k) * N + j) * TILE + lane];
              }
              Pr[0][j] = j1r[0] * p1[0] + j1r[1] * p1[1] + j1r[2] * p1[2] + j2r[0] * p3[0] + j2r[1] * p3[1] + j2r[2] * p3[2];
            }
          }
          // (A P) A^T + Q within each row
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            double rw[N];
#pragma unroll
            for (int j = 0; j < N; ++j) rw[j] = Pr[q][j];
            const int g = q * RS + r;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              Pr[q][j] = (rw[j] + rw[6 + j] * dt) + __ldg(&Q[g * N + j]);
              const double sacc = rw[3] * J1[j][0] + rw[4] * J1[j][1] + rw[5] * J1[j][2] + rw[9] * J2[j][0] + rw[10] * J2[j][1] + rw[11] * J2[j][2];
              Pr[q][3 + j] = sacc + __ldg(&Q[g * N + 3 + j]);
              Pr[q][6 + j] = rw[6 + j] + __ldg(&Q[g * N + 6 + j]);
              Pr[q][9 + j] = rw[9 + j] + __ldg(&Q[g * N + 9 + j]);
            }
          }
        }
      }
      if (TYPE == ANGULAR_VELOCITIES) __syncthreads();   // warps 3..5 have read the original rows 3..5 / 9..11

      // publish the predicted measured row r and x'[r] in place
      if (upd) {
        st[(LY::F_X + r) * TILE + lane] = xr[0];
#pragma unroll
        for (int c = 0; c < N; ++c) st[(LY::F_P + r * N + c) * TILE + lane] = Pr[0][c];
      }
      __syncthreads();

      // ---- phase B: S = P'[0:M,0:M] + R, v, this warp's columns of W -> Wbuf ---------------------------
      double Wc[M][CW];
      if (upd) {
        Chol<M> ch;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (j <= i) ch.at(i, j) = st[(LY::F_P + i * N + j) * TILE + lane] + __ldg(&R[i * M + j]);
        ch.factor();
        double v[M];
#pragma unroll
        for (int k = 0; k < M; ++k) v[k] = ybuf[k * TILE + lane] - st[(LY::F_X + k) * TILE + lane];
        ch.solve(v);
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          const int c = r * CW + cc;
          double col[M];
#pragma unroll
          for (int k = 0; k < M; ++k) col[k] = st[(LY::F_P + k * N + c) * TILE + lane];
          ch.solve(col);
#pragma unroll
          for (int k = 0; k < M; ++k) Wc[k][cc] = col[k];
        }
        // x += P'[rows,0:M] v (K (y - C x'), src/kalman.cpp:93) here, so that v is dead before phase C
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) sacc += Pr[q][k] * v[k];
          xr[q] += sacc;
        }
      }
      if (!WSEP) __syncthreads();   // in place: every warp has read S and its columns of P'[0:M,:]
      if (upd) {
#pragma unroll
        for (int cc = 0; cc < CW; ++cc)
#pragma unroll
          for (int k = 0; k < M; ++k) Wbuf[(k * N + r * CW + cc) * TILE + lane] = Wc[k][cc];
      }
      __syncthreads();

      // ---- phase C: own rows: P[rows,:] -= P'[rows,0:M] W ((I - K C) P, src/kalman.cpp:94) ----------------
      if (upd) {
        double Pk[RPT][M];
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int k = 0; k < M; ++k) Pk[q][k] = Pr[q][k];
#pragma unroll
        for (int c = N - 1; c >= 0; --c) {
          double w[M];
#pragma unroll
          for (int k = 0; k < M; ++k) w[k] = Wbuf[(k * N + c) * TILE + lane];
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            double sacc = Pr[q][c];
#pragma unroll
            for (int k = 0; k < M; ++k) sacc -= Pk[q][k] * w[k];
            Pr[q][c] = sacc;
          }
        }
      }
      if (!WSEP) __syncthreads();   // in place: every warp has read W before rows 0..M-1 are rewritten

      // ---- phase D: own rows back into the stage (nobody reads the published top rows after phase B) ------
      if (active) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int g = q * RS + r;
          st[(LY::F_X + g) * TILE + lane] = xr[q];
#pragma unroll
          for (int c = 0; c < N; ++c) st[(LY::F_P + g * N + c) * TILE + lane] = Pr[q][c];
        }
        if (r == 0) {
          st[LY::F_T * TILE + lane] = st[LY::F_T * TILE + lane] + dt;   // updateTime
          if (upd) {
            long long* nm = reinterpret_cast<long long*>(st + LY::F_NMEAS * TILE + lane);
            *nm = *nm + 1;                                             // updateMeasurement
          }
          if (a.clear_action) a.action[slot] = 0;
        }
      }
      if (a.pos_out && valid && r < 3) a.pos_out[(size_t)slot * 3 + r] = xr[0];
      fence_proxy_async();
      __syncthreads();   // also orders this tile's Wbuf / ybuf reads before the next tile's writes
      if (producer) {
        bulk_s2g(a.tiles + (size_t)tile * LY::TILE_DOUBLES, st, LY::TILE_BYTES);
        bulk_commit();
        if (a.clear_action) a.tile_flag[tile] = 0;
      }
    } else {
      if (a.pos_out && valid && r < 3) a.pos_out[(size_t)slot * 3 + r] = st[(LY::F_X + r) * TILE + lane];
      __syncthreads();   // the stage may be refilled by the producer in the next iteration
    }
  }
  if (producer) bulk_wait<0>();
}

}  // namespace te
