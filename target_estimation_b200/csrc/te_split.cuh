// te_split.cuh -- row-split KF step kernel for the large-state models (AR n=18, AV n=12).
//
// One CTA of RS warps works on one tile of 32 targets: lane = target, warp r owns the rows
// {q*RS + r : q < N/RS} of every target's covariance (AR: RS=6 -> rows r, 6+r, 12+r = one row of each
// kinematic block; AV: RS=3 -> rows r, 3+r, 6+r, 9+r), held in registers.  The tile itself is staged in
// shared memory by one TMA bulk copy and written back by one bulk store, as in kf_step_kernel; what the
// row owners must exchange -- the top M rows of the predicted covariance and W = S^-1 P'[0:M,:] -- goes
// through the staged tile in place, with CTA barriers between the phases:
//
//   A  predict own rows (x' = A x | f(x), P' = A P A^T + Q); publish rows 0..M-1 of P', x'[0:M], y      | barrier
//   B  every warp factors S = P'[0:M,0:M] + R (in-register Cholesky), solves v = S^-1 (y - x'[0:M]) and
//      its N/RS columns of W                                                                             | barrier
//      W overwrites P'[0:M,:] in the stage                                                               | barrier
//   C  own rows: x += P'[rows,0:M] v ; P[rows,:] = P'[rows,:] - P'[rows,0:M] W                            | barrier
//   D  own rows back into the stage, t / n_meas bookkeeping, bulk store
//
// Same arithmetic as te_device.cuh's step_lane (predict_kinematic / predict_av / kf_update); only the
// ownership of the rows differs.  Reference: src/kalman.cpp:84-95,129-140, src/types/angular_rates.cpp:72-115,
// src/types/angular_velocities.cpp:80-151.
#pragma once
#include "te_kernels.cuh"

namespace te {

template <int TYPE> struct Split;
template <> struct Split<ANGULAR_RATES> { static constexpr int RS = 6; };
template <> struct Split<ANGULAR_VELOCITIES> { static constexpr int RS = 3; };

template <int TYPE> __host__ __device__ constexpr size_t split_smem_bytes(int stages) {
  return 1024 + (size_t)stages * stage_doubles<TYPE>() * 8;
}

__device__ __forceinline__ double sel3(const double a[3], int r) { return r == 0 ? a[0] : (r == 1 ? a[1] : a[2]); }

template <int TYPE, int STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(Split<TYPE>::RS * 32, MIN_CTAS) kf_step_split_kernel(const StepArgs a) {
  using MT = Model<TYPE>;
  using LY = Layout<TYPE>;
  constexpr int N = MT::N, M = MT::M, RS = Split<TYPE>::RS;
  constexpr int RPT = N / RS;        // rows per thread
  constexpr int MQ = M / RS;         // of which measured (top) rows
  constexpr int CW = N / RS;         // W columns per warp
  constexpr int STAGE_DOUBLES = stage_doubles<TYPE>();
  static_assert(N % RS == 0 && M % RS == 0, "row split must divide n and m");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  double* stage0 = reinterpret_cast<double*>(smem_raw + 1024);
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

  for (int it = 0; it < n_my; ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    const int tile = tile_of(it);
    const int slot = tile * TILE + lane;
    const bool valid = slot < a.n_slots;
    int act = ACT_NONE;
    double dt = a.dt;
    int cls = 0;
    if (valid) {
      act = a.action ? (int)a.action[slot] : a.default_action;
      if (a.dt_slot) dt = a.dt_slot[slot];
      cls = (int)a.cls[slot];
    }
    const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
    const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;
    const bool mt = use_meas_tma(tile);

    if (producer && it + STAGES - 1 < n_my) {
      bulk_wait_read<0>();   // the previous iteration's store has drained its stage
      issue(it + STAGES - 1);
    }
    double* st = stage0 + (size_t)s * STAGE_DOUBLES;
    double* ymeas = st + LY::TILE_DOUBLES;   // [7][32] after warp 0 rewrites it: y[0..5]
    mbar_wait(&bars[s], parity);

    const int any = __syncthreads_or(act != ACT_NONE);
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

      // ---- phase A: measurement conversion (warp 0) + predict of the own rows -----------------------
      if (r == 0) {   // warp-uniform: every lane of warp 0 takes part in the __syncwarp()s
        double m7[7];
        if (upd) {
          const double* mp = mt ? (ymeas + lane * a.meas_stride) : (a.meas + (size_t)slot * a.meas_stride);
#pragma unroll
          for (int k = 0; k < 7; ++k) m7[k] = mp[k];
          double prev[3], un[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) prev[k] = st[(LY::F_PREV + k) * TILE + lane];
          meas_to_unwrapped_rpy(m7 + 3, prev, un);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            st[(LY::F_PREV + k) * TILE + lane] = un[k];   // meas_rpy_internal_ = unwrapped
            m7[3 + k] = un[k];
          }
        }
        // y goes into the stage's measurement block as [k][lane]: AoS -> SoA in place, after every lane's read
        __syncwarp();
        if (upd) {
#pragma unroll
          for (int k = 0; k < 6; ++k) ymeas[k * TILE + lane] = m7[k];
        }
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
          // EKF (angular_velocities.cpp:116-140): A linearised at the previous posterior, read from the stage
          double s_r, c_r, s_p, c_p;
          sincos(st[(LY::F_X + 3) * TILE + lane], &s_r, &c_r);
          sincos(st[(LY::F_X + 4) * TILE + lane], &s_p, &c_p);
          const double wx = st[(LY::F_X + 9) * TILE + lane], wy = st[(LY::F_X + 10) * TILE + lane], wz = st[(LY::F_X + 11) * TILE + lane];
          double J1[3][3], J2[3][3], E[3][3];
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
          E[0][0] = 1; E[0][1] = (s_p * s_r) / c_p; E[0][2] = (c_r * s_p) / c_p;
          E[1][0] = 0; E[1][1] = c_r;               E[1][2] = -s_r;
          E[2][0] = 0; E[2][1] = s_r / c_p;         E[2][2] = c_r / c_p;
          double j1r[3], j2r[3], er[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double c0[3] = {J1[0][j], J1[1][j], J1[2][j]}, c1[3] = {J2[0][j], J2[1][j], J2[2][j]}, c2[3] = {E[0][j], E[1][j], E[2][j]};
            j1r[j] = sel3(c0, r); j2r[j] = sel3(c1, r); er[j] = sel3(c2, r);
          }
          xr[0] = xr[0] + dt * xr[2];
          xr[1] = xr[1] + ((dt * er[0]) * wx + (dt * er[1]) * wy + (dt * er[2]) * wz);
          // A P : row r += dt * row 6+r ; row 3+r = J1[r,:] rows 3..5 + J2[r,:] rows 9..11 (original rows, from the stage)
#pragma unroll
          for (int j = 0; j < N; ++j) {
            Pr[0][j] = Pr[0][j] + dt * Pr[2][j];
            double p1[3], p3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              p1[i] = st[(LY::F_P + (3 + i) * N + j) * TILE + lane];
              p3[i] = st[(LY::F_P + (9 + i) * N + j) * TILE + lane];
            }
            Pr[1][j] = j1r[0] * p1[0] + j1r[1] * p1[1] + j1r[2] * p1[2] + j2r[0] * p3[0] + j2r[1] * p3[1] + j2r[2] * p3[2];
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
      if (TYPE == ANGULAR_VELOCITIES) __syncthreads();   // every warp has read the original rows 3..5 / 9..11

      // publish the predicted top rows and x'[0:M]
      if (upd) {
#pragma unroll
        for (int q = 0; q < MQ; ++q) {
          const int g = q * RS + r;
          st[(LY::F_X + g) * TILE + lane] = xr[q];
#pragma unroll
          for (int c = 0; c < N; ++c) st[(LY::F_P + g * N + c) * TILE + lane] = Pr[q][c];
        }
      }
      __syncthreads();

      // ---- phase B: S = P'[0:M,0:M] + R, v, this warp's columns of W ---------------------------------
      double v[M];
      double Wc[M][CW];
      if (upd) {
        Chol<M> ch;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) ch.at(i, j) = st[(LY::F_P + i * N + j) * TILE + lane] + __ldg(&R[i * M + j]);
        ch.factor();
#pragma unroll
        for (int k = 0; k < M; ++k) v[k] = ymeas[k * TILE + lane] - st[(LY::F_X + k) * TILE + lane];
        ch.solve(v);
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
          double col[M];
#pragma unroll
          for (int k = 0; k < M; ++k) col[k] = st[(LY::F_P + k * N + (r * CW + cc)) * TILE + lane];
          ch.solve(col);
#pragma unroll
          for (int k = 0; k < M; ++k) Wc[k][cc] = col[k];
        }
      }
      __syncthreads();   // all reads of P'[0:M,:] done
      if (upd) {
#pragma unroll
        for (int cc = 0; cc < CW; ++cc)
#pragma unroll
          for (int k = 0; k < M; ++k) st[(LY::F_P + k * N + (r * CW + cc)) * TILE + lane] = Wc[k][cc];
      }
      __syncthreads();

      // ---- phase C: own rows ---------------------------------------------------------------------------
      if (upd) {
        double Pk[RPT][M];
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) {
            Pk[q][k] = Pr[q][k];
            sacc += Pr[q][k] * v[k];
          }
          xr[q] += sacc;
        }
#pragma unroll
        for (int c = N - 1; c >= 0; --c) {
          double w[M];
#pragma unroll
          for (int k = 0; k < M; ++k) w[k] = st[(LY::F_P + k * N + c) * TILE + lane];
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            double sacc = Pr[q][c];
#pragma unroll
            for (int k = 0; k < M; ++k) sacc -= Pk[q][k] * w[k];
            Pr[q][c] = sacc;
          }
        }
      }
      __syncthreads();   // all reads of W done

      // ---- phase D: write back ---------------------------------------------------------------------------
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
      __syncthreads();
      if (producer) {
        bulk_s2g(a.tiles + (size_t)tile * LY::TILE_DOUBLES, st, LY::TILE_BYTES);
        bulk_commit();
        if (a.clear_action) a.tile_flag[tile] = 0;
      }
    } else {
      if (a.pos_out && valid && r < 3) a.pos_out[(size_t)slot * 3 + r] = st[(LY::F_X + r) * TILE + lane];
      __syncthreads();
    }
  }
  if (producer) bulk_wait<0>();
}

}  // namespace te
