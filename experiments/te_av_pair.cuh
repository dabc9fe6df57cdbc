// experiments/te_av_pair.cuh -- NOT part of the product build.  Round-2 experiment, kept for the record:
// angular-velocities EKF (n = 12, m = 6), TWO LANES PER TARGET, symmetric (packed) covariance.  It passed every parity test on the
// first run (18 tests incl. grid-capped multi-tile runs) and ran at 0.84 ms per tick of 1 Mi targets against 0.445 ms for the
// one-lane kernel of te_direct.cuh / te_av_sym.cuh: 128 registers per lane -> 1.3 KB of spill traffic per lane (ptxas: 632-byte
// stack frame, 1330 B spill stores; 168 registers: 636 B), 1170 DFMA per lane against 1401 for the whole target in one lane (what
// both lanes must do alike -- measurement conversion, sincos, Jacobians, Cholesky -- does not shrink), and the register file holds
// the same 256 targets per SM either way.  To build it again: include it from csrc/te_step.cu and launch
// kf_step_av_pair_kernel<8, 2> with av_pair_smem_bytes(8) of dynamic shared memory and grid = min(2 * SMs, ceil(2 * tiles / 8)).
//
// Idea.  A target is shared by the lane pair (2q, 2q + 1) of a warp, a warp takes half a tile (16 targets) per trip, every lane
// carries 39 covariance entries and 6 state entries.  Both lanes run the SAME instruction stream (divergent code would serialise);
// what differs is data, chosen by h = lane & 1.  The state splits into the translational group (p, v) -- lane 0 -- and the
// rotational group (r, w) -- lane 1; "top" = p | r, "bottom" = v | w.  In that grouping the EKF transition is block diagonal,
// A = G_0 (+) G_1 with G_h = [A1 A2; 0 I], (A1, A2) = (I, dt I) for h = 0, (J1, J2) for h = 1 (angular_velocities.cpp:116-140),
// and I, dt I have the sparsity of J1 = [a b 0; c 1 0; d e 1], J2 = [dt f g; 0 m n; 0 q r], so ONE piece of code predicts both
// groups (lane 0 multiplies by exact 0 / 1: bit-identical to the specialised form).
//   own group (21 entries: TT, TB, BB)      P'_gg = G_h P_gg G_h^T + Q                         local to the lane
//   cross block C = P[(p,v), (r,w)] (36)    C' = G_0 C G_1^T + Q: rows p on lane 0 (they need rows v: one shuffle per entry),
//                                           rows v on lane 1; both lanes right-multiply by the true J1, J2
//   then pw <-> vr change lanes (9 shuffles): lane 0 holds the cross COLUMNS r (pr, vr), lane 1 the columns w (pw, vw), so that
//   update (src/kalman.cpp:135-140, C = [I6 0]):  S = P'[0:6,0:6] + R is assembled on both lanes (15 shuffles) and factored
//   redundantly; lane h computes the Z = L^-1 P'[0:6,:] columns of ITS state entries from entries it holds (+ pr), adds Z^T u to
//   its state entries, parks its Z columns in shared memory (36 doubles per lane: 9 KB per warp); the rank-1 downdates
//   P -= z_k z_k^T read the partner's columns from there (conflict-free: lane-permuted addresses).
#pragma once
#include "te_kernels.cuh"

namespace te {

constexpr int AVP_Z_FIELDS = 36;   // Z[k][c], k = 0..5 (row), c = 0..5 (own column: top 0..2, bottom 0..2) at zs[(k * 6 + c) * 32 + lane]
__host__ __device__ constexpr size_t av_pair_smem_bytes(int warps) { return (size_t)warps * AVP_Z_FIELDS * TILE * 8; }

struct Sym3 {   // upper triangle of a symmetric 3 x 3 block
  double v[6];
  __device__ __forceinline__ double& operator()(int i, int j) { return i <= j ? v[i * 3 - (i * (i - 1)) / 2 + (j - i)] : v[j * 3 - (j * (j - 1)) / 2 + (i - j)]; }
};

template <int WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) kf_step_av_pair_kernel(const StepArgs a) {
  using LY = Layout<ANGULAR_VELOCITIES>;
  constexpr int N = 12, M = 6;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = lane & 1, q = lane >> 1;
  const bool odd = h != 0;
  double* zs = reinterpret_cast<double*>(smem_raw) + (size_t)warp * AVP_Z_FIELDS * TILE;
  const int n_work = a.d_nwork ? *a.d_nwork : a.n_tiles;
  const int gw = blockIdx.x * WARPS + warp, GW = gridDim.x * WARPS;
  for (int w = gw; w < 2 * n_work; w += GW) {
    const int tile = a.tile_list ? a.tile_list[w >> 1] : a.tile_begin + (w >> 1);
    const int col = (w & 1) * 16 + q;   // the pair's target: column `col` of the tile
    const int slot = tile * TILE + col;
    const bool valid = slot < a.n_slots;
    int act = ACT_NONE, cls = 0, dst = -1;
    double dt = a.dt;
    if (valid) {
      act = a.action ? (int)a.action[slot] : a.default_action;
      if (a.dt_slot) dt = a.dt_slot[slot];
      cls = (int)a.cls[slot];
      if (a.dst_tiles) {
        if (a.dst_alive[slot]) dst = a.dst_pos[slot];
        else act = ACT_NONE;
      }
    }
    const unsigned m_act = __ballot_sync(FULL, act != ACT_NONE);
    const unsigned m_upd = __ballot_sync(FULL, act == ACT_UPDATE);
    const double* in = a.tiles + (size_t)tile * LY::TILE_DOUBLES + col;
    double* out = a.dst_tiles ? (dst >= 0 ? a.dst_tiles + (size_t)(dst / TILE) * LY::TILE_DOUBLES + (dst % TILE) : nullptr)
                              : a.tiles + (size_t)tile * LY::TILE_DOUBLES + col;
    if (act != ACT_NONE) {
      const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
      const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;
      // own group: state index of top_i = 3h + i, of bottom_i = 6 + 3h + i; entry (sa, sb) of the group sits 39 h fields after
      // the same entry of the (p, v) group
      const double* ing = in + (size_t)(39 * h) * TILE;
      double top[3], bot[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        top[i] = in[(LY::F_X + 3 * h + i) * TILE];
        bot[i] = in[(LY::F_X + 6 + 3 * h + i) * TILE];
      }
      Sym3 TT, BB;
      double TB[3][3], CA[3][3], CB[3][3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (i <= j) {
            TT(i, j) = ing[(LY::F_P + i * N + j) * TILE];
            BB(i, j) = ing[(LY::F_P + (6 + i) * N + 6 + j) * TILE];
          }
          TB[i][j] = ing[(LY::F_P + i * N + 6 + j) * TILE];
          // cross block, by rows: lane 0 rows p (CA = pr, CB = pw), lane 1 rows v (CA = vr -- stored as (r_j, v_i) --, CB = vw)
          CA[i][j] = in[(LY::F_P + (odd ? (3 + j) * N + 6 + i : i * N + 3 + j)) * TILE];
          CB[i][j] = in[(LY::F_P + (odd ? (6 + i) * N + 9 + j : i * N + 9 + j)) * TILE];
        }
      double t_in = 0.0;
      long long nm_in = 0;
      if (!odd) {
        t_in = in[LY::F_T * TILE];
        nm_in = reinterpret_cast<const long long*>(in)[LY::F_NMEAS * TILE];
      }
      double prev[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) prev[k] = in[(LY::F_PREV + k) * TILE];
      double meas[7];
      if (act == ACT_UPDATE) {
        const double* mp = a.meas + (size_t)slot * a.meas_stride;
#pragma unroll
        for (int k = 0; k < 7; ++k) meas[k] = __ldg(mp + k);
      }
      // measurement conversion (angular_velocities.cpp:87-96), on both lanes
      double un[3] = {0.0, 0.0, 0.0};
      if (act == ACT_UPDATE) {
        meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
        for (int k = 0; k < 3; ++k) prev[k] = un[k];
      }
      if (odd) {
#pragma unroll
        for (int k = 0; k < 3; ++k) out[(LY::F_PREV + k) * TILE] = prev[k];
      }
      // the rotational state on both lanes, the Jacobians at the previous posterior (geometry.hpp:394-426)
      double rs0, rs1, w0, w1, w2;
      {
        const double o0 = __shfl_xor_sync(m_act, top[0], 1), o1 = __shfl_xor_sync(m_act, top[1], 1);
        const double b0 = __shfl_xor_sync(m_act, bot[0], 1), b1 = __shfl_xor_sync(m_act, bot[1], 1), b2 = __shfl_xor_sync(m_act, bot[2], 1);
        rs0 = odd ? top[0] : o0; rs1 = odd ? top[1] : o1;
        w0 = odd ? bot[0] : b0; w1 = odd ? bot[1] : b1; w2 = odd ? bot[2] : b2;
      }
      double j10, j11, j12, j13, j14, j20, j21, j22, j23, j24, j25;
      double B[3][3];   // top' = top + B bottom: dt I (lane 0), dt EarBaseInv(rpy) (lane 1)
      {
        double s_r, c_r, s_p, c_p;
        sincos(rs0, &s_r, &c_r);
        sincos(rs1, &s_p, &c_p);
        j10 = (dt * (w1 * c_r * s_p - w2 * s_p * s_r)) / c_p + 1;
        j11 = (dt * (w2 * c_r + w1 * s_r)) / (c_p * c_p);
        j12 = -dt * (w2 * c_r + w1 * s_r);
        j13 = (dt * (w1 * c_r - w2 * s_r)) / c_p;
        j14 = (dt * s_p * (w2 * c_r + w1 * s_r)) / (c_p * c_p);
        j20 = (dt * s_p * s_r) / c_p;
        j21 = (dt * c_r * s_p) / c_p;
        j22 = dt * c_r;
        j23 = -dt * s_r;
        j24 = (dt * s_r) / c_p;
        j25 = (dt * c_r) / c_p;
        const double E01 = (s_p * s_r) / c_p, E02 = (c_r * s_p) / c_p, E11 = c_r, E12 = -s_r, E21 = s_r / c_p, E22 = c_r / c_p;
        B[0][0] = dt * 1.0;             B[0][1] = odd ? dt * E01 : 0.0; B[0][2] = odd ? dt * E02 : 0.0;
        B[1][0] = odd ? dt * 0.0 : 0.0; B[1][1] = odd ? dt * E11 : dt;  B[1][2] = odd ? dt * E12 : 0.0;
        B[2][0] = odd ? dt * 0.0 : 0.0; B[2][1] = odd ? dt * E21 : 0.0; B[2][2] = odd ? dt * E22 : dt;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) top[i] = top[i] + (B[i][0] * bot[0] + B[i][1] * bot[1] + B[i][2] * bot[2]);
      const double g10 = odd ? j10 : 1.0, g11 = odd ? j11 : 0.0, g12 = odd ? j12 : 0.0, g13 = odd ? j13 : 0.0, g14 = odd ? j14 : 0.0;
      const double g20 = odd ? j20 : 0.0, g21 = odd ? j21 : 0.0, g22 = odd ? j22 : dt, g23 = odd ? j23 : 0.0, g24 = odd ? j24 : 0.0,
                   g25 = odd ? j25 : dt;
      auto J1 = [&](int i, int k) -> double {
        return i == 0 ? (k == 0 ? j10 : (k == 1 ? j11 : 0.0)) : (i == 1 ? (k == 0 ? j12 : (k == 1 ? 1.0 : 0.0)) : (k == 0 ? j13 : (k == 1 ? j14 : 1.0)));
      };
      auto J2 = [&](int i, int k) -> double {
        return i == 0 ? (k == 0 ? dt : (k == 1 ? j20 : j21)) : (i == 1 ? (k == 0 ? 0.0 : (k == 1 ? j22 : j23)) : (k == 0 ? 0.0 : (k == 1 ? j24 : j25)));
      };
      auto G1 = [&](int i, int k) -> double {
        return i == 0 ? (k == 0 ? g10 : (k == 1 ? g11 : 0.0)) : (i == 1 ? (k == 0 ? g12 : (k == 1 ? 1.0 : 0.0)) : (k == 0 ? g13 : (k == 1 ? g14 : 1.0)));
      };
      auto G2 = [&](int i, int k) -> double {
        return i == 0 ? (k == 0 ? dt : (k == 1 ? g20 : g21)) : (i == 1 ? (k == 0 ? 0.0 : (k == 1 ? g22 : g23)) : (k == 0 ? 0.0 : (k == 1 ? g24 : g25)));
      };
      // covariance predict, own group: P'_gg = G P_gg G^T + Q
      const double* __restrict__ Qg = Q + 39 * h;
      {
        double trr[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += G1(i, k) * TT(k, j);
#pragma unroll
            for (int k = 0; k < 3; ++k) s += G2(i, k) * TB[j][k];   // P(bottom_k, top_j)
            trr[i][j] = s;
          }
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // TB <- G1 TB + G2 BB, one column at a time
          double c[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += G1(i, k) * TB[k][j];
#pragma unroll
            for (int k = 0; k < 3; ++k) s += G2(i, k) * BB(k, j);
            c[i] = s;
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) TB[i][j] = c[i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (i <= j) {
              const double s = trr[i][0] * G1(j, 0) + trr[i][1] * G1(j, 1) + trr[i][2] * G1(j, 2) + TB[i][0] * G2(j, 0) + TB[i][1] * G2(j, 1) +
                               TB[i][2] * G2(j, 2);
              TT(i, j) = s + __ldg(&Qg[i * N + j]);
            }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            TB[i][j] = TB[i][j] + __ldg(&Qg[i * N + 6 + j]);
            if (i <= j) BB(i, j) = BB(i, j) + __ldg(&Qg[(6 + i) * N + 6 + j]);
          }
      }
      // cross block: X = G_0 C (rows p += dt rows v), C' = X G_1^T + Q (columns r mix through the true J1, J2)
      {
        const double dte = odd ? 0.0 : dt;
        const double* __restrict__ Qc = Q + 72 * h;   // rows v = rows p + 6 (Q is bitwise symmetric on this path)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double xa[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double oa = __shfl_xor_sync(m_act, CA[i][k], 1), ob = __shfl_xor_sync(m_act, CB[i][k], 1);
            xa[k] = CA[i][k] + dte * oa;
            CB[i][k] = CB[i][k] + dte * ob;
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double s = xa[0] * J1(j, 0) + xa[1] * J1(j, 1) + xa[2] * J1(j, 2) + CB[i][0] * J2(j, 0) + CB[i][1] * J2(j, 1) + CB[i][2] * J2(j, 2);
            CA[i][j] = s + __ldg(&Qc[i * N + 3 + j]);
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) CB[i][j] = CB[i][j] + __ldg(&Qc[i * N + 9 + j]);
        }
        // pw (lane 0) <-> vr (lane 1): from here lane 0 holds the cross columns r (CA = pr, CB = vr), lane 1 the columns w (CA = pw, CB = vw)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const double r = __shfl_xor_sync(m_act, odd ? CA[i][j] : CB[i][j], 1);
            CA[i][j] = odd ? r : CA[i][j];
            CB[i][j] = odd ? CB[i][j] : r;
          }
      }
      // update
      if (act == ACT_UPDATE) {
        double* zl = zs + lane;
        {
          Chol<M> ch;
          double PR[3][3];   // pr on both lanes
          {
            Sym3 TTo;
#pragma unroll
            for (int k = 0; k < 6; ++k) TTo.v[k] = __shfl_xor_sync(m_upd, TT.v[k], 1);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const double r = __shfl_xor_sync(m_upd, CA[i][j], 1);
                PR[i][j] = odd ? r : CA[i][j];
              }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
              for (int j = 0; j < M; ++j)
                if (j <= i) {
                  double s;
                  if (i < 3) s = odd ? TTo(j, i) : TT(j, i);                          // pp
                  else if (j < 3) s = PR[j][i - 3];                                   // P(r_{i-3}, p_j)
                  else s = odd ? TT(j - 3, i - 3) : TTo(j - 3, i - 3);                // rr
                  ch.at(i, j) = s + __ldg(&R[i * M + j]);
                }
          }
          ch.factor();
          double u[M];   // L^-1 (y - x'[0:6])
          {
            double xo[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) xo[k] = __shfl_xor_sync(m_upd, top[k], 1);
#pragma unroll
            for (int k = 0; k < M; ++k) {
              const double yk = k < 3 ? meas[k] : un[k - 3];
              const double xk = k < 3 ? (odd ? xo[k] : top[k]) : (odd ? top[k - 3] : xo[k - 3]);
              double s = yk - xk;
#pragma unroll
              for (int m = 0; m < M; ++m)
                if (m < k) s -= ch.L[k][m] * u[m];
              u[k] = s * ch.L[k][k];
            }
          }
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double z[M];
#pragma unroll
            for (int k = 0; k < M; ++k) {
              double s;
              if (c < 3) s = k < 3 ? (odd ? PR[k][c] : TT(k, c)) : (odd ? TT(k - 3, c) : PR[c][k - 3]);
              else s = k < 3 ? (odd ? CA[k][c - 3] : TB[k][c - 3]) : (odd ? TB[k - 3][c - 3] : CB[c - 3][k - 3]);
#pragma unroll
              for (int m = 0; m < M; ++m)
                if (m < k) s -= ch.L[k][m] * z[m];
              z[k] = s * ch.L[k][k];
            }
            double xs = c < 3 ? top[c] : bot[c - 3];
#pragma unroll
            for (int k = 0; k < M; ++k) {
              zl[(k * 6 + c) * TILE] = z[k];
              xs += z[k] * u[k];
            }
            if (c < 3) top[c] = xs;
            else bot[c - 3] = xs;
            if (c % 2 == 1) asm volatile("" ::: "memory");
          }
        }
        __syncwarp(m_upd);
        // P -= Z^T Z as six rank-1 downdates; the partner's columns come from its part of the scratch
        const double* pa1 = zs + (odd ? (lane ^ 1) : lane);              // CA row factor: lane 0 own top (p), lane 1 partner top (p)
        const double* pb1 = zs + (odd ? 3 * TILE + lane : (lane ^ 1));   // CA / CB column factor: lane 0 partner top (r), lane 1 own bottom (w)
        const double* pa2 = zs + 3 * TILE + (odd ? (lane ^ 1) : lane);   // CB row factor: lane 0 own bottom (v), lane 1 partner bottom (v)
#pragma unroll
        for (int k = 0; k < M; ++k) {
          double zo[6], a1[3], b1[3], a2[3];
#pragma unroll
          for (int c = 0; c < 6; ++c) zo[c] = zl[(k * 6 + c) * TILE];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            a1[i] = pa1[(k * 6 + i) * TILE];
            b1[i] = pb1[(k * 6 + i) * TILE];
            a2[i] = pa2[(k * 6 + i) * TILE];
          }
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (i <= j) {
                TT(i, j) -= zo[i] * zo[j];
                BB(i, j) -= zo[3 + i] * zo[3 + j];
              }
              TB[i][j] -= zo[i] * zo[3 + j];
              CA[i][j] -= a1[i] * b1[j];
              CB[i][j] -= a2[i] * b1[j];
            }
          asm volatile("" ::: "memory");
        }
        __syncwarp(m_upd);   // the scratch is rewritten by the next trip
      }
      // stores: state, own group, cross columns
      double* outg = out + (size_t)(39 * h) * TILE;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        out[(LY::F_X + 3 * h + i) * TILE] = top[i];
        out[(LY::F_X + 6 + 3 * h + i) * TILE] = bot[i];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (i <= j) {
            outg[(LY::F_P + i * N + j) * TILE] = TT(i, j);
            outg[(LY::F_P + (6 + i) * N + 6 + j) * TILE] = BB(i, j);
          }
          outg[(LY::F_P + i * N + 6 + j) * TILE] = TB[i][j];
          // lane 0: CA = pr (p_i, r_j), CB = vr (v_i, r_j) -> stored as (r_j, v_i); lane 1: CA = pw (p_i, w_j), CB = vw (v_i, w_j)
          out[(LY::F_P + (odd ? i * N + 9 + j : i * N + 3 + j)) * TILE] = CA[i][j];
          out[(LY::F_P + (odd ? (6 + i) * N + 9 + j : (3 + j) * N + 6 + i)) * TILE] = CB[i][j];
        }
      if (!odd) {
        out[LY::F_T * TILE] = t_in + dt;
        reinterpret_cast<long long*>(out)[LY::F_NMEAS * TILE] = nm_in + (act == ACT_UPDATE ? 1 : 0);
        if (a.clear_action) a.action[slot] = 0;
        if (a.pos_out) {
#pragma unroll
          for (int k = 0; k < 3; ++k) a.pos_out[(size_t)slot * 3 + k] = top[k];
        }
      }
    } else if (valid) {
      if (a.dst_tiles && dst >= 0) {
#pragma unroll 8
        for (int f = h; f < LY::NF; f += 2) __stcs(out + (size_t)f * TILE, __ldcs(in + (size_t)f * TILE));
      }
      if (a.pos_out && !odd) {
#pragma unroll
        for (int k = 0; k < 3; ++k) a.pos_out[(size_t)slot * 3 + k] = in[(LY::F_X + k) * TILE];
      }
    }
    if (a.clear_action && lane == 0) a.tile_flag[tile] = 0;
  }
}

}  // namespace te
