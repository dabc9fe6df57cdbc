// te_ar_pair.cuh -- packed-covariance step kernel of the angular-rates model (n = 18, m = 6): two lanes per target.
//
// The packed state of an AR target -- 18 + 171 doubles -- does not fit one thread, half of it does.  A warp therefore steps
// 16 targets at a time, lanes t and t + 16 sharing target t, and the symmetric 18 x 18 covariance is split by POSITION:
// A(dt) = I + dt E_6 + dt^2/2 E_12 couples only entries with the same (row mod 6, column mod 6) = (r, c), so the covariance
// is 21 independent 3 x 3 "macro" matrices m_rc[a][b] = P(6a + r, 6b + c), r <= c (6 unique entries when r = c), and the
// predict of each position is thread-local.  Half 0 owns positions (0,0) (3,3) (4,4) | (0,1..5) (3,4) (3,5) (4,5), half 1
// (1,1) (2,2) (5,5) | (1,2..5) (2,3..5) and one inactive duplicate: 3 diagonal + 8 off-diagonal slots = 90 doubles each,
// the same instruction stream for both halves (positions are data: two integers per slot).
//
// The update needs the top block B = P'[0:6,:] everywhere: every lane drops the entries it holds (macro row 0 and, by
// symmetry, macro column 0 of its positions) into a 6 x 18 shared-memory scratch per target, both halves factor
// S = B[:,0:6] + R (Cholesky, redundantly), each half forward-substitutes nine columns of Z = L^-1 B in place, and after one
// more warp barrier every position is downdated with the Z columns of its residues: m_rc[a][b] -= Z[:,6a+r] . Z[:,6b+c].
// State entries are split by parity.  Loads and stores go straight from / to HBM, upper triangle only (each half-warp moves
// a 128 B segment per instruction); the tile format and every result are those of the other kernels.
// Reference: src/types/angular_rates.cpp:72-115, src/kalman.cpp:84-95.
#pragma once
#include "te_kernels.cuh"

namespace te {

constexpr int AR_PAIR_BZ_DOUBLES = 6 * 18 * 16;          // B / Z per warp: [k][j][target of the pass]
constexpr int AR_PAIR_SCRATCH_DOUBLES = AR_PAIR_BZ_DOUBLES + 18 * 32;   // + the three diagonal positions of every lane: [slot][entry][lane]
__host__ __device__ constexpr size_t ar_pair_smem_bytes(int warps) { return (size_t)warps * AR_PAIR_SCRATCH_DOUBLES * 8; }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kf_step_ar_pair_kernel(const StepArgs a) {
  using LY = Layout<ANGULAR_RATES>;
  constexpr int N = 18, M = 6, FP = LY::F_P;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = lane >> 4, tl = lane & 15;
  double* sc = reinterpret_cast<double*>(smem_raw) + (size_t)warp * AR_PAIR_SCRATCH_DOUBLES + tl;   // entry (k, j) at sc[(k * N + j) * 16]
  // the diagonal positions (18 of a lane's 90 covariance entries) live in the lane's own shared-memory column between their
  // three uses: 72 + 18 doubles in registers left ptxas ~60 doubles short
  double* dsc = reinterpret_cast<double*>(smem_raw) + (size_t)warp * AR_PAIR_SCRATCH_DOUBLES + AR_PAIR_BZ_DOUBLES + lane;   // D[s][e] at dsc[(s * 6 + e) * 32]
  // positions of this half: residues (r, c) of the 3 diagonal and 8 off-diagonal slots; slot 7 of half 1 duplicates (2,5)
  int rD[3], rO[8], cO[8];
  rD[0] = h ? 1 : 0; rD[1] = h ? 2 : 3; rD[2] = h ? 5 : 4;
  rO[0] = h ? 1 : 0; cO[0] = h ? 2 : 1;
  rO[1] = h ? 1 : 0; cO[1] = h ? 3 : 2;
  rO[2] = h ? 1 : 0; cO[2] = h ? 4 : 3;
  rO[3] = h ? 1 : 0; cO[3] = h ? 5 : 4;
  rO[4] = h ? 2 : 0; cO[4] = h ? 3 : 5;
  rO[5] = h ? 2 : 3; cO[5] = h ? 4 : 4;
  rO[6] = h ? 2 : 3; cO[6] = 5;
  rO[7] = h ? 2 : 4; cO[7] = 5;
  const bool dup7 = h == 1;   // slot 7 of half 1 is computed but never published

  const int n_work = a.d_nwork ? *a.d_nwork : a.n_tiles;
  const int gw = blockIdx.x * WARPS + warp, GW = gridDim.x * WARPS;
  for (int w = gw; w < n_work; w += GW) {
    const int tile = a.tile_list ? a.tile_list[w] : a.tile_begin + w;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int tt = pass * 16 + tl;            // target within the tile
      const int slot = tile * TILE + tt;
      const bool valid = slot < a.n_slots;
      int act = ACT_NONE, cls = 0, dst = -1;
      double dt = a.dt;
      if (valid) {
        act = a.action ? (int)a.action[slot] : a.default_action;
        if (a.dt_slot) dt = a.dt_slot[slot];
        cls = (int)a.cls[slot];
        if (a.dst_tiles) {
          if (a.dst_alive[slot]) dst = a.dst_pos[slot];
          else act = ACT_NONE;   // erased at the end of this tick: its step is unobservable
        }
      }
      const double* in = a.tiles + (size_t)tile * LY::TILE_DOUBLES + tt;
      double* out = a.dst_tiles ? (dst >= 0 ? a.dst_tiles + (size_t)(dst / TILE) * LY::TILE_DOUBLES + (dst % TILE) : nullptr)
                                : a.tiles + (size_t)tile * LY::TILE_DOUBLES + tt;
      const bool active = act != ACT_NONE, upd = act == ACT_UPDATE;
      const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
      const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;

      // ---- measurement conversion first (angular_rates.cpp:79-88, both halves): its calls (atan2 / asin / fmod slow paths)
      //      would otherwise force the ~100 live covariance registers through local memory ----
      double prev[3] = {0.0, 0.0, 0.0}, y[M];
      if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) prev[k] = in[(LY::F_PREV + k) * TILE];
      }
      if (upd) {
        double meas[7], un[3];
        const double* mp = a.meas + (size_t)slot * a.meas_stride;
#pragma unroll
        for (int k = 0; k < 7; ++k) meas[k] = __ldg(mp + k);
        meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
        for (int k = 0; k < 3; ++k) { y[k] = meas[k]; y[3 + k] = un[k]; prev[k] = un[k]; }
      }

      // ---- loads: this half's state entries and positions, bookkeeping ----
      double xo[9], O[8][9];   // xo: the state entries of this half's parity, i = 2 ii + h (closed under the predict)
      double t_in = 0.0;
      long long nm_in = 0;
      if (active) {
#pragma unroll
        for (int ii = 0; ii < 9; ++ii) xo[ii] = in[(LY::F_X + 2 * ii + h) * TILE];
        t_in = in[LY::F_T * TILE];
        nm_in = reinterpret_cast<const long long*>(in)[LY::F_NMEAS * TILE];
        // upper-triangle field of macro entry (a, b) of position (r, c): a <= b -> (6a + r, 6b + c), else its mirror
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int base = FP + 19 * rD[s];
          constexpr int off[6] = {0, 6, 12, 114, 120, 228};
#pragma unroll
          for (int e = 0; e < 6; ++e) dsc[(s * 6 + e) * 32] = in[(base + off[e]) * TILE];
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int b1 = FP + 18 * rO[s] + cO[s], b2 = FP + 18 * cO[s] + rO[s];
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb)
              O[s][aa * 3 + bb] = aa <= bb ? in[(b1 + aa * 108 + bb * 6) * TILE] : in[(b2 + bb * 108 + aa * 6) * TILE];
        }
      }

      // ---- predict ----
      if (active) {
        const double hh = 0.5 * dt * dt;
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) {   // entries i, i + 6, i + 12 share their parity: xo[ii], xo[ii + 3], xo[ii + 6]
          xo[ii] = xo[ii] + dt * xo[ii + 3] + hh * xo[ii + 6];
          xo[ii + 3] = xo[ii + 3] + dt * xo[ii + 6];
        }
        auto predict_macro = [&](double m[3][3], int r, int c) {
#pragma unroll
          for (int b = 0; b < 3; ++b) {   // A P : rows
            m[0][b] = m[0][b] + dt * m[1][b] + hh * m[2][b];
            m[1][b] = m[1][b] + dt * m[2][b];
          }
#pragma unroll
          for (int aa = 0; aa < 3; ++aa) {   // (A P) A^T : columns
            m[aa][0] = m[aa][0] + dt * m[aa][1] + hh * m[aa][2];
            m[aa][1] = m[aa][1] + dt * m[aa][2];
          }
          const double* q = Q + 18 * r + c;
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) m[aa][bb] = m[aa][bb] + __ldg(q + aa * 108 + bb * 6);
        };
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          double d[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) d[e] = dsc[(s * 6 + e) * 32];
          double m[3][3] = {{d[0], d[1], d[2]}, {d[1], d[3], d[4]}, {d[2], d[4], d[5]}};
          predict_macro(m, rD[s], rD[s]);
          dsc[(s * 6 + 0) * 32] = m[0][0]; dsc[(s * 6 + 1) * 32] = m[0][1]; dsc[(s * 6 + 2) * 32] = m[0][2];
          dsc[(s * 6 + 3) * 32] = m[1][1]; dsc[(s * 6 + 4) * 32] = m[1][2]; dsc[(s * 6 + 5) * 32] = m[2][2];
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          double m[3][3];
#pragma unroll
          for (int e = 0; e < 9; ++e) m[e / 3][e % 3] = O[s][e];
          predict_macro(m, rO[s], cO[s]);
#pragma unroll
          for (int e = 0; e < 9; ++e) O[s][e] = m[e / 3][e % 3];
        }
      }

      // ---- update ----
      __syncwarp();   // the scratch of the previous pass has been read by everybody
      if (upd) {      // B = P'[0:6,:]: row r from macro row 0, row c (by symmetry) from macro column 0
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int b = 0; b < 3; ++b) sc[(rD[s] * N + b * 6 + rD[s]) * 16] = dsc[(s * 6 + b) * 32];
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            sc[(rO[s] * N + b * 6 + cO[s]) * 16] = O[s][b];
            sc[(cO[s] * N + b * 6 + rO[s]) * 16] = O[s][b * 3];
          }
      }
      __syncwarp();
      // predicted x'[0:6]: three entries are this half's, three the partner's
      double xp[6];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double mine = active ? xo[q] : 0.0;
        const double other = __shfl_xor_sync(0xffffffffu, mine, 16);
        xp[2 * q] = h == 0 ? mine : other;
        xp[2 * q + 1] = h == 0 ? other : mine;
      }
      double u[M];
      Chol<M> ch;
      if (upd) {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (j <= i) ch.at(i, j) = sc[(i * N + j) * 16] + __ldg(&R[i * M + j]);
        ch.factor();
#pragma unroll
        for (int k = 0; k < M; ++k) {   // u = L^-1 (y - x'[0:6])
          double s = y[k] - xp[k];
#pragma unroll
          for (int m = 0; m < M; ++m)
            if (m < k) s -= ch.L[k][m] * u[m];
          u[k] = s * ch.L[k][k];
        }
      }
      __syncwarp();   // both halves have read S = B[:,0:6] before anybody overwrites those columns with Z
      if (upd) {
#pragma unroll
        for (int jj = 0; jj < 9; ++jj) {   // Z = L^-1 B, in place: this half's nine columns
          const int j = 2 * jj + h;
          double z[M];
#pragma unroll
          for (int k = 0; k < M; ++k) {
            double s = sc[(k * N + j) * 16];
#pragma unroll
            for (int m = 0; m < M; ++m)
              if (m < k) s -= ch.L[k][m] * z[m];
            z[k] = s * ch.L[k][k];
          }
#pragma unroll
          for (int k = 0; k < M; ++k) sc[(k * N + j) * 16] = z[k];
        }
      }
      __syncwarp();
      if (upd) {
        auto zcol = [&](int j, double z[M]) {
#pragma unroll
          for (int k = 0; k < M; ++k) z[k] = sc[(k * N + j) * 16];
        };
        // state first (frees u): x_i += Z[:, i] . u, entries of this half's parity
#pragma unroll
        for (int ii = 0; ii < 9; ++ii) {
          const int i = 2 * ii + h;
          double z[M];
          zcol(i, z);
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < M; ++k) s += z[k] * u[k];
          xo[ii] += s;
        }
        asm volatile("" ::: "memory");
        // every position is downdated with the Z columns of its residues, one Z row at a time (six scratch loads feed nine
        // independent FMAs; the empty asm statements are scheduling fences that keep the front end from hoisting the loads of
        // all positions, which made ptxas spill)
#pragma unroll
        for (int s = 0; s < 3; ++s) {   // diagonal positions: m[a][b] -= Z[:,6a+r] . Z[:,6b+r], a <= b
          double d[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) d[e] = dsc[(s * 6 + e) * 32];
#pragma unroll
          for (int k = 0; k < M; ++k) {
            const double z0 = sc[(k * N + rD[s]) * 16], z1 = sc[(k * N + 6 + rD[s]) * 16], z2 = sc[(k * N + 12 + rD[s]) * 16];
            d[0] -= z0 * z0; d[1] -= z0 * z1; d[2] -= z0 * z2;
            d[3] -= z1 * z1; d[4] -= z1 * z2; d[5] -= z2 * z2;
          }
#pragma unroll
          for (int e = 0; e < 6; ++e) dsc[(s * 6 + e) * 32] = d[e];
          asm volatile("" ::: "memory");
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
#pragma unroll
          for (int k = 0; k < M; ++k) {
            double zr[3], zc[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) { zr[q] = sc[(k * N + 6 * q + rO[s]) * 16]; zc[q] = sc[(k * N + 6 * q + cO[s]) * 16]; }
#pragma unroll
            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) O[s][aa * 3 + bb] -= zr[aa] * zc[bb];
          }
          asm volatile("" ::: "memory");
        }
      }

      // ---- stores (upper triangle only when packed; this half's state entries; bookkeeping by half 0 / 1) ----
      if (active && out) {
#pragma unroll
        for (int ii = 0; ii < 9; ++ii) {
          const double v = xo[ii];
          out[(LY::F_X + 2 * ii + h) * TILE] = v;
          if (a.pos_out && 2 * ii + h < 3) a.pos_out[(size_t)slot * 3 + 2 * ii + h] = v;
        }
        if (h == 0) {
          out[LY::F_T * TILE] = t_in + dt;
          reinterpret_cast<long long*>(out)[LY::F_NMEAS * TILE] = nm_in + (upd ? 1 : 0);
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) out[(LY::F_PREV + k) * TILE] = prev[k];
        }
        const bool both = a.packed == 0;   // unpacked: the mirrored (lower-triangle) field is written too
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int base = FP + 19 * rD[s];
          double d[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) d[e] = dsc[(s * 6 + e) * 32];
          out[(base + 0) * TILE] = d[0];   out[(base + 6) * TILE] = d[1];   out[(base + 12) * TILE] = d[2];
          out[(base + 114) * TILE] = d[3]; out[(base + 120) * TILE] = d[4]; out[(base + 228) * TILE] = d[5];
          if (both) {   // mirrors of the three off-diagonal macro entries: (6b + r, 6a + r), a < b
            out[(base + 108) * TILE] = d[1]; out[(base + 216) * TILE] = d[2]; out[(base + 222) * TILE] = d[4];
          }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (s == 7 && dup7) continue;
          const int b1 = FP + 18 * rO[s] + cO[s], b2 = FP + 18 * cO[s] + rO[s];
#pragma unroll
          for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
              const double v = O[s][aa * 3 + bb];
              const int up = aa <= bb ? b1 + aa * 108 + bb * 6 : b2 + bb * 108 + aa * 6;
              const int lo = aa <= bb ? b2 + bb * 108 + aa * 6 : b1 + aa * 108 + bb * 6;
              out[up * TILE] = v;
              if (both) out[lo * TILE] = v;
            }
        }
        if (a.clear_action && h == 0) a.action[slot] = 0;
      } else if (valid) {
        if (a.dst_tiles && dst >= 0) {   // compacting tick: an untouched survivor still moves to its new slot (fields by parity)
#pragma unroll 4
          for (int f = h; f < LY::NF; f += 2) __stcs(out + (size_t)f * TILE, __ldcs(in + (size_t)f * TILE));
        }
        if (a.pos_out && h == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) a.pos_out[(size_t)slot * 3 + k] = in[(LY::F_X + k) * TILE];
        }
      }
    }
    if (a.clear_action && lane == 0) a.tile_flag[tile] = 0;
  }
}

}  // namespace te
