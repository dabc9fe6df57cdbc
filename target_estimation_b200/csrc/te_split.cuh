// te_split.cuh -- warp-specialised row/column-split KF step kernel for the large-state models (AR n=18, AV n=12).
//
// One CTA works on one tile of 32 targets (lane = target) with two groups of warps:
//
//  * 6*CS MAIN warps.  Warp (r, h) owns, for every target of the tile, the rows {q*6 + r} and the columns
//    {b*6 + 3h + cc} (CS = 2; all columns for CS = 1) of the covariance, in registers.  With that ownership the predict
//    of the kinematic models is thread-local (row r of each kinematic block, and the columns c, 6+c, 12+c that the
//    banded A(dt) couples), and the EKF predict of AV needs only the original rows 3..5 / 9..11 from the staged tile.
//  * NT CONVERTER warps, running up to STAGES-1 tiles AHEAD of the main warps: one Euler angle of the measurement
//    each (normalise quaternion -> rpy -> unwrap against the previous unwrapped angle: sqrt, divisions, atan2 / asin
//    and three fmods -- a ~4000-cycle dependent chain that used to sit on the critical path of every tile), and for AV
//    a fourth warp that evaluates the Euler-rate Jacobians J_rpy, J_w and E^-1 once per target.  They hand y (and J) to
//    the main warps through shared memory and a named barrier (bar.arrive / bar.sync) per stage.
//
// The tile is staged in shared memory by one TMA bulk copy (mbarrier completion; both groups wait on it) and written
// back by one bulk store, as in kf_step_kernel.  Main-warp phases, separated by named barriers of the main group:
//
//   A  predict own block (x' = A x | f(x), P' = A P A^T + Q); publish in place: row r of P', x'[r], P'[own rows, 0:6]  | barrier
//   B  warp 0: S = P'[0:6,0:6] + R (in-register Cholesky), v = S^-1 (y - x'[0:6]); warps 1..: factor S, columns of
//      W = S^-1 P'[0:6,:] -> W buffer; the last warp's lane 0 is the TMA producer and refills the other stage here   | barrier
//   C  own block: x += P'[rows, 0:6] v ; P[rows, cols] = P'[rows, cols] - P'[rows, 0:6] W[:, cols]
//   D  own block back into the stage, t / n_meas bookkeeping                                               | barrier, bulk store
// (AV: the r >= 3 warps read the original rows 3..5 / 9..11 in A; a named barrier among them orders those reads
//  before the in-place publish.)
//
// Same arithmetic as te_device.cuh's step_lane (predict_kinematic / predict_av / kf_update); only the ownership of
// the elements differs.  Reference: src/kalman.cpp:84-95,129-140, src/types/angular_rates.cpp:72-115,
// src/types/angular_velocities.cpp:80-151.
#pragma once
#include "te_kernels.cuh"

namespace te {

#ifdef TE_TIMELINE
// debug build only: CTA 0 / thread 0 records clock64() at the phase boundaries of its first 64 tiles
__device__ long long g_timeline[2 * 64 * 12];
#define TE_MARK(k) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 96) && it < 64) g_timeline[((threadIdx.x ? 1 : 0) * 64 + it) * 12 + (k)] = clock64(); } while (0)
#else
#define TE_MARK(k) do { } while (0)
#endif

#ifndef TE_BULK_CHUNK_DOUBLES
#define TE_BULK_CHUNK_DOUBLES (1 << 20)   // whole tile in one bulk copy (8 KB chunks measured slightly slower)
#endif
constexpr int BULK_CHUNK_DOUBLES = TE_BULK_CHUNK_DOUBLES;
#ifndef TE_SKIP
#define TE_SKIP 0   // timing experiments only (results become wrong): 1 skip phase B, 2 skip phase C, 4 skip predict, 16 skip converters
#endif
constexpr int SPLIT_RS = 6;     // row owners per target (= M for both angular models)
constexpr int JBUF_FIELDS = 14; // AV: J_rpy (5 non-trivial) + J_w (6) + predicted rpy (3)
constexpr int SPLIT_BAR_BYTES = 128;   // mbarriers in front of the stages

// converter warps: 2, so that 6 main + 2 = 8 warps = two per scheduler partition and every thread may keep 255 registers
// (a 9th warp would cap all of them at 168: registers are allocated per partition)
template <int TYPE> __host__ __device__ constexpr int split_nt() { return 2; }

// shared memory of one CTA: [mbarriers 128 B][STAGES x tile][W: 6 x N x 32][STAGES x y: 6 x 32][STAGES x jbuf: 14 x 32 (AV)]
// (AV, 2 stages: 112 128 B -> two CTAs per SM)
template <int TYPE> __host__ __device__ constexpr size_t split_smem_bytes(int stages) {
  return SPLIT_BAR_BYTES + ((size_t)stages * Layout<TYPE>::TILE_DOUBLES + (size_t)Model<TYPE>::M * Model<TYPE>::N * TILE +
                 (size_t)stages * (6 * TILE + (TYPE == ANGULAR_VELOCITIES ? JBUF_FIELDS * TILE : 0))) * 8;
}

// named barriers with immediate ids (a register id makes ptxas reserve all 16 hardware barriers, i.e. one CTA per SM)
template <int ID, int COUNT> __device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT> __device__ __forceinline__ void named_bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int BASE, int COUNT> __device__ __forceinline__ void stage_bar_sync(int s) {
  if (s == 0) named_bar_sync<BASE, COUNT>(); else if (s == 1) named_bar_sync<BASE + 1, COUNT>();
  else if (s == 2) named_bar_sync<BASE + 2, COUNT>(); else named_bar_sync<BASE + 3, COUNT>();
}
template <int BASE, int COUNT> __device__ __forceinline__ void stage_bar_arrive(int s) {
  if (s == 0) named_bar_arrive<BASE, COUNT>(); else if (s == 1) named_bar_arrive<BASE + 1, COUNT>();
  else if (s == 2) named_bar_arrive<BASE + 2, COUNT>(); else named_bar_arrive<BASE + 3, COUNT>();
}

// component k of quatToRpy (geometry.hpp:154-176), same expressions as quat_to_rpy()
__device__ __forceinline__ double quat_to_rpy_comp(const Quat& q, int k) {
  const double s = -2 * (q.x * q.z - q.w * q.y);
  const bool gimbal = (s > 0.9999) || (s < -0.9999);
  if (k == 0) return gimbal ? 0.0 : atan2(2 * (q.y * q.z + q.w * q.x), (q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z));
  if (k == 1) return s > 0.9999 ? TE_PI / 2 : (s < -0.9999 ? -TE_PI / 2 : asin(s));
  // one atan2 call site for both branches (the function is ~150 instructions inlined)
  const double ay = gimbal ? q.z : 2 * (q.x * q.y + q.w * q.z);
  const double ax = gimbal ? q.w : (q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z);
  const double at = atan2(ay, ax);
  return gimbal ? 2 * at : at;
}

template <int TYPE, int CS, int STAGES, int MIN_CTAS, bool COMPACT = false>
__global__ void __launch_bounds__((SPLIT_RS * CS + split_nt<TYPE>()) * 32, MIN_CTAS) kf_step_split_kernel(const StepArgs a) {
  using MT = Model<TYPE>;
  using LY = Layout<TYPE>;
  constexpr int N = MT::N, M = MT::M, RS = SPLIT_RS;
  constexpr int NW = RS * CS;        // warps per CTA
  constexpr int RPT = N / RS;        // rows per thread: AR 3, AV 2
  constexpr int NCOL = N / CS;       // columns per thread
  constexpr int BL = 6 / CS;         // own columns per block of 6
  constexpr int NT = split_nt<TYPE>();         // converter warps
  constexpr int NMAIN = NW * 32;               // threads of the main group
  constexpr int WPW = (N + NW - 2) / (NW - 1); // W columns per warp 1.. (warp 0 solves v)
  // named barriers: 0 = __syncthreads at start-up, 1 = AV r >= 3 group, BAR_MAIN = main group,
  // BAR_Y + s = converters -> main (y / J of the tile in stage s).  "tile in stage s finished" (main -> converters /
  // producer) is an mbarrier, done[s], so that the two converters can wait for different tiles independently.
  constexpr int BAR_MAIN = 2, BAR_Y = 3;
  constexpr int STAGE_DOUBLES = LY::TILE_DOUBLES;   // no measurement block: the converters read measurements from global memory
  static_assert(M == RS && (CS == 1 || CS == 2) && STAGES <= 4, "one measured row per row owner");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  double* stage0 = reinterpret_cast<double*>(smem_raw + SPLIT_BAR_BYTES);
  double* Wbuf = stage0 + (size_t)STAGES * STAGE_DOUBLES;   // [M][N][32]
  double* ybuf0 = Wbuf + (size_t)M * N * TILE;              // [STAGES][6][32]
  double* jbuf0 = ybuf0 + (size_t)STAGES * 6 * TILE;        // [STAGES][17][32] (AV)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = w % RS, h = (w / RS) % CS;
  const bool producer = threadIdx.x == (NW + 1) * 32;       // lane 0 of converter warp 1: every TMA issue / wait lives there
  auto colof = [&](int j) -> int { return CS == 1 ? j : (j / BL) * 6 + BL * h + (j % BL); };

  if (producer) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars[s], 1);                 // full[s]: tile landed (one expect_tx arrival + the copy's bytes)
      mbar_init(&bars[STAGES + s], NW);       // done[s]: every main warp finished the tile in stage s (lane 0 arrives)
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int n_work = a.d_nwork ? *a.d_nwork : a.n_tiles;
  const int n_my = (n_work > (int)blockIdx.x) ? (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int it) -> int {
    const int ww = blockIdx.x + it * gridDim.x;
    return a.tile_list ? a.tile_list[ww] : a.tile_begin + ww;
  };
  auto use_meas_tma = [&](int tile) -> bool { return a.meas_tma && (tile * TILE + TILE <= a.n_slots); };
  // Packed covariance (AR, a.packed): only the field ranges the step needs travel -- x + covariance row 0, the upper parts
  // of rows 1..N-2, the last diagonal entry + t / n_meas / prev_rpy: 194 of AR's 347 fields.  The lower triangle of the
  // staged tile stays unwritten; a row owner reads P(g, c), c < g, from the mirrored field P(c, g) instead, and only the
  // same ranges are stored back (each row's upper part comes from the warp that owns the row).
  const bool packed = (TYPE == ANGULAR_RATES) && a.packed != 0;
  auto seg_f0 = [&](int i) -> int { return i == 0 ? 0 : (i < N - 1 ? LY::F_P + i * (N + 1) : LY::F_P + N * N - 1); };
  auto seg_nf = [&](int i) -> int { return i == 0 ? 2 * N : (i < N - 1 ? N - i : 1 + (LY::NF - (LY::F_P + N * N))); };
  constexpr uint32_t PACKED_BYTES = (uint32_t)(N + N * (N + 1) / 2 + (LY::NF - (LY::F_P + N * N))) * TILE * 8u;
  auto issue = [&](int it) {
    const int tile = tile_of(it);
    const int s = it % STAGES;
    double* st = stage0 + (size_t)s * STAGE_DOUBLES;
    const double* src = a.tiles + (size_t)tile * LY::TILE_DOUBLES;
    if (packed) {
      mbar_expect_tx(&bars[s], PACKED_BYTES);
#pragma unroll 1
      for (int i = 0; i < N; ++i) bulk_g2s(st + (size_t)seg_f0(i) * TILE, src + (size_t)seg_f0(i) * TILE, (uint32_t)seg_nf(i) * TILE * 8u, &bars[s]);
      return;
    }
    mbar_expect_tx(&bars[s], (uint32_t)LY::TILE_BYTES);   // measurements are read from global memory by the converters
    // the tile travels as several bulk copies on one mbarrier: a single 40-90 KB copy is served at ~18 B/clk, several
    // smaller ones overlap in the copy engine
#pragma unroll 1
    for (int off = 0; off < LY::TILE_DOUBLES; off += BULK_CHUNK_DOUBLES) {
      const int nd = (LY::TILE_DOUBLES - off) < BULK_CHUNK_DOUBLES ? (LY::TILE_DOUBLES - off) : BULK_CHUNK_DOUBLES;
      bulk_g2s(st + off, src + off, (uint32_t)nd * 8u, &bars[s]);
    }
  };
  if (producer) {
    for (int pre = 0; pre < STAGES && pre < n_my; ++pre) issue(pre);
  }

  // per-lane control words, fetched one tile ahead so that their global-load latency never sits in front of a tile
  int act_n = ACT_NONE, cls_n = 0, dst_n = -1;   // dst_n: compacting tick, destination slot of this lane's target (-1 = erased)
  double dt_n = a.dt;
  auto load_ctrl = [&](int it) {
    act_n = ACT_NONE; cls_n = 0; dt_n = a.dt; dst_n = -1;
    if (it < n_my) {
      const int slot = tile_of(it) * TILE + lane;
      if (slot < a.n_slots) {
        act_n = a.action ? (int)a.action[slot] : a.default_action;
        if (a.dt_slot) dt_n = a.dt_slot[slot];
        cls_n = (int)a.cls[slot];
        if (COMPACT) {
          if (a.dst_alive[slot]) dst_n = a.dst_pos[slot];
          else act_n = ACT_NONE;   // erased at the end of this tick: its step is unobservable
        }
      }
    }
  };
  load_ctrl(0);

  // =========================== converter warps ===========================
  if (w >= NW) {
    const int tw = w - NW;
    // converter 1 is also the TMA producer: when the main warps have finished tile jd (named barrier BAR_DONE + stage) it
    // issues the bulk store, waits until the store has drained the stage and refills it with tile jd + STAGES -- the main
    // warps never wait on a copy
    unsigned anyh[4] = {0u, 0u, 0u, 0u};   // "tile has work" per stage, noted when the converter handled the tile
    auto wait_done = [&](int jd) {   // the main warps have finished tile jd (and with it the y / J buffers of its stage)
      if (jd >= 0 && jd < n_my) mbar_wait(&bars[STAGES + jd % STAGES], (uint32_t)(jd / STAGES) & 1u);
    };
    auto retire = [&](int jd) {      // converter 1 / producer: write tile jd back, refill its stage with tile jd + STAGES
      if (jd < 0 || jd >= n_my) return;
      const int sd = jd % STAGES;
      const int tile_d = tile_of(jd);
      const unsigned any_d = sd == 0 ? anyh[0] : (sd == 1 ? anyh[1] : (sd == 2 ? anyh[2] : anyh[3]));
      wait_done(jd);
      double* std_ = stage0 + (size_t)sd * STAGE_DOUBLES;
      if (COMPACT) {
        // compacting tick (te_pool_step_dense_expire): the main warps have already written the survivors' columns to their
        // destination slots in the other buffer -- no bulk store, and the source buffer is never written
        if (producer && jd + STAGES < n_my) issue(jd + STAGES);
        return;
      }
      if (producer) {
        if (any_d) {
          double* dstt = a.tiles + (size_t)tile_d * LY::TILE_DOUBLES;
          if (packed) {
#pragma unroll 1
            for (int i = 0; i < N; ++i) bulk_s2g(dstt + (size_t)seg_f0(i) * TILE, std_ + (size_t)seg_f0(i) * TILE, (uint32_t)seg_nf(i) * TILE * 8u);
          } else {
            bulk_s2g(dstt, std_, LY::TILE_BYTES);
          }
          bulk_commit();
          if (a.clear_action) a.tile_flag[tile_d] = 0;
        }
        if (jd + STAGES < n_my) {
          bulk_wait_read<0>();
          issue(jd + STAGES);
        }
      }
    };
    // The converters never touch the staged tile: measurement, previous unwrapped angles and (AV) the previous posterior
    // come straight from global memory -- this tick has not rewritten the tile yet -- so they run while the tile's bulk
    // copy is still in flight, up to STAGES tiles ahead of the main warps.  Their inputs are fetched one tile ahead too.
    double in_n[12];   // [0..6] measurement pose, [7..8] previous unwrapped angle(s) of this converter, AV: [6..11] reused below
    double xin_n[6];   // AV converter 0: x[3..5], x[9..11]
    auto load_inputs = [&](int it) {
#pragma unroll
      for (int k = 0; k < 12; ++k) in_n[k] = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) xin_n[k] = 0.0;
      if (it < n_my) {
        const int tile = tile_of(it);
        const int slot = tile * TILE + lane;
        if (slot < a.n_slots) {
          const double* gt = a.tiles + (size_t)tile * LY::TILE_DOUBLES + lane;   // this lane's column of the tile in HBM
          if (a.meas) {
            const double* mp = a.meas + (size_t)slot * a.meas_stride;
#pragma unroll
            for (int k = 0; k < 7; ++k) in_n[k] = __ldg(mp + k);
          }
          if (tw == 0) {
            in_n[7] = gt[(LY::F_PREV + 0) * TILE];
            in_n[8] = gt[(LY::F_PREV + 2) * TILE];
            if (TYPE == ANGULAR_VELOCITIES) {
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                xin_n[k] = gt[(LY::F_X + 3 + k) * TILE];
                xin_n[3 + k] = gt[(LY::F_X + 9 + k) * TILE];
              }
            }
          } else {
            in_n[7] = gt[(LY::F_PREV + 1) * TILE];
          }
        }
      }
    };
    load_inputs(0);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % STAGES;
      const int act = act_n;
      const double dt = dt_n;
      double in[9], xin[6];
#pragma unroll
      for (int k = 0; k < 9; ++k) in[k] = in_n[k];
#pragma unroll
      for (int k = 0; k < 6; ++k) xin[k] = xin_n[k];
      load_ctrl(it + 1);
      load_inputs(it + 1);
      // y / J buffers of stage s were last used by tile it - STAGES.  Converter 1 has already waited for that tile in its
      // retire(); converter 0 is not tied to the producer and may run up to STAGES tiles ahead of the main warps.
      if (tw == 0) wait_done(it - STAGES);
      double* ybuf = ybuf0 + (size_t)s * 6 * TILE;
      {
        const unsigned any_t = __ballot_sync(0xffffffffu, act != ACT_NONE);
        if (s == 0) anyh[0] = any_t; else if (s == 1) anyh[1] = any_t; else if (s == 2) anyh[2] = any_t; else anyh[3] = any_t;
      }
      if (act == ACT_UPDATE && !(TE_SKIP & 16)) {   // angular_rates.cpp:79-88 / angular_velocities.cpp:87-96
        // converter 0: roll and yaw (two atan2 chains; with the AV Jacobians a third independent chain, all interleaved by
        // the scheduler); converter 1, which is also the TMA producer and must react quickly, only pitch (asin)
        Quat qm{in[3], in[4], in[5], in[6]};
        quat_normalize(qm);
        if (tw == 0) {
          const double a0 = quat_to_rpy_comp(qm, 0), a2 = quat_to_rpy_comp(qm, 2);
          ybuf[3 * TILE + lane] = unwrap1(in[7], a0);
          ybuf[5 * TILE + lane] = unwrap1(in[8], a2);
          ybuf[0 * TILE + lane] = in[0];
          ybuf[2 * TILE + lane] = in[2];
        } else {
          const double a1 = quat_to_rpy_comp(qm, 1);
          ybuf[4 * TILE + lane] = unwrap1(in[7], a1);
          ybuf[1 * TILE + lane] = in[1];
        }
      }
      if (TYPE == ANGULAR_VELOCITIES && tw == 0) {
        if (act != ACT_NONE) {   // Jacobians at the previous posterior (angular_velocities.cpp:116-124, geometry.hpp:359-426)
          double s_r, c_r, s_p, c_p;
          const double x3 = xin[0], x4 = xin[1], x5 = xin[2];
          sincos(x3, &s_r, &c_r);
          sincos(x4, &s_p, &c_p);
          const double wx = xin[3], wy = xin[4], wz = xin[5];
          double* jb = jbuf0 + (size_t)s * JBUF_FIELDS * TILE + lane;
          jb[0 * TILE] = (dt * (wy * c_r * s_p - wz * s_p * s_r)) / c_p + 1;   // J1[0][0]
          jb[1 * TILE] = (dt * (wz * c_r + wy * s_r)) / (c_p * c_p);           // J1[0][1]
          jb[2 * TILE] = -dt * (wz * c_r + wy * s_r);                          // J1[1][0]
          jb[3 * TILE] = (dt * (wy * c_r - wz * s_r)) / c_p;                   // J1[2][0]
          jb[4 * TILE] = (dt * s_p * (wz * c_r + wy * s_r)) / (c_p * c_p);     // J1[2][1]
          jb[5 * TILE] = (dt * s_p * s_r) / c_p;                               // J2[0][1]
          jb[6 * TILE] = (dt * c_r * s_p) / c_p;                               // J2[0][2]
          jb[7 * TILE] = dt * c_r;                                             // J2[1][1]
          jb[8 * TILE] = -dt * s_r;                                            // J2[1][2]
          jb[9 * TILE] = (dt * s_r) / c_p;                                     // J2[2][1]
          jb[10 * TILE] = (dt * c_r) / c_p;                                    // J2[2][2]
          // f(x) for the Euler angles: rpy += dt * EarBaseInv(rpy) * w (angular_velocities.cpp:126-140, geometry.hpp:359-374)
          const double E01 = (s_p * s_r) / c_p, E02 = (c_r * s_p) / c_p, E11 = c_r, E12 = -s_r, E21 = s_r / c_p, E22 = c_r / c_p;
          jb[11 * TILE] = x3 + ((dt * 1.0) * wx + (dt * E01) * wy + (dt * E02) * wz);
          jb[12 * TILE] = x4 + ((dt * 0.0) * wx + (dt * E11) * wy + (dt * E12) * wz);
          jb[13 * TILE] = x5 + ((dt * 0.0) * wx + (dt * E21) * wy + (dt * E22) * wz);
        }
      }
      stage_bar_arrive<BAR_Y, NMAIN + NT * 32>(s);   // y / J of this tile are in shared memory
      if (tw == 1) retire(it - (STAGES - 1));   // write back the tile the main warps finish next, refill its stage
    }
    if (tw == 1) {
      for (int jd = n_my - (STAGES - 1); jd < n_my; ++jd) retire(jd);
      if (producer) bulk_wait<0>();
    }
    return;
  }

  // ============================= main warps =============================

  for (int it = 0; it < n_my; ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
    const int tile = tile_of(it);
    const int slot = tile * TILE + lane;
    const bool valid = slot < a.n_slots;
    const int act = act_n, cls = cls_n, dst = dst_n;
    const double dt = dt_n;
    load_ctrl(it + 1);
    // compacting tick: this warp moves the entries it owns (its rows; warp 0 also t / n_meas / prev_rpy) from the stage to
    // the target's destination slot -- entries it wrote itself or that nobody writes, so no barrier is needed
    auto compact_out = [&](const double* stg) {
      if (dst < 0) return;
      double* dr = a.dst_tiles + (size_t)(dst / TILE) * LY::TILE_DOUBLES + (dst % TILE);
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int g = q * RS + r;
        if (h == 0) __stcs(dr + (size_t)(LY::F_X + g) * TILE, stg[(LY::F_X + g) * TILE + lane]);
#pragma unroll
        for (int j = 0; j < NCOL; ++j)
          if (!packed || colof(j) >= g) __stcs(dr + (size_t)(LY::F_P + g * N + colof(j)) * TILE, stg[(LY::F_P + g * N + colof(j)) * TILE + lane]);
      }
      if (w == 0) {
#pragma unroll
        for (int f = LY::F_P + N * N; f < LY::NF; ++f) __stcs(dr + (size_t)f * TILE, stg[f * TILE + lane]);
      }
    };
    const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
    const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;
    // warp-uniform: every lane of the tile uses the class whose Q / R sit in the parameter constant bank
    // (only where the L1 is starved: with one CTA per SM the L1 keeps the tables and indexed LDCs were slower, AR 0.90 -> 0.69)
    const bool ctab = (MIN_CTAS > 1) && __all_sync(0xffffffffu, cls == a.cls_c);
    auto Qv = [&](int idx) -> double { return ctab ? a.Qc[idx] : __ldg(&Q[idx]); };
    auto Rv = [&](int idx) -> double { return ctab ? a.Rc[idx] : __ldg(&R[idx]); };

    TE_MARK(0);
    TE_MARK(1);
    double* st = stage0 + (size_t)s * STAGE_DOUBLES;
    double* ybuf = ybuf0 + (size_t)s * 6 * TILE;
    double* jbuf = jbuf0 + (size_t)s * JBUF_FIELDS * TILE;
    mbar_wait(&bars[s], parity);
    stage_bar_sync<BAR_Y, NMAIN + NT * 32>(s);   // the converters finished this tile (normally long ago)
    TE_MARK(2);

    // lane = target in every warp, so a warp ballot already is the tile-wide answer
    const unsigned any = __ballot_sync(0xffffffffu, act != ACT_NONE);
    if (any) {
      const bool active = act != ACT_NONE;
      const bool upd = act == ACT_UPDATE;

      double Pr[RPT][NCOL];
      double xr[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int g = q * RS + r;
        xr[q] = st[(LY::F_X + g) * TILE + lane];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
          const int col = colof(j);
          const int e = (packed && col < g) ? col * N + g : g * N + col;   // packed: lower-triangle entries from their mirror
          Pr[q][j] = st[(LY::F_P + e) * TILE + lane];
        }
      }

      // ---- phase A: predict of the own block ----
      if (active && !(TE_SKIP & 4)) {
        if (TYPE == ANGULAR_RATES) {
          const double hh = 0.5 * dt * dt;
          xr[0] = xr[0] + dt * xr[1] + hh * xr[2];
          xr[1] = xr[1] + dt * xr[2];
#pragma unroll
          for (int j = 0; j < NCOL; ++j) {   // A P on the three rows of this thread (one per kinematic block)
            Pr[0][j] = Pr[0][j] + dt * Pr[1][j] + hh * Pr[2][j];
            Pr[1][j] = Pr[1][j] + dt * Pr[2][j];
          }
#pragma unroll
          for (int q = 0; q < RPT; ++q) {   // (A P) A^T + Q within each row: columns c, 6+c, 12+c
#pragma unroll
            for (int c = 0; c < BL; ++c) {
              Pr[q][c] = Pr[q][c] + dt * Pr[q][BL + c] + hh * Pr[q][2 * BL + c];
              Pr[q][BL + c] = Pr[q][BL + c] + dt * Pr[q][2 * BL + c];
            }
#pragma unroll
            for (int j = 0; j < NCOL; ++j) Pr[q][j] = Pr[q][j] + Qv((q * RS + r) * N + colof(j));
          }
        } else {
          // EKF: thread r < 3 owns rows (p_r, v_r); thread r >= 3 owns rows (rpy_i, w_i), i = r - 3
          const double* jb = jbuf + lane;
          if (r < 3) {
            xr[0] = xr[0] + dt * xr[1];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) Pr[0][j] = Pr[0][j] + dt * Pr[1][j];
          } else {
            const int i = r - 3;
            // rows i of J_rpy, J_w, E^-1 (structural 0 / 1 / dt entries are not stored)
            double j1r[3], j2r[3];
            j1r[0] = i == 0 ? jb[0 * TILE] : (i == 1 ? jb[2 * TILE] : jb[3 * TILE]);
            j1r[1] = i == 0 ? jb[1 * TILE] : (i == 1 ? 1.0 : jb[4 * TILE]);
            j1r[2] = i == 2 ? 1.0 : 0.0;
            j2r[0] = i == 0 ? dt : 0.0;
            j2r[1] = i == 0 ? jb[5 * TILE] : (i == 1 ? jb[7 * TILE] : jb[9 * TILE]);
            j2r[2] = i == 0 ? jb[6 * TILE] : (i == 1 ? jb[8 * TILE] : jb[10 * TILE]);
            xr[0] = jb[(11 + i) * TILE];   // predicted Euler angle, evaluated by the converter warp
            // row 3+i of A P = J1[i,:] rows 3..5 + J2[i,:] rows 9..11 (original rows, from the stage)
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
              double p1[3], p3[3];
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                p1[k] = st[(LY::F_P + (3 + k) * N + colof(j)) * TILE + lane];
                p3[k] = st[(LY::F_P + (9 + k) * N + colof(j)) * TILE + lane];
              }
              Pr[0][j] = j1r[0] * p1[0] + j1r[1] * p1[1] + j1r[2] * p1[2] + j2r[0] * p3[0] + j2r[1] * p3[1] + j2r[2] * p3[2];
            }
          }
          // (A P) A^T + Q within each row; linear columns {0..2, 6..8} and angular columns {3..5, 9..11} are closed sets
          const bool lin = (CS == 1) || h == 0, angc = (CS == 1) || h == 1;
          double J1[3][3], J2[3][3];
          if (angc) {
            J1[0][0] = jb[0 * TILE]; J1[0][1] = jb[1 * TILE]; J1[0][2] = 0;
            J1[1][0] = jb[2 * TILE]; J1[1][1] = 1;            J1[1][2] = 0;
            J1[2][0] = jb[3 * TILE]; J1[2][1] = jb[4 * TILE]; J1[2][2] = 1;
            J2[0][0] = dt; J2[0][1] = jb[5 * TILE]; J2[0][2] = jb[6 * TILE];
            J2[1][0] = 0;  J2[1][1] = jb[7 * TILE]; J2[1][2] = jb[8 * TILE];
            J2[2][0] = 0;  J2[2][1] = jb[9 * TILE]; J2[2][2] = jb[10 * TILE];
          }
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            const int g = q * RS + r;
            if (CS == 1) {
              double rw[N];
#pragma unroll
              for (int j = 0; j < N; ++j) rw[j] = Pr[q][j];
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                Pr[q][j] = (rw[j] + rw[6 + j] * dt) + Qv(g * N + j);
                const double sacc = rw[3] * J1[j][0] + rw[4] * J1[j][1] + rw[5] * J1[j][2] + rw[9] * J2[j][0] + rw[10] * J2[j][1] + rw[11] * J2[j][2];
                Pr[q][3 + j] = sacc + Qv(g * N + 3 + j);
                Pr[q][6 + j] = rw[6 + j] + Qv(g * N + 6 + j);
                Pr[q][9 + j] = rw[9 + j] + Qv(g * N + 9 + j);
              }
            } else if (lin) {   // local 0..2 = cols 0..2, local 3..5 = cols 6..8
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                Pr[q][j] = (Pr[q][j] + Pr[q][3 + j] * dt) + Qv(g * N + j);
                Pr[q][3 + j] = Pr[q][3 + j] + Qv(g * N + 6 + j);
              }
            } else {            // local 0..2 = cols 3..5, local 3..5 = cols 9..11
              double rw[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) rw[j] = Pr[q][j];
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const double sacc = rw[0] * J1[j][0] + rw[1] * J1[j][1] + rw[2] * J1[j][2] + rw[3] * J2[j][0] + rw[4] * J2[j][1] + rw[5] * J2[j][2];
                Pr[q][j] = sacc + Qv(g * N + 3 + j);
                Pr[q][3 + j] = rw[3 + j] + Qv(g * N + 9 + j);
              }
            }
          }
        }
      }
      if (TYPE == ANGULAR_VELOCITIES && r >= 3) {
        // the r >= 3 warps have read the original rows 3..5 / 9..11; nobody else does.  Order those reads before the
        // in-place publish of rows 3..5 (and of P'[9..11, 0:6] when CS = 2) with a named barrier among them.
        asm volatile("bar.sync 1, %0;" ::"n"(3 * CS * 32) : "memory");
      }

      // publish in place: predicted measured row r and x'[r]; with CS = 2 also P'[own rows q >= 1, own part of 0:6]
      if (upd) {
        if (h == 0) st[(LY::F_X + r) * TILE + lane] = xr[0];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) st[(LY::F_P + r * N + colof(j)) * TILE + lane] = Pr[0][j];
        if (CS == 2) {
#pragma unroll
          for (int q = 1; q < RPT; ++q)
#pragma unroll
            for (int c = 0; c < BL; ++c) st[(LY::F_P + (q * RS + r) * N + BL * h + c) * TILE + lane] = Pr[q][c];
        }
      }
      TE_MARK(3);
      named_bar_sync<BAR_MAIN, NMAIN>();
      TE_MARK(4);
      TE_MARK(5);

      // ---- phase B: warp 0 factors S = P'[0:M,0:M] + R and solves v = S^-1 (y - x'[0:M]); warps 1..: factor S, columns
      //      of W -> Wbuf ----
      double Pk[RPT][M];   // P'[own rows, 0:M]: own part from registers, the partner's part from the stage
      if (upd && CS == 2) {   // (CS = 1: copied at the start of phase C, keeps 2*RPT*M registers free during the solves)
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int k = 0; k < M; ++k)
            Pk[q][k] = (k / BL == h) ? Pr[q][k % BL] : st[(LY::F_P + (q * RS + r) * N + k) * TILE + lane];
      }
      if (upd && !(TE_SKIP & 1)) {
        // one copy of the factorisation for every warp (code size: the kernel is I-cache bound enough as it is)
        Chol<M> ch;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int j = 0; j < M; ++j)
            if (j <= i) ch.at(i, j) = st[(LY::F_P + i * N + j) * TILE + lane] + Rv(i * M + j);
        ch.factor();
        TE_MARK(6);
        if (w == 0) {   // innovation: v = S^-1 (y - x'[0:M])
          double col[M];
#pragma unroll
          for (int k = 0; k < M; ++k) col[k] = ybuf[k * TILE + lane];
#pragma unroll
          for (int k = 0; k < 3; ++k) st[(LY::F_PREV + k) * TILE + lane] = col[3 + k];   // meas_rpy_internal_ = unwrapped rpy
#pragma unroll
          for (int k = 0; k < M; ++k) col[k] = col[k] - st[(LY::F_X + k) * TILE + lane];
          ch.solve(col);
#pragma unroll
          for (int k = 0; k < M; ++k) ybuf[k * TILE + lane] = col[k];   // y -> v in place (only this lane reads y[.][lane])
        } else if (MIN_CTAS > 1) {   // register-lean: one column at a time
#pragma unroll 1
          for (int cc = 0; cc < WPW; ++cc) {
            const int c = (w - 1) + cc * (NW - 1);
            if (c >= N) break;
            double col[M];
#pragma unroll
            for (int k = 0; k < M; ++k) col[k] = st[(LY::F_P + k * N + c) * TILE + lane];
            ch.solve(col);
#pragma unroll
            for (int k = 0; k < M; ++k) Wbuf[(k * N + c) * TILE + lane] = col[k];
          }
        } else {        // this warp's columns of W = S^-1 P'[0:M,:], solved together (independent chains interleave)
          double col[WPW][M];
#pragma unroll
          for (int cc = 0; cc < WPW; ++cc) {
            const int c = (w - 1) + cc * (NW - 1);
#pragma unroll
            for (int k = 0; k < M; ++k) col[cc][k] = (c < N) ? st[(LY::F_P + k * N + c) * TILE + lane] : 0.0;
          }
#pragma unroll
          for (int cc = 0; cc < WPW; ++cc) ch.solve(col[cc]);
#pragma unroll
          for (int cc = 0; cc < WPW; ++cc) {
            const int c = (w - 1) + cc * (NW - 1);
            if (c < N) {
#pragma unroll
              for (int k = 0; k < M; ++k) Wbuf[(k * N + c) * TILE + lane] = col[cc][k];
            }
          }
        }
      }
      TE_MARK(8);
      named_bar_sync<BAR_MAIN, NMAIN>();
      TE_MARK(9);

      // ---- phase C: own block: P[rows, cols] -= P'[rows,0:M] W[:, cols]  ((I - K C) P, src/kalman.cpp:94) ----
      if (upd && !(TE_SKIP & 2)) {
        if (CS == 1) {
#pragma unroll
          for (int q = 0; q < RPT; ++q)
#pragma unroll
            for (int k = 0; k < M; ++k) Pk[q][k] = Pr[q][k];
        }
        if (h == 0) {   // x += P'[rows,0:M] v  (K (y - C x'), src/kalman.cpp:93)
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) sacc += Pk[q][k] * ybuf[k * TILE + lane];
            xr[q] += sacc;
          }
        }
        if (MIN_CTAS == 1) {
          // rank-1 updates, k outermost: every step touches all RPT x NCOL owned elements independently (full ILP), the
          // dependent chain per element is the same k = 0..M-1 order as the reference's inner product
#pragma unroll
          for (int k = 0; k < M; ++k) {
            double wrow[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) wrow[j] = Wbuf[(k * N + colof(j)) * TILE + lane];
#pragma unroll
            for (int q = 0; q < RPT; ++q)
#pragma unroll
              for (int j = 0; j < NCOL; ++j) Pr[q][j] -= Pk[q][k] * wrow[j];
          }
        } else {
          // register-lean form for the 128-register budget of two CTAs per SM: one column at a time, two chains of three
#pragma unroll
          for (int jj = 0; jj < NCOL; ++jj) {
            const int j = NCOL - 1 - jj;
            double wv[M];
#pragma unroll
            for (int k = 0; k < M; ++k) wv[k] = Wbuf[(k * N + colof(j)) * TILE + lane];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
              double s0 = Pr[q][j], s1 = 0.0;
#pragma unroll
              for (int k = 0; k < M; k += 2) {
                s0 -= Pk[q][k] * wv[k];
                s1 -= Pk[q][k + 1] * wv[k + 1];
              }
              Pr[q][j] = s0 + s1;
            }
          }
        }
      }

      // ---- phase D: own block back into the stage (nobody reads the published entries after phase B) ----
      // (compacting tick: the stage is never stored, so a stepped target's rows go from the registers straight to its destination
      // slot below instead of through shared memory)
      if (active) {
        if (!COMPACT) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            const int g = q * RS + r;
            if (h == 0) st[(LY::F_X + g) * TILE + lane] = xr[q];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) st[(LY::F_P + g * N + colof(j)) * TILE + lane] = Pr[q][j];
          }
        }
        if (w == 0) {
          st[LY::F_T * TILE + lane] = st[LY::F_T * TILE + lane] + dt;   // updateTime
          if (upd) {
            long long* nm = reinterpret_cast<long long*>(st + LY::F_NMEAS * TILE + lane);
            *nm = *nm + 1;                                             // updateMeasurement
          }
          if (a.clear_action) a.action[slot] = 0;
        }
      }
      if (a.pos_out && valid && w < 3) a.pos_out[(size_t)slot * 3 + w] = xr[0];
      if (COMPACT && dst >= 0) {
        double* dr = a.dst_tiles + (size_t)(dst / TILE) * LY::TILE_DOUBLES + (dst % TILE);
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int g = q * RS + r;
          if (h == 0) __stcs(dr + (size_t)(LY::F_X + g) * TILE, active ? xr[q] : st[(LY::F_X + g) * TILE + lane]);
#pragma unroll
          for (int j = 0; j < NCOL; ++j)
            if (!packed || colof(j) >= g)
              __stcs(dr + (size_t)(LY::F_P + g * N + colof(j)) * TILE, active ? Pr[q][j] : st[(LY::F_P + g * N + colof(j)) * TILE + lane]);
        }
        if (w == 0) {   // t, n_meas (updated in the stage above), previous unwrapped angles
#pragma unroll
          for (int f = LY::F_P + N * N; f < LY::NF; ++f) __stcs(dr + (size_t)f * TILE, st[f * TILE + lane]);
        }
      }
      fence_proxy_async();   // generic-proxy writes of the stage -> visible to the producer's bulk store
      TE_MARK(10);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[STAGES + s]);   // no wait: the next tile lives in another stage, and W reuse is ordered
                                                       // by the next tile's first main barrier
    } else {
      if (a.pos_out && valid && w < 3) a.pos_out[(size_t)slot * 3 + w] = st[(LY::F_X + w) * TILE + lane];
      if (COMPACT) compact_out(st);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[STAGES + s]);
    }
  }
}


}  // namespace te
