// te_av_stream.cuh -- dense in-place tick of the angular-velocities EKF with the tile traffic on the TMA engine.
//
// The direct kernel (te_direct.cuh) issues ~100 LDG and ~95 STG per lane and tile from the very warps that carry the dependent
// FP64 chains of the step; with the register file full at eight warps per SM nothing hides them (profiles/r2_av_source_stalls.txt:
// 15 % of the warp time goes into issuing the loads, 15 % into waiting for them, 16 % into the stores, 7 % into the flag loads of
// the next tile, 15 % into the scalar chains of the measurement conversion and the Jacobians).  Here one lane per warp moves the
// warp's tile with 1-D bulk copies (cp.async.bulk, SASS UBLKCP / UBLKRED-free plain bulk stores):
//
//   main zone  (95 fields x 256 B per warp): x | upper triangle of P, row-packed | t | n_meas | prev_rpy -- 14 bulk loads land the
//              fields the step reads, the lanes pick their column up with LDS; during the update the P fields hold Z (P lives in
//              registers then); at the end the lanes put their results back with STS and 14 bulk stores write the zone out while
//              the warp goes on to its next tile (the next loads wait only until the stores have READ the zone).
//   front zone (8 fields + the tile's measurement block [32][7]): roll, pitch, body rates, previous unwrapped angles and the
//              measured poses of the warp's NEXT tile, fetched one tile ahead.  The measurement conversion (quat -> rpy -> unwrap)
//              and the trigonometry of the Jacobians -- the scalar chains of the step -- depend on nothing else, so they run while
//              the main zone of the tile is still in flight.
//
// Same arithmetic as the direct kernel (the shared pieces of te_av_sym.cuh).  Serves dense in-place packed ticks of one tick
// (no tile list, no per-slot dt, no compaction, 16-byte aligned measurements of stride 7); everything else stays on the direct kernel.
// Lanes that are not stepped (ACT_NONE, slots beyond the pool) keep the column that was loaded.
#pragma once
#include "te_av_sym.cuh"
#include "te_kernels.cuh"

namespace te {

constexpr int AVS_X = 0, AVS_P = 12, AVS_T = 90, AVS_NM = 91, AVS_PREV = 92, AVS_FIELDS = 95;
constexpr int AVF_RP = 0, AVF_W = 2, AVF_PREV = 5, AVF_FIELDS = 8;
constexpr int AVS_WARP_DOUBLES = (AVS_FIELDS + AVF_FIELDS + 7) * TILE;   // main zone | front zone | measurement block
constexpr int AVS_TAIL_BYTES = 128;   // per warp, behind the zones: two mbarriers (16 B) | the tile's action bytes (32 B) | its class ids (64 B)
__host__ __device__ constexpr size_t av_stream_smem_bytes(int warps) { return (size_t)warps * (AVS_WARP_DOUBLES * 8 + AVS_TAIL_BYTES); }

// QC: the pool has one class and its Q / R ride in the kernel-parameter constant bank (StepArgs::Qc / Rc): the 99 table
// reads per lane become constant operands
template <int WARPS, bool QC>
__global__ void __launch_bounds__(WARPS * 32, 1) kf_step_av_stream_kernel(const __grid_constant__ StepArgs a) {
  using LY = Layout<ANGULAR_VELOCITIES>;
  constexpr int N = 12, M = 6;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // (the warp index through a lane-0 broadcast: the compiler then knows that everything derived from it -- tile, zone and barrier
  //  addresses -- is warp-uniform and issues each bulk copy once; from threadIdx alone it wrapped every UBLKCP in a loop over the
  //  distinct values of its operands: R2UR + vote + branch, ~150 cycles per copy, 28 copies per tile)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  double* const Sw = reinterpret_cast<double*>(smem_raw) + (size_t)warp * AVS_WARP_DOUBLES;
  double* const S = Sw + lane;                                   // the lane's column of the main zone
  double* const Fz = Sw + AVS_FIELDS * TILE + lane;              // ... of the front zone
  double* const Mz = Sw + (AVS_FIELDS + AVF_FIELDS) * TILE;      // measurement block [lane][7]
  unsigned char* const tail = smem_raw + (size_t)WARPS * AVS_WARP_DOUBLES * 8 + (size_t)warp * AVS_TAIL_BYTES;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(tail);                  // [0] main, [1] front
  uint8_t* const Az = tail + 16;                                             // action bytes of the front zone's tile
  uint16_t* const Cz = reinterpret_cast<uint16_t*>(tail + 48);               // class ids
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int n_work = a.n_tiles;
  const int GW = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + warp;

  // A full tile takes its measurement block, action bytes and class ids through the front zone as well (a flag loaded into a
  // register a tile ahead was spilled at once -- 255 registers -- and the spill store waited for the load: a memory latency per
  // tile); the last, partial tile of a pool and unaligned caller arrays read them per lane.
  const bool act_tma = a.action != nullptr && (reinterpret_cast<uintptr_t>(a.action) & 15) == 0;
  auto full_tile = [&](int tile) -> bool { return (tile + 1) * TILE <= a.n_slots; };
  auto issue_front = [&](int tile) {   // lane 0
    const double* tb = a.tiles + (size_t)tile * LY::TILE_DOUBLES;
    const bool full = full_tile(tile);
    const bool mt = full && a.meas != nullptr, at = full && act_tma, ct = full && !QC;
    mbar_expect_tx(&bars[1], AVF_FIELDS * TILE * 8 + (mt ? 7 * TILE * 8 : 0) + (at ? TILE : 0) + (ct ? 2 * TILE : 0));
    bulk_g2s(Sw + (AVS_FIELDS + AVF_RP) * TILE, tb + (LY::F_X + 3) * TILE, 2 * TILE * 8, &bars[1]);
    bulk_g2s(Sw + (AVS_FIELDS + AVF_W) * TILE, tb + (LY::F_X + 9) * TILE, 3 * TILE * 8, &bars[1]);
    bulk_g2s(Sw + (AVS_FIELDS + AVF_PREV) * TILE, tb + LY::F_PREV * TILE, 3 * TILE * 8, &bars[1]);
    if (mt) bulk_g2s(Mz, a.meas + (size_t)tile * TILE * 7, 7 * TILE * 8, &bars[1]);
    if (at) bulk_g2s(Az, a.action + (size_t)tile * TILE, TILE, &bars[1]);
    if (ct) bulk_g2s(Cz, a.cls + (size_t)tile * TILE, 2 * TILE, &bars[1]);
  };

  if (w < n_work && lane == 0) issue_front(a.tile_begin + w);
  const double dt = a.dt;
  for (uint32_t it = 0; w < n_work; w += GW, ++it) {
    const int tile = a.tile_begin + w;
    const int slot = tile * TILE + lane;
    const bool valid = slot < a.n_slots;
    const bool full = full_tile(tile);
    double* const tb = a.tiles + (size_t)tile * LY::TILE_DOUBLES;
    if (lane == 0) {
      bulk_wait_read<0>();   // the previous tile's stores have read the zone
      mbar_expect_tx(&bars[0], (AVS_PREV - AVS_X) * TILE * 8);
      bulk_g2s(Sw + AVS_X * TILE, tb + LY::F_X * TILE, N * TILE * 8, &bars[0]);
      int pk = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        bulk_g2s(Sw + (AVS_P + pk) * TILE, tb + (LY::F_P + i * N + i) * TILE, (N - i) * TILE * 8, &bars[0]);
        pk += N - i;
      }
      bulk_g2s(Sw + AVS_T * TILE, tb + LY::F_T * TILE, 2 * TILE * 8, &bars[0]);
    }
    __syncwarp();   // (nobody writes into the zone before lane 0 has seen the stores drained)
    const int wn = w + GW;

    // ---- front end, under the main loads: measurement conversion (angular_velocities.cpp:87-96) and Jacobians ----
    mbar_wait(&bars[1], it & 1);
    int act = ACT_NONE, cls = 0;
    if (valid) {
      act = a.action ? (int)((full && act_tma) ? Az[lane] : a.action[slot]) : a.default_action;
      if (!QC) cls = (int)(full ? Cz[lane] : a.cls[slot]);
    }
    double ypos[3] = {0.0, 0.0, 0.0};
    AvFront F;
    {
      double prev[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) prev[k] = Fz[(AVF_PREV + k) * TILE];
      if (act == ACT_UPDATE) {
        double meas[7];
        if (full) {
#pragma unroll
          for (int k = 0; k < 7; ++k) meas[k] = Mz[lane * 7 + k];
        } else {
          const double* mp = a.meas + (size_t)slot * 7;
#pragma unroll
          for (int k = 0; k < 7; ++k) meas[k] = __ldg(mp + k);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) ypos[k] = meas[k];
        double un[3];
        meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
        for (int k = 0; k < 3; ++k) prev[k] = un[k];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) S[(AVS_PREV + k) * TILE] = prev[k];   // y[3..5] of the update = the new meas_rpy_internal_
      if (act != ACT_NONE) F.eval(Fz[(AVF_RP + 0) * TILE], Fz[(AVF_RP + 1) * TILE], Fz[(AVF_W + 0) * TILE], Fz[(AVF_W + 1) * TILE], Fz[(AVF_W + 2) * TILE], dt);
    }
    mbar_wait(&bars[0], it & 1);
    __syncwarp();   // every lane has read the front zone: it takes the next tile's inputs from here on
    if (lane == 0 && wn < n_work) issue_front(a.tile_begin + wn);

    if (act != ACT_NONE) {
      SymP<N> P;
#pragma unroll
      for (int k = 0; k < N * (N + 1) / 2; ++k) P.v[k] = S[(AVS_P + k) * TILE];
      {
        double x[N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = S[(AVS_X + i) * TILE];
        // f(x): p += dt v ; rpy += dt * EarBaseInv(rpy) * w
#pragma unroll
        for (int i = 0; i < 3; ++i) S[(AVS_X + i) * TILE] = x[i] + dt * x[6 + i];
        S[(AVS_X + 3) * TILE] = x[3] + F.d3;
        S[(AVS_X + 4) * TILE] = x[4] + F.d4;
        S[(AVS_X + 5) * TILE] = x[5] + F.d5;
      }
      const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
      const double* __restrict__ R = a.Rtab + (size_t)cls * M * M;
      av_predict_cov(P, F, [&](int idx) -> double { return QC ? a.Qc[idx] : __ldg(&Q[idx]); });
      if (act == ACT_UPDATE)
        av_update<2>(P, S + AVS_X * TILE, S + AVS_P * TILE, ypos, S + AVS_PREV * TILE, [&](int idx) -> double { return QC ? a.Rc[idx] : __ldg(&R[idx]); });
#pragma unroll
      for (int k = 0; k < N * (N + 1) / 2; ++k) S[(AVS_P + k) * TILE] = P.v[k];
      // updateTime (src/target_interface.cpp:148-152) / updateMeasurement (:142-146)
      S[AVS_T * TILE] = S[AVS_T * TILE] + dt;
      reinterpret_cast<long long*>(S)[AVS_NM * TILE] += (act == ACT_UPDATE ? 1 : 0);
    }
    if (a.pos_out && valid) {
#pragma unroll
      for (int k = 0; k < 3; ++k) a.pos_out[(size_t)slot * 3 + k] = S[(AVS_X + k) * TILE];
    }
    fence_proxy_async();   // the lanes' writes to the zone, before the bulk stores read it
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(tb + LY::F_X * TILE, Sw + AVS_X * TILE, N * TILE * 8);
      int pk = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        bulk_s2g(tb + (LY::F_P + i * N + i) * TILE, Sw + (AVS_P + pk) * TILE, (N - i) * TILE * 8);
        pk += N - i;
      }
      bulk_s2g(tb + LY::F_T * TILE, Sw + AVS_T * TILE, 5 * TILE * 8);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait<0>();
}

}  // namespace te
