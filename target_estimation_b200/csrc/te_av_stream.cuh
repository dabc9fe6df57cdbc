// te_av_stream.cuh -- dense in-place tick of the angular-velocities EKF with the tile loads on the TMA engine.
//
// The direct kernel (te_direct.cuh) issues ~100 LDG and ~95 STG per lane and tile from the very warps that carry the dependent
// FP64 chains of the step; with the register file full at eight warps per SM nothing hides them (profiles/r2_av_source_stalls.txt:
// 15 % of the warp time goes into issuing the loads, 15 % into waiting for them, 16 % into the stores, 7 % into the flag loads of
// the next tile, 15 % into the scalar chains of the measurement conversion and the Jacobians).  Here every warp owns a LANDING ZONE
// in shared memory: one lane moves the warp's tile into it with 1-D bulk copies (cp.async.bulk on one mbarrier, SASS UBLKCP) --
//
//   x | upper triangle of P, row-packed | t | n_meas | prev_rpy   (95 fields x 256 B, 14 copies: the fields the step reads)
//   the tile's measurement block [32][7], its action bytes, its class ids   (one copy each)
//
// -- the lanes pick their column up with LDS, and AT ONCE the zone takes the warp's NEXT tile: its copies travel under the whole
// step of this one.  That needs a step that uses no shared-memory scratch, so the update runs as six scalar updates of the whitened
// measurement (av_update_seq in te_av_sym.cuh: registers only; the joint form parks a 6 x 12 intermediate), and the results go
// from the registers straight to HBM.  Six fields beside the zone park the converted measurement of the tile being stepped
// (registers are the scarce resource during the covariance predict).
//
// A first form (round 2, 0.379 -> 0.365 ms per tick at 1 Mi targets against 0.443 of the direct kernel) kept the joint update with
// its Z scratch in the zone and wrote the results out of the zone with bulk stores; the zone was then busy from the load to the
// drain of the stores, and the load of the next tile -- 11 % of the warp time -- the store issue and the drain stayed on every
// warp's critical path.  This form: 0.315 ms.
//
// Predict and Jacobians are the direct kernel's (the shared pieces of te_av_sym.cuh).  Serves dense in-place packed ticks of one tick
// (no tile list, no per-slot dt, 16-byte aligned measurements of stride 7) of pools whose classes all have a
// Cholesky factor of R, in place or compacting (StepArgs::dst_*: every survivor's column goes to its compacted slot -- the results
// leave from registers by per-lane stores anyway); everything else stays on the direct kernel.  In place, lanes that are not stepped
// (ACT_NONE, slots beyond the pool) write nothing.
#pragma once
#include "te_av_sym.cuh"
#include "te_kernels.cuh"

namespace te {

constexpr int AVS_X = 0, AVS_P = 12, AVS_T = 90, AVS_NM = 91, AVS_PREV = 92, AVS_FIELDS = 95;   // field numbering of the zone
constexpr int AVS_TAIL_BYTES = 384;   // per warp, behind the zones: the mbarrier (16 B) | the tile's action bytes (32 B) | its class ids (64 B) |
                                      // 16 B unused | compacting tick: the tile's survivor flags (128 B) | their destination slots (128 B)
constexpr int AVL_PARK_FIELDS = 6;                                              // the measurement y of the tile being stepped
constexpr int AVL_WARP_BYTES = (AVS_FIELDS + 7 + AVL_PARK_FIELDS) * TILE * 8;   // landed fields + measurement block [32][7] + parking area
__host__ __device__ constexpr size_t av_stream_smem_bytes(int warps) { return (size_t)warps * (AVL_WARP_BYTES + AVS_TAIL_BYTES); }

// QC: the pool has one class and its Q / T ride in the kernel-parameter constant bank (StepArgs::Qc / Tc): the table reads
// become constant operands.  COMPACT: the compacting tick, an instantiation of its own (its extra registers cost the in-place
// tick 2 %).
template <int WARPS, bool QC, bool COMPACT>
__global__ void __launch_bounds__(WARPS * 32, 1) kf_step_av_stream_kernel(const __grid_constant__ StepArgs a) {
  using LY = Layout<ANGULAR_VELOCITIES>;
  constexpr int N = 12, M = 6;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // (the warp index through a lane-0 broadcast: the compiler then knows that everything derived from it -- tile, zone and barrier
  //  addresses -- is warp-uniform and issues each bulk copy once; from threadIdx alone it wrapped every UBLKCP in a loop over the
  //  distinct values of its operands: R2UR + vote + branch, ~150 cycles per copy)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  double* const Sw = reinterpret_cast<double*>(smem_raw + (size_t)warp * AVL_WARP_BYTES);
  const double* const S = Sw + lane;                             // the lane's column of the landed tile
  const double* const Mz = Sw + AVS_FIELDS * TILE;                // measurement block [lane][7]
  double* const Yp = Sw + (AVS_FIELDS + 7) * TILE + lane;         // the lane's column of the parking area
  unsigned char* const tail = smem_raw + (size_t)WARPS * AVL_WARP_BYTES + (size_t)warp * AVS_TAIL_BYTES;
  uint64_t* const bar = reinterpret_cast<uint64_t*>(tail);
  const uint8_t* const Az = tail + 16;
  const uint16_t* const Cz = reinterpret_cast<const uint16_t*>(tail + 48);
  const int* const Dalive = reinterpret_cast<const int*>(tail + 128);
  const int* const Dpos = reinterpret_cast<const int*>(tail + 256);
  constexpr bool compacting = COMPACT;   // te_pool_step_dense_expire / the fused mailbox tick: survivors go to their compacted slots (StepArgs::dst_*)
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const int n_work = a.n_tiles;
  const int GW = gridDim.x * WARPS;
  int w = blockIdx.x * WARPS + warp;
  // A full tile takes its measurement block, action bytes and class ids through the zone as well (a flag loaded into a register a
  // tile ahead was spilled at once -- 255 registers -- and the spill store waited for the load: a memory latency per tile); the
  // last, partial tile of a pool and unaligned caller arrays read them per lane.
  const bool act_tma = a.action != nullptr && (reinterpret_cast<uintptr_t>(a.action) & 15) == 0;
  auto full_tile = [&](int tile) -> bool { return (tile + 1) * TILE <= a.n_slots; };
  auto issue = [&](int tile) {   // lane 0: everything the step of `tile` reads
    const double* tb = a.tiles + (size_t)tile * LY::TILE_DOUBLES;
    const bool full = full_tile(tile);
    const bool mt = full && a.meas != nullptr, at = full && act_tma, ct = full && !QC;
    const bool dt_ = full && compacting;
    mbar_expect_tx(bar, AVS_FIELDS * TILE * 8 + (mt ? 7 * TILE * 8 : 0) + (at ? TILE : 0) + (ct ? 2 * TILE : 0) + (dt_ ? 8 * TILE : 0));
    bulk_g2s(Sw + AVS_X * TILE, tb + LY::F_X * TILE, N * TILE * 8, bar);
    int pk = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      bulk_g2s(Sw + (AVS_P + pk) * TILE, tb + (LY::F_P + i * N + i) * TILE, (N - i) * TILE * 8, bar);
      pk += N - i;
    }
    bulk_g2s(Sw + AVS_T * TILE, tb + LY::F_T * TILE, 5 * TILE * 8, bar);
    if (mt) bulk_g2s(Sw + AVS_FIELDS * TILE, a.meas + (size_t)tile * TILE * 7, 7 * TILE * 8, bar);
    if (at) bulk_g2s(tail + 16, a.action + (size_t)tile * TILE, TILE, bar);
    if (ct) bulk_g2s(tail + 48, a.cls + (size_t)tile * TILE, 2 * TILE, bar);
    if (dt_) {
      bulk_g2s(tail + 128, a.dst_alive + (size_t)tile * TILE, 4 * TILE, bar);
      bulk_g2s(tail + 256, a.dst_pos + (size_t)tile * TILE, 4 * TILE, bar);
    }
  };
  if (w < n_work && lane == 0) issue(a.tile_begin + w);
  const double dt = a.dt;
  for (uint32_t it = 0; w < n_work; w += GW, ++it) {
    const int tile = a.tile_begin + w;
    const int slot = tile * TILE + lane;
    const bool valid = slot < a.n_slots;
    const bool full = full_tile(tile);
    mbar_wait(bar, it & 1);
    int act = ACT_NONE, cls = 0, dst = -1;   // dst: compacting tick, destination slot of the lane's target (-1 = erased at the end of this tick)
    if (valid) {
      act = a.action ? (int)((full && act_tma) ? Az[lane] : a.action[slot]) : a.default_action;
      if (!QC) cls = (int)(full ? Cz[lane] : a.cls[slot]);
      if (compacting) {
        const int alive = full ? Dalive[lane] : a.dst_alive[slot];
        if (alive) dst = full ? Dpos[lane] : a.dst_pos[slot];
        else act = ACT_NONE;   // its step is unobservable
      }
    }
    // ---- front end while few registers are live: measurement conversion (angular_velocities.cpp:87-96), Jacobians ----
    double prev[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) prev[k] = S[(AVS_PREV + k) * TILE];
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
      double un[3];
      meas_to_unwrapped_rpy(meas + 3, prev, un);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Yp[k * TILE] = meas[k];
        Yp[(3 + k) * TILE] = un[k];
        prev[k] = un[k];   // the new meas_rpy_internal_
      }
    }
    AvFront F;
    if (act != ACT_NONE) F.eval(S[(AVS_X + 3) * TILE], S[(AVS_X + 4) * TILE], S[(AVS_X + 9) * TILE], S[(AVS_X + 10) * TILE], S[(AVS_X + 11) * TILE], dt);
    // ---- the registers take the tile; the zone takes the warp's next one ----
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = S[(AVS_X + i) * TILE];
    SymP<N> P;
#pragma unroll
    for (int k = 0; k < N * (N + 1) / 2; ++k) P.v[k] = S[(AVS_P + k) * TILE];
    const double t_in = S[AVS_T * TILE];
    const long long nm_in = reinterpret_cast<const long long*>(S)[AVS_NM * TILE];
    __syncwarp();   // every lane has its column
    const int wn = w + GW;
    if (lane == 0 && wn < n_work) issue(a.tile_begin + wn);

    double* const out = compacting ? a.dst_tiles + (size_t)((dst < 0 ? 0 : dst) / TILE) * LY::TILE_DOUBLES + ((dst < 0 ? 0 : dst) % TILE)
                                   : a.tiles + (size_t)tile * LY::TILE_DOUBLES + lane;
    if (compacting && dst >= 0 && act != ACT_UPDATE) {
      // the fields a stepped lane does not rewrite travel with the target: everything for an untouched survivor, the measurement
      // count and the previous angles for a predicted one
      if (act == ACT_NONE) out[LY::F_T * TILE] = t_in;
      reinterpret_cast<long long*>(out)[LY::F_NMEAS * TILE] = nm_in;
#pragma unroll
      for (int k = 0; k < 3; ++k) out[(LY::F_PREV + k) * TILE] = prev[k];
    }
    if (act != ACT_NONE) {
      // bookkeeping first (its values are final, its registers are wanted): updateTime (src/target_interface.cpp:148-152),
      // updateMeasurement (:142-146), meas_rpy_internal_
      out[LY::F_T * TILE] = t_in + dt;
      if (act == ACT_UPDATE) {
        reinterpret_cast<long long*>(out)[LY::F_NMEAS * TILE] = nm_in + 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) out[(LY::F_PREV + k) * TILE] = prev[k];
      }
      // f(x): p += dt v ; rpy += dt * EarBaseInv(rpy) * w
#pragma unroll
      for (int i = 0; i < 3; ++i) x[i] = x[i] + dt * x[6 + i];
      x[3] = x[3] + F.d3;
      x[4] = x[4] + F.d4;
      x[5] = x[5] + F.d5;
      const double* __restrict__ Q = a.Qtab + (size_t)cls * N * N;
      const double* __restrict__ T = a.Ttab + (size_t)cls * M * M;
      av_predict_cov(P, F, [&](int idx) -> double { return QC ? a.Qc[idx] : __ldg(&Q[idx]); });
      if (act == ACT_UPDATE) av_update_seq<TILE>(P, x, Yp, [&](int idx) -> double { return QC ? a.Tc[idx] : __ldg(&T[idx]); });
    }
    if (act != ACT_NONE || (compacting && dst >= 0)) {   // (compacting: an untouched survivor moves as it was loaded)
#pragma unroll
      for (int i = 0; i < N; ++i) out[(LY::F_X + i) * TILE] = x[i];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (i <= j) out[(LY::F_P + i * N + j) * TILE] = P(i, j);   // packed: the pool mirrors the upper triangle on demand
    }
    if (a.pos_out && valid) {
#pragma unroll
      for (int k = 0; k < 3; ++k) a.pos_out[(size_t)slot * 3 + k] = x[k];
    }
  }
}

}  // namespace te
