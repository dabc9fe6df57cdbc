"""-m gpu: oracle parity in the regime the bench times -- every warp / CTA of every default step kernel (and of its compacting
form) walks its persistent loop several times: the grid-stride loop of the direct kernels (te_direct.cuh) and the STAGES ring
of the warp-specialised split kernel (te_split.cuh: refill after the bulk-store drain, full[] / done[] mbarrier phase flips).

Two levers: pools of tens of thousands of targets (so that the launch picks the production shape: 8-warp CTAs, the whole
machine) and the grid cap test hook (te_pool_set_grid_cap / TE_GRID_CAP: a launch of at most k CTAs, so that each of them takes
many tiles).  The oracle side is orc.ShardedManager: the same single-threaded TargetManager arithmetic per target
(/root/reference/src/kalman.cpp:84-95 restated), the loop over independent targets spread over host threads.

Bar: 1e-9 relative on state and covariance (H2 norm, tests/synth.py), t / n_meas exact; every test also records the ratio under
SURVEY.md's strict 1e-6 floor (tests/report.py -> gpurun_out/parity_report.json)."""
import numpy as np
import pytest

from tests import orc, report, synth

pytestmark = pytest.mark.gpu

DT = 1.0 / 250.0
MODELS = ["uniform_velocity", "uniform_acceleration", "angular_velocities", "angular_rates"]
ACCEL = ("uniform_acceleration", "angular_rates")


def _setup(model, n, ticks, seed=0x7A26E7, variant=0, cap=0, id_stride=3, **stream_kw):
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(model)
    N, M = te.model_dims(mtype)
    meas, action, scale = synth.make_streams(n, ticks, DT, accel=model in ACCEL, angular=M == 6, seed=seed, **stream_kw)
    ids = (np.arange(n, dtype=np.uint32) * id_stride + 7)
    ref = orc.ShardedManager()
    ref.init_batch(mtype, ids, DT, Q, R, P0, meas[0], scale)
    pool = te.TargetPool(mtype)
    pool.set_variant(variant)
    pool.set_grid_cap(cap)
    assert pool.register_class(Q, R, P0) == 0
    assert pool.add(ids, meas[0], p0_scale=scale) == n
    _setup.scale = scale      # (for tests that build a second oracle on the same targets)
    return te, pool, ref, ids, meas, action, N, M


def _check(tag, pool, ref, ids, N, angular, worst):
    want = ref.states(ids, N)
    got = pool.read_state(ids)
    assert want["found"].all()
    x1, x2 = synth.compare_both(got["x"], want["x"])
    p1, p2 = synth.compare_both(got["P"], want["P"])
    worst["x"] = max(worst["x"], x1); worst["P"] = max(worst["P"], p1)
    worst["x_floor1e-6"] = max(worst["x_floor1e-6"], x2); worst["P_floor1e-6"] = max(worst["P_floor1e-6"], p2)
    assert np.array_equal(got["n_meas"], want["n_meas"]), tag
    assert np.array_equal(got["t"], want["t"]), tag
    if angular:
        assert synth.compare_h2(got["prev_rpy"], want["prev_rpy"]) <= 1.0, tag
    assert x1 <= 1.0 and p1 <= 1.0, (tag, worst)


def _worst():
    return {"x": 0.0, "P": 0.0, "x_floor1e-6": 0.0, "P_floor1e-6": 0.0}


# ---------------------------------------------------------------------------------------------------------------------
# 1. bench-size pools through the device-buffer entry point the bench times (te_pool_step_dense), production launch shape
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cap", [0, -1], ids=["whole_machine", "capped"])
@pytest.mark.parametrize("model", MODELS)
def test_bench_scale_dense(model, cap):
    """40 000 UV / UA, 20 000 AV / AR targets x 30 ticks.  capped: as many CTAs as give every warp (direct kernels) or CTA
    (split kernel) at least five tiles."""
    import torch
    n = 40000 + 17 if model.startswith("uniform") else 20000 + 17
    ticks = 30
    tiles = (n + 31) // 32
    if cap < 0:
        cap = max(1, tiles // (5 * 8)) if model != "angular_rates" else max(1, tiles // 17)
    te, pool, ref, ids, meas, action, N, M = _setup(model, n, ticks, cap=cap)
    d_meas = torch.from_numpy(meas).cuda()
    d_act = torch.from_numpy(action).cuda()
    worst = _worst()
    for k in range(ticks):
        pool.step_dense(DT, d_meas[k], 7, d_act[k])
        if k % 10 == 9:
            ref.step_ticks(ids, DT, meas[k - 9:k + 1], action[k - 9:k + 1])
            _check((model, k), pool, ref, ids, N, M == 6, worst)
    report.record("bench_scale_dense[%s,cap=%d]" % (model, cap), targets=n, ticks=ticks, **worst)
    pool.close(); ref.close()


# ---------------------------------------------------------------------------------------------------------------------
# 2. a handful of CTAs over ~100 tiles: every stage of every ring is reused tens of times; dense, sparse, packed, staged
# ---------------------------------------------------------------------------------------------------------------------
CASES = [("uniform_velocity", 0), ("uniform_velocity", 10), ("uniform_acceleration", 0), ("uniform_acceleration", 10), ("uniform_acceleration", 1),
         ("angular_velocities", 0), ("angular_velocities", 13), ("angular_velocities", 10), ("angular_velocities", 1), ("angular_rates", 0), ("angular_rates", 11)]


@pytest.mark.parametrize("cap", [1, 4])
@pytest.mark.parametrize("model,variant", CASES)
def test_ring_reuse_dense(model, variant, cap):
    n, ticks = 3013, 40
    te, pool, ref, ids, meas, action, N, M = _setup(model, n, ticks, variant=variant, cap=cap, seed=5)
    worst = _worst()
    for k in range(ticks):
        pool.step_dense_host(DT, meas[k], action[k])
        if k % 20 == 19:
            ref.step_ticks(ids, DT, meas[k - 19:k + 1], action[k - 19:k + 1])
            _check((model, variant, k), pool, ref, ids, N, M == 6, worst)
    report.record("ring_reuse_dense[%s,v%d,cap=%d]" % (model, variant, cap), targets=n, ticks=ticks, **worst)
    pool.close(); ref.close()


@pytest.mark.parametrize("model,variant", [("uniform_acceleration", 0), ("angular_velocities", 0), ("angular_rates", 0), ("angular_rates", 11)])
def test_ring_reuse_sparse_tile_list(model, variant):
    """te_pool_step_ids: a random 40 % of the ids per tick (sparse tile list in arbitrary order, per-slot dt), 3 CTAs"""
    n, ticks = 3013, 30
    te, pool, ref, ids, meas, action, N, M = _setup(model, n, ticks, variant=variant, cap=3, seed=6)
    rng = np.random.default_rng(1)
    worst = _worst()
    for k in range(ticks):
        pick = rng.random(n) < 0.4
        sub = np.flatnonzero(pick)
        rng.shuffle(sub)
        dts = np.where(rng.random(sub.size) < 0.5, DT, 2 * DT)
        act = action[k, sub]
        pool.step_ids(ids[sub], dts, meas[k, sub], act)
        for d in (DT, 2 * DT):       # the oracle applies the same ops, grouped by dt (targets are independent)
            g = sub[dts == d]
            ref.step_batch(ids[g], d, meas[k, g], action[k, g])
        if k % 10 == 9:
            want = ref.states(ids, N); got = pool.read_state(ids)
            x1, x2 = synth.compare_both(got["x"], want["x"]); p1, p2 = synth.compare_both(got["P"], want["P"])
            worst.update(x=max(worst["x"], x1), P=max(worst["P"], p1))
            worst["x_floor1e-6"] = max(worst["x_floor1e-6"], x2); worst["P_floor1e-6"] = max(worst["P_floor1e-6"], p2)
            assert np.array_equal(got["n_meas"], want["n_meas"])
            assert np.allclose(got["t"], want["t"], rtol=0, atol=1e-12)     # (dt and 2 dt added in a different order per target)
            assert x1 <= 1.0 and p1 <= 1.0, (k, worst)
    report.record("ring_reuse_sparse[%s,v%d]" % (model, variant), targets=n, ticks=ticks, **worst)
    pool.close(); ref.close()


# ---------------------------------------------------------------------------------------------------------------------
# 3. the compacting forms against the oracle (not only against the unfused GPU form): expiry fused into the step
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cap", [0, 3])
@pytest.mark.parametrize("model", MODELS)
def test_compacting_step_vs_oracle(model, cap):
    """te_pool_step_dense_expire, every tick: 2 % of the live ids fall silent, expire three ticks later by the reference's
    predicate (evaluated here in numpy FP64: no FMA, bit-exact), survivors are compacted by the step kernel itself.  The
    oracle steps the same ids and erases the same ones; erase lists, id order, state, covariance, t, n_meas compared."""
    import torch
    n, ticks = 6000 + 11, 24
    te, pool, ref, ids, meas, action, N, M = _setup(model, n, ticks, cap=cap, seed=8, id_stride=1)
    rng = np.random.default_rng(2)
    timeout = 3 * DT
    silent = np.zeros(ids.max() + 1, dtype=bool)
    last = np.zeros(ids.max() + 1)            # last_meas_time_ by id
    live = ids.copy()
    row = {int(i): k for k, i in enumerate(ids)}
    worst = _worst()
    n_erased = 0
    for k in range(ticks):
        ns = 1000 * 10 ** 9 + k * 4000000
        sec, nsec = ns // 10 ** 9, ns % 10 ** 9
        silent[rng.choice(live, size=max(1, live.size // 50), replace=False)] = True
        rows = np.array([row[int(i)] for i in live])
        act = np.where(silent[live], te.ACT_PREDICT, action[k, rows]).astype(np.uint8)
        m = np.ascontiguousarray(meas[k, rows])
        er = pool.step_dense_expire(DT, torch.from_numpy(m).cuda(), 7, torch.from_numpy(act).cuda(), te.ACT_UPDATE, (sec, nsec), (sec, nsec), timeout)
        # host model of the tick (src/target_manager_ros.cpp:59-72): stamp the updated ones, then the predicate
        now = float(sec) + 1e-9 * float(nsec)
        last[live[act == te.ACT_UPDATE]] = now
        expired = (last[live] > 0.0) & ((now - last[live]) >= timeout)
        assert np.array_equal(er, live[expired]), k
        n_erased += int(expired.sum())
        ref.step_batch(live[~expired], DT, m[~expired], act[~expired])     # (the step of a target erased this tick is unobservable)
        live = live[~expired]
        assert np.array_equal(pool.ids(), live), k
        if k % 8 == 7 or k == ticks - 1:
            _check((model, k), pool, ref, live, N, M == 6, worst)
    assert n_erased > n // 4
    report.record("compacting_step_vs_oracle[%s,cap=%d]" % (model, cap), targets=n, ticks=ticks, erased=n_erased, **worst)
    pool.close(); ref.close()


# ---------------------------------------------------------------------------------------------------------------------
# 4. SURVEY.md 8(d) parity protocol for the angular models at full size (UV / UA: tests/test_gpu_parity.py)
#
# Conditioning.  The angular-velocities EKF of the reference is not stable on every stream: its pitch STATE drifts away from the
# measured pitch (<= 0.55 rad here) and, on fast-rotating targets, through +-pi/2, where J_rpy and J_w divide by cos(pitch)^2; from
# then on the filter amplifies rounding-level noise exponentially.  Measured on the CPU oracle alone (the oracle run twice, the
# second time with every measurement quaternion component moved by ONE ULP; 4096 targets, worst |d| / bar over the run):
#     roll / yaw rates within +-2 rad/s (tests/synth.py default): after 2000 ticks 343 targets differ FROM THEMSELVES by more
#         than the 1e-9 bar, 692 by more than 1e-2 of it;
#     roll / yaw rates within +-0.5 rad/s (the rates SURVEY.md 8(d) specifies): none above 1e-1 of the bar, 38 above 1e-2.
# No implementation with another libm / summation order can track an amplifying target to 1e-9, the reference compiled against
# another libm included.  The protocol therefore runs the perturbed oracle beside the oracle and measures each target's
# amplification a_i = worst |d| / bar of the oracle against itself under the one-ulp perturbation:
#     well-conditioned (a_i <= 1e-2, i.e. one ulp of input moves the result by <= 1e-11 relative): the 1e-9 bar, no exception;
#     the others (the reference's own trajectory is not reproducible to the bar from one-ulp-different inputs; once a target
#         amplifies exponentially the difference saturates at O(1), so no multiple of a_i bounds it): counted, bounded in number,
#         reported with their worst ratio -- and excluded.
# On the SURVEY.md 8(d) stream (+-0.5 rad/s) NOTHING is excluded: all 4096 targets meet the bar at every checkpoint (measured worst
# ratio 0.06).  t and n_meas stay exact for all targets, everything stays finite.
# ---------------------------------------------------------------------------------------------------------------------
ILL = 1e-2


def _run_conditioned(tag, model, n, ticks, every, ill_frac_max, exclude=True, **stream_kw):
    te, pool, ref, ids, meas, action, N, M = _setup(model, n, ticks, **stream_kw)
    import target_estimation_b200 as te_
    mtype, _, Q, R, P0 = te_.load_model(model)
    scale = _setup.scale
    pert = orc.ShardedManager()
    pert.init_batch(mtype, ids, DT, Q, R, P0, meas[0], scale)
    meas_p = synth.perturb_quaternions(meas)
    amp = np.zeros(n)       # a_i so far (sticky: an amplifying target stays one)
    worst = {"x": 0.0, "P": 0.0, "x_all": 0.0, "P_all": 0.0, "rel_ill": 0.0}
    fail = None
    for k in range(ticks):
        pool.step_dense_host(DT, meas[k], action[k])
        if k % every == every - 1:
            ref.step_ticks(ids, DT, meas[k - every + 1:k + 1], action[k - every + 1:k + 1])
            pert.step_ticks(ids, DT, meas_p[k - every + 1:k + 1], action[k - every + 1:k + 1])
            want, wp, got = ref.states(ids, N), pert.states(ids, N), pool.read_state(ids)
            amp = np.maximum(amp, np.maximum(synth.ratio_per_target(wp["x"], want["x"]), synth.ratio_per_target(wp["P"], want["P"])))
            ill = amp > ILL
            rx, rP = synth.ratio_per_target(got["x"], want["x"]), synth.ratio_per_target(got["P"], want["P"])
            worst["x_all"] = max(worst["x_all"], float(rx.max())); worst["P_all"] = max(worst["P_all"], float(rP.max()))
            if (~ill).any():
                worst["x"] = max(worst["x"], float(rx[~ill].max())); worst["P"] = max(worst["P"], float(rP[~ill].max()))
            if ill.any():
                worst["rel_ill"] = max(worst["rel_ill"], float((np.maximum(rx, rP)[ill] / amp[ill]).max()))
            assert np.array_equal(got["n_meas"], want["n_meas"]) and np.array_equal(got["t"], want["t"]), k
            assert np.isfinite(got["x"]).all() and np.isfinite(got["P"]).all(), k
            ok = (worst["x"] <= 1.0 and worst["P"] <= 1.0) if exclude else (worst["x_all"] <= 1.0 and worst["P_all"] <= 1.0)
            if fail is None and not ok:
                fail = (tag, k, int(ill.sum()), dict(worst))
    pitch = float(np.abs(ref.states(ids, N)["x"][:, 4]).max()) if model == "angular_velocities" else None
    n_ill = int((amp > ILL).sum())
    report.record(tag, targets=n, ticks=ticks, ill_conditioned=n_ill, amplification_hist={("%g" % t): int((amp > t).sum()) for t in (1e-3, 1e-2, 1e-1, 1.0)},
                  max_state_pitch=pitch, x=worst["x"], P=worst["P"], ill_error_over_amplification=worst["rel_ill"],
                  x_incl_ill=worst["x_all"], P_incl_ill=worst["P_all"], first_failure=str(fail) if fail else None)
    pool.close(); ref.close(); pert.close()
    assert fail is None, fail
    assert n_ill <= ill_frac_max * n, (tag, n_ill)
    return n_ill


def test_step_parity_4096x2000_angular_velocities():
    """4096 targets x 2000 ticks, compared every 100 ticks, attitude rates as specified there (+-0.5 rad/s)"""
    _run_conditioned("step_parity_4096x2000[angular_velocities]", "angular_velocities", 4096, 2000, 100, 0.03, exclude=False, att_rate=0.5)


def test_step_parity_4096x2000_angular_velocities_fast_rotation():
    """the same with the +-2 rad/s roll / yaw rates every other test uses (several wraps per target): on these the reference's EKF
    itself amplifies one-ulp input noise past the bar on 8 % of the targets (CPU measurement above)"""
    _run_conditioned("step_parity_4096x2000[angular_velocities,fast]", "angular_velocities", 4096, 2000, 100, 0.25)


def test_step_parity_4096x2000_angular_rates():
    """(linear filter: no amplifying targets at all)"""
    assert _run_conditioned("step_parity_4096x2000[angular_rates]", "angular_rates", 4096, 2000, 100, 0.0, exclude=False) == 0


# ---------------------------------------------------------------------------------------------------------------------
# 5. how wide is the region in which the AV EKF meets the contract?  measured pitch within +-pitch_max (+-1.2)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pitch_max", [0.55, 0.8, 1.0, 1.2])
def test_av_pitch_range(pitch_max):
    """every well-conditioned target (see above) meets 1e-9 for every range; the report carries the number of amplifying ones
    per range -- the width of the region in which the reference's EKF is trackable at all"""
    _run_conditioned("av_pitch_range[%.2f]" % pitch_max, "angular_velocities", 2048, 600, 100, 0.25, seed=21, pitch0=pitch_max - 0.15, pitch_amp=0.15)


# ---------------------------------------------------------------------------------------------------------------------
# 6. the host-buffer tick the bench's e2e figure times (te_pool_tick_host: chunked H2D / step / D2H pipeline over three streams)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pipelined", [0, 1, 2], ids=["sync", "two_in_flight", "three_in_flight"])
@pytest.mark.parametrize("stride", [7, 3])
def test_tick_host_pipeline(stride, pipelined, monkeypatch):
    """300 000 uniform-acceleration targets in pipeline chunks of 2048 tiles (65 536 targets: four chunks + a ragged tail); pose [n][7] and
    xyz-only [n][3] measurements; state, covariance and the returned positions against the oracle.  two / three_in_flight: the same
    ticks through te_pool_tick_host_async / te_pool_tick_host_wait(lag = 1 / 2) (the copies of tick k + 1 under the kernels of tick k
    and the read-back of tick k - 1), every tick's returned positions checked against the oracle's."""
    import ctypes as C
    monkeypatch.setenv("TE_TICK_CHUNK_TILES", "2048")     # (the default chunk is the whole pool at this size)
    n, ticks = 300000 + 5, 6
    te, pool, ref, ids, meas, action, N, M = _setup("uniform_acceleration", n, ticks, seed=31)
    outs = [np.zeros((n, 3)) for _ in range(ticks)]
    worst = _worst()
    ms = [np.ascontiguousarray(meas[k][:, :stride]) for k in range(ticks)]
    acts = [np.ascontiguousarray(action[k]) for k in range(ticks)]
    for k in range(ticks):
        fn = te.lib.te_pool_tick_host_async if pipelined else te.lib.te_pool_tick_host
        rc = fn(pool._h, DT, ms[k].ctypes.data_as(C.c_void_p), stride, acts[k].ctypes.data_as(C.c_void_p), 2, outs[k].ctypes.data_as(C.c_void_p))
        assert rc == 0, te._lib.last_error()
        if pipelined:
            assert te.lib.te_pool_tick_host_wait(pool._h, pipelined) == 0
            if k >= pipelined:
                assert np.abs(outs[k - pipelined]).max() > 0  # tick k - lag has landed while the newer ones are in flight
    assert te.lib.te_pool_tick_host_wait(pool._h, 0) == 0
    # the oracle tick by tick: the returned positions of EVERY tick are that tick's estimated positions
    for k in range(ticks):
        ref.step_ticks(ids, DT, meas[k:k + 1], action[k:k + 1])
        want = ref.states(ids, N)["x"][:, :3]
        assert synth.compare_h2(outs[k], want) <= 1.0, k
    _check(("tick_host", stride, pipelined), pool, ref, ids, N, False, worst)
    assert np.array_equal(outs[-1], pool.read_state(ids)["x"][:, :3])          # the D2H record of the last tick = the estimated positions
    report.record("tick_host_pipeline[stride=%d,%s]" % (stride, "two_in_flight" if pipelined else "sync"), targets=n, ticks=ticks, **worst)
    pool.close(); ref.close()
