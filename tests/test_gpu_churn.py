"""-m gpu: add / erase churn and expiry -- track ids and erase decisions bit-exact against the oracle's
RosTargetManager-tick restatement (src/target_manager_ros.cpp:41-92), state parity after compaction."""
import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
DT = 1.0 / 250.0


def test_add_erase_order_and_state():
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model("uniform_velocity")
    pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
    ref = orc.Manager()
    rng = np.random.default_rng(0)
    live = set()
    next_meas = lambda n: np.hstack([rng.normal(size=(n, 3)), np.tile([0, 0, 0, 1.0], (n, 1))])
    for rnd in range(12):
        # add a random batch (some duplicates of live ids, unsorted, interleaved with existing ids)
        new = rng.choice(5000, size=rng.integers(1, 200), replace=False).astype(np.uint32)
        p0 = next_meas(new.size)
        added = pool.add(new, p0, t0=np.full(new.size, 0.1 * rnd))
        fresh = [int(i) for i in new if int(i) not in live]
        assert added == len(fresh)
        for k, i in enumerate(new):
            if int(i) not in live:
                ref.init_full(mtype, int(i), DT, 0.1 * rnd, Q, R, P0, p0[k]); live.add(int(i))
        ids = pool.ids()
        assert np.array_equal(ids, ref.ids()) and np.array_equal(ids, np.sort(ids))
        for _ in range(3):
            m = next_meas(ids.size); act = rng.integers(0, 3, ids.size).astype(np.uint8)
            pool.step_dense_host(DT, m, act); ref.step_batch(ids, DT, m, act)
        gone = rng.choice(ids, size=min(ids.size // 3, 150), replace=False)
        assert pool.erase(np.concatenate([gone, [9999999]])) == gone.size
        for g in gone:
            ref.erase(int(g)); live.discard(int(g))
        ids = pool.ids()
        assert np.array_equal(ids, ref.ids())
        got, want = pool.read_state(), ref.states(ids, 6)
        assert synth.compare_h2(got["x"], want["x"]) <= 1.0 and synth.compare_h2(got["P"], want["P"]) <= 1.0
        assert np.array_equal(got["n_meas"], want["n_meas"]) and np.array_equal(got["t"], want["t"])
    pool.close()


def test_reserve_after_merge_add():
    """te_pool_reserve once a merge flipped the buffers (the flipped buffer can be larger than the work arrays)"""
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model("uniform_velocity")
    pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
    p0 = np.tile([0, 0, 0, 0, 0, 0, 1.0], (100, 1))
    assert pool.add(np.arange(100, dtype=np.uint32) * 2 + 1000, p0) == 100
    assert pool.add(np.arange(100, dtype=np.uint32) * 2, p0) == 100          # below the existing ids: merge into the other buffer
    pool.reserve(250)
    assert pool.add(np.arange(50, dtype=np.uint32) * 2 + 1, p0[:50]) == 50
    assert len(pool) == 250 and np.array_equal(pool.ids(), np.sort(pool.ids()))
    pool.step_dense_host(DT, np.tile(p0[:1], (250, 1)))
    assert np.array_equal(pool.read_state()["n_meas"], np.ones(250, dtype=np.int64))
    pool.close()


@pytest.mark.parametrize("n0,ticks", [(2048, 48), (16384, 24)])
def test_expiry_bit_exact(n0, ticks):
    """(16 384 targets = the sub-run size SURVEY.md 8(d) names)
    SURVEY.md C3 sub-run: synthetic clock (sec, nsec) with epoch 1000 s, per tick 1 % of the live ids stop
    receiving measurements, expire after timeout = 8 dt by `last > 0 && now - last >= timeout`, and as many fresh ids
    appear.  The set and order of live ids and every tick's erase list must match the oracle exactly."""
    import target_estimation_b200 as te
    name = "angular_rates"
    mtype, _, Q, R, P0 = te.load_model(name)
    timeout = 8 * DT
    pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
    L = orc.lib()
    h = L.orc_tick_new(mtype, orc.ptr(orc.colmajor(Q)), Q.shape[0], orc.ptr(orc.colmajor(R)), R.shape[0], orc.ptr(orc.colmajor(P0)))
    L.orc_tick_set_expiration(h, timeout)
    rng = np.random.default_rng(17)
    next_id = n0
    active = list(range(n0))            # ids still producing measurements
    t_tick = 0.0
    def stamp(k):                        # clock at tick k: 1000 s + k * 4 ms, in integer nanoseconds
        ns = 1000 * 10 ** 9 + k * 4000000
        return ns // 10 ** 9, ns % 10 ** 9
    for k in range(ticks):
        sec, nsec = stamp(k)
        if k > 0:
            quit_ = rng.choice(len(active), size=max(1, len(active) // 100), replace=False)
            active = [a for j, a in enumerate(active) if j not in set(quit_.tolist())]
            fresh = list(range(next_id, next_id + len(quit_))); next_id += len(quit_)
            active += fresh
        ids = np.array(active, dtype=np.uint32)
        quat = synth.rpy_to_quat(rng.uniform(-1, 1, (ids.size, 3)))
        poses = np.hstack([rng.normal(size=(ids.size, 3)), quat])
        stamps = np.tile(np.array([sec, nsec], dtype=np.uint32), (ids.size, 1))
        # ---- oracle: /tf callback then the tick
        L.orc_tick_callback_ids(h, ids.size, orc.ptr(ids), orc.ptr(np.ascontiguousarray(stamps)), orc.ptr(np.ascontiguousarray(poses)))
        erased_ref = np.zeros(1 << 16, dtype=np.uint32)
        n_er = L.orc_tick_update(h, DT, sec, nsec, orc.ptr(erased_ref), erased_ref.size)
        # ---- device: the same tick expressed with the pool primitives
        have = set(pool.ids().tolist())
        new_mask = np.array([int(i) not in have for i in ids])
        if new_mask.any():                                   # init on first sight with the measurement as p0, t0 = t_
            pool.add(ids[new_mask], poses[new_mask], t0=np.full(int(new_mask.sum()), t_tick))
        pool.set_stamps(ids, stamps[:, 0], stamps[:, 1])
        # every live target: update if its mailbox has a (possibly stale) measurement -- the reference's new_meas_
        # flag is sticky (H10), so silent targets re-apply their last pose until they expire
        live = pool.ids()
        last = dict(zip(ids.tolist(), poses))
        if k == 0:
            mailbox = {}
        mailbox.update(last)
        m = np.array([mailbox[int(i)] for i in live])
        pool.step_ids(live, DT, m)
        erased = pool.expire(sec, nsec, timeout)
        for e in erased:
            mailbox.pop(int(e), None)
        t_tick = t_tick + DT
        assert n_er == erased.size and np.array_equal(erased, erased_ref[:n_er]), k
        ref_ids = np.zeros(max(L.orc_num_targets(h), 1), dtype=np.uint32)
        n_ref = L.orc_get_ids(h, orc.ptr(ref_ids), ref_ids.size)
        assert np.array_equal(pool.ids(), ref_ids[:n_ref]), k
    assert next_id > n0 + n0 // 100 * (ticks - 1) // 2 and len(pool) > 0
    # state parity of a sample of survivors after all the compaction
    live = pool.ids()
    sample = live[:: max(1, live.size // 64)]
    got = pool.read_state(sample)
    N = Q.shape[0]
    xs = np.zeros((sample.size, N)); Ps = np.zeros((sample.size, N, N))
    import ctypes as C
    for j, i in enumerate(sample):
        t = C.c_double(); nm = C.c_longlong()
        L.orc_get_state(h, int(i), orc.ptr(xs[j]), orc.ptr(Ps[j]), C.byref(t), C.byref(nm), None)
        assert got["n_meas"][j] == nm.value and got["t"][j] == t.value
    assert synth.compare_h2(got["x"], xs) <= 1.0 and synth.compare_h2(got["P"], Ps) <= 1.0
    pool.close()


@pytest.mark.parametrize("name,variant", [("uniform_velocity", 0), ("uniform_acceleration", 0), ("angular_velocities", 0), ("angular_rates", 0),
                                          ("uniform_acceleration", 10), ("angular_velocities", 10), ("angular_velocities", 13), ("angular_rates", 11)])
def test_fused_step_expire_is_bit_identical(name, variant):
    """te_pool_step_dense_expire (compaction fused into the step kernel: survivors' columns go straight to their compacted
    slots in the second buffer) == te_pool_step_dense + te_pool_stamp_dense + te_pool_expire, bit for bit: erased ids,
    surviving ids and order, x, P, t, n_meas, prev_rpy.  Pool size not a multiple of the tile, ticks with and without
    expiries, heavy (30 %) and light churn, fresh ids appended in between."""
    import torch
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(name)
    angular = name.startswith("angular")
    n0, ticks = 3000 + 13, 30
    timeout = 3 * DT
    meas_all, action_all, scale = synth.make_streams(n0 + 40 * ticks, ticks, DT, accel=True, angular=angular, seed=11)
    pools = []
    for _ in range(2):
        p = te.TargetPool(mtype); p.register_class(Q, R, P0)
        # 0: the default kernels (angular velocities: the TMA-streamed kernel in both forms); 10: the full-matrix kernels; 11: AR
        # row-split kernel, packed; 13: AV direct kernel in both forms
        p.set_variant(variant)
        p.add(np.arange(n0, dtype=np.uint32), meas_all[0, :n0], p0_scale=scale[:n0])
        pools.append(p)
    rng = np.random.default_rng(23)
    next_id = n0
    silent = np.zeros(n0 + 40 * ticks, dtype=bool)     # by id: no more measurements
    stamp = lambda k: ((1000 * 10 ** 9 + k * 4000000) // 10 ** 9, (1000 * 10 ** 9 + k * 4000000) % 10 ** 9)
    total = 0
    for k in range(ticks):
        ids = pools[0].ids()
        assert np.array_equal(ids, pools[1].ids())
        if k in (2, 9):        # a burst: 30 % fall silent at once
            silent[rng.choice(ids, size=ids.size * 3 // 10, replace=False)] = True
        elif k % 4 != 3:       # light churn; every fourth tick nobody new falls silent
            silent[rng.choice(ids, size=max(1, ids.size // 100), replace=False)] = True
        act = np.where(silent[ids], te.ACT_PREDICT, action_all[k % ticks, ids]).astype(np.uint8)
        act[silent[ids] & (rng.random(ids.size) < 0.1)] = te.ACT_NONE
        m = torch.from_numpy(np.ascontiguousarray(meas_all[k % ticks, ids])).cuda()
        a = torch.from_numpy(act).cuda()
        sec, nsec = stamp(k)
        pools[0].step_dense(DT, m, 7, a, te.ACT_UPDATE)
        pools[0].stamp_dense(sec, nsec, a)
        er0 = pools[0].expire(sec, nsec, timeout)
        er1 = pools[1].step_dense_expire(DT, m, 7, a, te.ACT_UPDATE, (sec, nsec), (sec, nsec), timeout)
        assert np.array_equal(er0, er1), k
        total += er0.size
        s0, s1 = pools[0].read_state(), pools[1].read_state()
        assert np.array_equal(pools[0].ids(), pools[1].ids())
        for f in ("x", "P", "t", "n_meas") + (("prev_rpy",) if angular else ()):
            assert np.array_equal(s0[f].view(np.uint64) if s0[f].dtype == np.float64 else s0[f],
                                  s1[f].view(np.uint64) if s1[f].dtype == np.float64 else s1[f]), (k, f)
        # fresh ids (appended: larger than every live id), first measurement as p0
        nf = 40
        fresh = np.arange(next_id, next_id + nf, dtype=np.uint32); next_id += nf
        for p in pools:
            p.add(fresh, meas_all[k % ticks, fresh], t0=np.full(nf, k * DT), p0_scale=scale[fresh])
    assert total > n0 // 3
    for p in pools:
        p.close()
