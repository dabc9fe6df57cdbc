"""-m gpu: device-resident mailboxes (te_pool_mailbox_ingest / te_pool_mailbox_tick) against the oracle's RosTargetManager
restatement (oracle::TickTargetManager, itself bit-identical to the reference's src/target_manager_ros.cpp compiled
unmodified: tests/test_ref_ros_tick.py).  Per tick: live ids, erase list (target-less mailboxes included), mailbox count
bit-exact; state, covariance at the 1e-9 bar; time and measurement counts exact."""
import ctypes as C

import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
DT = 1.0 / 250.0


def _pair(name):
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(name)
    pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
    L = orc.lib()
    h = L.orc_tick_new(mtype, orc.ptr(orc.colmajor(Q)), Q.shape[0], orc.ptr(orc.colmajor(R)), R.shape[0], orc.ptr(orc.colmajor(P0)))
    return pool, L, h, Q.shape[0]


def _deliver(pool, L, h, ids, stamps, poses):
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    st = np.ascontiguousarray(stamps, dtype=np.uint32).reshape(-1, 2)
    ps = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 7)
    L.orc_tick_callback_ids(h, ids.size, orc.ptr(ids), orc.ptr(st), orc.ptr(ps))
    pool.mailbox_ingest(ids, st[:, 0].copy(), st[:, 1].copy(), ps)


def _tick(pool, L, h, k, t_tick, now_ns, timeout):
    sec, nsec = now_ns // 10**9, now_ns % 10**9
    erased_ref = np.zeros(1 << 16, dtype=np.uint32)
    n_er = L.orc_tick_update(h, DT, sec, nsec, orc.ptr(erased_ref), erased_ref.size)
    erased, added = pool.mailbox_tick(DT, t_tick, (sec, nsec), timeout)
    assert np.array_equal(erased, erased_ref[:n_er]), (k, erased, erased_ref[:n_er])
    ref_ids = np.zeros(max(L.orc_num_targets(h), 1), dtype=np.uint32)
    n_ref = L.orc_get_ids(h, orc.ptr(ref_ids), ref_ids.size)
    assert np.array_equal(pool.ids(), ref_ids[:n_ref]), k
    assert pool.mailbox_count() == L.orc_tick_mailboxes(h), k
    return erased, added


def _compare_states(pool, L, h, N, ids=None):
    live = pool.ids() if ids is None else np.asarray(ids, dtype=np.uint32)
    if live.size == 0:
        return
    got = pool.read_state(live)
    xs = np.zeros((live.size, N)); Ps = np.zeros((live.size, N, N)); mp = np.zeros((live.size, 7))
    for j, i in enumerate(live):
        t = C.c_double(); nm = C.c_longlong()
        assert L.orc_get_state(h, int(i), orc.ptr(xs[j]), orc.ptr(Ps[j]), C.byref(t), C.byref(nm), None)
        assert got["n_meas"][j] == nm.value and got["t"][j] == t.value, (int(i), got["n_meas"][j], nm.value)
        assert L.orc_get_measured_pose(h, int(i), orc.ptr(mp[j]))
    assert synth.compare_h2(got["x"], xs) <= 1.0 and synth.compare_h2(got["P"], Ps) <= 1.0
    assert np.array_equal(got["measured_pose"], mp)


@pytest.mark.parametrize("grid_cap", [0, 2])
@pytest.mark.parametrize("name", ["uniform_velocity", "uniform_acceleration", "angular_velocities", "angular_rates"])
def test_mailbox_churn_matches_oracle_tick(name, grid_cap, monkeypatch):
    """(grid_cap 2: the fused compacting tick on two CTAs -- every CTA of the split kernel takes ~12 tiles through its two-stage
    ring, every warp of the direct kernels several tiles)
    a /tf message per tick for a changing set of ids in shuffled order: newcomers (merged anywhere in the id order), ids
    falling silent (sticky re-application of their last pose, then expiry), stale stamps (stored, predict-only), an id named
    twice in one message, a second message before some ticks."""
    if grid_cap:
        monkeypatch.setenv("TE_GRID_CAP", str(grid_cap))     # read by te_pool_create
    pool, L, h, N = _pair(name)
    timeout = 8 * DT
    L.orc_tick_set_expiration(h, timeout)
    rng = np.random.default_rng(len(name))
    n0, ticks = 700, 60
    universe = rng.choice(200000, size=n0 + 16 * ticks, replace=False).astype(np.uint32)
    streams, _, _ = synth.make_streams(universe.size, ticks, DT, accel=True, angular=name.startswith("angular"), seed=17)
    live = list(range(n0)); nxt = n0
    t_tick = 0.0
    n_erased = n_added = 0
    for k in range(ticks):
        now = 1000 * 10**9 + k * 4_000_000
        gone = set(j for j in live if rng.random() < 0.01)
        live = [j for j in live if j not in gone] + list(range(nxt, nxt + 16)); nxt += 16
        speak = np.array([j for j in live if rng.random() < 0.93])
        rng.shuffle(speak)
        stale = rng.random(speak.size) < 0.05
        st_ns = now - np.where(stale, 12_000_000, 0)
        stamps = np.stack([st_ns // 10**9, st_ns % 10**9], axis=1)
        poses = streams[k, speak]
        ids = universe[speak]
        # one id twice in the same message: the second record (same stamp -> not newer) makes the mailbox unreadable
        if k % 7 == 3:
            ids = np.concatenate([ids, ids[:2]]); stamps = np.concatenate([stamps, stamps[:2]]); poses = np.concatenate([poses, poses[:2] + 0.5])
        _deliver(pool, L, h, ids, stamps, poses)
        if k % 5 == 2:   # a second message before the tick, newer stamps for a few ids
            sub = speak[:40]
            st2 = np.tile([(now + 1_000_000) // 10**9, (now + 1_000_000) % 10**9], (sub.size, 1))
            _deliver(pool, L, h, universe[sub], st2, streams[k, sub] + 0.01)
        erased, added = _tick(pool, L, h, k, t_tick, now, timeout)
        n_erased += erased.size; n_added += added
        t_tick = t_tick + DT
        if k % 10 == 9:
            live_ids = pool.ids()
            _compare_states(pool, L, h, N, live_ids[:: max(1, live_ids.size // 48)])
    assert n_erased > 100 and n_added > n0 + 500
    live_ids = pool.ids()
    _compare_states(pool, L, h, N, live_ids[:: max(1, live_ids.size // 96)])
    pool.close(); L.orc_manager_delete(h)


def test_mailbox_quirks_match_oracle_tick():
    """mailboxes without a target: a stamp of 0 (never newer than the initial stamp), a new record followed by an older one
    before the tick (unreadable, but carries a last_meas_time and expires); expiry one rounding error short of the boundary; an expired
    id coming back as a new target; erase lists include the target-less mailboxes, as the reference's erase does."""
    pool, L, h, N = _pair("uniform_acceleration")
    timeout = 0.1
    L.orc_tick_set_expiration(h, timeout)
    rng = np.random.default_rng(3)
    pose = lambda n=1: np.hstack([rng.normal(size=(n, 3)), np.tile([0, 0, 0, 1.0], (n, 1))])
    t_tick = 0.0
    base = 10 * 10**9
    S = lambda ns: (ns // 10**9, ns % 10**9)
    # message 0: id 5 normal; id 7 stamp 0; id 9 new then older within one message; id 11 normal
    _deliver(pool, L, h, [5, 7, 9, 9, 11], [S(base), (0, 0), S(base), S(base - 1000), S(base)], pose(5))
    events = {}
    for k in range(80):
        now = base + k * 4_000_000
        if k == 3:      # 9 becomes readable again -> target on this tick; 7 gets a real stamp
            _deliver(pool, L, h, [9, 7], [S(now), S(now)], pose(2))
        if k == 10:     # a fresh id whose only record is then overwritten by an older one in a second message
            _deliver(pool, L, h, [13], [S(now)], pose(1))
            _deliver(pool, L, h, [13], [S(now - 5)], pose(1))
        if k == 40:     # 5 expired long ago: comes back as a new target
            _deliver(pool, L, h, [5], [S(now)], pose(1))
        if k in (20, 21, 22):   # keep 11 alive a little longer than the others
            _deliver(pool, L, h, [11], [S(now)], pose(1))
        erased, added = _tick(pool, L, h, k, t_tick, now, timeout)
        t_tick = t_tick + DT
        if erased.size or added:
            events[k] = (erased.tolist(), added)
        _compare_states(pool, L, h, N)
    assert events[0] == ([], 2)                  # 5 and 11; 7 and 9 are mailboxes without targets
    assert events[3] == ([], 2)                  # 9 and 7 promoted
    assert events[26][0] == [5]                  # at k = 25 the clock reads 10.1 s, and 10.1 - 10.0 = 0.09999999999999964 < 0.1
    assert 13 in events[35][0]                   # the target-less mailbox of 13 expires 0.1 s after its accepted stamp
    assert events[40] == ([], 1)                 # 5 is back
    assert pool.mailbox_count() == 0 and len(pool) == 0
    pool.close(); L.orc_manager_delete(h)


def test_mailboxes_and_by_hand_calls_interplay_like_the_reference():
    """TargetManager::erase / init by hand on a manager whose mailboxes live on the device: an erased target keeps its mailbox
    (still readable -> the next tick re-creates the target from the stored pose); a target created by hand has no mailbox and
    is not touched by the tick; one created by hand for an id that already has a target-less mailbox is fed by it (here an
    unreadable one: predicted every tick).  The mailboxes follow their slots through te_pool_erase_batch and both
    te_pool_add_batch paths (merge, append)."""
    import target_estimation_b200 as te
    pool, L, h, N = _pair("uniform_velocity")
    mtype, _, Q, R, P0 = te.load_model("uniform_velocity")
    ref = orc.Manager.__new__(orc.Manager); ref.L = L; ref.h = h      # the by-hand API of the oracle's tick manager
    L.orc_tick_set_expiration(h, 1000.0)
    rng = np.random.default_rng(1)
    pose = lambda n: np.hstack([rng.normal(size=(n, 3)), np.tile([0, 0, 0, 1.0], (n, 1))])
    ids = np.arange(100, 400, 3, dtype=np.uint32)
    p0 = pose(ids.size)
    _deliver(pool, L, h, np.concatenate([ids, [5000]]), [(50, 0)] * ids.size + [(0, 0)], np.concatenate([p0, pose(1)]))
    _tick(pool, L, h, 0, 0.0, 50 * 10**9, 1000.0)
    assert len(pool) == ids.size and pool.mailbox_count() == ids.size + 1
    gone = ids[5:20]
    assert pool.erase(gone) == gone.size
    for g in gone:
        assert ref.erase(int(g))
    assert pool.mailbox_count() == ids.size + 1 == L.orc_tick_mailboxes(h)
    mid = np.array([101, 251, 252, 5000, 5001], dtype=np.uint32); pm = pose(mid.size)
    assert pool.add(mid, pm, t0=np.full(mid.size, DT)) == mid.size
    for j, i in enumerate(mid):
        ref.init_full(mtype, int(i), DT, DT, Q, R, P0, pm[j])
    t_tick = DT
    for k in range(1, 7):
        if k == 4:   # a record for a by-hand target creates its mailbox
            _deliver(pool, L, h, [251], [(50, k * 4_000_000)], pose(1))
        erased, added = _tick(pool, L, h, k, t_tick, 50 * 10**9 + k * 4_000_000, 1000.0)
        assert added == (gone.size if k == 1 else 0) and erased.size == 0
        t_tick = t_tick + DT
        _compare_states(pool, L, h, N)
    got = pool.read_state(mid)
    assert got["n_meas"].tolist() == [0, 3, 0, 0, 0] and got["t"][0] == DT and got["t"][3] > 6 * DT    # 5000 predicted, 101 untouched
    with pytest.raises(te.TeError):
        pool.step_dense_expire(DT, None, 7, None, te.ACT_PREDICT, (50, 0), (50, 0), 1.0)
    pool.close(); L.orc_manager_delete(h); ref.h = None


@pytest.mark.parametrize("name", ["uniform_velocity", "uniform_acceleration", "angular_velocities", "angular_rates"])
def test_fused_mailbox_tick_is_bit_identical_to_rebuild_then_step(name, monkeypatch):
    """the default tick lets the step kernel move the survivors to their merged slots and gives the new targets their first
    update in a sparse follow-up launch; TE_MB_UNFUSED=1 selects rebuild-then-step.  Same ids, same erase lists, same bits."""
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(name)
    results = []
    for unfused in (False, True):
        if unfused:
            monkeypatch.setenv("TE_MB_UNFUSED", "1")
        else:
            monkeypatch.delenv("TE_MB_UNFUSED", raising=False)
        pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
        if name == "angular_velocities":
            # rebuild-then-step steps in place, where the default is the TMA-streamed kernel; the compacting step of the fused tick is
            # the direct kernel.  Same arithmetic in the source, but ptxas contracts multiply-adds per kernel: bits are compared between
            # the two forms of ONE kernel (variant 13 = the direct kernel for every launch)
            pool.set_variant(13)
        rng = np.random.default_rng(11)
        universe = rng.choice(100000, size=3000, replace=False).astype(np.uint32)
        streams, _, _ = synth.make_streams(universe.size, 30, DT, accel=True, angular=name.startswith("angular"), seed=5)
        live = list(range(900)); nxt = 900
        log = []
        for k in range(30):
            now = 1000 * 10**9 + k * 4_000_000
            gone = set(j for j in live if rng.random() < 0.02)
            live = [j for j in live if j not in gone] + list(range(nxt, nxt + 40)); nxt += 40
            speak = np.array([j for j in live if rng.random() < 0.95])
            if k % 4 == 1:
                speak = speak[:0]                    # a tick without any message: nothing appears, maybe something expires
            st = np.tile([now // 10**9, now % 10**9], (speak.size, 1)).astype(np.uint32)
            if speak.size:
                pool.mailbox_ingest(universe[speak], st[:, 0].copy(), st[:, 1].copy(), streams[k, speak])
            erased, added = pool.mailbox_tick(DT, k * DT, (now // 10**9, now % 10**9), 6 * DT, want_added=True)
            log.append((erased.copy(), added.copy()))
        results.append((pool.ids(), pool.read_state(), log))
        pool.close()
    (ids_a, st_a, log_a), (ids_b, st_b, log_b) = results
    assert np.array_equal(ids_a, ids_b) and ids_a.size > 1000
    for (ea, aa), (eb, ab) in zip(log_a, log_b):
        assert np.array_equal(ea, eb) and np.array_equal(aa, ab)
    assert sum(e.size for e, _ in log_a) > 100
    N = Q.shape[0]
    iu = np.triu_indices(N)
    for key in ("x", "t", "n_meas", "prev_rpy", "measured_pose"):
        assert np.array_equal(st_a[key], st_b[key]), key
    assert np.array_equal(st_a["P"][:, iu[0], iu[1]], st_b["P"][:, iu[0], iu[1]])   # (the packed kernels keep the upper triangle)


def test_device_resident_messages_equal_host_messages():
    """te_pool_mailbox_ingest_dev (records already in device memory, used in place; unknown ids read back for the host queue)
    against te_pool_mailbox_ingest on the same shuffled, duplicated, partly stale messages: same ids, erase lists, bits."""
    import torch
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model("angular_rates")
    results = []
    for on_device in (False, True):
        pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
        rng = np.random.default_rng(21)
        universe = rng.choice(50000, size=2500, replace=False).astype(np.uint32)
        streams, _, _ = synth.make_streams(universe.size, 24, DT, accel=True, angular=True, seed=8)
        live = list(range(600)); nxt = 600
        log = []
        for k in range(24):
            now = 1000 * 10**9 + k * 4_000_000
            gone = set(j for j in live if rng.random() < 0.03)
            live = [j for j in live if j not in gone] + list(range(nxt, nxt + 30)); nxt += 30
            speak = np.array([j for j in live if rng.random() < 0.9]); rng.shuffle(speak)
            speak = np.concatenate([speak, speak[:5]])                      # five ids twice
            st_ns = now - np.where(rng.random(speak.size) < 0.05, 12_000_000, 0)
            sec = (st_ns // 10**9).astype(np.uint32); nsec = (st_ns % 10**9).astype(np.uint32)
            ids = universe[speak]; poses = np.ascontiguousarray(streams[k, speak])
            if on_device:
                d = [torch.from_numpy(a.view(np.int32)).cuda() for a in (ids, sec, nsec)] + [torch.from_numpy(poses).cuda()]
                torch.cuda.synchronize()
                pool.mailbox_ingest_dev(ids.size, *d)
            else:
                pool.mailbox_ingest(ids, sec, nsec, poses)
            log.append(pool.mailbox_tick(DT, k * DT, (now // 10**9, now % 10**9), 6 * DT, want_added=True))
        results.append((pool.ids(), pool.read_state(), log, pool.mailbox_count()))
        pool.close()
    (ids_a, st_a, log_a, mc_a), (ids_b, st_b, log_b, mc_b) = results
    assert np.array_equal(ids_a, ids_b) and ids_a.size > 500 and mc_a == mc_b
    for (ea, aa), (eb, ab) in zip(log_a, log_b):
        assert np.array_equal(ea, eb) and np.array_equal(aa, ab)
    assert sum(e.size for e, _ in log_a) > 50
    for key in ("x", "P", "t", "n_meas", "prev_rpy", "measured_pose"):
        assert np.array_equal(st_a[key], st_b[key]), key


def test_prefetched_messages_equal_synchronous_ingest():
    """te_pool_mailbox_prefetch + te_pool_mailbox_ingest_prefetched: message k + 1 starts its copy BEFORE tick k and takes effect AFTER
    it (looked up against the slots tick k left behind) -- same ids, erase / add lists and bits as te_pool_mailbox_ingest called
    after tick k; shuffled, duplicated, partly stale messages, first sights, an empty message."""
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model("uniform_acceleration")
    rng = np.random.default_rng(33)
    universe = rng.choice(50000, size=2500, replace=False).astype(np.uint32)
    ticks = 24
    streams, _, _ = synth.make_streams(universe.size, ticks, DT, accel=True, angular=False, seed=9)
    live = list(range(600)); nxt = 600
    msgs = []
    for k in range(ticks):
        now = 1000 * 10**9 + k * 4_000_000
        gone = set(j for j in live if rng.random() < 0.03)
        live = [j for j in live if j not in gone] + list(range(nxt, nxt + 30)); nxt += 30
        speak = np.array([j for j in live if rng.random() < 0.9]); rng.shuffle(speak)
        speak = np.concatenate([speak, speak[:5]])
        if k == 7:
            speak = speak[:0]                                                  # an empty message
        st_ns = now - np.where(rng.random(speak.size) < 0.05, 12_000_000, 0)
        msgs.append((universe[speak], (st_ns // 10**9).astype(np.uint32), (st_ns % 10**9).astype(np.uint32), np.ascontiguousarray(streams[k, speak]), now))
    results = []
    for prefetch in (False, True):
        pool = te.TargetPool(mtype); pool.register_class(Q, R, P0)
        log = []
        if prefetch:
            pool.mailbox_prefetch(*msgs[0][:4])
        for k in range(ticks):
            ids, sec, nsec, poses, now = msgs[k]
            if prefetch:
                pool.mailbox_ingest_prefetched()
                if k + 1 < ticks:
                    pool.mailbox_prefetch(*msgs[k + 1][:4])                    # in flight during the tick below
            else:
                pool.mailbox_ingest(ids, sec, nsec, poses)
            log.append(pool.mailbox_tick(DT, k * DT, (now // 10**9, now % 10**9), 6 * DT, want_added=True))
        results.append((pool.ids(), pool.read_state(), log, pool.mailbox_count()))
        pool.close()
    (ids_a, st_a, log_a, mc_a), (ids_b, st_b, log_b, mc_b) = results
    assert np.array_equal(ids_a, ids_b) and ids_a.size > 500 and mc_a == mc_b
    for (ea, aa), (eb, ab) in zip(log_a, log_b):
        assert np.array_equal(ea, eb) and np.array_equal(aa, ab)
    assert sum(e.size for e, _ in log_a) > 50
    for key in ("x", "P", "t", "n_meas", "measured_pose"):
        assert np.array_equal(st_a[key], st_b[key]), key
