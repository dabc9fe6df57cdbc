"""-m gpu: the reference-facing C-ABI (include/target_manager_c.h, lib/libtarget_c.so) against the oracle's
TargetManager on identical inputs -- reads like the reference's own test/target_manager_test.cpp, plus the paths
that test never touches (erase, predict-only, unknown ids, stale getter scratch, batched extensions)."""
import os

import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 1.0 / 250.0
CASES = [(0, "uniform_velocity"), (1, "uniform_acceleration"), (2, "angular_rates"), (3, "angular_velocities")]


def _yaml(name):
    return os.path.join(ROOT, "models", "model_%s_params.yaml" % name)


def _reftest(n_points):
    meas = np.zeros((4, n_points, 7)); real = np.zeros((4, n_points, 7))
    orc.lib().orc_reftest_streams(DT, n_points, 4, orc.ptr(meas), orc.ptr(real))
    return meas


@pytest.mark.parametrize("idx,name", CASES)
def test_reference_scenario_through_c_abi(idx, name):
    """BASELINE configs[0]: target_manager_test (1 target, 10 000 steps, libstdc++ noise stream), per-id calls
    update_meas -> get_est_pose -> get_est_twist every step, against the oracle and the reference tolerances."""
    from target_estimation_b200.manager import TargetManagerC
    n = 10000
    meas = _reftest(n)[idx]
    mgr = TargetManagerC(_yaml(name))
    ref = orc.Manager(_yaml(name))
    mgr.init(idx, DT, meas[0], 0.0)
    ref.init_default(idx, DT, meas[0], 0.0)
    pose = np.zeros((n, 7)); twist = np.zeros((n, 6))
    rpose = np.zeros((n, 7)); rtwist = np.zeros((n, 6))
    N = orc.load_yaml(_yaml(name))["Q"].shape[0]
    xs = np.zeros((n // 500, N)); Ps = np.zeros((n // 500, N, N))
    orc.lib().orc_run_stream(ref.h, idx, n, DT, orc.ptr(np.ascontiguousarray(meas)), None, 500, orc.ptr(xs), orc.ptr(Ps), orc.ptr(rpose),
                             orc.ptr(rtwist))
    k_rec = 0
    for k in range(n):
        mgr.update_meas(idx, DT, meas[k])
        ok1, _ = mgr.get_est_pose(idx, pose[k])
        ok2, _ = mgr.get_est_twist(idx, twist[k])
        assert ok1 and ok2
        if (k + 1) % 500 == 0:
            st = mgr.state(idx)
            assert synth.compare_h2(st["x"][None], xs[k_rec][None]) <= 1.0, k
            assert synth.compare_h2(st["P"][None], Ps[k_rec][None]) <= 1.0, k
            k_rec += 1
    assert mgr.get_n_measurements(idx) == n == ref.n_measurements(idx)
    # reference tolerances (test/target_manager_test.cpp:179-189, 335-340)
    goal = np.array([0.2, 0.3, 0.4])
    assert np.all(np.abs(pose[-1, :3] - goal) < 0.01)
    assert np.all(np.abs(twist[:, :3].mean(axis=0) - goal / (n * DT)) < 0.01)
    if name == "angular_velocities":
        omega = np.array([3.0, 0.01, 0.1])
        assert np.all(np.abs(twist[:, 3:].mean(axis=0) - omega) < 0.05) and np.all(np.abs(twist[-1, 3:] - omega) < 0.01)
    # and against the oracle, every step: positions / twists 1e-9 relative (scale = max |ref|)
    assert np.abs(pose[:, :3] - rpose[:, :3]).max() <= 1e-9 * max(1.0, np.abs(rpose[:, :3]).max())
    assert np.abs(twist - rtwist).max() <= 1e-9 * max(1.0, np.abs(rtwist).max())
    # quaternion outputs: sign-insensitive
    dq = np.minimum(np.abs(pose[:, 3:] - rpose[:, 3:]).max(axis=1), np.abs(pose[:, 3:] + rpose[:, 3:]).max(axis=1))
    assert dq.max() <= 1e-9
    mgr.close()


def test_registry_semantics():
    """init twice, unknown ids, erase, predict-only, ascending ids, getter scratch (src/target_manager.cpp:144-295,
    src/target_manager_c.cpp:37-59)."""
    from target_estimation_b200.manager import TargetManagerC
    name = "uniform_acceleration"
    mgr = TargetManagerC(_yaml(name)); ref = orc.Manager(_yaml(name))
    rng = np.random.default_rng(11)
    p = rng.normal(size=(6, 7)); p[:, 3:] = [0, 0, 0, 1]
    for i, id_ in enumerate([42, 7, 1000000, 3]):
        mgr.init(id_, DT, p[i], 0.5 * i); ref.init_default(id_, DT, p[i], 0.5 * i)
    mgr.init(7, DT, p[5], 9.0); ref.init_default(7, DT, p[5], 9.0)            # "already exists": no-op
    assert list(mgr.ids()) == list(ref.ids()) == [3, 7, 42, 1000000]
    for k in range(30):
        for id_ in (42, 7, 3):
            m = p[0] + 0.01 * k
            if (k + id_) % 4 == 0:
                mgr.update(id_, DT); ref.update(id_, DT)                         # predict only
            else:
                mgr.update_meas(id_, DT, m); ref.update_meas(id_, DT, m)
        mgr.update_meas(555, DT, p[0])                                           # unknown id: skipped
        assert not ref.update_meas(555, DT, p[0])
    for id_ in (3, 7, 42, 1000000):
        st, rs = mgr.state(id_), ref.state(id_, 9)
        assert synth.compare_h2(st["x"][None], rs["x"][None]) <= 1.0
        assert synth.compare_h2(st["P"][None], rs["P"][None]) <= 1.0
        assert st["t"] == rs["t"]
        assert mgr.get_n_measurements(id_) == ref.n_measurements(id_)
    # unknown id: false and the previous successful value stays in the output
    ok, good = mgr.get_est_pose(42)
    ok2, stale = mgr.get_est_pose(999)
    assert ok and not ok2 and np.array_equal(good, stale)
    assert mgr.get_n_measurements(999) == 0
    # update(dt) for all targets
    mgr.update_all(DT); ref.update_all(DT)
    for id_ in (3, 1000000):
        assert synth.compare_h2(mgr.state(id_)["P"][None], ref.state(id_, 9)["P"][None]) <= 1.0
    # erase
    assert mgr.erase(7) and ref.erase(7)
    assert not mgr.erase(7) and not ref.erase(7)
    assert list(mgr.ids()) == list(ref.ids()) == [3, 42, 1000000]
    assert mgr.state(7) is None
    st, rs = mgr.state(42), ref.state(42, 9)
    assert synth.compare_h2(st["P"][None], rs["P"][None]) <= 1.0
    # re-create the erased id: fresh filter
    mgr.init(7, DT, p[2], 1.0); ref.init_default(7, DT, p[2], 1.0)
    assert mgr.get_n_measurements(7) == 0
    assert synth.compare_h2(mgr.state(7)["x"][None], ref.state(7, 9)["x"][None]) <= 1.0
    mgr.close()


@pytest.mark.parametrize("name", ["uniform_velocity", "angular_rates"])
def test_batched_extensions(name):
    from target_estimation_b200.manager import TargetManagerC
    y = orc.load_yaml(_yaml(name))
    N = y["Q"].shape[0]
    n, ticks = 300, 25
    meas, action, _ = synth.make_streams(n, ticks, DT, accel=name == "angular_rates", angular=y["R"].shape[0] == 6, seed=5)
    rng = np.random.default_rng(3)
    ids = rng.permutation(np.arange(n, dtype=np.uint32) * 5 + 1)               # unsorted ids
    mgr = TargetManagerC(_yaml(name)); ref = orc.Manager(_yaml(name))
    t0 = rng.uniform(0, 2, n)
    assert mgr.init_batch(ids, DT, meas[0], t0) == n
    assert mgr.init_batch(ids[:10], DT, meas[0][:10]) == 0                      # all exist
    for k in range(n):
        ref.init_default(int(ids[k]), DT, meas[0, k], float(t0[k]))
    assert list(mgr.ids()) == sorted(ids.tolist())
    for k in range(ticks):
        act = action[k].copy()
        act[rng.uniform(size=n) < 0.1] = 0                                       # some targets untouched this tick
        assert mgr.update_batch(ids, DT, meas[k], act) == int((act != 0).sum())
        ref.step_batch(ids, DT, meas[k], act)
    order = np.argsort(ids)
    got_pose, got_twist, got_acc, found = mgr.get_estimates_batch(ids)
    assert found.all()
    for j in order[:: max(1, n // 40)]:
        st, rs = mgr.state(int(ids[j])), ref.state(int(ids[j]), N)
        assert synth.compare_h2(st["x"][None], rs["x"][None]) <= 1.0
        assert synth.compare_h2(st["P"][None], rs["P"][None]) <= 1.0
        ok, rp = ref.pose(int(ids[j])); _, rt = ref.twist(int(ids[j])); _, ra = ref.acc(int(ids[j]))
        assert np.abs(got_pose[j, :3] - rp[:3]).max() <= 1e-9 * max(1, np.abs(rp[:3]).max())
        assert min(np.abs(got_pose[j, 3:] - rp[3:]).max(), np.abs(got_pose[j, 3:] + rp[3:]).max()) <= 1e-9
        assert np.abs(got_twist[j] - rt).max() <= 1e-9 * max(1, np.abs(rt).max())
        assert np.abs(got_acc[j] - ra).max() <= 1e-9 * max(1, np.abs(ra).max())
    # erase a batch (with unknown ids mixed in) and compare the surviving id set
    gone = np.concatenate([ids[::7], np.array([999999], dtype=np.uint32)])
    assert mgr.erase_batch(gone) == len(ids[::7])
    for g in ids[::7]:
        ref.erase(int(g))
    assert np.array_equal(mgr.ids(), ref.ids())
    gone_set = set(ids[::7].tolist())
    j = next(int(o) for o in order if int(ids[o]) not in gone_set)
    st, rs = mgr.state(int(ids[j])), ref.state(int(ids[j]), N)
    assert synth.compare_h2(st["P"][None], rs["P"][None]) <= 1.0
    mgr.close()


@pytest.mark.parametrize("name", ["uniform_acceleration", "angular_rates"])
def test_update_batch_with_repeated_ids(name):
    """target_manager_update_batch with an id named several times in one batch = the reference's sequential update() calls:
    every record applied, in order (the batch is cut in front of each repeat)."""
    from target_estimation_b200.manager import TargetManagerC
    y = orc.load_yaml(_yaml(name))
    N = y["Q"].shape[0]
    n, ticks = 64, 6
    meas, action, _ = synth.make_streams(n, 4 * ticks, DT, accel=True, angular=y["R"].shape[0] == 6, seed=9)
    ids = np.arange(n, dtype=np.uint32) * 2 + 3
    mgr = TargetManagerC(_yaml(name)); ref = orc.Manager(_yaml(name))
    assert mgr.init_batch(ids, DT, meas[0]) == n
    for k in range(n):
        ref.init_default(int(ids[k]), DT, meas[0, k], 0.0)
    rng = np.random.default_rng(4)
    for k in range(ticks):
        # three records for a third of the ids, two for another third, shuffled into one batch
        reps = np.where(np.arange(n) % 3 == 0, 3, np.where(np.arange(n) % 3 == 1, 2, 1))
        rows = np.repeat(np.arange(n), reps)
        occ = np.concatenate([np.arange(r) for r in reps])
        order = rng.permutation(rows.size)
        order = order[np.argsort(occ[order], kind="stable")] if k % 2 else order    # odd ticks: all first records, then the repeats
        b_ids = ids[rows[order]]
        b_meas = np.stack([meas[4 * k + occ[j], rows[j]] for j in order])
        b_act = np.where(rng.random(order.size) < 0.2, 1, 2).astype(np.uint8)
        assert mgr.update_batch(b_ids, DT, b_meas, b_act) == order.size
        ref.step_batch(b_ids, DT, b_meas, b_act)
    for j in range(n):
        st, rs = mgr.state(int(ids[j])), ref.state(int(ids[j]), N)
        assert mgr.get_n_measurements(int(ids[j])) == rs["n_meas"] and st["t"] == rs["t"]
        assert synth.compare_h2(st["x"][None], rs["x"][None]) <= 1.0 and synth.compare_h2(st["P"][None], rs["P"][None]) <= 1.0
    mgr.close()


@pytest.mark.parametrize("name", ["uniform_acceleration", "angular_rates"])
def test_sampled_logging_and_text_dumps(name, tmp_path):
    """target_manager_watch / target_manager_log / target_manager_write_log: the five quantities the reference publishes per
    target under LOGGER_ON (measured_pose_, pose_internal_, twist_, acceleration_, P_; src/target_interface.cpp:32-40) sampled
    once per log() call, against the oracle's getters; dumps in writeTxtFile's format under the names matlab/plot_*.m load."""
    from target_estimation_b200.manager import TargetManagerC
    y = orc.load_yaml(_yaml(name)); N = y["Q"].shape[0]
    mgr = TargetManagerC(_yaml(name)); ref = orc.Manager(_yaml(name))
    ids = np.array([4, 9, 17], dtype=np.uint32)
    ticks = 25
    meas, action, _ = synth.make_streams(ids.size, ticks, DT, accel=True, angular=name.startswith("angular"), seed=3)
    watched = np.array([9, 555, 4], dtype=np.uint32)      # 555 never exists; 4 is erased half way
    mgr.watch(watched)
    for j, i in enumerate(ids):
        mgr.init(int(i), DT, meas[0, j], 0.0); ref.init_default(int(i), DT, meas[0, j], 0.0)
    for k in range(ticks):
        if k == 12:
            assert mgr.erase(4) and ref.erase(4)
        live = [(j, int(i)) for j, i in enumerate(ids) if not (i == 4 and k >= 12)]
        for j, i in live:
            if action[k, j] == 2:
                mgr.update_meas(i, DT, meas[k, j]); ref.update_meas(i, DT, meas[k, j])
            else:
                mgr.update(i, DT); ref.update(i, DT)
        mgr.log()
        assert mgr.log_samples() == k + 1
        for jw, i in enumerate(watched):
            s = mgr.log_sample(k, jw)
            if i == 555 or (i == 4 and k >= 12):
                assert s is None
                continue
            st = ref.state(int(i), N)
            assert s["t"] == st["t"]
            assert np.array_equal(s["measured_pose"], ref.measured_pose(int(i))[1])
            for key, want in (("pose_internal", ref.pose_internal(int(i))[1]), ("twist", ref.twist(int(i))[1]), ("acceleration", ref.acc(int(i))[1])):
                assert np.abs(s[key] - want).max() <= 1e-9 * max(1.0, np.abs(want).max()), (k, i, key)
            assert synth.compare_h2(s["P"][None], st["P"][None]) <= 1.0
    folder = str(tmp_path) + "/"
    assert mgr.write_log(folder) == 12                    # 6 files for each of the two ids that ever existed
    t9 = np.loadtxt(folder + "time_9"); p9 = np.loadtxt(folder + "est_pose_9"); m4 = np.loadtxt(folder + "meas_pose_4")
    assert t9.shape == (ticks,) and p9.shape == (ticks, 6) and m4.shape == (12, 7) and not os.path.exists(folder + "time_555")
    assert np.loadtxt(folder + "cov_diag_9").shape == (ticks, N) and np.loadtxt(folder + "est_twist_9").shape == (ticks, 6)
    assert np.allclose(p9[-1, :3], ref.pose_internal(9)[1][:3], rtol=1e-5, atol=1e-6)      # 6 significant digits in the text form
    mgr.watch([])                                         # stop: the series is dropped
    mgr.log()
    assert mgr.log_samples() == 0
    mgr.close()


@pytest.mark.parametrize("stride", [7, 3])
def test_dense_tick_through_c_abi(stride):
    """target_manager_update_dense(_async): record k = the k-th id of target_manager_get_dense_ids, no ids travel; state against the
    oracle, the returned positions against the getters, the pipelined form against the synchronous one (bit-identical)"""
    from target_estimation_b200.manager import TargetManagerC
    name = "uniform_acceleration"
    y = orc.load_yaml(_yaml(name))
    n, ticks = 5000 + 3, 8
    meas, action, _ = synth.make_streams(n, ticks, DT, accel=True, angular=False, seed=41)
    ids = np.random.default_rng(4).permutation(np.arange(n, dtype=np.uint32) * 3 + 2)
    order = np.argsort(ids)                       # dense order of a plain manager = ascending ids
    ref = orc.ShardedManager()
    ref.init_batch(y["type"], ids, DT, y["Q"], y["R"], y["P"], meas[0])
    outs = {}
    for pipelined in (False, True):
        mgr = TargetManagerC(_yaml(name))
        assert mgr.init_batch(ids, DT, meas[0]) == n
        assert np.array_equal(mgr.dense_ids(), ids[order])
        pos = [np.zeros((n, 3)) for _ in range(ticks)]
        for k in range(ticks):
            m = np.ascontiguousarray(meas[k][order][:, :stride])
            a = np.ascontiguousarray(action[k][order])
            assert mgr.update_dense(DT, m, a, pos[k], pipelined=pipelined) == n
            if pipelined:
                mgr.update_dense_wait(1)
        mgr.update_dense_wait(0)
        pose, _, _, found = mgr.get_estimates_batch(ids[order])
        assert found.all() and np.array_equal(pose[:, :3], pos[-1])
        outs[pipelined] = (pos, mgr.state(int(ids[5])))
        mgr.close()
    for k in range(ticks):
        assert np.array_equal(outs[False][0][k], outs[True][0][k]), k
        ref.step_batch(ids, DT, meas[k], action[k])
        want = ref.states(ids[order], y["Q"].shape[0])["x"][:, :3]
        assert synth.compare_h2(outs[False][0][k], want) <= 1.0, k
    st = outs[True][1]
    want = ref.states(ids[5:6], y["Q"].shape[0])
    assert synth.compare_h2(st["x"][None], want["x"]) <= 1.0 and synth.compare_h2(st["P"][None], want["P"]) <= 1.0
    ref.close()
