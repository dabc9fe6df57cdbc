"""-m gpu: the tick front-end (TickTargetManager: /tf mailboxes, init on first sight, sticky measurements, stale
stamps, expiry, frame-name parsing incl. the loop-breaking "<token>_filt_<id>" frames) against the oracle's
restatement of RosTargetManager (src/target_manager_ros.cpp:26-92, target_manager_ros.hpp:74-134)."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 1.0 / 250.0


def _stamp(k, extra_ns=0):
    ns = 1000 * 10 ** 9 + k * 4000000 + extra_ns
    return ns // 10 ** 9, ns % 10 ** 9


@pytest.mark.parametrize("name", ["uniform_acceleration", "angular_velocities"])
def test_tick_semantics(name):
    from target_estimation_b200.manager import TickManagerC
    path = os.path.join(ROOT, "models", "model_%s_params.yaml" % name)
    y = orc.load_yaml(path)
    N = y["Q"].shape[0]
    L = orc.lib()
    h = L.orc_tick_new(y["type"], orc.ptr(orc.colmajor(y["Q"])), N, orc.ptr(orc.colmajor(y["R"])), y["R"].shape[0], orc.ptr(orc.colmajor(y["P"])))
    mgr = TickManagerC(path)
    timeout = 6 * DT
    L.orc_tick_set_expiration(h, timeout); mgr.set_expiration(timeout)
    rng = np.random.default_rng(5)
    n_ids = 40
    base = np.hstack([rng.normal(size=(n_ids, 3)), synth.rpy_to_quat(rng.uniform(-0.4, 0.4, (n_ids, 3)))])
    total_erased = 0
    for k in range(60):
        sec, nsec = _stamp(k)
        # which ids speak this tick: everybody early on; later some fall silent for good, some send STALE stamps
        frames, stamps, poses = [], [], []
        for i in range(n_ids):
            silent = (i % 5 == 0 and k > 10 + i % 7) and not (i == 10 and k > 45)     # id 10 comes back after expiring
            if silent:
                continue
            stale = (i % 4 == 1 and k % 3 == 0 and k > 0)                              # resend with an old stamp -> predict only
            s = _stamp(k - 1) if stale else (sec, nsec)
            frames.append("target_%d" % i); stamps.append(s)
            poses.append(base[i] + np.r_[0.01 * k, -0.005 * k, 0.002 * k, 0, 0, 0, 0] + np.r_[rng.normal(0, 0.01, 3), 0, 0, 0, 0])
            if i == 20 and k % 9 == 4:       # the node's own output frame: contains the token, 3 parts -> breaks the loop
                frames.append("target_filt_%d" % i); stamps.append((sec, nsec)); poses.append(base[i])
        if k % 6 == 5:                       # unrelated frames are ignored
            frames.insert(0, "camera_link"); stamps.insert(0, (sec, nsec)); poses.insert(0, base[0])
        stamps = np.array(stamps, dtype=np.uint32); poses = np.array(poses)
        L.orc_tick_callback(h, len(frames), "\n".join(frames).encode(), orc.ptr(np.ascontiguousarray(stamps)), orc.ptr(np.ascontiguousarray(poses)))
        mgr.callback_frames(frames, stamps[:, 0], stamps[:, 1], poses)
        er_ref = np.zeros(256, dtype=np.uint32)
        n_er = L.orc_tick_update(h, DT, sec, nsec, orc.ptr(er_ref), 256)
        er = mgr.tick(DT, sec, nsec)
        assert n_er == er.size and np.array_equal(er, er_ref[:n_er]), k
        total_erased += n_er
        ref_ids = np.zeros(256, dtype=np.uint32)
        n_ref = L.orc_get_ids(h, orc.ptr(ref_ids), 256)
        assert np.array_equal(mgr.ids(), ref_ids[:n_ref]), k
        assert mgr.mailboxes() == L.orc_tick_mailboxes(h)
        assert mgr.time() == L.orc_tick_time(h)
        pub_ids, pub_poses = mgr.published()
        assert np.array_equal(pub_ids, ref_ids[:n_ref])
        for j in range(0, n_ref, 5):
            p = np.zeros(7)
            L.orc_get_est_pose(h, int(ref_ids[j]), orc.ptr(p))
            assert np.abs(pub_poses[j, :3] - p[:3]).max() <= 1e-9 * max(1.0, np.abs(p[:3]).max())
            assert min(np.abs(pub_poses[j, 3:] - p[3:]).max(), np.abs(pub_poses[j, 3:] + p[3:]).max()) <= 1e-9
    assert total_erased >= 6
    # filter state of the survivors
    ids = mgr.ids()
    for i in ids[::3]:
        x = np.zeros(N); P = np.zeros((N, N)); t = C.c_double(); nm = C.c_longlong()
        L.orc_get_state(h, int(i), orc.ptr(x), orc.ptr(P), C.byref(t), C.byref(nm), None)
        st = mgr.state(int(i))
        assert synth.compare_h2(st["x"][None], x[None]) <= 1.0 and synth.compare_h2(st["P"][None], P[None]) <= 1.0
        assert st["t"] == t.value and mgr.get_n_measurements(int(i)) == nm.value
    mgr.close()


def test_tick_manager_by_hand_erase_and_init_like_the_reference():
    """target_manager_erase / target_manager_init on a tick manager: the erased target's mailbox survives (sticky, readable ->
    the target is re-created on the next tick from the stored pose, with t0 = the manager clock); a target created by hand is
    not stepped by the tick until its first /tf record."""
    from target_estimation_b200.manager import TickManagerC
    path = os.path.join(ROOT, "models", "model_uniform_acceleration_params.yaml")
    y = orc.load_yaml(path)
    N = y["Q"].shape[0]
    L = orc.lib()
    h = L.orc_tick_new(y["type"], orc.ptr(orc.colmajor(y["Q"])), N, orc.ptr(orc.colmajor(y["R"])), y["R"].shape[0], orc.ptr(orc.colmajor(y["P"])))
    ref = orc.Manager.__new__(orc.Manager); ref.L = L; ref.h = h
    mgr = TickManagerC(path)
    rng = np.random.default_rng(9)
    ids = np.array([3, 8, 21, 34, 55], dtype=np.uint32)
    pose = lambda n: np.hstack([rng.normal(size=(n, 3)), np.tile([0, 0, 0, 1.0], (n, 1))])

    def deliver(which, k):
        st = np.tile(np.array(_stamp(k), dtype=np.uint32), (len(which), 1)); ps = pose(len(which))
        w = np.ascontiguousarray(which, dtype=np.uint32)
        L.orc_tick_callback_ids(h, w.size, orc.ptr(w), orc.ptr(np.ascontiguousarray(st)), orc.ptr(ps))
        mgr.callback_ids(w, st[:, 0].copy(), st[:, 1].copy(), ps)

    for k in range(12):
        if k < 3 or k == 9:
            deliver(ids if k < 3 else [13], k)
        if k == 4:      # by hand: erase 8 and 34 (their mailboxes stay), create 13 (no mailbox until k = 9)
            for i in (8, 34):
                assert mgr.erase(i) and ref.erase(i)
            p13 = pose(1)[0]
            mgr.init(13, DT, p13, 0.25)
            ref.init_full(y["type"], 13, DT, 0.25, y["Q"], y["R"], y["P"], p13)
        sec, nsec = _stamp(k)
        er_ref = np.zeros(64, dtype=np.uint32)
        n_er = L.orc_tick_update(h, DT, sec, nsec, orc.ptr(er_ref), 64)
        er = mgr.tick(DT, sec, nsec)
        assert n_er == er.size == 0
        ref_ids = np.zeros(64, dtype=np.uint32)
        n_ref = L.orc_get_ids(h, orc.ptr(ref_ids), 64)
        assert np.array_equal(mgr.ids(), ref_ids[:n_ref]), k
        assert mgr.mailboxes() == L.orc_tick_mailboxes(h), k
        for i in ref_ids[:n_ref]:
            x = np.zeros(N); P = np.zeros((N, N)); t = C.c_double(); nm = C.c_longlong()
            L.orc_get_state(h, int(i), orc.ptr(x), orc.ptr(P), C.byref(t), C.byref(nm), None)
            st = mgr.state(int(i))
            assert synth.compare_h2(st["x"][None], x[None]) <= 1.0 and synth.compare_h2(st["P"][None], P[None]) <= 1.0, (k, i)
            assert st["t"] == t.value and mgr.get_n_measurements(int(i)) == nm.value, (k, i)
    assert mgr.ids().tolist() == [3, 8, 13, 21, 34, 55] and mgr.get_n_measurements(8) == 8 and mgr.get_n_measurements(13) == 3
    mgr.close(); ref.h = None; L.orc_manager_delete(h)
