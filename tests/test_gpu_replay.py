"""-m gpu: offline replay of the reference's recording (test/test_multiple_targets.bag, rebuilt here from the committed
golden records) through target_tick_manager_replay_bag -- the node's loop of src/target_node.cpp:36-44 on the bag's clock
-- against the same loop driven tick by tick through the oracle's restatement of RosTargetManager."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import bagfile, orc, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "bag_tf_records.npz")


@pytest.mark.parametrize("name,freq,timeout", [("uniform_acceleration", 250.0, 1.5), ("angular_rates", 100.0, 0.6), ("angular_velocities", 50.0, 30.0)])
def test_replay_recording(tmp_path, name, freq, timeout):
    from target_estimation_b200.manager import TickManagerC
    rec = np.load(GOLDEN)
    bag = str(tmp_path / "rebuilt.bag")
    msgs = bagfile.messages_from_records(rec)
    bagfile.write_bag(bag, msgs)
    path = os.path.join(ROOT, "models", "model_%s_params.yaml" % name)
    y = orc.load_yaml(path)
    N = y["Q"].shape[0]
    L = orc.lib()
    h = L.orc_tick_new(y["type"], orc.ptr(orc.colmajor(y["Q"])), N, orc.ptr(orc.colmajor(y["R"])), y["R"].shape[0], orc.ptr(orc.colmajor(y["P"])))
    mgr = TickManagerC(path)
    L.orc_tick_set_expiration(h, timeout); mgr.set_expiration(timeout)
    extra = 25
    stats = mgr.replay_bag(bag, freq, extra_ticks=extra)

    # the same loop on the oracle
    dt = 1.0 / freq
    period = int(round(1e9 / freq))
    t0 = msgs[0][0][0] * 10 ** 9 + msgs[0][0][1]
    nxt, k, left, erased_ref, ticks = 0, 0, extra, 0, 0
    er = np.zeros(16, dtype=np.uint32)
    while True:
        now = t0 + k * period
        erased_ref += L.orc_tick_update(h, dt, now // 10 ** 9, now % 10 ** 9, orc.ptr(er), 16)
        ticks += 1
        while nxt < len(msgs) and msgs[nxt][0][0] * 10 ** 9 + msgs[nxt][0][1] <= now:
            trs = msgs[nxt][1]
            stamps = np.array([[tr[1], tr[2]] for tr in trs], dtype=np.uint32)
            poses = np.array([tr[5] for tr in trs], dtype=np.float64)
            L.orc_tick_callback(h, len(trs), "\n".join(tr[4] for tr in trs).encode(), orc.ptr(np.ascontiguousarray(stamps)), orc.ptr(np.ascontiguousarray(poses)))
            nxt += 1
        if nxt >= len(msgs):
            if left <= 0:
                break
            left -= 1
        k += 1
    assert stats["ticks"] == ticks and stats["messages"] == len(msgs) == 572 and stats["transforms"] == 572
    assert stats["erased"] == erased_ref
    if timeout < 5:
        assert erased_ref >= 1          # target_2 is seen only ten times in the recording
    ref_ids = np.zeros(16, dtype=np.uint32)
    n_ref = L.orc_get_ids(h, orc.ptr(ref_ids), 16)
    ids = mgr.ids()
    assert np.array_equal(ids, ref_ids[:n_ref]) and n_ref >= 1
    assert mgr.time() == L.orc_tick_time(h) and mgr.mailboxes() == L.orc_tick_mailboxes(h)
    pub_ids, pub_poses = mgr.published()
    assert np.array_equal(pub_ids, ids)
    for j, i in enumerate(ids):
        x = np.zeros(N); P = np.zeros((N, N)); t = C.c_double(); nm = C.c_longlong()
        L.orc_get_state(h, int(i), orc.ptr(x), orc.ptr(P), C.byref(t), C.byref(nm), None)
        st = mgr.state(int(i))
        assert synth.compare_h2(st["x"][None], x[None]) <= 1.0 and synth.compare_h2(st["P"][None], P[None]) <= 1.0, (name, int(i))
        assert st["t"] == t.value and mgr.get_n_measurements(int(i)) == nm.value
        p = np.zeros(7)
        L.orc_get_est_pose(h, int(i), orc.ptr(p))
        assert np.abs(pub_poses[j, :3] - p[:3]).max() <= 1e-9 * max(1.0, np.abs(p[:3]).max())
        assert min(np.abs(pub_poses[j, 3:] - p[3:]).max(), np.abs(pub_poses[j, 3:] + p[3:]).max()) <= 1e-9
    mgr.close()
