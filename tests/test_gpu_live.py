"""-m gpu: live launches (te_pool_live_*): one resident launch holds every target of a small pool in registers and applies ticks as
they are released -- BASELINE configs[1]'s 250 Hz loop over 10 000 uniform-acceleration targets without a launch per tick.  Bit-identical
to the same ticks through te_pool_step_dense, every tick's published positions included; against the oracle at the 1e-9 bar."""
import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
DT = 1.0 / 250.0


def _pools(model, n, ticks, seed):
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(model)
    N, M = te.model_dims(mtype)
    meas, action, scale = synth.make_streams(n, ticks, DT, accel=model == "uniform_acceleration", angular=False, seed=seed)
    action[3:, ::7] = 0          # some targets untouched on some ticks
    ids = np.arange(n, dtype=np.uint32) * 2 + 5
    pools = []
    for _ in range(2):
        p = te.TargetPool(mtype); p.register_class(Q, R, P0); p.add(ids, meas[0], p0_scale=scale)
        pools.append(p)
    return te, pools, ids, meas, action, scale, (mtype, Q, R, P0, N)


@pytest.mark.parametrize("model,n,stride", [("uniform_acceleration", 10000, 7), ("uniform_velocity", 3001, 3), ("uniform_acceleration", 20011, 3)])
def test_live_ticks_equal_sequential_ticks(model, n, stride):
    """closed loop: push tick k (host arrays -> rings, release), wait for it, read its positions; then the state"""
    import torch
    ticks = 30
    te, (seq, live), ids, meas, action, scale, (mtype, Q, R, P0, N) = _pools(model, n, ticks, 51)
    d_meas = torch.zeros((ticks + 5, n, stride), dtype=torch.float64, device="cuda")
    d_act = torch.zeros((ticks + 5, n), dtype=torch.uint8, device="cuda")
    d_pos = torch.zeros((ticks + 5, n, 3), dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()          # (created BEFORE the launch: creating a stream waits for the device)
    torch.cuda.synchronize()
    live.live_begin(ticks + 5, DT, d_meas, stride, d_act, te.ACT_UPDATE, d_pos)
    with pytest.raises(te.TeError):
        live.step_dense(DT, d_meas[0], stride, d_act[0])          # the pool is held by the live launch
    snaps = {}
    # (nothing that synchronises the whole device may run while the launch is resident -- a cudaFree or a stream creation would
    #  wait for it forever -- so the sequential pool takes its ticks afterwards)
    for k in range(ticks):
        assert live.live_push(np.ascontiguousarray(meas[k][:, :stride]), action[k]) == k + 1
        assert live.live_wait(k + 1) >= k + 1
        if k % 7 == 3:                                           # tick k's positions are complete while the launch keeps running
            with torch.cuda.stream(side):
                snaps[k] = d_pos[k].cpu().numpy()
    assert live.live_end() == ticks                              # five unreleased ticks are skipped
    for k in range(ticks):
        seq.step_dense(DT, torch.from_numpy(np.ascontiguousarray(meas[k][:, :stride])).cuda(), stride, torch.from_numpy(action[k]).cuda())
        if k in snaps:
            assert np.array_equal(snaps[k], seq.read_state(ids)["x"][:, :3]), k
    a, b = seq.read_state(ids), live.read_state(ids)
    for key in ("x", "P", "t", "n_meas"):
        assert np.array_equal(a[key], b[key]), key
    mgr = orc.ShardedManager()
    mgr.init_batch(mtype, ids, DT, Q, R, P0, meas[0], scale)
    mgr.step_ticks(ids, DT, meas[:ticks], action[:ticks])
    ref = mgr.states(ids, N)
    assert synth.compare_h2(b["x"], ref["x"]) <= 1.0 and synth.compare_h2(b["P"], ref["P"]) <= 1.0
    # the pool is an ordinary pool again
    live.step_dense(DT, torch.from_numpy(np.ascontiguousarray(meas[0][:, :stride])).cuda(), stride, None)
    for p in (seq, live):
        p.close()
    mgr.close()


def test_live_release_ahead_and_limits():
    """ticks whose blocks the caller wrote itself, released several at once; argument checks"""
    import torch
    ticks = 12
    te, (seq, live), ids, meas, action, scale, _ = _pools("uniform_acceleration", 2000, ticks, 52)
    d_meas = torch.from_numpy(np.ascontiguousarray(meas[:ticks])).cuda()
    d_act = torch.from_numpy(np.ascontiguousarray(action[:ticks])).cuda()
    torch.cuda.synchronize()
    live.live_begin(ticks, DT, d_meas, 7, d_act, te.ACT_UPDATE, None)
    with pytest.raises(te.TeError):
        live.live_begin(ticks, DT, d_meas, 7, d_act, te.ACT_UPDATE, None)      # one at a time
    with pytest.raises(te.TeError):
        live.live_release(ticks + 1)
    live.live_release(5)
    assert live.live_wait(5) >= 5
    live.live_release(ticks)
    assert live.live_wait(ticks) == ticks
    assert live.live_end() == ticks
    for k in range(ticks):
        seq.step_dense(DT, d_meas[k], 7, d_act[k])
    a, b = seq.read_state(ids), live.read_state(ids)
    for key in ("x", "P", "t", "n_meas"):
        assert np.array_equal(a[key], b[key]), key
    # an angular pool cannot go live (its step does not keep the target in registers)
    mtype, _, Q, R, P0 = te.load_model("angular_rates")
    p = te.TargetPool(mtype); p.register_class(Q, R, P0); p.add(np.arange(10, dtype=np.uint32), np.tile([0, 0, 0, 0, 0, 0, 1.0], (10, 1)))
    with pytest.raises(te.TeError):
        p.live_begin(4, DT, None, 7, None, te.ACT_PREDICT, None)
    for q in (seq, live, p):
        q.close()
