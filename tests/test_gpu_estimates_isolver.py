"""-m gpu: derived outputs (getEstimatedPose/Twist/Acceleration([t1]), measured pose, pose_internal) and the
batched IntersectionSolver against the oracle."""
import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu
DT = 1.0 / 250.0
MODELS = ["uniform_velocity", "uniform_acceleration", "angular_velocities", "angular_rates"]


def _setup(name, n, ticks, seed=9):
    import target_estimation_b200 as te
    mtype, _, Q, R, P0 = te.load_model(name)
    meas, action, scale = synth.make_streams(n, ticks, DT, accel=name in ("uniform_acceleration", "angular_rates"), angular=R.shape[0] == 6,
                                             seed=seed)
    ids = np.arange(n, dtype=np.uint32) * 2 + 10
    ref = orc.Manager()
    for k in range(n):
        ref.init_full(mtype, int(ids[k]), DT, 0.0, Q, R, scale[k] * P0, meas[0, k])
    pool = te.TargetPool(mtype)
    pool.register_class(Q, R, P0)
    pool.add(ids, meas[0], p0_scale=scale)
    for k in range(ticks):
        ref.step_batch(ids, DT, meas[k], action[k])
        pool.step_ids(ids, DT, meas[k], action[k])
    return te, pool, ref, ids, meas, action


def _close(a, b, tol=1e-9):
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def _qclose(a, b, tol=1e-9):
    return min(np.abs(a - b).max(), np.abs(a + b).max()) <= tol


@pytest.mark.parametrize("name", MODELS)
def test_estimates(name):
    te, pool, ref, ids, meas, action = _setup(name, 64, 40)
    cur = pool.read_estimates(ids)
    st = pool.read_state(ids)
    t1 = st["t"] + 0.37
    fut = pool.read_estimates(ids, t1=t1)
    for j, i in enumerate(ids):
        i = int(i)
        _, p = ref.pose(i); _, tw = ref.twist(i); _, ac = ref.acc(i); _, p6 = ref.pose_internal(i); _, mp = ref.measured_pose(i)
        assert _close(cur["pose"][j, :3], p[:3]) and _qclose(cur["pose"][j, 3:], p[3:])
        assert _close(cur["twist"][j], tw) and _close(cur["acc"][j], ac) and _close(cur["pose6"][j], p6)
        assert np.array_equal(st["measured_pose"][j], mp)
        _, p = ref.pose_at(i, float(t1[j])); _, tw = ref.twist_at(i, float(t1[j])); _, ac = ref.acc_at(i, float(t1[j]))
        assert _close(fut["pose"][j, :3], p[:3]) and _qclose(fut["pose"][j, 3:], p[3:])
        assert _close(fut["twist"][j], tw) and _close(fut["acc"][j], ac)
    unk = pool.read_estimates(np.array([1, 10, 3], dtype=np.uint32))
    assert list(unk["found"]) == [0, 1, 0]
    pool.close()


@pytest.mark.parametrize("name", MODELS)
def test_intersection_solver(name):
    """IntersectionSolver semantics incl. H9: UV / AV always -1 (zero t^4 coefficient); one solver object shares
    its filters across every id queried through it (stream 0 for all queries, one query per call)."""
    te, pool, ref, ids, meas, action = _setup(name, 24, 30, seed=21)
    solver = te.IntersectionSolver(pool, n_streams=1, filters_length=8)
    rs = orc.lib().orc_isolver_new(ref.h, 8)
    rng = np.random.default_rng(2)
    st = pool.read_state(ids)
    n_found = 0
    for rep in range(3):
        for j, i in enumerate(ids):
            # put the sphere ahead of the target so that an interception exists for accelerating models
            ok, p = ref.pose_at(int(i), float(st["t"][j]) + 0.3)
            origin = p[:3] + rng.normal(0, 0.05, 3)
            radius = float(rng.uniform(0.1, 0.5))
            t1 = float(st["t"][j])
            rp = np.zeros(7)
            rt = orc.lib().orc_isolver_time(rs, int(i), t1, orc.ptr(origin), radius)
            rc = orc.lib().orc_isolver_pose(rs, int(i), t1, 0.05, 0.1, orc.ptr(origin), radius, orc.ptr(rp))
            d, pose, conv = solver.query([i], t1, origin, radius, 0.05, 0.1, stream=[0])
            if rt < 0:
                assert d[0] == -1.0
            else:
                n_found += 1
                assert abs(d[0] - rt) <= 1e-9 * max(1.0, abs(rt)), (d[0], rt)
            assert _close(pose[0, :3], rp[:3], 1e-8) and _qclose(pose[0, 3:], rp[3:], 1e-8)
            assert bool(conv[0]) == bool(rc)
    if name in ("uniform_velocity", "angular_velocities"):
        assert n_found == 0
    else:
        assert n_found > 10
    # unknown id -> -1 / identity pose / not converged
    d, pose, conv = solver.query([5], 0.0, [0, 0, 0], 1.0, 0.05, 0.1, stream=[0])
    assert d[0] == -1.0 and list(pose[0]) == [0, 0, 0, 0, 0, 0, 1] and not conv[0]
    orc.lib().orc_isolver_delete(rs)
    solver.close(); pool.close()


def test_intersection_batched_streams():
    """one query per target per call, one solver stream per target == one reference solver object per target"""
    te, pool, ref, ids, meas, action = _setup("uniform_acceleration", 50, 20, seed=4)
    n = len(ids)
    solver = te.IntersectionSolver(pool, n_streams=n, filters_length=5)
    refs = [orc.lib().orc_isolver_new(ref.h, 5) for _ in range(n)]
    rng = np.random.default_rng(8)
    for rep in range(9):
        st = pool.read_state(ids)
        origin = np.zeros((n, 3)); radius = rng.uniform(0.2, 0.6, n)
        for j, i in enumerate(ids):
            _, p = ref.pose_at(int(i), float(st["t"][j]) + 0.25)
            origin[j] = p[:3] + rng.normal(0, 0.02, 3)
        d, pose, conv = solver.query(ids, st["t"], origin, radius, 0.5, 0.5)
        for j, i in enumerate(ids):
            rp = np.zeros(7)
            rc = orc.lib().orc_isolver_pose(refs[j], int(i), float(st["t"][j]), 0.5, 0.5, orc.ptr(origin[j]), float(radius[j]), orc.ptr(rp))
            assert _close(pose[j, :3], rp[:3], 1e-8) and bool(conv[j]) == bool(rc), (rep, j)
        k = 20 + rep
        m = meas[k % meas.shape[0]]
        ref.step_batch(ids, DT, m, action[0]); pool.step_ids(ids, DT, m, action[0])
    for r in refs:
        orc.lib().orc_isolver_delete(r)
    solver.close(); pool.close()


def test_dense_entry_points_match_id_entry_points():
    """te_isolver_query_dense / te_pool_stamp_dense (device-resident forms) == the id-based host forms"""
    import torch
    te, pool, ref, ids, meas, action = _setup("uniform_acceleration", 70, 15, seed=31)
    n = len(ids)
    s_id = te.IntersectionSolver(pool, n_streams=n, filters_length=7)
    s_dn = te.IntersectionSolver(pool, n_streams=n, filters_length=7)
    rng = np.random.default_rng(4)
    st = pool.read_state(ids)
    origin = np.zeros((n, 3)); radius = rng.uniform(0.2, 0.6, n)
    for j, i in enumerate(ids):
        _, p = ref.pose_at(int(i), float(st["t"][j]) + 0.25)
        origin[j] = p[:3] + rng.normal(0, 0.02, 3)
    d_o, d_r = torch.from_numpy(origin).cuda(), torch.from_numpy(radius).cuda()
    d_delta = torch.empty(n, dtype=torch.float64, device="cuda"); d_pose = torch.empty((n, 7), dtype=torch.float64, device="cuda")
    d_conv = torch.empty(n, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        d, pose, conv = s_id.query(ids, st["t"], origin, radius, 0.5, 0.5)
        s_dn.query_dense(d_o, d_r, 0.5, 0.5, None, d_delta, d_pose, d_conv)
        pool.sync()
        assert np.array_equal(d, d_delta.cpu().numpy()) and np.array_equal(pose, d_pose.cpu().numpy())
        assert np.array_equal(conv, d_conv.cpu().numpy())
    # stamp_dense + expire == set_stamps + expire
    act = np.where(np.arange(n) % 3 == 0, 1, 2).astype(np.uint8)
    pool.stamp_dense(1000, 0, torch.from_numpy(act).cuda())
    gone = pool.expire(1000, 50000000, 0.04)     # 1000.05 - 1000.0 >= 0.04 (1000.04 - 1000.0 rounds to just below 0.04)
    assert np.array_equal(gone, ids[act == 2])
    s_id.close(); s_dn.close(); pool.close()
