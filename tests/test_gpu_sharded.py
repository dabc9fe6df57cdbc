"""-m gpu: the C++ ShardedTargetManager (target_manager_new_sharded of libtarget_c.so) against the oracle: one TargetManager per
shard, owner(id) = id mod G, batched calls routed on the host, per-id calls to the owner, the optional all-gather of estimate
records between the devices (NCCL between distinct devices; same-device shards -- the one-GPU form of this test -- use device
copies).  The world-size-2 gloo tests (tests/test_dist_gloo.py) cover the same routing rule in the torchrun harness.

The reference has one TargetManager (/root/reference/include/target_estimation/target_manager.hpp:66-203); every observable of
the sharded one -- ids in ascending order, per-target state, estimates, n_meas, "does not exist" -- must equal that of a single
manager fed the same calls."""
import numpy as np
import pytest

from tests import orc, synth

pytestmark = pytest.mark.gpu

DT = 1.0 / 250.0


def _yaml(name):
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "models", "model_%s_params.yaml" % name)


def _devices(n_shards, distinct):
    import torch
    nd = torch.cuda.device_count()
    if distinct:
        if nd < n_shards:
            pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (n_shards, n_shards))
        return list(range(n_shards))
    return [0] * n_shards


def _run(name, n_shards, devices, n=5003, ticks=12):
    from target_estimation_b200.manager import ShardedManagerC, TargetManagerC
    y = orc.load_yaml(_yaml(name))
    N = y["Q"].shape[0]
    angular = y["R"].shape[0] == 6
    meas, action, _ = synth.make_streams(n, ticks, DT, accel=name in ("uniform_acceleration", "angular_rates"), angular=angular, seed=77)
    rng = np.random.default_rng(3)
    ids = rng.permutation(np.arange(n, dtype=np.uint32) * 5 + 11)       # arbitrary order: routing must not depend on sorted input
    mgr = ShardedManagerC(_yaml(name), n_shards, devices)
    assert mgr.shards() == n_shards
    ref = orc.ShardedManager()
    ref.init_batch(y["type"], ids, DT, y["Q"], y["R"], y["P"], meas[0])
    assert mgr.init_batch(ids, DT, meas[0]) == n
    assert mgr.init_batch(ids[:100], DT, meas[0][:100]) == 0             # "already exists" in every shard
    assert np.array_equal(mgr.ids(), np.sort(ids))                        # ascending over all shards
    for k in range(ticks):
        order = rng.permutation(n)                                        # a differently ordered batch every tick
        assert mgr.update_batch(ids[order], DT, meas[k][order], action[k][order]) == n
        ref.step_batch(ids[order], DT, meas[k][order], action[k][order])
    # unknown ids are skipped, like the reference's "does not exist"
    extra = np.concatenate([ids[:50], np.array([3, 4, 2 ** 31 + 1], dtype=np.uint32)])
    assert mgr.update_batch(extra, DT, np.tile(meas[0][:1], (extra.size, 1)), np.full(extra.size, 1, dtype=np.uint8)) == 50
    ref.step_batch(ids[:50], DT, meas[0][:50], np.full(50, 1, dtype=np.uint8))
    # per-id calls go to the owner; a repeated id inside one batch is applied in order
    from target_estimation_b200 import ACT_UPDATE
    twice = np.array([ids[7], ids[8], ids[7]], dtype=np.uint32)
    m3 = np.stack([meas[1][7], meas[1][8], meas[2][7]])
    assert mgr.update_batch(twice, DT, m3, np.full(3, ACT_UPDATE, dtype=np.uint8)) == 3
    ref.step_batch(twice[:2], DT, m3[:2], np.full(2, ACT_UPDATE, dtype=np.uint8))
    ref.step_batch(twice[2:], DT, m3[2:], np.full(1, ACT_UPDATE, dtype=np.uint8))
    mgr.update_meas(int(ids[9]), DT, meas[3][9])
    mgr.update(int(ids[10]), DT)
    ref.step_batch(ids[9:10], DT, meas[3][9:10], np.full(1, ACT_UPDATE, dtype=np.uint8))
    ref.step_batch(ids[10:11], DT, meas[3][10:11], np.full(1, 1, dtype=np.uint8))
    want = ref.states(ids, N)
    for j in (0, 7, 8, 9, 10, n - 1):
        st = mgr.state(int(ids[j]))
        assert synth.compare_h2(st["x"][None], want["x"][j][None]) <= 1.0 and synth.compare_h2(st["P"][None], want["P"][j][None]) <= 1.0, j
        assert st["t"] == want["t"][j]
        assert mgr.get_n_measurements(int(ids[j])) == want["n_meas"][j]
    # estimates through the routed batch getter == a single manager on one device fed the same calls
    one = TargetManagerC(_yaml(name), device=devices[0])
    assert one.init_batch(ids, DT, meas[0]) == n
    for k in range(ticks):
        one.update_batch(ids, DT, meas[k], action[k])
    one.update_batch(ids[:50], DT, meas[0][:50], np.full(50, 1, dtype=np.uint8))
    one.update_batch(twice, DT, m3, np.full(3, ACT_UPDATE, dtype=np.uint8))
    one.update_meas(int(ids[9]), DT, meas[3][9]); one.update(int(ids[10]), DT)
    q = np.concatenate([ids[::3], np.array([1, 2], dtype=np.uint32)])
    t1 = 0.2
    pa, ta, aa, fa = mgr.get_estimates_batch(q, t1)
    pb, tb, ab, fb = one.get_estimates_batch(q, t1)
    assert np.array_equal(fa, fb) and fa[:-2].all() and not fa[-2:].any()
    assert np.array_equal(pa, pb) and np.array_equal(ta, tb) and np.array_equal(aa, ab)      # same kernels, same order per target: bit-identical
    # the all-gather of [pose7 | twist6] records: every target exactly once, shard-major, records = the getters' values
    g_ids, g_rec = mgr.gather_estimates(publisher=n_shards - 1)
    assert g_ids.size == n and np.array_equal(np.sort(g_ids), np.sort(ids))
    owner = g_ids % n_shards
    assert np.all(np.diff(owner) >= 0)                                     # shard-major
    for r in range(n_shards):
        part = g_ids[owner == r]
        assert np.all(np.diff(part.astype(np.int64)) > 0)                  # ascending inside a shard
    pc, tc, _, fc = one.get_estimates_batch(g_ids)
    assert fc.all() and np.array_equal(g_rec[:, :7], pc) and np.array_equal(g_rec[:, 7:], tc)
    assert mgr.last_gather_ms() >= 0.0
    # dense tick: shard-major records (dense_ids order), every shard's slice on its own device; == the same tick through update_batch
    d_ids = mgr.dense_ids()
    assert np.array_equal(d_ids, g_ids)
    row = {int(i): k for k, i in enumerate(ids)}
    sel = np.array([row[int(i)] for i in d_ids])
    pos = np.zeros((n, 3))
    ka, kb = ticks - 2, ticks - 1
    assert mgr.update_dense(DT, meas[ka][sel], action[ka][sel], pos) == n
    one.update_batch(ids, DT, meas[ka], action[ka])
    pd, _, _, _ = one.get_estimates_batch(d_ids)
    assert np.array_equal(pos, pd[:, :3])
    assert mgr.update_dense(DT, meas[kb][sel], action[kb][sel], pos, pipelined=True) == n
    mgr.update_dense_wait(0)
    one.update_batch(ids, DT, meas[kb], action[kb])
    pd, _, _, _ = one.get_estimates_batch(d_ids)
    assert np.array_equal(pos, pd[:, :3])
    # erase through the routed batch call, then the id lists agree again
    gone = ids[::4]
    assert mgr.erase_batch(gone) == gone.size
    assert np.array_equal(mgr.ids(), np.sort(np.setdiff1d(ids, gone)))
    assert not mgr.erase(int(gone[0])) and mgr.erase(int(ids[1]))
    used_nccl = mgr.gather_uses_nccl()
    mgr.close(); one.close(); ref.close()
    return used_nccl


@pytest.mark.parametrize("name", ["uniform_acceleration", "angular_rates"])
def test_sharded_manager_same_device(name):
    """three shards on cuda:0: the host side of the sharded manager on a one-GPU box (exchange by device copies)"""
    assert _run(name, 3, _devices(3, False)) is False


def test_sharded_manager_single_shard():
    assert _run("uniform_velocity", 1, [0], n=1001, ticks=5) is False


@pytest.mark.parametrize("name", ["uniform_acceleration", "angular_velocities"])
def test_sharded_manager_two_gpus_nccl(name):
    """two shards on two GPUs, the exchange over NCCL (skipped on a one-GPU box)"""
    assert _run(name, 2, _devices(2, True)) is True


def test_ragged_gather_two_gpus_nccl():
    """shards of different sizes: the ragged form of the all-gather (one broadcast per shard inside one NCCL group)"""
    from target_estimation_b200.manager import ShardedManagerC
    dev = _devices(2, True)
    mgr = ShardedManagerC(_yaml("uniform_acceleration"), 2, dev)
    ids = np.concatenate([np.arange(0, 4000, 2), np.arange(1, 1201, 2)]).astype(np.uint32)     # 2000 on shard 0, 600 on shard 1
    p0 = np.zeros((ids.size, 7)); p0[:, 0] = ids; p0[:, 6] = 1.0
    assert mgr.init_batch(ids, DT, p0) == ids.size
    mgr.update_batch(ids, DT, p0)
    g_ids, g_rec = mgr.gather_estimates(publisher=1)
    assert np.array_equal(g_ids, ids) and mgr.gather_uses_nccl()
    assert np.allclose(g_rec[:, 0], ids, rtol=0, atol=1e-6)
    mgr.close()
