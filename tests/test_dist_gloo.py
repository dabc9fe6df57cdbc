"""not gpu: world_size-2 gloo run of the multi-GPU host logic -- id sharding (owner = id mod G), per-shard tick
routing, and the variable-length all-gather of estimate records back into global ascending-id order.  Each rank
runs the ORACLE as its stand-in filter (the CUDA pool needs a GPU); what is under test is the plumbing, which is the
same code bench.py and a multi-GPU caller use."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from target_estimation_b200 import sharding
    from tests import orc, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = 1.0 / 250.0
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_uniform_acceleration_params.yaml"))
    n, ticks = 101, 12
    meas, action, scale = synth.make_streams(n, ticks, dt, angular=False, seed=77)
    ids = (np.arange(n, dtype=np.uint32) * 7 + 3)
    mine = sharding.route(ids, world, np.arange(n))[rank]
    my_ids, my_idx = mine
    assert np.all(sharding.owner(my_ids, world) == rank)
    mgr = orc.Manager()
    for i, k in zip(my_ids, my_idx):
        mgr.init_full(y["type"], int(i), dt, 0.0, y["Q"], y["R"], scale[k] * y["P"], meas[0, k])
    for t in range(ticks):
        shard = sharding.route(ids, world, meas[t], action[t])[rank]     # host routes the tick's batch by owner
        mgr.step_batch(shard[0], dt, shard[1], shard[2])
    rec = np.zeros((len(my_ids), 13))
    for j, i in enumerate(my_ids):
        rec[j, :7] = mgr.pose(int(i))[1]
        rec[j, 7:] = mgr.twist(int(i))[1]
    g_ids, g_rec = sharding.all_gather_records(torch.from_numpy(my_ids.astype(np.int64)), torch.from_numpy(rec), dist)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), ids=g_ids.numpy(), rec=g_rec.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_shard_route_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from tests import orc, synth
    dt = 1.0 / 250.0
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_uniform_acceleration_params.yaml"))
    n, ticks = 101, 12
    meas, action, scale = synth.make_streams(n, ticks, dt, angular=False, seed=77)
    ids = (np.arange(n, dtype=np.uint32) * 7 + 3)
    mgr = orc.Manager()      # single-process reference: one manager holding every target
    for k, i in enumerate(ids):
        mgr.init_full(y["type"], int(i), dt, 0.0, y["Q"], y["R"], scale[k] * y["P"], meas[0, k])
    for t in range(ticks):
        mgr.step_batch(ids, dt, meas[t], action[t])
    want = np.zeros((n, 13))
    for k, i in enumerate(ids):
        want[k, :7] = mgr.pose(int(i))[1]
        want[k, 7:] = mgr.twist(int(i))[1]
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["ids"], ids.astype(np.int64))          # globally ascending, nothing lost or duplicated
        assert np.array_equal(got["rec"], want)                          # sharding does not change a single bit


def test_route_and_merge_are_inverse():
    sys.path.insert(0, ROOT)
    from target_estimation_b200 import sharding
    rng = np.random.default_rng(0)
    ids = np.sort(rng.choice(10 ** 6, 5000, replace=False).astype(np.uint32))
    payload = rng.normal(size=(ids.size, 7))
    for world in (1, 2, 4, 8):
        shards = sharding.route(ids, world, payload)
        assert sum(len(s[0]) for s in shards) == ids.size
        for r, s in enumerate(shards):
            assert np.all(s[0] % world == r) and np.all(np.diff(s[0].astype(np.int64)) > 0)
        m_ids, m_payload = sharding.merge_sorted(shards)
        assert np.array_equal(m_ids, ids) and np.array_equal(m_payload, payload)


def _tick_stream(ticks=40, n0=60, seed=4):
    """/tf messages of a churning id set: per tick (ids, stamps[n][2], poses[n][7]) in shuffled arrival order, some ids named
    twice, some stamps stale; plus the clock of every tick"""
    sys.path.insert(0, ROOT)
    rng = np.random.default_rng(seed)
    universe = rng.choice(100000, size=n0 + 4 * ticks, replace=False).astype(np.uint32)
    live = list(range(n0)); nxt = n0
    out = []
    for k in range(ticks):
        now = 1000 * 10 ** 9 + k * 4000000
        gone = set(j for j in live if rng.random() < 0.03)
        live = [j for j in live if j not in gone] + list(range(nxt, nxt + 4)); nxt += 4
        speak = np.array([j for j in live if rng.random() < 0.9])
        rng.shuffle(speak)
        st = now - np.where(rng.random(speak.size) < 0.06, 12000000, 0)
        ids = universe[speak]
        stamps = np.stack([st // 10 ** 9, st % 10 ** 9], axis=1).astype(np.uint32)
        poses = np.hstack([rng.normal(size=(speak.size, 3)), np.tile([0, 0, 0, 1.0], (speak.size, 1))])
        if k % 6 == 2:
            ids = np.concatenate([ids, ids[:3]]); stamps = np.concatenate([stamps, stamps[:3]]); poses = np.concatenate([poses, poses[:3] + 0.25])
        out.append((ids, stamps, poses, (now // 10 ** 9, now % 10 ** 9)))
    return out


def _tick_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from target_estimation_b200 import sharding
    from tests import orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dt = 1.0 / 250.0
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_uniform_acceleration_params.yaml"))
    L = orc.lib()
    N = y["Q"].shape[0]
    h = L.orc_tick_new(y["type"], orc.ptr(orc.colmajor(y["Q"])), N, orc.ptr(orc.colmajor(y["R"])), y["R"].shape[0], orc.ptr(orc.colmajor(y["P"])))
    L.orc_tick_set_expiration(h, 6 * dt)
    erased_log, id_log = [], []
    for ids, stamps, poses, now in _tick_stream():
        r_ids, r_st, r_po = sharding.route(ids, world, stamps, poses)[rank]     # arrival order kept inside the shard
        r_ids = np.ascontiguousarray(r_ids); r_st = np.ascontiguousarray(r_st); r_po = np.ascontiguousarray(r_po)
        if r_ids.size:
            L.orc_tick_callback_ids(h, r_ids.size, orc.ptr(r_ids), orc.ptr(r_st), orc.ptr(r_po))
        er = np.zeros(4096, dtype=np.uint32)
        n_er = L.orc_tick_update(h, dt, now[0], now[1], orc.ptr(er), er.size)
        live = np.zeros(max(L.orc_num_targets(h), 1), dtype=np.uint32)
        n_live = L.orc_get_ids(h, orc.ptr(live), live.size)
        assert np.all(sharding.owner(live[:n_live], world) == rank)
        # the publishing rank needs every shard's erase list and live ids of the tick: variable-length gather, merged ascending
        g_er, _ = sharding.all_gather_records(torch.from_numpy(er[:n_er].astype(np.int64)), torch.zeros((n_er, 1), dtype=torch.float64), dist)
        g_live, _ = sharding.all_gather_records(torch.from_numpy(live[:n_live].astype(np.int64)), torch.zeros((n_live, 1), dtype=torch.float64), dist)
        erased_log.append(g_er.numpy()); id_log.append(g_live.numpy())
    np.savez(os.path.join(out_dir, "tick_rank%d.npz" % rank), n=len(erased_log), **{"er%d" % k: e for k, e in enumerate(erased_log)},
             **{"id%d" % k: e for k, e in enumerate(id_log)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_tick_records_world2(tmp_path):
    """the node loop sharded by id: every /tf message is routed by owner(id) = id mod G (arrival order kept per shard), each
    rank runs its own tick (oracle stand-in for the device mailboxes), and the per-tick erase lists / live ids gathered to
    every rank equal those of ONE manager seeing the whole stream -- no cross-shard state in the mailbox / expiry logic."""
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mp.spawn(_tick_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from tests import orc
    dt = 1.0 / 250.0
    y = orc.load_yaml(os.path.join(ROOT, "models", "model_uniform_acceleration_params.yaml"))
    L = orc.lib()
    N = y["Q"].shape[0]
    h = L.orc_tick_new(y["type"], orc.ptr(orc.colmajor(y["Q"])), N, orc.ptr(orc.colmajor(y["R"])), y["R"].shape[0], orc.ptr(orc.colmajor(y["P"])))
    L.orc_tick_set_expiration(h, 6 * dt)
    got = [np.load(os.path.join(str(tmp_path), "tick_rank%d.npz" % r)) for r in range(world)]
    total = 0
    for k, (ids, stamps, poses, now) in enumerate(_tick_stream()):
        ids = np.ascontiguousarray(ids); stamps = np.ascontiguousarray(stamps); poses = np.ascontiguousarray(poses)
        L.orc_tick_callback_ids(h, ids.size, orc.ptr(ids), orc.ptr(stamps), orc.ptr(poses))
        er = np.zeros(4096, dtype=np.uint32)
        n_er = L.orc_tick_update(h, dt, now[0], now[1], orc.ptr(er), er.size)
        live = np.zeros(max(L.orc_num_targets(h), 1), dtype=np.uint32)
        n_live = L.orc_get_ids(h, orc.ptr(live), live.size)
        total += n_er
        for g in got:
            assert np.array_equal(g["er%d" % k], er[:n_er].astype(np.int64)), k
            assert np.array_equal(g["id%d" % k], live[:n_live].astype(np.int64)), k
    assert total > 20
    L.orc_manager_delete(h)
