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
