#!/usr/bin/env python3
"""tools/bench_sharded.py -- ONE process, G GPUs, through the C++ ShardedTargetManager of libtarget_c.so (target_manager_new_sharded):

    python tools/bench_sharded.py [--gpus 1,2,4,8] [--targets-per-gpu 4194304] [--steps 20]

per G: (a) the dense host tick target_manager_update_dense_async / _wait(2) -- shard-major pinned host arrays in, every target's
estimated position out, every shard's slice copied / stepped / read back on its own device, two ticks in flight; (b) the routed
batch call target_manager_update_batch (ids in arbitrary order, routed to the owners on the host by the shard worker threads);
(c) the all-gather of [pose7 | twist6] records between the devices, issued from C++ over NCCL (target_manager_gather_estimates with
no host output).  One JSON line per G."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

DT = 1.0 / 250.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2")
    ap.add_argument("--targets-per-gpu", type=int, default=4 << 20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--model", default="uniform_acceleration")
    args = ap.parse_args()
    import torch
    from target_estimation_b200.manager import ShardedManagerC
    yaml = os.path.join(ROOT, "models", "model_%s_params.yaml" % args.model)
    nd = torch.cuda.device_count()
    for G in [int(g) for g in args.gpus.split(",")]:
        if G > nd:
            print(json.dumps({"n_gpus": G, "skipped": "only %d devices" % nd}), flush=True)
            continue
        n = args.targets_per_gpu * G
        mgr = ShardedManagerC(yaml, G)
        rng = np.random.default_rng(5)
        t0 = time.perf_counter()
        chunk = 1 << 20
        for s in range(0, n, chunk):
            ids = np.arange(s, min(n, s + chunk), dtype=np.uint32)
            p0 = np.zeros((ids.size, 7)); p0[:, :3] = rng.uniform(-5, 5, (ids.size, 3)); p0[:, 6] = 1.0
            assert mgr.init_batch(ids, DT, p0) == ids.size
        t_init = time.perf_counter() - t0
        d_ids = mgr.dense_ids()
        assert d_ids.size == n
        stride = 3
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = [pin(rng.uniform(-5, 5, (n, stride))) for _ in range(2)]
        h_act = [pin(np.where(rng.random(n) < 0.05, 1, 2).astype(np.uint8)) for _ in range(2)]
        h_out = [pin(np.zeros((n, 3))) for _ in range(3)]
        K = args.steps

        def one(k):
            assert mgr.update_dense(DT, h_in[k % 2].numpy(), h_act[k % 2].numpy(), h_out[k % 3].numpy(), pipelined=True) == n
            mgr.update_dense_wait(2)
        for k in range(4):
            one(k)
        mgr.update_dense_wait(0)
        t0 = time.perf_counter()
        for k in range(K):
            one(k)
        mgr.update_dense_wait(0)
        t_dense = (time.perf_counter() - t0) / K
        # routed batch: ids in arbitrary order, pose measurements, from pinned host arrays
        perm = rng.permutation(n).astype(np.uint32)
        b_ids = pin(perm)
        b_meas = pin(np.concatenate([rng.uniform(-5, 5, (n, 3)), np.tile([0.0, 0.0, 0.0, 1.0], (n, 1))], axis=1))
        Kb = max(2, K // 4)
        mgr.update_batch(b_ids.numpy(), DT, b_meas.numpy())
        t0 = time.perf_counter()
        for k in range(Kb):
            assert mgr.update_batch(b_ids.numpy(), DT, b_meas.numpy()) == n
        mgr.flush()
        mgr.get_est_pose(0)
        t_batch = (time.perf_counter() - t0) / Kb
        # the all-gather of estimate records between the devices (nothing comes to the host)
        mgr.gather_estimates(fetch=False)
        ms = []
        for _ in range(5):
            assert mgr.gather_estimates(fetch=False) == n
            ms.append(mgr.last_gather_ms())
        ag = float(np.median(ms))
        out = {"tool": "bench_sharded", "n_gpus": G, "model": args.model, "targets_per_gpu": args.targets_per_gpu, "targets_total": n,
               "init_s": t_init,
               "dense_tick": {"ms_per_step": 1e3 * t_dense, "target_steps_per_s": n / t_dense, "h2d_bytes_per_step": n * (stride * 8 + 1), "d2h_bytes_per_step": n * 24,
                              "api": "target_manager_update_dense_async + _wait(2), shard-major pinned arrays, one process"},
               "routed_batch": {"ms_per_step": 1e3 * t_batch, "target_steps_per_s": n / t_batch,
                                "api": "target_manager_update_batch(ids in arbitrary order, meas[n][7]): host routing by id mod G on the shard worker threads"},
               "allgather": {"ms": ag, "uses_nccl": mgr.gather_uses_nccl(), "bytes_per_rank": args.targets_per_gpu * 104,
                             "bus_gbs": (G - 1) * args.targets_per_gpu * 108 / (ag * 1e-3) / 1e9 if G > 1 and ag > 0 else None,
                             "records": "ids + pose7 | twist6 per target, every device receives all shards"}}
        print(json.dumps(out), flush=True)
        mgr.close()


if __name__ == "__main__":
    main()
