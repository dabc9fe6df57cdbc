#!/usr/bin/env python3
"""tools/e2e_probe.py -- where the host-buffer tick (te_pool_tick_host / _async) spends its time: 4 Mi UA targets, the chunk size of
the copy / step / read-back pipeline (TE_TICK_CHUNK_TILES, one process per setting), copies in one direction only, both, pipelined
across ticks or not.  Prints one JSON line per setting."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    import target_estimation_b200 as te
    n = 4 << 20
    mtype, _, Q, R, P0 = te.load_model("uniform_acceleration")
    pool = te.TargetPool(mtype)
    pool.register_class(Q, R, P0)
    pool.reserve(n)
    rng = np.random.default_rng(1)
    for s in range(0, n, 1 << 20):
        p0 = np.zeros((1 << 20, 7)); p0[:, :3] = rng.uniform(-5, 5, (1 << 20, 3)); p0[:, 6] = 1
        pool.add(np.arange(s, s + (1 << 20), dtype=np.uint32), p0)
    h3 = [torch.from_numpy(rng.uniform(-5, 5, (n, 3))).pin_memory() for _ in range(2)]
    act = [torch.full((n,), 2, dtype=torch.uint8).pin_memory() for _ in range(2)]
    out = [torch.empty((n, 3), dtype=torch.float64).pin_memory() for _ in range(3)]
    lag = int(os.environ.get("TE_PROBE_LAG", "2"))   # ticks left in flight by the loop (the pool stages three)
    L = te.lib
    res = {"chunk_tiles": os.environ.get("TE_TICK_CHUNK_TILES", "default"), "lag": lag}

    def run(name, pipelined, with_in, with_out, K=20):
        def one(k):
            fn = L.te_pool_tick_host_async if pipelined else L.te_pool_tick_host
            rc = fn(pool._h, 0.004, h3[k % 2].data_ptr() if with_in else None, 3, act[k % 2].data_ptr() if with_in else None, 2 if with_in else 1,
                    out[k % 3].data_ptr() if with_out else None)
            assert rc == 0, te._lib.last_error()
            if pipelined:
                L.te_pool_tick_host_wait(pool._h, lag)
        for k in range(4):
            one(k)
        L.te_pool_tick_host_wait(pool._h, 0)
        t0 = time.perf_counter()
        for k in range(K):
            one(k)
        L.te_pool_tick_host_wait(pool._h, 0)
        res[name] = round((time.perf_counter() - t0) * 1e3 / K, 3)
    run("sync_in_out", False, True, True)
    run("pipe_in_out", True, True, True)
    run("sync_in_only", False, True, False)
    run("pipe_in_only", True, True, False)
    run("sync_out_only", False, False, True)
    run("pipe_out_only", True, False, True)
    run("sync_none", False, False, False)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for c, lag in (("8192", "2"), ("32768", "1"), ("32768", "2"), ("131072", "1"), ("131072", "2")):
            env = dict(os.environ, TE_TICK_CHUNK_TILES=c, TE_PROBE_LAG=lag)
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
            print(pr.stdout.strip().splitlines()[-1] if pr.stdout.strip() else pr.stderr[-500:], flush=True)
